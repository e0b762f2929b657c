// stereomatch_b200/csrc/pms.cu — PatchMatch 3D-label proposals over the forest (north-star item 5).
//
// Reference: MSTCostAggregationAndLabelUpdate src/Stereo3DMST.cpp:160-186, called from MST_PMS :546-629
// once per proposal.  Here a whole injected proposal list is evaluated: proposals are grouped by tree
// (order preserved), and one CTA per tree evaluates them KB at a time, the proposal index playing the
// role the label index plays in the dense kernel (scratch layout [node][KB], proposal-minor).
// Work item = (node of the current level, proposal); per-item arithmetic is the reference's:
//   cost  = compute3DLabelCost (:103-118, two gathers from the node's label row + lerp, fp32)
//   up    = (((0 + w_k A[c_k]) + ...) + w_1 A[c_1]) + cost          fp64, children in reverse order
//   down  = w * A[parent] + w2 * A_up
//   update: sequential over the proposals in list order, strict '<'  (:173-185)
// Bytes per node-visit: 2x4 cost gathers + 8 write + 8 read scratch + 8 min_cost.
#include <float.h>

#include <algorithm>
#include <vector>

#include "hd_math.h"
#include "internal.h"

#define PMS_KB 32

struct PmsArgs {
    int W, D, Dp;
    float oob;
    const int* unit_tree;
    const int* tree_start;
    const int* tree_depth;
    const int* lvl_start;
    const NodeUp* node_up;
    const int* parent;
    const uint16_t* pw;
    const int* node_pixel;
    const float* cost;
    const double* lut_w;
    const double* lut_w2;
    const int* prop_off;   // [T+1] into labels, by tree
    const float* labels;   // [n][3] grouped by tree, original order inside a tree
    double* scr;           // [N][PMS_KB]
    double* min_cost;      // [N] pixel order
    float* abc;            // [N][3] pixel order
    int cap;
};

__global__ void __launch_bounds__(256) k_pms(PmsArgs A) {
    extern __shared__ double s_lvl[];  // [2][cap][PMS_KB]
    __shared__ float s_lab[PMS_KB * 3];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int t = A.unit_tree[blockIdx.x];
    const int p_lo = A.prop_off[t], p_hi = A.prop_off[t + 1];
    if (p_lo == p_hi) return;
    const int base = A.tree_start[t], end = A.tree_start[t + 1];
    const int* lvl = A.lvl_start + base + t;
    const int depth = A.tree_depth[t];
    const int cap = A.cap;
    double* cur = s_lvl;
    double* prev = s_lvl + (size_t)cap * PMS_KB;

    for (int b0 = p_lo; b0 < p_hi; b0 += PMS_KB) {
        const int K = min(PMS_KB, p_hi - b0);
        for (int i = tid; i < 3 * K; i += nt) s_lab[i] = A.labels[3 * (size_t)b0 + i];
        __syncthreads();
        // ---- leaf -> root
        for (int L = depth - 1; L >= 0; --L) {
            const int ls = lvl[L], le = lvl[L + 1];
            const int items = (le - ls) * K;
            for (int idx = tid; idx < items; idx += nt) {
                const int i = idx / K, k = idx - i * K;
                const int v = ls + i;
                const NodeUp nu = A.node_up[v];
                double acc = 0.0;
                for (int c = (nu.child_count & 7) - 1; c >= 0; --c) {
                    const int ch = nu.child_begin + c;
                    const int j = ch - le;
                    const uint32_t iw = ((c & 2) ? nu.cw23 : nu.cw01) >> ((c & 1) * 16) & 0xFFFFu;
                    const double val = j < cap ? prev[(size_t)j * PMS_KB + k] : A.scr[(size_t)ch * PMS_KB + k];
                    acc = S3_DADD(acc, S3_DMUL(__ldg(A.lut_w + iw), val));
                }
                const int pix = A.node_pixel[v];
                const float cst = s3_label_cost(A.cost + (size_t)v * A.Dp, s_lab[3 * k], s_lab[3 * k + 1], s_lab[3 * k + 2],
                                                pix % A.W, pix / A.W, A.D, A.oob);
                acc = S3_DADD(acc, (double)cst);
                A.scr[(size_t)v * PMS_KB + k] = acc;
                if (i < cap) cur[(size_t)i * PMS_KB + k] = acc;
            }
            __syncthreads();
            double* tmp = cur; cur = prev; prev = tmp;
        }
        // ---- root -> leaf
        for (int L = 0; L < depth; ++L) {
            const int ls = lvl[L], le = lvl[L + 1];
            const int ps = L > 0 ? lvl[L - 1] : 0;
            const int items = (le - ls) * K;
            for (int idx = tid; idx < items; idx += nt) {
                const int i = idx / K, k = idx - i * K;
                const int v = ls + i;
                double fin = A.scr[(size_t)v * PMS_KB + k];
                if (L > 0) {
                    const int p = A.parent[v];
                    const int j = p - ps;
                    const uint32_t iw = A.pw[v];
                    const double pv = j < cap ? prev[(size_t)j * PMS_KB + k] : A.scr[(size_t)p * PMS_KB + k];
                    fin = S3_DADD(S3_DMUL(__ldg(A.lut_w + iw), pv), S3_DMUL(__ldg(A.lut_w2 + iw), fin));
                    A.scr[(size_t)v * PMS_KB + k] = fin;
                }
                if (i < cap) cur[(size_t)i * PMS_KB + k] = fin;
            }
            __syncthreads();
            double* tmp = cur; cur = prev; prev = tmp;
        }
        // ---- label update, proposals in list order (:173-185)
        for (int v = base + tid; v < end; v += nt) {
            const int pix = A.node_pixel[v];
            double m = A.min_cost[pix];
            int lab = -1;
            for (int k = 0; k < K; k++) {
                const double a = A.scr[(size_t)v * PMS_KB + k];
                if (a < m) { m = a; lab = k; }
            }
            if (lab >= 0) {
                A.min_cost[pix] = m;
                A.abc[3 * (size_t)pix] = s_lab[3 * lab];
                A.abc[3 * (size_t)pix + 1] = s_lab[3 * lab + 1];
                A.abc[3 * (size_t)pix + 2] = s_lab[3 * lab + 2];
            }
        }
        __syncthreads();
    }
}

int s3_pms_apply(s3dmst_ctx* ctx, int view, const int32_t* h_tree_ids, const float* h_labels, size_t n) {
    View& V = ctx->v[view];
    if (!V.forest_ready || !V.cost_ready || !V.labels_ready)
        return s3_fail(ctx, S3DMST_E_STATE, "pms_apply: forest, cost volume and labels required");
    if (n == 0) return 0;
    const int T = V.T;
    std::vector<int> off(T + 1, 0);
    for (size_t i = 0; i < n; i++) {
        if (h_tree_ids[i] < 0 || h_tree_ids[i] >= T) return s3_fail(ctx, S3DMST_E_ARG, "pms_apply: tree id %d out of range", h_tree_ids[i]);
        off[h_tree_ids[i] + 1]++;
    }
    for (int t = 0; t < T; t++) off[t + 1] += off[t];
    std::vector<float> lab(3 * n);
    {
        std::vector<int> cur(off.begin(), off.end() - 1);
        for (size_t i = 0; i < n; i++) {
            const int pos = cur[h_tree_ids[i]]++;
            lab[3 * (size_t)pos] = h_labels[3 * i];
            lab[3 * (size_t)pos + 1] = h_labels[3 * i + 1];
            lab[3 * (size_t)pos + 2] = h_labels[3 * i + 2];
        }
    }
    const size_t scr_bytes = (size_t)ctx->N * 64 * sizeof(double);  // [node][64] (dataflow kernel) / [node][PMS_KB] (simple kernel)
    const size_t lab_bytes = (3 * n * sizeof(float) + 255) / 256 * 256;
    const size_t off_bytes = ((T + 1) * sizeof(int) + 255) / 256 * 256;
    const size_t lab_cap = std::max<size_t>(lab_bytes, 1 << 20) * 2;  // head-room: the proposal count changes from round to round
    const size_t need = scr_bytes + lab_cap + off_bytes;
    if (ctx->pms_scratch_cap < scr_bytes + lab_bytes + off_bytes) {
        if (ctx->pms_scratch) S3_CUDA(cudaFree(ctx->pms_scratch));
        ctx->pms_scratch = nullptr; ctx->pms_scratch_cap = 0;
        S3_CUDA(cudaMalloc(&ctx->pms_scratch, need));
        ctx->pms_scratch_cap = need;
    }
    char* basep = (char*)ctx->pms_scratch;
    double* scr = (double*)basep;
    float* d_lab = (float*)(basep + scr_bytes);
    int* d_off = (int*)(basep + ctx->pms_scratch_cap - off_bytes);
    S3_CUDA(cudaMemcpyAsync(d_lab, lab.data(), 3 * n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    S3_CUDA(cudaMemcpyAsync(d_off, off.data(), (T + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));

    static const bool simple = getenv("S3_PMS_SIMPLE") && atoi(getenv("S3_PMS_SIMPLE")) != 0;
    if (!simple) {
        const int r = s3_pms_apply_flow(ctx, view, off.data(), d_off, d_lab, scr);
        if (r != 1) return r;
    }
    PmsArgs A;
    A.W = ctx->W; A.D = V.D; A.Dp = V.Dp; A.oob = ctx->P.oob_cost;
    A.unit_tree = V.unit_tree; A.tree_start = V.tree_start; A.tree_depth = V.tree_depth; A.lvl_start = V.lvl_start;
    A.node_up = V.node_up; A.parent = V.parent; A.pw = V.pw; A.node_pixel = V.node_pixel; A.cost = V.cost;
    A.lut_w = ctx->lut_w; A.lut_w2 = ctx->lut_w2; A.prop_off = d_off; A.labels = d_lab; A.scr = scr;
    A.min_cost = V.min_cost; A.abc = V.abc;
    A.cap = ctx->P.agg_cache_nodes > 0 ? ctx->P.agg_cache_nodes : 32;
    const size_t smem = 2 * (size_t)A.cap * PMS_KB * sizeof(double);
    S3_CUDA(cudaFuncSetAttribute(k_pms, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    S3_EV_BEGIN(S3DMST_T_PMS, view);
    k_pms<<<T, 256, smem, ctx->stream>>>(A);
    S3_LAUNCH_CHECK();
    S3_EV_END(S3DMST_T_PMS, view);
    S3_CUDA(cudaStreamSynchronize(ctx->stream));  // host staging vectors go out of scope
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Host-side label initialisation and proposal generation (a6, a12).  These are the two spots of the
// path that draw random numbers; the numbers come from the same libstdc++ objects the reference uses,
// so the initial labels are bit-identical to the reference's (src/Stereo3DMST.cpp:390-430: one
// default_random_engine per view behind uniform_real_distribution<float>(0,1), raster order, rejection
// sampling of the normal in the unit disc).
#include <cmath>
#include <functional>
#include <random>

int s3_init_labels(s3dmst_ctx* ctx, int view, int Dmax) {
    View& V = ctx->v[view];
    const int W = ctx->W, H = ctx->H;
    if (ctx->N == 0 || Dmax <= 0) return s3_fail(ctx, S3DMST_E_ARG, "init_labels: images and Dmax > 0 required");
    std::vector<float> abc(3 * (size_t)ctx->N);
    std::default_random_engine engine;  // fresh engine per view (:390): both views draw the same stream
    std::uniform_real_distribution<float> unit(0.0f, 1.0f);
    size_t o = 0;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++, o += 3) {
            const float d = unit(engine) * Dmax;
            float u, v, s;
            do {  // (u, v) uniform in the positive quadrant of the unit disc (Q7)
                u = unit(engine);
                v = unit(engine);
                s = u * u + v * v;
            } while (!(s < 1.0f));
            const float k = std::sqrt(1.0f - u * u - v * v);
            const float nx = 2.0f * u * k, ny = 2.0f * v * k;
            const float nz = std::sqrt(1.0f - nx * nx - ny * ny);
            abc[o] = -nx / nz;
            abc[o + 1] = -ny / nz;
            abc[o + 2] = (nx * x + ny * y + nz * d) / nz;
        }
    S3_CUDA(cudaMemcpyAsync(V.abc, abc.data(), abc.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    std::vector<double> big(ctx->N, DBL_MAX);  // :820-821
    S3_CUDA(cudaMemcpyAsync(V.min_cost, big.data(), big.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    V.labels_ready = true;
    return 0;
}

// n_iter rounds of MST_PMS (:546-629) with the library's own generator.  Per round and per tree (ascending id):
// one proposal per neighbouring tree (ascending id) = the label of a random pixel of that tree, then the
// refinement ladder max_d = Dmax/2, /2, ... > refine_floor around a random pixel of the tree itself.  The "dice"
// stream is a fresh default engine behind U(-1,1) in every round, as in the reference (std::bind copies, Q8).
// Defined deviation (parity for this stage is by proposal injection): proposals of a round read the
// labels as they were when the round started (the serial reference lets tree t see what trees < t changed in
// the same round; its own OpenMP build already races on exactly that, Q13), and the refinement pixel comes from
// a seeded mt19937 instead of the unseeded process-global std::rand() (Q6/Q9).
int s3_pms_iterate(s3dmst_ctx* ctx, int view, int n_iter, unsigned seed, const std::vector<int>& adj_ptr, const std::vector<int>& adj) {
    View& V = ctx->v[view];
    if (!V.forest_ready || !V.cost_ready || !V.labels_ready)
        return s3_fail(ctx, S3DMST_E_STATE, "pms_iterate: forest, cost volume and labels required");
    const int N = ctx->N, W = ctx->W, T = V.T, Dmax = V.D;
    std::vector<int> tid(N);
    S3_CUDA(cudaMemcpyAsync(tid.data(), V.tree_id, sizeof(int) * N, cudaMemcpyDeviceToHost, ctx->stream));
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<int> tstart(T + 1, 0), tpix(N);  // pixels of every tree in raster order (:352-367)
    for (int p = 0; p < N; p++) tstart[tid[p] + 1]++;
    for (int t = 0; t < T; t++) tstart[t + 1] += tstart[t];
    {
        std::vector<int> cur(tstart.begin(), tstart.end() - 1);
        for (int p = 0; p < N; p++) tpix[cur[tid[p]]++] = p;
    }
    std::vector<float> abc(3 * (size_t)N), labels;
    std::vector<int32_t> trees;
    std::mt19937 pick(seed);
    for (int it = 0; it < n_iter; it++) {
        S3_CUDA(cudaMemcpyAsync(abc.data(), V.abc, abc.size() * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        S3_CUDA(cudaStreamSynchronize(ctx->stream));
        std::default_random_engine engine;
        std::uniform_real_distribution<float> sym(-1.0f, 1.0f);
        trees.clear();
        labels.clear();
        auto push = [&](int t, float a, float b, float c) {
            trees.push_back(t);
            labels.push_back(a); labels.push_back(b); labels.push_back(c);
        };
        for (int t = 0; t < T; t++) {
            for (int k = adj_ptr[t]; k < adj_ptr[t + 1]; k++) {
                const int nb = adj[k], sz = tstart[nb + 1] - tstart[nb];
                const int idx = std::min(sz - 1, (int)((sym(engine) + 1.0f) * 0.5f * sz));  // Q11 clamp
                const float* l = &abc[3 * (size_t)tpix[tstart[nb] + idx]];
                push(t, l[0], l[1], l[2]);
            }
            const int sz = tstart[t + 1] - tstart[t];
            const int p = tpix[tstart[t] + (int)(pick() % (unsigned)sz)];
            const float px = (float)(p % W), py = (float)(p / W);
            const float* l = &abc[3 * (size_t)p];
            const float nz = 1.0f / std::sqrt(l[0] * l[0] + l[1] * l[1] + 1.0f);
            const float nx = -l[0] * nz, ny = -l[1] * nz;
            const float d = px * l[0] + py * l[1] + l[2];
            float span_n = 1.0f;
            for (float span_d = 0.5f * Dmax; span_d > ctx->P.refine_floor; span_d *= 0.5f, span_n *= 0.5f) {
                const float rd = d + sym(engine) * span_d;
                if (rd < 0.0f || rd > (float)Dmax) continue;  // draws 1 number, then 3 more only when in range (A11)
                float rx = nx + sym(engine) * span_n;
                float ry = ny + sym(engine) * span_n;
                float rz = nz + sym(engine) * span_n;
                const float inv = 1.0f / std::sqrt(rx * rx + ry * ry + rz * rz);
                rx *= inv; ry *= inv; rz = std::fabs(rz * inv);
                push(t, -rx / rz, -ry / rz, (rx * px + ry * py + rz * rd) / rz);
            }
        }
        S3_TRY(s3_pms_apply(ctx, view, trees.data(), labels.data(), trees.size()));
    }
    return 0;
}
