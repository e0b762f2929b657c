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
#define CNT_ADJ_ERR (S3_MAX_ROUNDS - 40)   // View::counters slot of the adjacency builder's overflow flag

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

// ---- data term of params.pms_cost_mode = 1: gradients of both views (pm.cpp:70-88: cvtColor BGR2GRAY, Sobel 3x3 / 8)
__global__ void k_pm_gradients(int W, int H, const uint8_t* __restrict__ bgr, float* __restrict__ grad) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= W * H) return;
    const int x = p % W, y = p / W;
    const int xs[3] = {s3_reflect101(x - 1, W), x, s3_reflect101(x + 1, W)}, ys[3] = {s3_reflect101(y - 1, H), y, s3_reflect101(y + 1, H)};
    int g[3][3];
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const uint8_t* c = bgr + 3 * ((size_t)ys[j] * W + xs[i]);
            g[j][i] = s3_cv_gray(c[0], c[1], c[2]);
        }
    const int gx = (g[0][2] + 2 * g[1][2] + g[2][2]) - (g[0][0] + 2 * g[1][0] + g[2][0]);
    const int gy = (g[2][0] + 2 * g[2][1] + g[2][2]) - (g[0][0] + 2 * g[0][1] + g[0][2]);
    grad[2 * (size_t)p] = (float)gx / 8.f;
    grad[2 * (size_t)p + 1] = (float)gy / 8.f;
}

int s3_prepare_plane_cost(s3dmst_ctx* ctx, int Dmax) {
    const int N = ctx->N;
    for (int view = 0; view < 2; view++) {
        View& V = ctx->v[view];
        if (V.cost_ready && V.D != Dmax) return s3_fail(ctx, S3DMST_E_ARG, "prepare_plane_cost: the view holds a cost volume of %d labels", V.D);
        if (!V.pgrad) S3_CUDA(cudaMalloc(&V.pgrad, sizeof(float) * 2 * (size_t)N));
        k_pm_gradients<<<(N + 255) / 256, 256, 0, ctx->stream>>>(ctx->W, ctx->H, V.bgr, V.pgrad);
        S3_LAUNCH_CHECK();
        if (!V.cost_ready) { V.D = Dmax; V.Dp = (Dmax + 3) / 4 * 4; }
        V.plane_ready = true;
    }
    return 0;
}

// scratch of the proposal kernels: [N][64] doubles + the label list + the offsets
static int pms_scratch(s3dmst_ctx* ctx, size_t n_labels, int T, double** scr, float** d_lab, int** d_off) {
    const size_t scr_bytes = (size_t)ctx->N * 64 * sizeof(double);  // [node][64] (dataflow kernel) / [node][PMS_KB] (simple kernel)
    const size_t lab_bytes = (3 * n_labels * sizeof(float) + 255) / 256 * 256;
    const size_t off_bytes = ((size_t)(T + 1) * sizeof(int) + 255) / 256 * 256;
    if (ctx->pms_scratch_cap < scr_bytes + lab_bytes + off_bytes) {
        const size_t lab_cap = std::max<size_t>(lab_bytes, 1 << 20) * 2;  // head-room: the proposal count changes from round to round
        const size_t need = scr_bytes + lab_cap + off_bytes;
        if (ctx->pms_scratch) S3_CUDA(cudaFree(ctx->pms_scratch));
        ctx->pms_scratch = nullptr; ctx->pms_scratch_cap = 0;
        S3_CUDA(cudaMalloc(&ctx->pms_scratch, need));
        ctx->pms_scratch_cap = need;
    }
    char* basep = (char*)ctx->pms_scratch;
    *scr = (double*)basep;
    *d_lab = (float*)(basep + scr_bytes);
    *d_off = (int*)(basep + ctx->pms_scratch_cap - off_bytes);
    return 0;
}

int s3_pms_apply(s3dmst_ctx* ctx, int view, const int32_t* h_tree_ids, const float* h_labels, size_t n) {
    View& V = ctx->v[view];
    const bool plane = ctx->P.pms_cost_mode == 1;
    if (!V.forest_ready || !(plane ? ctx->v[0].plane_ready && ctx->v[1].plane_ready : V.cost_ready) || !V.labels_ready)
        return s3_fail(ctx, S3DMST_E_STATE, "pms_apply: forest, %s and labels required", plane ? "s3dmst_prepare_plane_cost" : "cost volume");
    if (plane && ctx->P.agg_kernel == 1) return s3_fail(ctx, S3DMST_E_ARG, "pms_apply: the plane cost runs in the dataflow kernel only");
    if (n == 0) return 0;
    S3_TRY(s3_forest_finish_host(ctx));
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
    double* scr; float* d_lab; int* d_off;
    S3_TRY(pms_scratch(ctx, n, T, &scr, &d_lab, &d_off));
    S3_TRY(s3_h2d_staged(ctx, d_lab, lab.data(), 3 * n * sizeof(float)));
    S3_TRY(s3_h2d_staged(ctx, d_off, off.data(), (T + 1) * sizeof(int)));

    if (ctx->P.agg_kernel != 1) {  // default: the dataflow kernel's proposal mode
        PmsFlowPlan plan;
        const int r = s3_pms_flow_plan(ctx, view, off.data(), scr, &plan);
        if (r == 0) return s3_pms_flow_launch(ctx, view, &plan, d_off, d_lab, 0, 0u, 0u);
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
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Tree adjacency graph on the device (tree_g, Stereo3DMST.cpp:377-384: two trees are neighbours iff a grid edge
// joins them; adjacent_vertices iterates a setS, i.e. ascending tree id).  Built once per forest:
//   1. every boundary grid edge inserts its two directed pairs into an open-addressing hash set (duplicates collapse)
//      and a first insertion counts towards the degree of its source tree;
//   2. exclusive scan of the degrees -> CSR offsets;  3. the set's entries are dealt to their rows;
//   4. every row is rank-sorted (keys are unique).
// Scratch: the forest kernel's live-edge lists (idle after the forest is built).
__device__ __forceinline__ void adj_insert(unsigned long long* table, unsigned mask, int a, int b, int* deg, int* err) {
    const unsigned long long key = (((unsigned long long)(unsigned)a << 32) | (unsigned)b) + 1ull;  // 0 = empty slot
    unsigned h = (unsigned)(s3_mix64(key) >> 32) & mask;
    for (int probe = 0; probe < 4096; probe++) {
        const unsigned long long old = atomicCAS(table + h, 0ull, key);
        if (old == 0ull) { atomicAdd(deg + a, 1); return; }
        if (old == key) return;
        h = (h + 1) & mask;
    }
    *err = 1;
}
__global__ void k_adj_insert(int W, int H, const int* __restrict__ tree_id, unsigned long long* table, unsigned mask, int* deg, int* err) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= W * H) return;
    const int x = p % W, y = p / W, t = tree_id[p];
    // a pair repeated along a straight boundary is inserted once per run (the previous pixel of the row / column saw it)
    if (x < W - 1) {
        const int u = tree_id[p + 1];
        if (u != t && !(y > 0 && tree_id[p - W] == t && tree_id[p - W + 1] == u)) { adj_insert(table, mask, t, u, deg, err); adj_insert(table, mask, u, t, deg, err); }
    }
    if (y < H - 1) {
        const int u = tree_id[p + W];
        if (u != t && !(x > 0 && tree_id[p - 1] == t && tree_id[p + W - 1] == u)) { adj_insert(table, mask, t, u, deg, err); adj_insert(table, mask, u, t, deg, err); }
    }
}
// single-CTA exclusive scan of deg[0..T) into ptr[0..T]; also resets deg (reused as the fill cursor)
__global__ void __launch_bounds__(1024) k_adj_scan(int T, int* deg, int* ptr) {
    __shared__ int s_w[32];
    __shared__ int s_run;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_run = 0;
    __syncthreads();
    for (int b = 0; b < T; b += 1024) {
        const int i = b + tid;
        const int v = i < T ? deg[i] : 0;
        int inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) s_w[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int w = s_w[lane], winc = w;
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += u;
            }
            s_w[lane] = winc - w;
        }
        __syncthreads();
        const int run = s_run;
        if (i < T) { ptr[i] = run + s_w[wid] + inc - v; deg[i] = 0; }
        __syncthreads();
        if (tid == 1023) s_run = run + s_w[31] + inc;
        __syncthreads();
    }
    if (tid == 0) ptr[T] = s_run;
}
__global__ void k_adj_fill(unsigned cap, const unsigned long long* __restrict__ table, const int* __restrict__ ptr, int* cursor, int* out) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    const unsigned long long key = table[i];
    if (!key) return;
    const int a = (int)((key - 1) >> 32), b = (int)(unsigned)(key - 1);
    out[ptr[a] + atomicAdd(cursor + a, 1)] = b;
}
// one CTA per row: rank sort (the keys of a row are distinct)
__global__ void __launch_bounds__(128) k_adj_sort(const int* __restrict__ ptr, const int* __restrict__ in, int* __restrict__ out) {
    const int lo = ptr[blockIdx.x], n = ptr[blockIdx.x + 1] - lo;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int x = in[lo + i];
        int r = 0;
        for (int j = 0; j < n; j++) r += in[lo + j] < x;
        out[lo + r] = x;
    }
}

int s3_tree_adjacency(s3dmst_ctx* ctx, int view) {
    View& V = ctx->v[view];
    if (!V.forest_ready) return s3_fail(ctx, S3DMST_E_STATE, "tree adjacency: no forest");
    if (V.adj_ready) return 0;
    S3_TRY(s3_forest_finish_host(ctx));
    const int N = ctx->N, W = ctx->W, H = ctx->H, T = V.T;
    unsigned cap = 1;
    while ((size_t)cap * 2 * sizeof(unsigned long long) <= 32 * (size_t)N && cap < (1u << 30)) cap *= 2;  // largest power of two inside fh_ent[0]
    cap = std::max(cap, 64u);  // (the buffers carry S3_FH_MAX_CTAS * 16 KB of slack: tiny images still fit)
    unsigned long long* table = reinterpret_cast<unsigned long long*>(V.fh_ent[0]);
    int* tmp = reinterpret_cast<int*>(V.fh_ent[1]);          // unsorted rows
    int* deg = V.e_ra;                                        // [T] degrees, then fill cursors
    int* err = V.counters + CNT_ADJ_ERR;
    if (V.adj_ptr_cap < (size_t)T + 1) {
        if (V.adj_ptr) S3_CUDA(cudaFree(V.adj_ptr));
        V.adj_ptr = nullptr; V.adj_ptr_cap = 0;
        S3_CUDA(cudaMalloc(&V.adj_ptr, sizeof(int) * ((size_t)T + 1)));
        V.adj_ptr_cap = (size_t)T + 1;
    }
    S3_CUDA(cudaMemsetAsync(table, 0, (size_t)cap * sizeof(unsigned long long), ctx->stream));
    S3_CUDA(cudaMemsetAsync(deg, 0, sizeof(int) * (size_t)T, ctx->stream));
    S3_CUDA(cudaMemsetAsync(err, 0, sizeof(int), ctx->stream));
    k_adj_insert<<<(N + 255) / 256, 256, 0, ctx->stream>>>(W, H, V.tree_id, table, cap - 1, deg, err);
    S3_LAUNCH_CHECK();
    k_adj_scan<<<1, 1024, 0, ctx->stream>>>(T, deg, V.adj_ptr);
    S3_LAUNCH_CHECK();
    int h[2] = {0, 0};
    S3_CUDA(cudaMemcpyAsync(&h[0], V.adj_ptr + T, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    S3_CUDA(cudaMemcpyAsync(&h[1], err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    S3_CUDA(cudaStreamSynchronize(ctx->stream));  // once per forest: the size of the CSR
    if (h[1]) return s3_fail(ctx, S3DMST_E_LIMIT, "tree adjacency: hash set overflow");
    if ((size_t)h[0] * sizeof(int) > 16 * (size_t)N + 1024) return s3_fail(ctx, S3DMST_E_LIMIT, "tree adjacency: %d entries exceed the scratch", h[0]);
    V.n_adj = h[0];
    if (V.adj_cap < (size_t)std::max(1, V.n_adj)) {
        if (V.adj) S3_CUDA(cudaFree(V.adj));
        V.adj = nullptr; V.adj_cap = 0;
        S3_CUDA(cudaMalloc(&V.adj, sizeof(int) * (size_t)std::max(1, V.n_adj)));
        V.adj_cap = (size_t)std::max(1, V.n_adj);
    }
    k_adj_fill<<<(cap + 255) / 256, 256, 0, ctx->stream>>>(cap, table, V.adj_ptr, deg, tmp);
    S3_LAUNCH_CHECK();
    k_adj_sort<<<T, 128, 0, ctx->stream>>>(V.adj_ptr, tmp, V.adj);
    S3_LAUNCH_CHECK();
    V.adj_ready = true;
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

__global__ void k_fill_f64(int n, double v, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}

int s3_init_labels(s3dmst_ctx* ctx, int view, int Dmax) {
    View& V = ctx->v[view];
    const int W = ctx->W, H = ctx->H;
    if (ctx->N == 0 || Dmax <= 0) return s3_fail(ctx, S3DMST_E_ARG, "init_labels: images and Dmax > 0 required");
    // The stream depends only on (W, H, Dmax) — a fresh default engine per view (:390): both views, and every later
    // call, draw the same labels.  The serial host loop (rejection sampling: the draw count per pixel is data
    // dependent) therefore runs once per size and the result is kept on the device.
    if (!ctx->abc_init || ctx->abc_init_w != W || ctx->abc_init_h != H || ctx->abc_init_d != Dmax) {
        std::vector<float> abc(3 * (size_t)ctx->N);
        std::default_random_engine engine;
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
        if (ctx->abc_init) S3_CUDA(cudaFree(ctx->abc_init));
        ctx->abc_init = nullptr;
        S3_CUDA(cudaMalloc(&ctx->abc_init, abc.size() * sizeof(float)));
        S3_CUDA(cudaMemcpyAsync(ctx->abc_init, abc.data(), abc.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        S3_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->abc_init_w = W; ctx->abc_init_h = H; ctx->abc_init_d = Dmax;
    }
    S3_CUDA(cudaMemcpyAsync(V.abc, ctx->abc_init, 3 * (size_t)ctx->N * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    k_fill_f64<<<(ctx->N + 255) / 256, 256, 0, ctx->stream>>>(ctx->N, DBL_MAX, V.min_cost);  // :820-821
    S3_LAUNCH_CHECK();
    V.labels_ready = true;
    ctx->pms_round[view] = 0;
    return 0;
}

// SPATIAL PROPAGATION proposals of one round (:563-574): for every tree, the label of one random pixel of each neighbouring
// tree (ascending id), read from the labels as they stand when the round starts.  One thread per CSR entry.
__global__ void k_pms_gen_prop(int T, const int* __restrict__ adj_ptr, const int* __restrict__ adj, const int* __restrict__ tree_start,
                               const int* __restrict__ node_pixel, const float* __restrict__ abc, uint32_t seed, uint32_t round, float* __restrict__ labels) {
    const int t = blockIdx.x;
    const int lo = adj_ptr[t], n = adj_ptr[t + 1] - lo;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const int nb = adj[lo + j];
        const int b = tree_start[nb], sz = tree_start[nb + 1] - b;
        const int pix = node_pixel[b + s3_sample_index(s3_rng(seed, round, (uint32_t)t, (uint32_t)j), sz)];
        labels[3 * (size_t)(lo + j)] = abc[3 * (size_t)pix];
        labels[3 * (size_t)(lo + j) + 1] = abc[3 * (size_t)pix + 1];
        labels[3 * (size_t)(lo + j) + 2] = abc[3 * (size_t)pix + 2];
    }
}

// n_iter rounds of MST_PMS (:546-629) with the library's own generator, entirely on the device: per round one small
// kernel gathers the propagation proposals (one per neighbouring tree, ascending id = the label of a random pixel of
// that tree), then the proposal kernel evaluates them tree by tree and, still inside the kernel, generates and
// evaluates the refinement ladder max_d = Dmax/2, /2, ... > refine_floor around a random pixel of the tree itself,
// read AFTER the tree's propagation proposals were applied, as in the reference (:584-595).  No host synchronisation
// between rounds.  Defined deviations (parity for this stage is by proposal injection): propagation
// proposals read the labels as they were when the round started (the serial reference lets tree t see what trees < t
// changed in the same round; its own OpenMP build races on exactly that, Q13); the random numbers come from a
// counter-based generator (hd_math.h: s3_rng) instead of the sequential minstd_rand0 / unseeded std::rand() streams
// (Q6/Q8/Q9), and a "random pixel of a tree" indexes the tree's BFS order instead of its raster order.
int s3_pms_iterate(s3dmst_ctx* ctx, int view, int n_iter, unsigned seed) {
    View& V = ctx->v[view];
    const bool plane = ctx->P.pms_cost_mode == 1;
    if (!V.forest_ready || !(plane ? ctx->v[0].plane_ready && ctx->v[1].plane_ready : V.cost_ready) || !V.labels_ready)
        return s3_fail(ctx, S3DMST_E_STATE, "pms_iterate: forest, %s and labels required", plane ? "s3dmst_prepare_plane_cost" : "cost volume");
    if (!ctx->P.exact) return s3_fail(ctx, S3DMST_E_ARG, "pms_iterate: proposals are evaluated in the exact mode only");
    S3_TRY(s3_forest_finish_host(ctx));
    S3_TRY(s3_tree_adjacency(ctx, view));
    double* scr; float* d_lab; int* d_off;
    S3_TRY(pms_scratch(ctx, (size_t)V.n_adj, V.T, &scr, &d_lab, &d_off));
    PmsFlowPlan plan;
    S3_TRY(s3_pms_flow_plan(ctx, view, nullptr, scr, &plan));
    for (int it = 0; it < n_iter; it++) {
        const uint32_t round = ctx->pms_round[view]++;
        k_pms_gen_prop<<<V.T, 64, 0, ctx->stream>>>(V.T, V.adj_ptr, V.adj, V.tree_start, V.node_pixel, V.abc, seed, round, d_lab);
        S3_LAUNCH_CHECK();
        S3_TRY(s3_pms_flow_launch(ctx, view, &plan, V.adj_ptr, d_lab, 1, seed, round));
    }
    return 0;
}
