// stereomatch_b200/csrc/aggregate.cu — two-pass tree-filter aggregation + WTA (north-star items 4, 5-WTA).
//
// Reference (src/Stereo3DMST.cpp):  aggregateCostFromChildren :120-138 (leaf->root),
// aggregateCostFromParent :141-158 (root->leaf), label update :173-185; dense-label mode per
// SURVEY A13 (cost = C[d][p], strict '<', ascending d).
//
// Data layout in HBM: cost and the running sums are node-major, label-minor: row v (BFS node index,
// trees concatenated) holds the Dp labels of that node contiguously, so one warp moves one node's
// labels with fully coalesced 8/16-byte-per-lane accesses.  A lane owns label pairs
//   label(h, e) = d0 + slice*64*HV + h*64 + 2*lane + e      (h < HV, e < 2)
// for the whole kernel, so every value a lane reads back (children's sums on the way up, the
// parent's final value on the way down) was produced for the same labels by a lane of its own CTA.
//
// Work unit = (tree, label slice); one CTA per unit, units sorted by decreasing tree size.  Inside a
// CTA the tree is walked level by level (BFS levels are contiguous node ranges), warps striding over
// the nodes of a level, one __syncthreads() per level.  The previous level's values are handed over
// through shared memory (first `cap` nodes of a level; wider levels fall back to the copy in HBM/L2).
//
// Exact mode arithmetic is FP64 in the reference's association order, without FMA contraction:
//   up:   A[v] = (((0 + w_k A[c_k]) + ... ) + w_1 A[c_1]) + cost(v)     children in reverse BFS order
//   down: A[c] = w_c * A[parent] + w2_c * A_up[c]
// which makes the result bit-identical to the serial host code by construction.
//
// Algorithmic HBM bytes per pixel-label: 4 (cost) + 8 (write A_up) + 8 (read A_up) = 20 in exact mode;
// the roofline in bench.py is quoted against the 12-byte fp32 model of SURVEY §8d.
#include <float.h>

#include <algorithm>

#include "hd_math.h"
#include "internal.h"

struct AggArgs {
    int n_units, n_slices;
    const int* unit_tree;
    const int* tree_start;
    const int* tree_depth;
    const int* lvl_start;
    const NodeUp* node_up;
    const int* parent;
    const uint16_t* pw;
    const int* node_pixel;
    const float* cost;
    double* aup;
    int Dp, d0, d1;
    const double* lut_w;
    const double* lut_w2;
    int cap;
    int keep;
    int N;
    int32_t* disp;   // pixel order (n_slices == 1) or [slice][node] partials
    double* best;
};

template <int HV>
__global__ void __launch_bounds__(512) k_agg_dense(AggArgs A) {
    extern __shared__ double2 s_buf[];  // [2][cap][HV][32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NW = blockDim.x >> 5;
    const int unit = blockIdx.x;
    const int slice = unit % A.n_slices;
    const int t = A.unit_tree[unit / A.n_slices];
    const int base = A.tree_start[t];
    const int* lvl = A.lvl_start + base + t;
    const int depth = A.tree_depth[t];
    const int cap = A.cap;
    double2* cur = s_buf;
    double2* prev = s_buf + (size_t)cap * HV * 32;

    int lab[HV];
    bool act[HV];
#pragma unroll
    for (int h = 0; h < HV; h++) {
        lab[h] = A.d0 + slice * 64 * HV + h * 64 + 2 * lane;
        act[h] = lab[h] < A.d1;  // d1 - d0 may be odd: the pad label is masked in the WTA only
    }
    const size_t Dp = A.Dp;

    // ------------------------------------------------------------------ leaf -> root
    for (int L = depth - 1; L >= 0; --L) {
        const int ls = lvl[L], le = lvl[L + 1];
        for (int i = warp; i < le - ls; i += NW) {
            const int v = ls + i;
            const NodeUp nu = A.node_up[v];
            double2 acc[HV];
#pragma unroll
            for (int h = 0; h < HV; h++) acc[h] = make_double2(0.0, 0.0);
            for (int k = (nu.child_count & 7) - 1; k >= 0; --k) {
                const int ch = nu.child_begin + k;
                const int j = ch - le;
                const uint32_t iw = ((k & 2) ? nu.cw23 : nu.cw01) >> ((k & 1) * 16) & 0xFFFFu;
                const double w = __ldg(A.lut_w + iw);
#pragma unroll
                for (int h = 0; h < HV; h++) {
                    if (!act[h]) continue;
                    double2 cv;
                    if (j < cap)
                        cv = prev[((size_t)j * HV + h) * 32 + lane];
                    else
                        cv = *reinterpret_cast<const double2*>(A.aup + (size_t)ch * Dp + lab[h]);
                    acc[h].x = S3_DADD(acc[h].x, S3_DMUL(w, cv.x));
                    acc[h].y = S3_DADD(acc[h].y, S3_DMUL(w, cv.y));
                }
            }
#pragma unroll
            for (int h = 0; h < HV; h++) {
                if (!act[h]) continue;
                const float2 c = __ldg(reinterpret_cast<const float2*>(A.cost + (size_t)v * Dp + lab[h]));
                acc[h].x = S3_DADD(acc[h].x, (double)c.x);
                acc[h].y = S3_DADD(acc[h].y, (double)c.y);
                *reinterpret_cast<double2*>(A.aup + (size_t)v * Dp + lab[h]) = acc[h];
                if (i < cap) cur[((size_t)i * HV + h) * 32 + lane] = acc[h];
            }
        }
        __syncthreads();
        double2* tmp = cur; cur = prev; prev = tmp;
    }

    // ------------------------------------------------------------------ root -> leaf, WTA folded in
    for (int L = 0; L < depth; ++L) {
        const int ls = lvl[L], le = lvl[L + 1];
        const int ps = L > 0 ? lvl[L - 1] : 0;
        for (int i = warp; i < le - ls; i += NW) {
            const int v = ls + i;
            double2 fin[HV];
            if (L == 0) {
#pragma unroll
                for (int h = 0; h < HV; h++)
                    if (act[h]) fin[h] = *reinterpret_cast<const double2*>(A.aup + (size_t)v * Dp + lab[h]);
            } else {
                const int p = A.parent[v];
                const int j = p - ps;
                const uint32_t iw = A.pw[v];
                const double w = __ldg(A.lut_w + iw), w2 = __ldg(A.lut_w2 + iw);
#pragma unroll
                for (int h = 0; h < HV; h++) {
                    if (!act[h]) continue;
                    double2 pv;
                    if (j < cap)
                        pv = prev[((size_t)j * HV + h) * 32 + lane];
                    else
                        pv = *reinterpret_cast<const double2*>(A.aup + (size_t)p * Dp + lab[h]);
                    const double2 au = *reinterpret_cast<const double2*>(A.aup + (size_t)v * Dp + lab[h]);
                    fin[h].x = S3_DADD(S3_DMUL(w, pv.x), S3_DMUL(w2, au.x));
                    fin[h].y = S3_DADD(S3_DMUL(w, pv.y), S3_DMUL(w2, au.y));
                }
            }
            // hand over to the next level + running arg-min over this lane's labels (ascending)
            double bc = DBL_MAX;  // the oracle's initial best (cost < DBL_MAX is required to win)
            int bd = 0x7fffffff;
#pragma unroll
            for (int h = 0; h < HV; h++) {
                if (!act[h]) continue;
                if (i < cap) cur[((size_t)i * HV + h) * 32 + lane] = fin[h];
                if (i >= cap || A.keep) *reinterpret_cast<double2*>(A.aup + (size_t)v * Dp + lab[h]) = fin[h];
                if (fin[h].x < bc) { bc = fin[h].x; bd = lab[h]; }
                if (lab[h] + 1 < A.d1 && fin[h].y < bc) { bc = fin[h].y; bd = lab[h] + 1; }
            }
            // warp arg-min, ties -> lowest label (strict '<' in ascending order, Stereo3DMST.cpp:177)
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const double oc = __shfl_xor_sync(0xffffffffu, bc, o);
                const int od = __shfl_xor_sync(0xffffffffu, bd, o);
                if (oc < bc || (oc == bc && od < bd)) { bc = oc; bd = od; }
            }
            if (lane == 0) {
                if (A.n_slices == 1) {
                    const int pix = A.node_pixel[v];
                    A.disp[pix] = bd;
                    A.best[pix] = bc;
                } else {
                    A.disp[(size_t)slice * A.N + v] = bd;
                    A.best[(size_t)slice * A.N + v] = bc;
                }
            }
        }
        __syncthreads();
        double2* tmp = cur; cur = prev; prev = tmp;
    }
}

// combine per-slice partial minima (node order) into pixel order
__global__ void k_wta_finish(int N, int n_slices, const int* __restrict__ node_pixel, const int32_t* __restrict__ pdisp,
                             const double* __restrict__ pbest, int32_t* __restrict__ disp, double* __restrict__ best) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N) return;
    double bc = pbest[v];
    int bd = pdisp[v];
    for (int s = 1; s < n_slices; s++) {
        const double c = pbest[(size_t)s * N + v];
        const int d = pdisp[(size_t)s * N + v];
        if (c < bc || (c == bc && d < bd)) { bc = c; bd = d; }
    }
    const int pix = node_pixel[v];
    disp[pix] = bd;
    best[pix] = bc;
}

int s3_aggregate_dense(s3dmst_ctx* ctx, int view, int d0, int d1) {
    View& V = ctx->v[view];
    S3_TRY(s3_forest_finish_host(ctx));
    if (!V.forest_ready || !V.cost_ready) return s3_fail(ctx, S3DMST_E_STATE, "aggregate_dense: forest and cost volume required");
    if (d0 < 0 || d1 > V.D || d0 >= d1 || (d0 & 1)) return s3_fail(ctx, S3DMST_E_ARG, "aggregate_dense: need 0 <= d0 < d1 <= D and d0 even");
    const int nl = d1 - d0;
    const int HV = nl > 64 ? 2 : 1;
    const int SW = 64 * HV;
    const int n_slices = (nl + SW - 1) / SW;
    int threads = ctx->P.agg_threads > 0 ? ctx->P.agg_threads : 256;
    threads = std::max(32, std::min(512, threads / 32 * 32));
    int cap = ctx->P.agg_cache_nodes > 0 ? ctx->P.agg_cache_nodes : 16;
    const size_t smem = 2 * (size_t)cap * HV * 32 * sizeof(double2);

    AggArgs A;
    A.n_units = V.T * n_slices; A.n_slices = n_slices;
    A.unit_tree = V.unit_tree; A.tree_start = V.tree_start; A.tree_depth = V.tree_depth; A.lvl_start = V.lvl_start;
    A.node_up = V.node_up; A.parent = V.parent; A.pw = V.pw; A.node_pixel = V.node_pixel;
    A.cost = V.cost; A.aup = V.aup; A.Dp = V.Dp; A.d0 = d0; A.d1 = d1;
    A.lut_w = ctx->lut_w; A.lut_w2 = ctx->lut_w2; A.cap = cap; A.keep = ctx->P.keep_aggregated; A.N = ctx->N;
    int32_t* pdisp = nullptr;
    double* pbest = nullptr;
    if (n_slices == 1) {
        A.disp = V.disp_i; A.best = V.best;
    } else {
        const size_t need = (size_t)n_slices * ctx->N * (sizeof(double) + sizeof(int32_t));
        if (ctx->pms_scratch_cap < need) {
            if (ctx->pms_scratch) S3_CUDA(cudaFree(ctx->pms_scratch));
            ctx->pms_scratch = nullptr; ctx->pms_scratch_cap = 0;
            S3_CUDA(cudaMalloc(&ctx->pms_scratch, need));
            ctx->pms_scratch_cap = need;
        }
        pbest = (double*)ctx->pms_scratch;
        pdisp = (int32_t*)(pbest + (size_t)n_slices * ctx->N);
        A.disp = pdisp; A.best = pbest;
    }
    S3_EV_BEGIN(S3DMST_T_AGG, view);
    if (HV == 2) {
        S3_CUDA(cudaFuncSetAttribute(k_agg_dense<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_agg_dense<2><<<A.n_units, threads, smem, ctx->stream>>>(A);
    } else {
        S3_CUDA(cudaFuncSetAttribute(k_agg_dense<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_agg_dense<1><<<A.n_units, threads, smem, ctx->stream>>>(A);
    }
    S3_LAUNCH_CHECK();
    if (n_slices > 1) {
        k_wta_finish<<<(ctx->N + 255) / 256, 256, 0, ctx->stream>>>(ctx->N, n_slices, V.node_pixel, pdisp, pbest, V.disp_i, V.best);
        S3_LAUNCH_CHECK();
    }
    S3_EV_END(S3DMST_T_AGG, view);
    V.agg_ready = true;
    V.agg_d0 = d0; V.agg_d1 = d1;
    return 0;
}
