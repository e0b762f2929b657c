// stereomatch_b200/csrc/cost.cu — matching-cost volume (north-star item 1) and volume (re)layout.
//
//   k_cost_adgrad     truncated colour + gradient absolute difference for every pixel and label
//                     (formula: PatchMatchStereoGPU.cu:1482-1550, restated in oracle orc_cost_adgrad),
//                     written straight into the node-major, label-minor layout the aggregation reads;
//                     the a2 ingest (Stereo3DMST.cpp:785-803) is fused in.
//   k_from_dmajor     external volume float[D][H][W] (mc-cnn left.bin/right.bin, :769-773) -> node-major
//   k_to_dmajor       inverse, for parity dumps
// Bound: HBM writes, 4 B per pixel-label (+ 8 B/pixel of image reads served from L2).
#include "hd_math.h"
#include "internal.h"

struct CostArgs {
    int W, H, D, Dp;
    int view;  // 0: left volume (match = this pixel, ref = right image at x-d); 1: right volume
    const uchar4* left4;
    const uchar4* right4;
    const float* lgray;
    const float* rgray;
    const int* pixel_node;
    float* cost;
    int ingest;
    float cap, offset, scale;
};

// One CTA per 128-pixel row segment (8 warps x 16 pixels); a lane owns labels lane + 32*j of a pixel, so that a warp
// reads 32 consecutive window entries per shared-memory access (no bank conflicts) and stores 128 contiguous bytes.
// Every label of the segment reads the OTHER view's image inside one contiguous window of the same row (left volume:
// right pixels x0-(D-1) .. x0+128; right volume: left pixels x0 .. x0+128+D), so the window (packed BGR + gray) is
// staged into shared memory once — with two bulk copies (TMA, cp.async.bulk + mbarrier) when the window lies inside the
// row and the row pitch keeps it 16-byte aligned, with plain coalesced loads at the image borders — and the inner loop
// touches shared memory only.  The colour term depends on the integer L1 distance alone and saturates at L1 = 21: a
// 22-entry table built with the oracle's exact expression lives in the lanes' registers (one shuffle per label) and
// replaces the int->double->float chain; the L1 distance itself is two SIMD-in-word instructions.
#define CV_TP 128       // pixels per CTA
#define CV_MAXWIN 1216  // window entries: CV_TP + Dp + 8 (Dp <= 1080)
__device__ __forceinline__ uint32_t cv_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int VIEW>
__global__ void __launch_bounds__(256) k_cost_adgrad(CostArgs A) {
    __shared__ __align__(16) uint32_t s_win4[CV_MAXWIN];
    __shared__ __align__(16) float s_wing[CV_MAXWIN];
    __shared__ __align__(8) unsigned long long s_bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = A.W, D = A.D, Dp = A.Dp;
    const int segs = (W + CV_TP - 1) / CV_TP;
    const int y = blockIdx.x / segs, x0 = (blockIdx.x % segs) * CV_TP;
    const int rowbase = y * W;
    const uchar4* other4 = VIEW == 0 ? A.right4 : A.left4;
    const float* otherg = VIEW == 0 ? A.rgray : A.lgray;
    const uchar4* me4 = VIEW == 0 ? A.left4 : A.right4;
    const float* meg = VIEW == 0 ? A.lgray : A.rgray;
    // window [xs, xs + nwin) of the other image's row: every (x, d) of this segment plus the right neighbour
    const int Dr = (D + 3) & ~3;
    const int xs = VIEW == 0 ? x0 - Dr : x0;
    const int nwin = CV_TP + Dr + 4;  // multiple of 4
    const bool bulk = (W & 3) == 0 && xs >= 0 && xs + nwin <= W;
    if (bulk) {
        const uint32_t bar = cv_smem(&s_bar);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            const uint32_t bytes = (uint32_t)nwin * 4u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2u * bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(cv_smem(s_win4)),
                         "l"(other4 + rowbase + xs), "r"(bytes), "r"(bar)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(cv_smem(s_wing)),
                         "l"(otherg + rowbase + xs), "r"(bytes), "r"(bar)
                         : "memory");
        }
    } else {
        for (int i = tid; i < nwin; i += 256) {
            const int x = xs + i;
            const bool in = x >= 0 && x < W;
            const uchar4 o = in ? other4[rowbase + x] : make_uchar4(0, 0, 0, 0);
            s_win4[i] = (uint32_t)o.x | ((uint32_t)o.y << 8) | ((uint32_t)o.z << 16);
            s_wing[i] = in ? otherg[rowbase + x] : 0.0f;
        }
    }
    // colour term of L1 = lane (lanes >= 21 hold the saturated value): 0.11f * min((float)((double)l1 * 0.33333333333), 7.0f)
    float ct_lane = (float)S3_DMUL((double)(float)lane, 0.33333333333);
    ct_lane = ct_lane < 7.0f ? ct_lane : 7.0f;
    ct_lane = S3_FMUL(0.11f, ct_lane);
    __syncthreads();
    if (bulk) {
        const uint32_t bar = cv_smem(&s_bar);
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.b32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar) : "memory");
    }
#pragma unroll 1
    for (int k = 0; k < CV_TP / 8; k++) {
        const int x = x0 + warp * (CV_TP / 8) + k;
        if (x >= W) break;
        const int p = rowbase + x;
        float* row = A.cost + (size_t)A.pixel_node[p] * Dp;
        const uchar4 m4 = me4[p];
        const uint32_t me = (uint32_t)m4.x | ((uint32_t)m4.y << 8) | ((uint32_t)m4.z << 16);
        const float me_g = meg[p];
        const bool has_next = x + 1 < W;
        const float me_gn = has_next ? meg[p + 1] : 0.0f;
        // labels with a defined cost: left volume 0 <= d <= x (and x+1 < W); right volume d <= W-2-x
        const int dvalid = VIEW == 0 ? (has_next ? x + 1 : 0) : W - 1 - x;
        for (int d0 = 0; d0 < Dp; d0 += 32) {  // warp-uniform trip count: the colour-term shuffle needs every lane
            const int d = d0 + lane;
            // left volume at (d, x): ref = right(x-d), match = left(x); right volume at (d, x): ref = right(x), match = left(x+d)
            const bool valid = d < D && d < dvalid;
            const int wi = valid ? (VIEW == 0 ? x - d : x + d) - xs : 0;
            const uint32_t o = s_win4[wi];
            const float og = s_wing[wi], ogn = s_wing[wi + 1];
            const int l1 = (int)__dp4a(__vabsdiffu4(o, me), 0x00010101u, 0u);
            const float ctv = __shfl_sync(0xffffffffu, ct_lane, min(l1, 21));
            // g = (match_gray - ref_gray) + (ref_gray_next - match_gray_next)
            const float g = VIEW == 0 ? S3_FADD(S3_FSUB(me_g, og), S3_FSUB(ogn, me_gn)) : S3_FADD(S3_FSUB(og, me_g), S3_FSUB(me_gn, ogn));
            const float ag = fabsf(g);
            const float gterm = ag < 2.0f ? ag : 2.0f;
            float c = 3.0f;  // bad_cost, and the value of never-written left entries (Q19)
            if (valid) c = S3_FADD(ctv, S3_FMUL(0.89f, gterm));
            if (A.ingest) c = s3_ingest(c, A.cap, A.offset, A.scale);
            if (d >= D) c = 0.0f;  // row padding
            if (d < Dp)
                row[d] = c;
        }
    }
}

// [D][N] (pixel order) -> [node][Dp]; 32x32 tile through shared memory, both sides coalesced
__global__ void __launch_bounds__(256) k_from_dmajor(const float* __restrict__ in, int N, int D, int Dp,
                                                     const int* __restrict__ pixel_node, float* __restrict__ out,
                                                     int ingest, float cap, float offset, float scale) {
    __shared__ float tile[32][33];
    const int p0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int d = d0 + r, p = p0 + tx;
        float v = 0.0f;
        if (d < D && p < N) {
            v = in[(size_t)d * N + p];
            if (ingest) v = s3_ingest(v, cap, offset, scale);
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int p = p0 + r, d = d0 + tx;
        if (p < N && d < Dp) out[(size_t)pixel_node[p] * Dp + d] = tile[tx][r];
    }
}

__global__ void __launch_bounds__(256) k_to_dmajor(const float* __restrict__ in, int N, int D, int Dp,
                                                   const int* __restrict__ pixel_node, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int p0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int p = p0 + r, d = d0 + tx;
        tile[r][tx] = (p < N && d < D) ? in[(size_t)pixel_node[p] * Dp + d] : 0.0f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int d = d0 + r, p = p0 + tx;
        if (d < D && p < N) out[(size_t)d * N + p] = tile[tx][r];
    }
}

int s3_ensure_volume(s3dmst_ctx* ctx, int view, int D, bool need_cost) {
    View& V = ctx->v[view];
    if (D < 1 || D > 4096) return s3_fail(ctx, S3DMST_E_ARG, "D=%d out of range", D);
    const int Dp = (D + 3) / 4 * 4;
    const size_t need = (size_t)ctx->N * Dp;
    if (need_cost && V.cost_cap < need) {
        if (V.cost) S3_CUDA(cudaFree(V.cost));
        V.cost = nullptr; V.cost_cap = 0;
        S3_CUDA(cudaMalloc(&V.cost, need * sizeof(float)));
        V.cost_cap = need;
    }
    if (V.aup_cap < need) {
        if (V.aup) S3_CUDA(cudaFree(V.aup));
        V.aup = nullptr; V.aup_cap = 0;
        S3_CUDA(cudaMalloc(&V.aup, need * sizeof(double)));
        V.aup_cap = need;
    }
    V.D = D;
    V.Dp = Dp;
    return 0;
}

// Dense runs (s3dmst_run_dense / _batch / aggregate_dense_sharded) on the library's own AD+gradient cost do not need the
// volume at all: the dataflow kernel computes a node's matching cost where it would have read the row (aggregate3.cu:
// FUSE) — 8 of the 24.75 bytes per pixel·label of a step (4 written here, 4 read there) and the whole kernel below
// disappear.  params.fuse_cost = -1 (or S3_FUSE=0 in the environment) keeps the volume.
bool s3_want_fused_cost(const s3dmst_ctx* ctx) {
    static const int env = getenv("S3_FUSE") ? atoi(getenv("S3_FUSE")) : 1;
    return env != 0 && ctx->P.fuse_cost >= 0 && ctx->P.exact != 0 && ctx->P.agg_kernel == 0;
}
int s3_fused_prepare(s3dmst_ctx* ctx, int D) {
    for (int view = 0; view < 2; view++) {
        if (!ctx->v[view].forest_ready) return s3_fail(ctx, S3DMST_E_STATE, "dense run: both forests required");
        S3_TRY(s3_ensure_volume(ctx, view, D, false));
        ctx->v[view].cost_ready = false;  // whatever volume the view held belongs to other images or another D
    }
    ctx->fused_D = D;
    return 0;
}
int s3_materialize_cost(s3dmst_ctx* ctx) {
    if (ctx->fused_D <= 0 || (ctx->v[0].cost_ready && ctx->v[1].cost_ready)) return 0;
    const int D = ctx->fused_D;
    ctx->fused_D = 0;
    return s3_cost_adgrad(ctx, D, 0);
}

int s3_cost_adgrad(s3dmst_ctx* ctx, int D, int apply_ingest) {
    for (int view = 0; view < 2; view++)
        if (!ctx->v[view].forest_ready) return s3_fail(ctx, S3DMST_E_STATE, "build_cost_volume: both forests required");
    S3_EV_BEGIN(S3DMST_T_COST, 0);
    for (int view = 0; view < 2; view++) {
        S3_TRY(s3_ensure_volume(ctx, view, D));
        View& V = ctx->v[view];
        CostArgs A;
        A.W = ctx->W; A.H = ctx->H; A.D = D; A.Dp = V.Dp; A.view = view;
        A.left4 = ctx->v[0].raw4; A.right4 = ctx->v[1].raw4; A.lgray = ctx->v[0].gray; A.rgray = ctx->v[1].gray;
        A.pixel_node = V.pixel_node; A.cost = V.cost;
        A.ingest = apply_ingest; A.cap = ctx->P.cost_cap; A.offset = ctx->P.cost_offset; A.scale = ctx->P.cost_scale;
        if (V.Dp + CV_TP + 8 > CV_MAXWIN) return s3_fail(ctx, S3DMST_E_ARG, "build_cost_volume: D up to %d", CV_MAXWIN - CV_TP - 8);
        const int segs = (ctx->W + CV_TP - 1) / CV_TP;
        if (view == 0) k_cost_adgrad<0><<<(unsigned)(segs * ctx->H), 256, 0, ctx->stream>>>(A);
        else k_cost_adgrad<1><<<(unsigned)(segs * ctx->H), 256, 0, ctx->stream>>>(A);
        S3_LAUNCH_CHECK();
        V.cost_ready = true;
        V.agg_ready = false;
    }
    S3_EV_END(S3DMST_T_COST, 0);
    return 0;
}

int s3_cost_from_dmajor(s3dmst_ctx* ctx, int view, const float* dev_dmajor, int D, int apply_ingest) {
    View& V = ctx->v[view];
    if (!V.forest_ready) return s3_fail(ctx, S3DMST_E_STATE, "set_cost_volume: forest required");
    S3_TRY(s3_ensure_volume(ctx, view, D));
    dim3 grid((ctx->N + 31) / 32, (V.Dp + 31) / 32);
    k_from_dmajor<<<grid, 256, 0, ctx->stream>>>(dev_dmajor, ctx->N, D, V.Dp, V.pixel_node, V.cost, apply_ingest,
                                                 ctx->P.cost_cap, ctx->P.cost_offset, ctx->P.cost_scale);
    S3_LAUNCH_CHECK();
    V.cost_ready = true;
    V.agg_ready = false;
    return 0;
}

int s3_cost_to_dmajor(s3dmst_ctx* ctx, int view, float* dev_dmajor) {
    View& V = ctx->v[view];
    if (!V.cost_ready) return s3_fail(ctx, S3DMST_E_STATE, "get_cost_volume: no volume");
    dim3 grid((ctx->N + 31) / 32, (V.D + 31) / 32);
    k_to_dmajor<<<grid, 256, 0, ctx->stream>>>(V.cost, ctx->N, V.D, V.Dp, V.pixel_node, dev_dmajor);
    S3_LAUNCH_CHECK();
    return 0;
}
