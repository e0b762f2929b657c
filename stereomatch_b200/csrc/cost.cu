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

// one warp per pixel; lane owns labels 4*lane + 128*it .. +3 (16-byte stores, 512 B per warp-store)
__global__ void __launch_bounds__(256) k_cost_adgrad(CostArgs A) {
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int N = A.W * A.H;
    if (warp_global >= N) return;
    const int p = warp_global;
    const int x = p % A.W, rowbase = p - x;
    float* row = A.cost + (size_t)A.pixel_node[p] * A.Dp;
    // the pixel of this view
    const uchar4 me = A.view == 0 ? A.left4[p] : A.right4[p];
    const float me_g = A.view == 0 ? A.lgray[p] : A.rgray[p];
    const bool has_next = x + 1 < A.W;
    const float me_gn = has_next ? (A.view == 0 ? A.lgray[p + 1] : A.rgray[p + 1]) : 0.0f;
    for (int d4 = 4 * lane; d4 < A.Dp; d4 += 128) {
        float out[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int d = d4 + e;
            float c = 3.0f;  // bad_cost, and the value of never-written left entries (Q19)
            if (d < A.D) {
                if (A.view == 0) {
                    // left volume at (d, x): ref = right(x-d), match = left(x); valid iff x-d >= 0 and x+1 < W
                    const int xr = x - d;
                    if (xr >= 0 && has_next) {
                        const uchar4 r = A.right4[rowbase + xr];
                        c = s3_adgrad(r.x, r.y, r.z, A.rgray[rowbase + xr], A.rgray[rowbase + xr + 1], me.x, me.y, me.z,
                                      me_g, me_gn);
                    }
                } else {
                    // right volume at (d, x): ref = right(x), match = left(x+d); valid iff x+d+1 < W
                    const int xl = x + d;
                    if (xl + 1 < A.W) {
                        const uchar4 m = A.left4[rowbase + xl];
                        c = s3_adgrad(me.x, me.y, me.z, me_g, me_gn, m.x, m.y, m.z, A.lgray[rowbase + xl],
                                      A.lgray[rowbase + xl + 1]);
                    }
                }
                if (A.ingest) c = s3_ingest(c, A.cap, A.offset, A.scale);
            } else
                c = 0.0f;  // row padding
            out[e] = c;
        }
        *reinterpret_cast<float4*>(row + d4) = make_float4(out[0], out[1], out[2], out[3]);
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

int s3_ensure_volume(s3dmst_ctx* ctx, int view, int D) {
    View& V = ctx->v[view];
    if (D < 1 || D > 4096) return s3_fail(ctx, S3DMST_E_ARG, "D=%d out of range", D);
    const int Dp = (D + 3) / 4 * 4;
    const size_t need = (size_t)ctx->N * Dp;
    if (V.cost_cap < need) {
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
        const long long threads = (long long)ctx->N * 32;
        k_cost_adgrad<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(A);
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
