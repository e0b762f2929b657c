// stereomatch_b200/csrc/postfilter.cu — the post-filters the reference's author ran around the same outputs (SURVEY §8f
// rank 4):
//  * weighted median of the disparity map on the pixels the left-right check invalidated
//    (weightedMedianFilter, src/PatchMatchStereoGPU.cu:2436-2599): window (2r+1)^2 (r = 10 there), weight of a window
//    pixel = exp(-sqrt(|dR|+|dG|+|dB|) * gamma) against the centre pixel's colour (gamma = 0.1 for 0..255 intensities),
//    entries outside the image carry weight 0 and disparity 0; entries are stably sorted by disparity and the filter
//    returns the first disparity at which the running sum of weight / weight_sum reaches 0.5;
//  * the normalisation factor of the tree filter, 1 / (tree filter of the all-ones volume)
//    (ComputeMSTCostNormFactor / cost_norm_factor, PatchMatchStereoGPU.cu:5333-5429, :5898-5919).
// Defined where the reference is racy: it filters in place, so a thread may read a neighbour another thread has
// already replaced; here every pixel reads the map as it was before the call.  Every floating-point step keeps the
// reference's order (weight_sum in window order, the running sum in sorted order, both sequential fp32), and the
// weights come from a host-computed 766-entry table (s3dmst_wmf_table: glibc expf/sqrtf, not the device's), so the
// result is bit-identical to the CPU restatement in oracle/postfilter_oracle.py.
#include <float.h>
#include <math.h>

#include <vector>

#include "hd_math.h"
#include "internal.h"

#define WMF_MAX_R 15
#define WMF_WARPS 4

__global__ void k_wmf_compact(int N, const uint8_t* __restrict__ mask, int* __restrict__ list, int* __restrict__ count) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool m = p < N && mask[p] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, m);
    if (!bal) return;
    const int lane = threadIdx.x & 31;
    int off = 0;
    if (lane == 0) off = atomicAdd(count, __popc(bal));
    off = __shfl_sync(0xffffffffu, off, 0);
    if (m) list[off + __popc(bal & ((1u << lane) - 1))] = p;
}

// one warp per invalid pixel; n2 = window entries padded to a power of two
__global__ void __launch_bounds__(32 * WMF_WARPS) k_weighted_median(int W, int H, int r, int n2, const int* __restrict__ list, const int* __restrict__ count,
                                                                     const uint8_t* __restrict__ img, const float* __restrict__ tab, const float* __restrict__ din,
                                                                     float* __restrict__ dout) {
    extern __shared__ unsigned long long s_wmf[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned long long* key = s_wmf + (size_t)w * n2;                                           // (orderable disparity << 32) | window index
    float* wt = reinterpret_cast<float*>(s_wmf + (size_t)WMF_WARPS * n2) + (size_t)w * n2;      // weight by window index, later weight / weight_sum
    const int ws = 2 * r + 1, n = ws * ws;
    const int total = *count;
    for (int it = blockIdx.x * WMF_WARPS + w; it < total; it += gridDim.x * WMF_WARPS) {
        const int p = list[it];
        const int x = p % W, y = p / W;
        const int b0 = img[3 * (size_t)p], g0 = img[3 * (size_t)p + 1], r0 = img[3 * (size_t)p + 2];  // the view's image as set_images stored it
        for (int i = lane; i < n2; i += 32) {
            unsigned long long k = ~0ull;  // padding sorts last
            float wgt = 0.0f;
            if (i < n) {
                const int yy = y + i / ws - r, xx = x + i % ws - r;
                float d = 0.0f;
                if (xx >= 0 && xx < W && yy >= 0 && yy < H) {
                    const int q = yy * W + xx;
                    const uint8_t* c = img + 3 * (size_t)q;
                    wgt = tab[abs((int)c[0] - b0) + abs((int)c[1] - g0) + abs((int)c[2] - r0)];
                    d = din[q];
                }
                const unsigned b = __float_as_uint(d);
                k = ((unsigned long long)((b >> 31) ? ~b : (b | 0x80000000u)) << 32) | (unsigned)i;
            }
            key[i] = k;
            wt[i] = wgt;
        }
        __syncwarp();
        float wsum = 0.0f;
        if (lane == 0)
            for (int i = 0; i < n; i++) wsum = S3_FADD(wsum, wt[i]);  // window order; out-of-image entries add exactly 0
        wsum = __shfl_sync(0xffffffffu, wsum, 0);
        // bitonic sort of the keys (distinct: the window index breaks ties, which makes the order the stable one)
        for (int k2 = 2; k2 <= n2; k2 <<= 1)
            for (int j = k2 >> 1; j > 0; j >>= 1) {
                for (int i = lane; i < n2; i += 32) {
                    const int l = i ^ j;
                    if (l > i) {
                        const unsigned long long a = key[i], b = key[l];
                        const bool up = (i & k2) == 0;
                        if ((a > b) == up) { key[i] = b; key[l] = a; }
                    }
                }
                __syncwarp();
            }
        for (int i = lane; i < n; i += 32) wt[i] = wt[i] / wsum;   // IEEE division, as weightedMask[i] / weight_sum
        __syncwarp();
        if (lane == 0) {
            float acc = 0.0f;
            for (int i = 0; i < n; i++) {
                const unsigned long long k = key[i];
                acc = S3_FADD(acc, wt[(unsigned)k & 0xFFFFu]);
                if (acc >= 0.5f) {
                    const unsigned ob = (unsigned)(k >> 32);
                    dout[p] = __uint_as_float((ob >> 31) ? (ob & 0x7fffffffu) : ~ob);
                    break;
                }
            }
        }
        __syncwarp();
    }
}

extern "C" void s3dmst_wmf_table(float gamma, float* tab766) {
    for (int i = 0; i < S3_NUM_W; i++) tab766[i] = expf(-sqrtf((float)i) * gamma);
}

int s3_weighted_median(s3dmst_ctx* ctx, int view, int radius, float gamma, const uint8_t* h_mask) {
    View& V = ctx->v[view];
    const int N = ctx->N;
    if (N == 0) return s3_fail(ctx, S3DMST_E_STATE, "weighted_median: no images");
    if (radius < 1 || radius > WMF_MAX_R) return s3_fail(ctx, S3DMST_E_ARG, "weighted_median: radius must be 1..%d", WMF_MAX_R);
    if (!h_mask && view != 0) return s3_fail(ctx, S3DMST_E_ARG, "weighted_median: the left-right check only masks the left view; pass a mask for the right one");
    const int n = (2 * radius + 1) * (2 * radius + 1);
    int n2 = 32;
    while (n2 < n) n2 *= 2;
    // scratch: weight table, pixel list + counter, snapshot of the map
    const size_t need = 4096 + sizeof(int) * ((size_t)N + 64) + sizeof(float) * (size_t)N;
    if (ctx->pms_scratch_cap < need) {
        if (ctx->pms_scratch) S3_CUDA(cudaFree(ctx->pms_scratch));
        ctx->pms_scratch = nullptr; ctx->pms_scratch_cap = 0;
        S3_CUDA(cudaMalloc(&ctx->pms_scratch, need));
        ctx->pms_scratch_cap = need;
    }
    char* base = reinterpret_cast<char*>(ctx->pms_scratch);
    float* tab = reinterpret_cast<float*>(base);
    int* count = reinterpret_cast<int*>(base + 4096 - 64);
    int* list = reinterpret_cast<int*>(base + 4096);
    float* snap = reinterpret_cast<float*>(base + 4096 + sizeof(int) * ((size_t)N + 64));
    float h_tab[S3_NUM_W];
    s3dmst_wmf_table(gamma, h_tab);
    S3_TRY(s3_h2d_staged(ctx, tab, h_tab, sizeof h_tab));
    if (h_mask) S3_TRY(s3_h2d_staged(ctx, V.lr_mask, h_mask, (size_t)N));
    S3_CUDA(cudaMemsetAsync(count, 0, sizeof(int), ctx->stream));
    S3_CUDA(cudaMemcpyAsync(snap, V.disp_f, sizeof(float) * (size_t)N, cudaMemcpyDeviceToDevice, ctx->stream));
    S3_EV_BEGIN(S3DMST_T_POST, 1);
    k_wmf_compact<<<(N + 255) / 256, 256, 0, ctx->stream>>>(N, V.lr_mask, list, count);
    S3_LAUNCH_CHECK();
    const size_t smem = (size_t)WMF_WARPS * n2 * (sizeof(unsigned long long) + sizeof(float));
    S3_CUDA(cudaFuncSetAttribute(k_weighted_median, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_weighted_median<<<ctx->num_sms * 4, 32 * WMF_WARPS, smem, ctx->stream>>>(ctx->W, ctx->H, radius, n2, list, count, V.bgr, tab, snap, V.disp_f);
    S3_LAUNCH_CHECK();
    S3_EV_END(S3DMST_T_POST, 1);
    return 0;
}

// 1 / (tree filter of the all-ones volume), per pixel (PatchMatchStereoGPU.cu:5333-5429, :5898-5919): the tree filter is
// run on a 4-label volume of ones through the dense kernel, in a scratch pair of volumes, and read back from `best`.
__global__ void k_fill_f32(size_t n, float v, float* out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}
__global__ void k_reciprocal(int n, const double* __restrict__ in, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __ddiv_rn(1.0, in[i]);
}

int s3_norm_factor(s3dmst_ctx* ctx, int view, double* h_out) {
    View& V = ctx->v[view];
    const int N = ctx->N;
    if (!V.forest_ready) return s3_fail(ctx, S3DMST_E_STATE, "norm_factor: no forest");
    if (!ctx->P.exact) return s3_fail(ctx, S3DMST_E_ARG, "norm_factor: exact mode only");
    // swap a 4-label scratch volume into the view, aggregate, swap back
    float* ones = nullptr;
    double* aup = nullptr;
    S3_CUDA(cudaMalloc(&ones, sizeof(float) * 4 * (size_t)N));
    cudaError_t e = cudaMalloc(&aup, sizeof(double) * 4 * (size_t)N);
    if (e != cudaSuccess) { cudaFree(ones); return s3_fail(ctx, S3DMST_E_CUDA, "norm_factor: %s", cudaGetErrorString(e)); }
    k_fill_f32<<<(unsigned)((4 * (size_t)N + 255) / 256), 256, 0, ctx->stream>>>(4 * (size_t)N, 1.0f, ones);
    float* cost0 = V.cost; double* aup0 = V.aup;
    const int D0 = V.D, Dp0 = V.Dp;
    const bool ready0 = V.cost_ready, agg0 = V.agg_ready;
    const int a0 = V.agg_d0, a1 = V.agg_d1;
    V.cost = ones; V.aup = aup; V.D = 4; V.Dp = 4; V.cost_ready = true;
    int rc = s3_aggregate_flow(ctx, 1 << view, 0, 4);
    V.cost = cost0; V.aup = aup0; V.D = D0; V.Dp = Dp0; V.cost_ready = ready0;
    if (rc == 0) {
        k_reciprocal<<<(N + 255) / 256, 256, 0, ctx->stream>>>(N, V.best, aup);   // (aup is free again: reuse it for the output)
        ctx->launches++;
        if (cudaMemcpyAsync(h_out, aup, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = s3_fail(ctx, S3DMST_E_CUDA, "norm_factor: D2H failed");
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ones);
    cudaFree(aup);
    V.agg_ready = false;  // `best` / `disp_i` now hold the scratch run
    (void)agg0; (void)a0; (void)a1;
    return rc;
}
