// stereomatch_b200/csrc/rectify.cu — rectification front-end.
//
// The reference's only caller rectifies the raw camera pair before it calls the path (src/stereo_Yin.cpp:139-144):
//     initUndistortRectifyMap(M, D, R, P, size, CV_16SC2, map1, map2);   // once per camera
//     remap(img, imgr, map1, map2, INTER_LINEAR);                        // per frame
// Here the per-frame half runs on the GPU, straight from the uploaded raw pair into the context's image buffers
// (no rectified image ever crosses PCIe).  The maps stay on the device; they are the caller's CV_16SC2 / CV_16UC1
// pair, unchanged.  Arithmetic = OpenCV's fixed-point bilinear remap (imgproc, not under /root/reference; restated
// in oracle/remap_oracle.py and pinned against cv2.remap): source corner (sx, sy) from map1, fractions in 1/32 from
// map2, four int16 weights per fraction pair that sum to 32768, value = (sum w * tap + 16384) >> 15, taps outside
// the source = 0 (BORDER_CONSTANT), footprints completely outside = 0.  Bit-identical results are asserted in
// tests/test_gpu_parity.py::test_remap_matches_opencv.
// Bound: HBM, 4 taps x 3 B (L2-resident) + 6 B of map + 3 B written per pixel — microseconds per frame.
#include <math.h>

#include <algorithm>
#include <vector>

#include "internal.h"

#define RM_TAB 32  // INTER_TAB_SIZE

// The weight table as imgwarp.cpp builds it (initInterTab2D, fixed point): products of the two 1-D linear kernels in
// float, rounded to int16 with saturation; if the four do not sum to 32768 the difference goes to the largest (or
// smallest) weight found by a scan that — for this 2x2 kernel — looks at elements [3, 4, 5, 6] counted from the
// entry's start, i.e. past the entry's end into entries that are still zero at that moment.  The only entries the
// fix-up touches are the ones holding a saturated 32767 (a fraction pair with a weight of exactly 1).
static void build_remap_tab(std::vector<int16_t>& out) {
    const int n = RM_TAB * RM_TAB * 4;
    std::vector<int> flat(n + 8, 0);
    const float scale = 1.0f / RM_TAB;
    float t1[RM_TAB][2];
    for (int i = 0; i < RM_TAB; i++) {
        const float x = (float)i * scale;
        t1[i][0] = 1.0f - x;
        t1[i][1] = x;
    }
    for (int i = 0; i < RM_TAB; i++)
        for (int j = 0; j < RM_TAB; j++) {
            int* e = flat.data() + (i * RM_TAB + j) * 4;
            int isum = 0;
            for (int k1 = 0; k1 < 2; k1++)
                for (int k2 = 0; k2 < 2; k2++) {
                    const float v = t1[i][k1] * t1[j][k2];
                    int iv = (int)lrint((double)v * 32768.0);
                    iv = std::max(-32768, std::min(32767, iv));
                    e[k1 * 2 + k2] = iv;
                    isum += iv;
                }
            if (isum != 32768) {
                const int diff = isum - 32768;
                int mk = 3, Mk = 3;
                for (int k1 = 1; k1 < 3; k1++)
                    for (int k2 = 1; k2 < 3; k2++) {
                        const int idx = k1 * 2 + k2;
                        if (e[idx] < e[mk]) mk = idx;
                        else if (e[idx] > e[Mk]) Mk = idx;
                    }
                int& tgt = diff < 0 ? e[Mk] : e[mk];
                tgt = (int16_t)(tgt - diff);
            }
        }
    out.resize(n);
    for (int i = 0; i < n; i++) out[i] = (int16_t)flat[i];
}

void s3_remap_table(int16_t* tab) {
    std::vector<int16_t> t;
    build_remap_tab(t);
    std::copy(t.begin(), t.end(), tab);
}

struct RemapArgs {
    const uint8_t* src;       // [Hs][Ws][3]
    int Ws, Hs;
    const short2* map_xy;     // [N]
    const uint16_t* map_fxy;  // [N]
    const short4* tab;        // [1024]
    int N;
    uint8_t* dst;             // [N][3]
};

__global__ void __launch_bounds__(256) k_remap_bilinear(RemapArgs A0, RemapArgs A1) {
    const RemapArgs& A = blockIdx.y ? A1 : A0;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= A.N) return;
    const short2 xy = A.map_xy[p];
    const int sx = xy.x, sy = xy.y;
    const short4 w = __ldg(A.tab + (A.map_fxy[p] & (RM_TAB * RM_TAB - 1)));
    int acc[3] = {0, 0, 0};
    const bool outside = sx >= A.Ws || sx + 1 < 0 || sy >= A.Hs || sy + 1 < 0;
    if (!outside) {
        const int wk[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int xx = sx + (k & 1), yy = sy + (k >> 1);
            if (xx >= 0 && xx < A.Ws && yy >= 0 && yy < A.Hs) {
                const uint8_t* s = A.src + 3 * ((size_t)yy * A.Ws + xx);
                acc[0] += wk[k] * (int)s[0];
                acc[1] += wk[k] * (int)s[1];
                acc[2] += wk[k] * (int)s[2];
            }
        }
    }
    uint8_t* d = A.dst + 3 * (size_t)p;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const int v = (acc[c] + (1 << 14)) >> 15;
        d[c] = (uint8_t)min(max(v, 0), 255);
    }
}

int s3_set_rectify_maps(s3dmst_ctx* ctx, int view, const int16_t* map_xy, const uint16_t* map_fxy, int W, int H) {
    const size_t n = (size_t)W * H;
    if (ctx->map_w[view] != W || ctx->map_h[view] != H) {
        if (ctx->map_xy[view]) S3_CUDA(cudaFree(ctx->map_xy[view]));
        if (ctx->map_fxy[view]) S3_CUDA(cudaFree(ctx->map_fxy[view]));
        ctx->map_xy[view] = nullptr; ctx->map_fxy[view] = nullptr; ctx->map_w[view] = ctx->map_h[view] = 0;
        S3_CUDA(cudaMalloc(&ctx->map_xy[view], n * 2 * sizeof(int16_t)));
        S3_CUDA(cudaMalloc(&ctx->map_fxy[view], n * sizeof(uint16_t)));
        ctx->map_w[view] = W; ctx->map_h[view] = H;
    }
    if (!ctx->remap_tab) {
        std::vector<int16_t> tab;
        build_remap_tab(tab);
        S3_CUDA(cudaMalloc(&ctx->remap_tab, tab.size() * sizeof(int16_t)));
        S3_CUDA(cudaMemcpyAsync(ctx->remap_tab, tab.data(), tab.size() * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    S3_CUDA(cudaMemcpyAsync(ctx->map_xy[view], map_xy, n * 2 * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
    S3_CUDA(cudaMemcpyAsync(ctx->map_fxy[view], map_fxy, n * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
    S3_CUDA(cudaStreamSynchronize(ctx->stream));  // the caller's (and the table's) host buffers may go away
    return 0;
}

// both raw images -> the context's two image buffers (ctx->W x ctx->H == the maps' size), one launch
int s3_remap_raw_pair(s3dmst_ctx* ctx, const uint8_t* left_raw, const uint8_t* right_raw, int sw, int sh, int stride) {
    const size_t img = 3 * (size_t)sw * sh;
    if (ctx->raw_stage_cap < 2 * img) {
        if (ctx->raw_stage) S3_CUDA(cudaFree(ctx->raw_stage));
        ctx->raw_stage = nullptr; ctx->raw_stage_cap = 0;
        S3_CUDA(cudaMalloc(&ctx->raw_stage, 2 * img));
        ctx->raw_stage_cap = 2 * img;
    }
    const uint8_t* src[2] = {left_raw, right_raw};
    RemapArgs A[2];
    for (int i = 0; i < 2; i++) {
        S3_CUDA(cudaMemcpy2DAsync(ctx->raw_stage + i * img, 3 * (size_t)sw, src[i], stride, 3 * (size_t)sw, sh, cudaMemcpyHostToDevice, ctx->stream));
        A[i].src = ctx->raw_stage + i * img; A[i].Ws = sw; A[i].Hs = sh;
        A[i].map_xy = reinterpret_cast<const short2*>(ctx->map_xy[i]); A[i].map_fxy = ctx->map_fxy[i];
        A[i].tab = reinterpret_cast<const short4*>(ctx->remap_tab); A[i].N = ctx->N; A[i].dst = ctx->v[i].bgr;
    }
    k_remap_bilinear<<<dim3((unsigned)((ctx->N + 255) / 256), 2), 256, 0, ctx->stream>>>(A[0], A[1]);
    S3_LAUNCH_CHECK();
    return 0;
}

void s3_rectify_free(s3dmst_ctx* ctx) {
    for (int i = 0; i < 2; i++) {
        if (ctx->map_xy[i]) cudaFree(ctx->map_xy[i]);
        if (ctx->map_fxy[i]) cudaFree(ctx->map_fxy[i]);
        ctx->map_xy[i] = nullptr; ctx->map_fxy[i] = nullptr; ctx->map_w[i] = ctx->map_h[i] = 0;
    }
    if (ctx->remap_tab) cudaFree(ctx->remap_tab);
    if (ctx->raw_stage) cudaFree(ctx->raw_stage);
    ctx->remap_tab = nullptr; ctx->raw_stage = nullptr; ctx->raw_stage_cap = 0;
}
