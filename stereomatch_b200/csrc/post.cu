// stereomatch_b200/csrc/post.cu — label -> disparity, left-right consistency check, scan-line fill,
// and the helper of the label-sharded min-loc reduction.
//
// Reference: LabelToDisp src/Stereo3DMST.cpp:189-201 (+ the *(Dmax-1) at :900-902),
// leftRightConsistencyCheck :632-710.
#include <float.h>

#include "hd_math.h"
#include "internal.h"

__global__ void k_label_to_disp(int W, int N, int max_disp, const float* __restrict__ abc, float* __restrict__ disp) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    disp[p] = s3_label_disp(abc[3 * p], abc[3 * p + 1], abc[3 * p + 2], p % W, p / W, max_disp);
}

__global__ void k_int_to_disp(int N, const int32_t* __restrict__ di, float* __restrict__ disp) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < N) disp[p] = (float)di[p];
}

// pass 1 (:639-664): in place on left; right is read only, so there is no cross-thread hazard
__global__ void k_lr_pass1(int W, int N, int max_disp, float* __restrict__ left, const float* __restrict__ right,
                           uint8_t* __restrict__ mask) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const int x = p % W;
    const float d_f = left[p];
    const float r = roundf(d_f);  // std::round: half away from zero
    bool ok = false;
    if (r >= -2147483648.0f && r < 2147483648.0f) {  // (int) of NaN/huge is INT_MIN on x86 => invalid
        const int d = (int)r;
        if (x - d >= 0 && d >= 0 && d < max_disp) ok = !(fabsf(d_f - right[p - d]) > 1.0f);
    }
    mask[p] = ok ? 0 : 1;
    if (!ok) left[p] = 0.0f;
}

// pass 2 (:668-709).  The reference scans each row left to right; for a run of invalid pixels between
// valid A (left) and valid B (right) every pixel of the run ends up min(A, B) (A if there is no B, B if
// there is no A, 0 if neither): the filled left neighbour equals min(A,B) already, so the sequential
// dependence collapses to "nearest originally-valid pixel on each side".  One CTA per row; nearest-valid
// indices via a max-scan / reverse min-scan over the row held in shared memory.
__global__ void __launch_bounds__(256) k_lr_fill(int W, float* __restrict__ left, const uint8_t* __restrict__ mask) {
    extern __shared__ int s_near[];  // [2][W]: nearest valid index to the left / right
    int* s_l = s_near;
    int* s_r = s_near + W;
    __shared__ int s_carry[256];
    float* row = left + (size_t)blockIdx.x * W;
    const uint8_t* m = mask + (size_t)blockIdx.x * W;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int chunk = (W + nt - 1) / nt;
    const int x0 = tid * chunk, x1 = min(W, x0 + chunk);
    // forward: last valid index <= x
    int last = -1;
    for (int x = x0; x < x1; x++) {
        if (m[x] == 0) last = x;
        s_l[x] = last;
    }
    s_carry[tid] = last;
    __syncthreads();
    if (tid == 0) {
        int run = -1;
        for (int i = 0; i < nt; i++) {
            const int v = s_carry[i];
            s_carry[i] = run;  // best index strictly before this thread's chunk
            if (v >= 0) run = v;
        }
    }
    __syncthreads();
    {
        const int carry = s_carry[tid];
        for (int x = x0; x < x1; x++)
            if (s_l[x] < 0) s_l[x] = carry;
    }
    __syncthreads();
    // backward: first valid index >= x
    int nxt = -1;
    for (int x = x1 - 1; x >= x0; x--) {
        if (m[x] == 0) nxt = x;
        s_r[x] = nxt;
    }
    s_carry[tid] = nxt;
    __syncthreads();
    if (tid == 0) {
        int run = -1;
        for (int i = nt - 1; i >= 0; i--) {
            const int v = s_carry[i];
            s_carry[i] = run;
            if (v >= 0) run = v;
        }
    }
    __syncthreads();
    {
        const int carry = s_carry[tid];
        for (int x = x0; x < x1; x++)
            if (s_r[x] < 0) s_r[x] = carry;
    }
    __syncthreads();
    // valid pixels keep their value; read all sources before anyone writes
    float outv[32];
    int cnt = 0;
    for (int x = x0; x < x1 && cnt < 32; x++, cnt++) {
        float v = row[x];
        if (m[x]) {
            const int li = s_l[x], ri = s_r[x];
            if (li >= 0) {
                v = row[li];
                if (ri >= 0 && row[ri] < v) v = row[ri];
            } else if (ri >= 0)
                v = row[ri];
        }
        outv[cnt] = v;
    }
    __syncthreads();
    cnt = 0;
    for (int x = x0; x < x1 && cnt < 32; x++, cnt++) row[x] = outv[cnt];
}

int s3_label_to_disp(s3dmst_ctx* ctx, int view) {
    View& V = ctx->v[view];
    if (!V.labels_ready || V.D <= 0) return s3_fail(ctx, S3DMST_E_STATE, "label_to_disp: labels and a cost volume (Dmax) required");
    k_label_to_disp<<<(ctx->N + 255) / 256, 256, 0, ctx->stream>>>(ctx->W, ctx->N, V.D, V.abc, V.disp_f);
    S3_LAUNCH_CHECK();
    return 0;
}

int s3_dense_to_disp(s3dmst_ctx* ctx, int view) {
    View& V = ctx->v[view];
    if (!V.agg_ready) return s3_fail(ctx, S3DMST_E_STATE, "dense_to_disparity: aggregate_dense first");
    k_int_to_disp<<<(ctx->N + 255) / 256, 256, 0, ctx->stream>>>(ctx->N, V.disp_i, V.disp_f);
    S3_LAUNCH_CHECK();
    return 0;
}

int s3_lr_check(s3dmst_ctx* ctx, int fill) {
    View& L = ctx->v[0];
    View& R = ctx->v[1];
    const int D = L.D > 0 ? L.D : R.D;
    if (D <= 0) return s3_fail(ctx, S3DMST_E_STATE, "lr_check: Dmax unknown (no cost volume set)");
    if (ctx->W > 256 * 32) return s3_fail(ctx, S3DMST_E_ARG, "lr_check: W > 8192 unsupported");
    S3_EV_BEGIN(S3DMST_T_POST, 0);
    k_lr_pass1<<<(ctx->N + 255) / 256, 256, 0, ctx->stream>>>(ctx->W, ctx->N, D, L.disp_f, R.disp_f, L.lr_mask);
    S3_LAUNCH_CHECK();
    if (fill) {
        const size_t smem = 2 * (size_t)ctx->W * sizeof(int);  // up to 64 KB at W = 8192: above the 48 KB default limit
        S3_CUDA(cudaFuncSetAttribute(k_lr_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_lr_fill<<<ctx->H, 256, smem, ctx->stream>>>(ctx->W, L.disp_f, L.lr_mask);
        S3_LAUNCH_CHECK();
    }
    S3_EV_END(S3DMST_T_POST, 0);
    return 0;
}

// ---- the step after the path in its only caller (src/stereo_Yin.cpp:218-243): left disparities
// below a floor are raised to it, cv::reprojectImageTo3D(disp, xyz, Q, handleMissingValues = true), and the RGB point
// cloud is assembled.  Arithmetic of OpenCV's reprojectImageTo3D for a CV_32F map (checked bit for bit against cv2 by
// tests/test_gpu_parity.py): P = Q * (x, y, d, 1) accumulated left to right in double; each of P.x, P.y, P.z is first
// rounded to float, then scaled by the double 1 / P.w and rounded to float again; where |d - min(d)| <= FLT_EPSILON
// the depth is the "missing value" marker 10000.
__global__ void k_floor_min(int N, float floor_v, float* __restrict__ disp, unsigned* __restrict__ min_bits) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    float d = 3.0e38f;
    if (p < N) {
        d = disp[p];
        if (d < floor_v) disp[p] = d = floor_v;  // :218-222
    }
    // non-negative floats order like their bit patterns
    unsigned b = __float_as_uint(d < 0.0f ? 0.0f : d);
    b = __reduce_min_sync(0xffffffffu, b);
    if ((threadIdx.x & 31) == 0) atomicMin(min_bits, b);
}
struct QMat {
    double q[16];
};
__global__ void k_reproject(int W, int N, QMat Q, int handle_missing, const float* __restrict__ disp, const unsigned* __restrict__ min_bits,
                            const uint8_t* __restrict__ bgr, float* __restrict__ xyz, uint32_t* __restrict__ rgb) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const double x = (double)(p % W), y = (double)(p / W), d = (double)disp[p];
    double P[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
        P[i] = S3_DADD(S3_DADD(S3_DADD(S3_DMUL(Q.q[4 * i], x), S3_DMUL(Q.q[4 * i + 1], y)), S3_DMUL(Q.q[4 * i + 2], d)), Q.q[4 * i + 3]);
    const double iw = __ddiv_rn(1.0, P[3]);
    float o[3];
#pragma unroll
    for (int i = 0; i < 3; i++) o[i] = (float)S3_DMUL((double)(float)P[i], iw);
    if (handle_missing && fabs(d - (double)__uint_as_float(*min_bits)) <= (double)FLT_EPSILON) o[2] = 10000.0f;
    if (xyz) {
        xyz[3 * (size_t)p] = o[0];
        xyz[3 * (size_t)p + 1] = o[1];
        xyz[3 * (size_t)p + 2] = o[2];
    }
    if (rgb) {
        const uint8_t* c = bgr + 3 * (size_t)p;  // (b, g, r) as set_images stored them; pcl::PointXYZRGB packing of stereo_Yin.cpp:240-241
        rgb[p] = (uint32_t)c[2] * 0x10000u + (uint32_t)c[1] * 0x100u + (uint32_t)c[0];
    }
}

int s3_reproject(s3dmst_ctx* ctx, const double* Q16, float disp_floor, int handle_missing, float* h_xyz, uint32_t* h_rgb) {
    View& L = ctx->v[0];
    const int N = ctx->N;
    const size_t need = 16 + (size_t)N * 16;
    if (ctx->pms_scratch_cap < need) {
        if (ctx->pms_scratch) S3_CUDA(cudaFree(ctx->pms_scratch));
        ctx->pms_scratch = nullptr; ctx->pms_scratch_cap = 0;
        S3_CUDA(cudaMalloc(&ctx->pms_scratch, need));
        ctx->pms_scratch_cap = need;
    }
    unsigned* min_bits = reinterpret_cast<unsigned*>(ctx->pms_scratch);
    float* d_xyz = reinterpret_cast<float*>(reinterpret_cast<char*>(ctx->pms_scratch) + 16);
    uint32_t* d_rgb = reinterpret_cast<uint32_t*>(d_xyz + 3 * (size_t)N);
    S3_CUDA(cudaMemsetAsync(min_bits, 0xFF, sizeof(unsigned), ctx->stream));
    S3_EV_BEGIN(S3DMST_T_POST, 1);
    k_floor_min<<<(N + 255) / 256, 256, 0, ctx->stream>>>(N, disp_floor, L.disp_f, min_bits);
    S3_LAUNCH_CHECK();
    QMat Q;
    for (int i = 0; i < 16; i++) Q.q[i] = Q16[i];
    k_reproject<<<(N + 255) / 256, 256, 0, ctx->stream>>>(ctx->W, N, Q, handle_missing, L.disp_f, min_bits, L.bgr, h_xyz ? d_xyz : nullptr, h_rgb ? d_rgb : nullptr);
    S3_LAUNCH_CHECK();
    S3_EV_END(S3DMST_T_POST, 1);
    if (h_xyz) S3_CUDA(cudaMemcpyAsync(h_xyz, d_xyz, sizeof(float) * 3 * (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
    if (h_rgb) S3_CUDA(cudaMemcpyAsync(h_rgb, d_rgb, sizeof(uint32_t) * (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
