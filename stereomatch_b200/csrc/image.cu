// stereomatch_b200/csrc/image.cu — image stage of the forest build.
//
//   k_prep          3x3 median per channel (cv::medianBlur, Stereo3DMST.cpp:226-228) + raw uchar4 / gray planes
//   k_edge_weights  w = |dR|+|dG|+|dB| for the right and down neighbour (:83-91, :244-262) + weight histogram
//   k_bucket_*      counting sort of edge ids by integer weight (the reference's std::sort by (w,a,b),
//                   segment-graph.h:57; only the grouping by w is needed here, see forest.cu)
// All HBM-bound streaming kernels: ~4 B read + 12 B written per pixel.
#include "hd_math.h"
#include "internal.h"

__global__ void k_prep(const uint8_t* __restrict__ bgr, int W, int H, int do_median, uchar4* __restrict__ raw4,
                       float* __restrict__ gray, uchar4* __restrict__ med) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= W * H) return;
    const int x = p % W, y = p / W;
    const uint8_t* c = bgr + 3 * (size_t)p;
    const int b = c[0], g = c[1], r = c[2];
    raw4[p] = make_uchar4(b, g, r, 0);
    gray[p] = s3_gray(b, g, r);
    if (!do_median) {
        med[p] = make_uchar4(b, g, r, 0);
        return;
    }
    int v[3][9];
    int k = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; dy++) {
        const int yy = min(H - 1, max(0, y + dy));
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
            const int xx = min(W - 1, max(0, x + dx));
            const uint8_t* q = bgr + 3 * ((size_t)yy * W + xx);
            v[0][k] = q[0];
            v[1][k] = q[1];
            v[2][k] = q[2];
            k++;
        }
    }
    uchar4 m;
    m.x = s3_median9(v[0][0], v[0][1], v[0][2], v[0][3], v[0][4], v[0][5], v[0][6], v[0][7], v[0][8]);
    m.y = s3_median9(v[1][0], v[1][1], v[1][2], v[1][3], v[1][4], v[1][5], v[1][6], v[1][7], v[1][8]);
    m.z = s3_median9(v[2][0], v[2][1], v[2][2], v[2][3], v[2][4], v[2][5], v[2][6], v[2][7], v[2][8]);
    m.w = 0;
    med[p] = m;
}

__device__ __forceinline__ int l1_diff(uchar4 a, uchar4 b) {
    return abs((int)a.x - (int)b.x) + abs((int)a.y - (int)b.y) + abs((int)a.z - (int)b.z);
}

__global__ void k_edge_weights(const uchar4* __restrict__ med, int W, int H, uint16_t* __restrict__ ew,
                               int* __restrict__ hist) {
    __shared__ int sh[S3_NUM_W];
    for (int i = threadIdx.x; i < S3_NUM_W; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < W * H) {
        const int x = p % W, y = p / W;
        const uchar4 c = med[p];
        uint32_t wr = S3_NO_EDGE, wd = S3_NO_EDGE;
        if (x < W - 1) {
            wr = l1_diff(c, med[p + 1]);
            atomicAdd(&sh[wr], 1);
        }
        if (y < H - 1) {
            wd = l1_diff(c, med[p + W]);
            atomicAdd(&sh[wd], 1);
        }
        reinterpret_cast<uint32_t*>(ew)[p] = wr | (wd << 16);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S3_NUM_W; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// single block: exclusive scan of the 766-bin histogram
__global__ void k_bucket_offsets(const int* __restrict__ hist, int* __restrict__ lvl_off, int* __restrict__ cursor) {
    __shared__ int s[1024];
    const int t = threadIdx.x;
    s[t] = t < S3_NUM_W ? hist[t] : 0;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        int v = t >= off ? s[t - off] : 0;
        __syncthreads();
        s[t] += v;
        __syncthreads();
    }
    if (t < S3_NUM_W) {
        const int ex = s[t] - hist[t];
        lvl_off[t] = ex;
        cursor[t] = ex;
    }
    if (t == S3_NUM_W - 1) lvl_off[S3_NUM_W] = s[t];
}

// scatter edge ids into their weight bucket.  One CTA takes BS_EDGES consecutive edge ids: a shared-memory histogram
// reserves one range per (CTA, weight) with a single global atomic, then the edges are placed with shared-memory
// cursors — ~10x fewer same-address global atomics than one per warp and weight (weights are heavily skewed on
// natural images and spread over ~500 values on the random-dot workload).
#define BS_THREADS 256
#define BS_PER_THREAD 32
#define BS_EDGES (BS_THREADS * BS_PER_THREAD)
__global__ void __launch_bounds__(BS_THREADS) k_bucket_scatter(const uint16_t* __restrict__ ew, int E2, int* __restrict__ cursor,
                                                               uint32_t* __restrict__ elist) {
    __shared__ int s_cnt[S3_NUM_W];
    __shared__ int s_base[S3_NUM_W];
    for (int i = threadIdx.x; i < S3_NUM_W; i += BS_THREADS) s_cnt[i] = 0;
    __syncthreads();
    const int e0 = blockIdx.x * BS_EDGES;
    uint16_t w[BS_PER_THREAD];
#pragma unroll
    for (int k = 0; k < BS_PER_THREAD; k++) {
        const int e = e0 + k * BS_THREADS + threadIdx.x;
        w[k] = e < E2 ? ew[e] : (uint16_t)S3_NO_EDGE;
        if (w[k] != S3_NO_EDGE) atomicAdd(&s_cnt[w[k]], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S3_NUM_W; i += BS_THREADS) {
        const int c = s_cnt[i];
        s_base[i] = c ? atomicAdd(&cursor[i], c) : 0;
        s_cnt[i] = 0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BS_PER_THREAD; k++) {
        if (w[k] == S3_NO_EDGE) continue;
        const int e = e0 + k * BS_THREADS + threadIdx.x;
        elist[s_base[w[k]] + atomicAdd(&s_cnt[w[k]], 1)] = (uint32_t)e;
    }
}

int s3_image_stage(s3dmst_ctx* ctx, int view) {
    View& V = ctx->v[view];
    const int N = ctx->N, W = ctx->W, H = ctx->H;
    const int TB = 256;
    k_prep<<<(N + TB - 1) / TB, TB, 0, ctx->stream>>>(V.bgr, W, H, ctx->P.median != 0, V.raw4, V.gray, V.med);
    S3_LAUNCH_CHECK();
    S3_CUDA(cudaMemsetAsync(V.hist, 0, sizeof(int) * S3_NUM_W, ctx->stream));
    k_edge_weights<<<(N + TB - 1) / TB, TB, 0, ctx->stream>>>(V.med, W, H, V.ew, V.hist);
    S3_LAUNCH_CHECK();
    k_bucket_offsets<<<1, 1024, 0, ctx->stream>>>(V.hist, V.lvl_off, V.lvl_cursor);
    S3_LAUNCH_CHECK();
    k_bucket_scatter<<<(2 * N + BS_EDGES - 1) / BS_EDGES, BS_THREADS, 0, ctx->stream>>>(V.ew, 2 * N, V.lvl_cursor, V.elist);
    S3_LAUNCH_CHECK();
    return 0;
}
