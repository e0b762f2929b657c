// micro-benchmark: FP64 DMUL/DADD latency and throughput, LDS latency, bar.sync cost on this GPU
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int mode) {
    double a = threadIdx.x * 1e-3 + 1.0, b = 1.0000001, c = 0.5, d = 0.25, e2 = 0.125;
    __shared__ double sm[1024];
    sm[threadIdx.x] = a;
    __syncthreads();
    long long t0 = clock64();
    if (mode == 0) {  // dependent chain: 256 x (DMUL, DADD)
        for (int i = 0; i < 256; i++) { a = __dmul_rn(a, b); a = __dadd_rn(a, c); }
    } else if (mode == 1) {  // 4 independent chains
        for (int i = 0; i < 256; i++) {
            a = __dmul_rn(a, b); c = __dmul_rn(c, b); d = __dmul_rn(d, b); e2 = __dmul_rn(e2, b);
            a = __dadd_rn(a, b); c = __dadd_rn(c, b); d = __dadd_rn(d, b); e2 = __dadd_rn(e2, b);
        }
    } else if (mode == 2) {  // dependent LDS chain
        int idx = threadIdx.x;
        for (int i = 0; i < 256; i++) { a += sm[idx]; idx = (idx + (int)a) & 1023; }
    } else if (mode == 3) {  // bar.sync
        for (int i = 0; i < 256; i++) __syncthreads();
    } else if (mode == 4) {  // dependent FFMA chain
        float x = a, y = b;
        for (int i = 0; i < 512; i++) x = __fmaf_rn(x, y, 0.5f);
        a = x;
    } else if (mode == 5) {  // dependent IMAD chain
        int x = threadIdx.x;
        for (int i = 0; i < 512; i++) x = x * 3 + i;
        a = x;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + c + d + e2;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[mode] = t1 - t0;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 1024 * 148 * 4); cudaMalloc(&cyc, 64);
    const char* names[] = {"dep DMUL+DADD x256 (per pair)", "4-way ILP 8 FP64 x256 (per iter)", "dep LDS x256", "bar.sync x256", "dep FFMA x512", "dep IMAD x512"};
    int div[] = {256, 256, 256, 256, 512, 512};
    for (int threads : {32, 576}) for (int blocks : {1, 296}) {
        printf("threads %d blocks %d\n", threads, blocks);
        for (int m = 0; m < 6; m++) {
            k<<<blocks, threads>>>(out, cyc, m); cudaDeviceSynchronize();
            long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
            printf("  %-36s %8.1f cycles\n", names[m], (double)h[m] / div[m]);
        }
    }
    return 0;
}
