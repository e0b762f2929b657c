// stereomatch_b200/csrc/comm.cu — label-range sharding of ONE large pair over the GPUs of a box (BASELINE
// config C5): every rank builds the (deterministic) forests itself, aggregates its own label range, and the only
// exchange is a per-pixel MIN-LOC of (aggregated cost, disparity) — the reduction that replaces the serial
// `if (agg < min_cost)` over ascending labels of the reference (Stereo3DMST.cpp:173-185; dense mode: SURVEY A13).
//
// Default: ONE kernel over peer memory (k_minloc_p2p below: every rank's result buffers mapped through CUDA IPC).  The
// fall-back, and the transport of s3dmst_reduce_minloc before the first sharded call, is NCCL — which has no MINLOC, and
// the exact-mode cost is fp64, so it is two all-reduces on the context's communication stream:
//   1. ncclAllReduce(MIN, f64) of the best cost                       -> global minimum per pixel
//   2. disparity := INT32_MAX where local cost != global minimum      (k_minloc_mask)
//      ncclAllReduce(MIN, i32) of the disparity                       -> lowest d attaining the minimum
// = the reference's tie rule (strict '<' over ascending d).  No host synchronisation; the reduction of the left view
// runs on the communication stream while the main stream aggregates the right view.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy the process already holds — e.g. torch's — or the
// system's), so the library has no link-time dependency on it and contexts that never shard never load it.
#include <dlfcn.h>
#include <float.h>
#include <limits.h>
#include <string.h>

#include "internal.h"
#include "../../include/s3dmst.h"

// the slice of nccl.h this file needs (stable since NCCL 2.0)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } s3_ncclUniqueId;
enum { s3_ncclInt8 = 0, s3_ncclInt32 = 2, s3_ncclFloat64 = 8, s3_ncclMin = 3 };
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(s3_ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, s3_ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string err;
};
static NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return &api;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) {
        api.err = std::string("NCCL not found: ") + dlerror();
        return &api;
    }
#define S3_SYM(field, name)                                                     \
    do {                                                                        \
        *(void**)(&api.field) = dlsym(api.handle, name);                        \
        if (!api.field) api.err = std::string("NCCL symbol missing: ") + name;  \
    } while (0)
    S3_SYM(GetUniqueId, "ncclGetUniqueId");
    S3_SYM(CommInitRank, "ncclCommInitRank");
    S3_SYM(CommDestroy, "ncclCommDestroy");
    S3_SYM(AllReduce, "ncclAllReduce");
    S3_SYM(AllGather, "ncclAllGather");
    S3_SYM(GetErrorString, "ncclGetErrorString");
#undef S3_SYM
    return &api;
}
#define S3_NCCL(call)                                                                                                     \
    do {                                                                                                                  \
        const int r__ = (call);                                                                                           \
        if (r__ != 0) return s3_fail(ctx, S3DMST_E_COMM, "%s:%d %s -> %s", __FILE__, __LINE__, #call, api->GetErrorString(r__)); \
    } while (0)

__global__ void k_minloc_mask(int N, const double* __restrict__ best, const double* __restrict__ gmin, int32_t* __restrict__ disp) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < N && best[p] != gmin[p]) disp[p] = INT_MAX;
}
__global__ void k_minloc_identity(int N, double* __restrict__ best, int32_t* __restrict__ disp) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < N) { best[p] = DBL_MAX; disp[p] = INT_MAX; }
}

// ------------------------------------------------------------------------------------------------
// MIN-LOC over peer memory.  NCCL needs two all-reduces and a mask kernel per view (it has no MINLOC) — 2 x (8 + 4) bytes
// per pixel through its ring/tree protocols plus four launches.  With every rank's (best, disparity) buffers mapped into
// every process (CUDA IPC; NVSwitch gives each GPU full bandwidth to each peer) it is ONE kernel: rank r owns pixel
// slice r, reads that slice from all ranks (NVLink loads), takes the minimum with the lowest disparity on ties (ranks
// hold ascending label ranges) and stores the result into all ranks' buffers.  Two flag exchanges frame it: "my partial
// result is complete" before the loads, "my stores have landed" after them; a one-thread kernel behind it waits for the
// peers' second flag, so the stream continues only when this rank's buffers hold the global result.  Every wait has a
// time-out (a rank that never arrives must not hang the GPU); it raises a flag in mapped host memory.
struct P2PArgs {
    int N, rank, nranks, epoch;
    double* best[S3_P2P_MAX];
    int32_t* disp[S3_P2P_MAX];
    int* flags[S3_P2P_MAX];   // [2][S3_P2P_MAX] per rank: READY then DONE words, indexed by the signalling rank
    int* counter;
    int* err;
};
#define S3_P2P_TIMEOUT_NS 4000000000ull
__device__ __forceinline__ unsigned long long p2p_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void p2p_signal(int* p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int p2p_peek(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// waits until every rank's word in flags[kind] of THIS rank has reached epoch; false on time-out
__device__ bool p2p_wait_all(const int* own_flags, int kind, int nranks, int epoch, int* err) {
    const unsigned long long t0 = p2p_now();
    for (int r = 0; r < nranks; r++) {
        while (p2p_peek(own_flags + kind * S3_P2P_MAX + r) < epoch) {
            if (p2p_now() - t0 > S3_P2P_TIMEOUT_NS) { *err = 1 + kind; return false; }
            __nanosleep(200);
        }
    }
    return true;
}
__global__ void __launch_bounds__(256) k_minloc_p2p(P2PArgs A) {
    const int tid = threadIdx.x;
    __shared__ int s_ok;
    if (tid == 0) {
        if (blockIdx.x == 0) {  // everything this GPU wrote before this kernel (the aggregation) is visible to the peers first
            __threadfence_system();
            for (int r = 0; r < A.nranks; r++) p2p_signal(A.flags[r] + 0 * S3_P2P_MAX + A.rank, A.epoch);
        }
        s_ok = p2p_wait_all(A.flags[A.rank], 0, A.nranks, A.epoch, A.err) ? 1 : 0;
    }
    __syncthreads();
    if (s_ok) {
        // pixel slice of this rank, in pairs (16-byte loads of the costs); system-scope accesses: never this SM's L1
        const int pairs = (A.N + 1) / 2, per = (pairs + A.nranks - 1) / A.nranks;
        const int lo = A.rank * per, hi = min(pairs, lo + per);
        for (int q = lo + blockIdx.x * blockDim.x + tid; q < hi; q += gridDim.x * blockDim.x) {
            const int i = 2 * q;
            const bool two = i + 1 < A.N;
            double b0 = DBL_MAX, b1 = DBL_MAX;
            int d0 = INT_MAX, d1 = INT_MAX;
            for (int r = 0; r < A.nranks; r++) {
                double x0, x1 = DBL_MAX;
                int y0, y1 = INT_MAX;
                if (two) {
                    asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(x0), "=d"(x1) : "l"(A.best[r] + i) : "memory");
                    asm volatile("ld.relaxed.sys.global.v2.s32 {%0, %1}, [%2];" : "=r"(y0), "=r"(y1) : "l"(A.disp[r] + i) : "memory");
                } else {
                    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(x0) : "l"(A.best[r] + i) : "memory");
                    asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(y0) : "l"(A.disp[r] + i) : "memory");
                }
                if (x0 < b0 || (x0 == b0 && y0 < d0)) { b0 = x0; d0 = y0; }
                if (x1 < b1 || (x1 == b1 && y1 < d1)) { b1 = x1; d1 = y1; }
            }
            for (int r = 0; r < A.nranks; r++) {
                if (two) {
                    asm volatile("st.relaxed.sys.global.v2.f64 [%0], {%1, %2};" ::"l"(A.best[r] + i), "d"(b0), "d"(b1) : "memory");
                    asm volatile("st.relaxed.sys.global.v2.s32 [%0], {%1, %2};" ::"l"(A.disp[r] + i), "r"(d0), "r"(d1) : "memory");
                } else {
                    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(A.best[r] + i), "d"(b0) : "memory");
                    asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(A.disp[r] + i), "r"(d0) : "memory");
                }
            }
        }
    }
    // this rank's stores have landed everywhere -> DONE to every rank (by the last CTA)
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        if (atomicAdd(A.counter, 1) == (int)gridDim.x - 1) {
            *A.counter = 0;
            __threadfence_system();
            for (int r = 0; r < A.nranks; r++) p2p_signal(A.flags[r] + 1 * S3_P2P_MAX + A.rank, A.epoch);
        }
    }
}
__global__ void k_p2p_wait_done(P2PArgs A) {
    if (threadIdx.x == 0) p2p_wait_all(A.flags[A.rank], 1, A.nranks, A.epoch, A.err);
}

// step 2 alone, for callers that run the two all-reduces themselves (s3dmst_minloc_mask)
int s3_minloc_mask(s3dmst_ctx* ctx, int view, const double* global_min_dev) {
    View& V = ctx->v[view];
    if (ctx->N == 0) return s3_fail(ctx, S3DMST_E_STATE, "minloc_mask: no dense result");
    S3_CUDA(cudaSetDevice(ctx->device));
    k_minloc_mask<<<(ctx->N + 255) / 256, 256, 0, ctx->stream>>>(ctx->N, V.best, global_min_dev, V.disp_i);
    S3_LAUNCH_CHECK();
    return 0;
}

// contiguous label shard [d0, d1) of `rank`: boundaries are multiples of 4 labels (16-byte rows); ranks beyond the
// number of blocks get an empty range
static void label_range(int D, int nranks, int rank, int* d0, int* d1) {
    const int blocks = (D + 3) / 4, per = blocks / nranks, extra = blocks % nranks;
    const int b0 = rank * per + (rank < extra ? rank : extra), b1 = b0 + per + (rank < extra ? 1 : 0);
    *d0 = b0 * 4 < D ? b0 * 4 : D;
    *d1 = b1 * 4 < D ? b1 * 4 : D;
}

static void p2p_close(s3dmst_ctx* ctx);

extern "C" {

int s3dmst_comm_unique_id(void* id128) {
    NcclApi* api = nccl_api();
    if (!id128 || !api->err.empty()) return S3DMST_E_COMM;
    s3_ncclUniqueId id;
    if (api->GetUniqueId(&id) != 0) return S3DMST_E_COMM;
    memcpy(id128, &id, sizeof id);
    return 0;
}

int s3dmst_comm_init(s3dmst_ctx* ctx, const void* id128, int rank, int nranks) {
    if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return s3_fail(ctx, S3DMST_E_ARG, "comm_init: bad arguments");
    NcclApi* api = nccl_api();
    if (!api->err.empty()) return s3_fail(ctx, S3DMST_E_COMM, "comm_init: %s", api->err.c_str());
    if (ctx->comm) return s3_fail(ctx, S3DMST_E_STATE, "comm_init: the context already has a communicator");
    S3_CUDA(cudaSetDevice(ctx->device));
    s3_ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclComm_t comm = nullptr;
    S3_NCCL(api->CommInitRank(&comm, nranks, id, rank));
    ctx->comm = comm;
    ctx->comm_rank = rank;
    ctx->comm_nranks = nranks;
    if (!ctx->comm_stream) S3_CUDA(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++)
        if (!ctx->ev_comm[i]) S3_CUDA(cudaEventCreateWithFlags(&ctx->ev_comm[i], cudaEventDisableTiming));
    for (int i = 0; i < 4; i++)
        if (!ctx->ev_comm_t[i]) S3_CUDA(cudaEventCreate(&ctx->ev_comm_t[i]));
    return 0;
}

int s3dmst_comm_destroy(s3dmst_ctx* ctx) {
    if (!ctx->comm) return 0;
    NcclApi* api = nccl_api();
    cudaSetDevice(ctx->device);
    if (ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    s3_comm_before_free(ctx);   // (collective, like this call) nobody frees result buffers a peer still has mapped
    p2p_close(ctx);
    ctx->p2p_tried = false;
    if (ctx->p2p_xbuf) { cudaFree(ctx->p2p_xbuf); ctx->p2p_xbuf = nullptr; }
    if (ctx->p2p_counter) { cudaFree(ctx->p2p_counter); ctx->p2p_counter = nullptr; }
    if (ctx->p2p_err_host) { cudaFreeHost(ctx->p2p_err_host); ctx->p2p_err_host = nullptr; ctx->p2p_err_dev = nullptr; }
    api->CommDestroy(reinterpret_cast<ncclComm_t>(ctx->comm));
    ctx->comm = nullptr;
    ctx->comm_nranks = 0;
    return 0;
}

int s3dmst_comm_label_range(const s3dmst_ctx* ctx, int D, int* d0, int* d1) {
    if (!ctx->comm || !d0 || !d1) return S3DMST_E_ARG;
    label_range(D, ctx->comm_nranks, ctx->comm_rank, d0, d1);
    return 0;
}

}  // extern "C"

static void p2p_close(s3dmst_ctx* ctx) {
    for (int r = 0; r < S3_P2P_MAX; r++) {
        if (r == ctx->comm_rank) continue;
        for (int view = 0; view < 2; view++) {
            if (ctx->p2p_best[view][r]) cudaIpcCloseMemHandle(ctx->p2p_best[view][r]);
            if (ctx->p2p_disp[view][r]) cudaIpcCloseMemHandle(ctx->p2p_disp[view][r]);
        }
        if (ctx->p2p_flags[r]) cudaIpcCloseMemHandle(ctx->p2p_flags[r]);
    }
    if (ctx->comm_rank >= 0 && ctx->comm_rank < S3_P2P_MAX && ctx->p2p_flags[ctx->comm_rank]) cudaFree(ctx->p2p_flags[ctx->comm_rank]);
    memset(ctx->p2p_best, 0, sizeof ctx->p2p_best);
    memset(ctx->p2p_disp, 0, sizeof ctx->p2p_disp);
    memset(ctx->p2p_flags, 0, sizeof ctx->p2p_flags);
    cudaGetLastError();
    ctx->p2p_ok = false;
    ctx->p2p_N = 0;
}

// Collective (only when the peer-memory mapping is live): every rank closes its mappings of the others' buffers, and
// nobody goes on to free its own before all have — freeing memory another process still has mapped is undefined.
int s3_comm_before_free(s3dmst_ctx* ctx) {
    if (!ctx->comm || !ctx->p2p_ok) return 0;
    NcclApi* api = nccl_api();
    ncclComm_t comm = reinterpret_cast<ncclComm_t>(ctx->comm);
    S3_CUDA(cudaSetDevice(ctx->device));
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    S3_CUDA(cudaStreamSynchronize(ctx->comm_stream));
    int* own_flags = ctx->p2p_flags[ctx->comm_rank];
    ctx->p2p_flags[ctx->comm_rank] = nullptr;   // freed below, after the barrier
    p2p_close(ctx);
    int* w = reinterpret_cast<int*>(reinterpret_cast<char*>(ctx->p2p_xbuf));
    S3_CUDA(cudaMemsetAsync(w, 0, sizeof(int), ctx->comm_stream));
    S3_NCCL(api->AllReduce(w, w, 1, s3_ncclInt32, s3_ncclMin, comm, ctx->comm_stream));
    S3_CUDA(cudaStreamSynchronize(ctx->comm_stream));
    if (own_flags) cudaFree(own_flags);
    ctx->p2p_tried = false;   // the next sharded call maps the new buffers
    return 0;
}

// Collective: maps every rank's result buffers and flag array into this process (again whenever the image size, hence the
// buffers, changed).  All ranks agree on the outcome (an all-reduce of "it worked here"): either everybody uses the
// peer-memory kernel or everybody stays with NCCL.
static int p2p_setup(s3dmst_ctx* ctx) {
    NcclApi* api = nccl_api();
    ncclComm_t comm = reinterpret_cast<ncclComm_t>(ctx->comm);
    const int R = ctx->comm_nranks, me = ctx->comm_rank;
    ctx->p2p_tried = true;
    p2p_close(ctx);
    if (ctx->P.comm_p2p < 0 || R > S3_P2P_MAX) return 0;
    struct Pack { cudaIpcMemHandle_t h[5]; };
    if (!ctx->p2p_xbuf) S3_CUDA(cudaMalloc(&ctx->p2p_xbuf, sizeof(Pack) * S3_P2P_MAX + 64));
    if (!ctx->p2p_counter) {
        S3_CUDA(cudaMalloc(&ctx->p2p_counter, 64));
        S3_CUDA(cudaMemset(ctx->p2p_counter, 0, 64));
    }
    if (!ctx->p2p_err_host) {
        S3_CUDA(cudaHostAlloc(&ctx->p2p_err_host, 64, cudaHostAllocMapped));
        *ctx->p2p_err_host = 0;
        S3_CUDA(cudaHostGetDevicePointer(&ctx->p2p_err_dev, ctx->p2p_err_host, 0));
    }
    int ok = 1;
    int* flags = nullptr;
    S3_CUDA(cudaMalloc(&flags, sizeof(int) * 2 * S3_P2P_MAX));
    S3_CUDA(cudaMemset(flags, 0, sizeof(int) * 2 * S3_P2P_MAX));
    ctx->p2p_flags[me] = flags;
    ctx->p2p_epoch = 0;
    Pack mine;
    memset(&mine, 0, sizeof mine);
    void* ptrs[5] = {ctx->v[0].best, ctx->v[0].disp_i, ctx->v[1].best, ctx->v[1].disp_i, flags};
    for (int i = 0; i < 5; i++)
        if (cudaIpcGetMemHandle(&mine.h[i], ptrs[i]) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    char* xb = reinterpret_cast<char*>(ctx->p2p_xbuf);
    S3_CUDA(cudaMemcpyAsync(xb + sizeof(Pack) * me, &mine, sizeof mine, cudaMemcpyHostToDevice, ctx->comm_stream));
    S3_NCCL(api->AllGather(xb + sizeof(Pack) * me, xb, sizeof(Pack), s3_ncclInt8, comm, ctx->comm_stream));
    std::vector<Pack> all(R);
    S3_CUDA(cudaMemcpyAsync(all.data(), xb, sizeof(Pack) * R, cudaMemcpyDeviceToHost, ctx->comm_stream));
    S3_CUDA(cudaStreamSynchronize(ctx->comm_stream));
    for (int r = 0; r < R && ok; r++) {
        if (r == me) {
            ctx->p2p_best[0][r] = ctx->v[0].best; ctx->p2p_disp[0][r] = ctx->v[0].disp_i;
            ctx->p2p_best[1][r] = ctx->v[1].best; ctx->p2p_disp[1][r] = ctx->v[1].disp_i;
            continue;
        }
        void* p[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        for (int i = 0; i < 5 && ok; i++)
            if (cudaIpcOpenMemHandle(&p[i], all[r].h[i], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; }
        ctx->p2p_best[0][r] = (double*)p[0]; ctx->p2p_disp[0][r] = (int32_t*)p[1];
        ctx->p2p_best[1][r] = (double*)p[2]; ctx->p2p_disp[1][r] = (int32_t*)p[3];
        ctx->p2p_flags[r] = (int*)p[4];
    }
    // does every rank see every other?  (MIN over the ranks of "ok here")
    int* okd = reinterpret_cast<int*>(xb + sizeof(Pack) * S3_P2P_MAX);
    S3_CUDA(cudaMemcpyAsync(okd, &ok, sizeof(int), cudaMemcpyHostToDevice, ctx->comm_stream));
    S3_NCCL(api->AllReduce(okd, okd, 1, s3_ncclInt32, s3_ncclMin, comm, ctx->comm_stream));
    S3_CUDA(cudaMemcpyAsync(&ok, okd, sizeof(int), cudaMemcpyDeviceToHost, ctx->comm_stream));
    S3_CUDA(cudaStreamSynchronize(ctx->comm_stream));
    if (!ok) { p2p_close(ctx); return 0; }
    ctx->p2p_ok = true;
    ctx->p2p_N = ctx->N;
    return 0;
}

// MIN-LOC all-reduce of view's (best, disp_i) over the communicator, queued on `st` (no host synchronisation)
static int reduce_minloc_on(s3dmst_ctx* ctx, int view, cudaStream_t st) {
    NcclApi* api = nccl_api();
    View& V = ctx->v[view];
    if (ctx->p2p_ok && ctx->p2p_N == ctx->N) {
        if (*ctx->p2p_err_host) return s3_fail(ctx, S3DMST_E_COMM, "peer-memory MIN-LOC: a wait for the other ranks timed out (code %d)", *ctx->p2p_err_host);
        P2PArgs A;
        memset(&A, 0, sizeof A);
        A.N = ctx->N; A.rank = ctx->comm_rank; A.nranks = ctx->comm_nranks; A.epoch = ++ctx->p2p_epoch;
        for (int r = 0; r < ctx->comm_nranks; r++) { A.best[r] = ctx->p2p_best[view][r]; A.disp[r] = ctx->p2p_disp[view][r]; A.flags[r] = ctx->p2p_flags[r]; }
        A.counter = ctx->p2p_counter; A.err = ctx->p2p_err_dev;
        const int pairs = (ctx->N + 1) / 2, per = (pairs + ctx->comm_nranks - 1) / ctx->comm_nranks;
        const int grid = std::max(1, std::min(ctx->num_sms, (per + 1023) / 1024));
        k_minloc_p2p<<<grid, 256, 0, st>>>(A);
        S3_LAUNCH_CHECK();
        k_p2p_wait_done<<<1, 32, 0, st>>>(A);
        S3_LAUNCH_CHECK();
        return 0;
    }
    ncclComm_t comm = reinterpret_cast<ncclComm_t>(ctx->comm);
    const size_t N = ctx->N;
    if (ctx->gmin_cap < N) {
        if (ctx->gmin) S3_CUDA(cudaFree(ctx->gmin));
        ctx->gmin = nullptr; ctx->gmin_cap = 0;
        S3_CUDA(cudaMalloc(&ctx->gmin, sizeof(double) * N));
        ctx->gmin_cap = N;
    }
    S3_NCCL(api->AllReduce(V.best, ctx->gmin, N, s3_ncclFloat64, s3_ncclMin, comm, st));
    k_minloc_mask<<<(unsigned)((N + 255) / 256), 256, 0, st>>>((int)N, V.best, ctx->gmin, V.disp_i);
    S3_LAUNCH_CHECK();
    S3_NCCL(api->AllReduce(V.disp_i, V.disp_i, N, s3_ncclInt32, s3_ncclMin, comm, st));
    S3_CUDA(cudaMemcpyAsync(V.best, ctx->gmin, sizeof(double) * N, cudaMemcpyDeviceToDevice, st));  // the context holds the global minimum next to the global disparity
    return 0;
}

extern "C" {

int s3dmst_reduce_minloc(s3dmst_ctx* ctx, int view) {
    if (view < 0 || view > 1) return s3_fail(ctx, S3DMST_E_ARG, "bad view");
    if (!ctx->comm) return s3_fail(ctx, S3DMST_E_STATE, "reduce_minloc: s3dmst_comm_init first");
    if (ctx->N == 0) return s3_fail(ctx, S3DMST_E_STATE, "reduce_minloc: no dense result");
    S3_CUDA(cudaSetDevice(ctx->device));
    return reduce_minloc_on(ctx, view, ctx->stream);
}

int s3dmst_aggregate_dense_sharded(s3dmst_ctx* ctx, int D) {
    if (!ctx->comm) return s3_fail(ctx, S3DMST_E_STATE, "aggregate_dense_sharded: s3dmst_comm_init first");
    S3_CUDA(cudaSetDevice(ctx->device));
    int d0, d1;
    label_range(D, ctx->comm_nranks, ctx->comm_rank, &d0, &d1);
    // Without a volume of D labels on both views the matching cost is computed inside the aggregation kernel: a rank then
    // holds no cost volume at all, only the running sums of the labels it aggregates... (rows keep their full pitch).
    const bool have_vol = ctx->v[0].cost_ready && ctx->v[1].cost_ready && ctx->v[0].D == D && ctx->v[1].D == D;
    const bool fuse = !have_vol && s3_want_fused_cost(ctx);
    if (fuse) S3_TRY(s3_fused_prepare(ctx, D));
    if (!ctx->p2p_tried || (ctx->p2p_ok && ctx->p2p_N != ctx->N)) S3_TRY(p2p_setup(ctx));  // collective, first call / new image size
    for (int view = 0; view < 2; view++) {
        View& V = ctx->v[view];
        if (!V.forest_ready || !(fuse || have_vol)) return s3_fail(ctx, S3DMST_E_STATE, "aggregate_dense_sharded: forests and a cost volume of D labels required");
        if (d1 > d0) {
            int rc = ctx->P.agg_kernel == 1 ? 1 : s3_aggregate_flow(ctx, 1 << view, d0, d1, fuse);
            if (rc == 1) rc = s3_aggregate_dense(ctx, view, d0, d1);
            if (rc) return rc;
        } else {  // more ranks than label blocks: contribute the identity of MIN-LOC
            k_minloc_identity<<<(ctx->N + 255) / 256, 256, 0, ctx->stream>>>(ctx->N, V.best, V.disp_i);
            S3_LAUNCH_CHECK();
            V.agg_ready = true; V.agg_d0 = V.agg_d1 = 0;
        }
        // the reduction of this view goes to the communication stream; the main stream moves on to the next view
        S3_CUDA(cudaEventRecord(ctx->ev_comm[0], ctx->stream));
        S3_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_comm[0], 0));
        S3_CUDA(cudaEventRecord(ctx->ev_comm_t[2 * view], ctx->comm_stream));
        S3_TRY(reduce_minloc_on(ctx, view, ctx->comm_stream));
        S3_CUDA(cudaEventRecord(ctx->ev_comm_t[2 * view + 1], ctx->comm_stream));
        V.agg_d0 = 0; V.agg_d1 = D;
    }
    S3_CUDA(cudaEventRecord(ctx->ev_comm[1], ctx->comm_stream));
    S3_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_comm[1], 0));
    ctx->comm_timed = true;
    return 0;
}

int s3dmst_comm_transport(const s3dmst_ctx* ctx) { return !ctx || !ctx->comm ? -1 : (ctx->p2p_ok ? 1 : 0); }

double s3dmst_comm_minloc_ms(s3dmst_ctx* ctx) {
    if (!ctx->comm_timed) return -1.0;
    double total = 0.0;
    for (int view = 0; view < 2; view++) {
        float ms = 0.f;
        if (cudaEventSynchronize(ctx->ev_comm_t[2 * view + 1]) != cudaSuccess) { cudaGetLastError(); return -1.0; }
        if (cudaEventElapsedTime(&ms, ctx->ev_comm_t[2 * view], ctx->ev_comm_t[2 * view + 1]) != cudaSuccess) { cudaGetLastError(); return -1.0; }
        total += ms;
    }
    return total;
}

}  // extern "C"
