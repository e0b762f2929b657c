// stereomatch_b200/csrc/comm.cu — label-range sharding of ONE large pair over the GPUs of a box (BASELINE
// config C5): every rank builds the (deterministic) forests itself, aggregates its own label range, and the only
// exchange is a per-pixel MIN-LOC of (aggregated cost, disparity) — the reduction that replaces the serial
// `if (agg < min_cost)` over ascending labels of the reference (Stereo3DMST.cpp:173-185; dense mode: SURVEY A13).
//
// NCCL has no MINLOC and the exact-mode cost is fp64, so it is two all-reduces on the context's communication stream:
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
enum { s3_ncclInt32 = 2, s3_ncclFloat64 = 8, s3_ncclMin = 3 };
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(s3_ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, s3_ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
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

// MIN-LOC all-reduce of view's (best, disp_i) over the communicator, queued on `st` (no host synchronisation)
static int reduce_minloc_on(s3dmst_ctx* ctx, int view, cudaStream_t st) {
    NcclApi* api = nccl_api();
    View& V = ctx->v[view];
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
