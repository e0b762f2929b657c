// stereomatch_b200/csrc/api.cu — the C ABI declared in include/s3dmst.h: context, buffers, staging copies
// and the stage orchestration of stereo3dmst() (src/Stereo3DMST.cpp:714-912).  Device work only; the few
// host-side loops here are metadata (tree order, uploaded-forest bookkeeping, parity dumps).
#include <float.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "internal.h"

int s3_fail(s3dmst_ctx* c, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}


// Small host -> device metadata copies (unit lists, tree offsets, proposal lists) go through a pinned arena owned by the
// context: the copy is truly asynchronous and its source may die when this returns.  Two halves, each guarded by an
// event recorded behind the last copy queued from it; a half is reused only after its event has completed.
int s3_h2d_staged(s3dmst_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
    if (!bytes) return 0;
    const size_t need = (bytes + 255) / 256 * 256;
    if (need > ctx->stage_half) {  // (re)allocate: everything queued so far has to finish first
        S3_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->stage) S3_CUDA(cudaFreeHost(ctx->stage));
        ctx->stage = nullptr;
        size_t half = std::max<size_t>(1 << 20, ctx->stage_half);
        while (half < need) half *= 2;
        S3_CUDA(cudaHostAlloc(&ctx->stage, 2 * half, cudaHostAllocDefault));
        ctx->stage_half = half;
        ctx->stage_used = 0;
        ctx->stage_cur = 0;
        ctx->stage_pending[0] = ctx->stage_pending[1] = false;
        for (int i = 0; i < 2; i++)
            if (!ctx->stage_ev[i]) S3_CUDA(cudaEventCreateWithFlags(&ctx->stage_ev[i], cudaEventDisableTiming));
    }
    if (ctx->stage_used + need > ctx->stage_half) {  // this half is full: mark it, move to the other one once it has drained
        S3_CUDA(cudaEventRecord(ctx->stage_ev[ctx->stage_cur], ctx->stream));
        ctx->stage_pending[ctx->stage_cur] = true;
        ctx->stage_cur ^= 1;
        if (ctx->stage_pending[ctx->stage_cur]) {
            S3_CUDA(cudaEventSynchronize(ctx->stage_ev[ctx->stage_cur]));
            ctx->stage_pending[ctx->stage_cur] = false;
        }
        ctx->stage_used = 0;
    }
    char* p = ctx->stage + (size_t)ctx->stage_cur * ctx->stage_half + ctx->stage_used;
    memcpy(p, src_host, bytes);
    ctx->stage_used += need;
    S3_CUDA(cudaMemcpyAsync(dst_dev, p, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

static thread_local std::string g_create_err;

template <class T>
static cudaError_t dalloc(T** p, size_t n) {
    return cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T));
}
#define DFREE(p)          \
    do {                  \
        if (p) cudaFree(p); \
        p = nullptr;      \
    } while (0)

static void free_view(View& V) {
    DFREE(V.bgr); DFREE(V.raw4); DFREE(V.med); DFREE(V.gray); DFREE(V.ew); DFREE(V.pgrad); V.plane_ready = false;
    DFREE(V.match8); DFREE(V.uf_comp); DFREE(V.uf_parent); DFREE(V.adjw); DFREE(V.bfs_front); DFREE(V.fh_ent[0]); DFREE(V.fh_ent[1]); DFREE(V.uf_resv);
    DFREE(V.mask); DFREE(V.elist); DFREE(V.e_ra); DFREE(V.e_rb); DFREE(V.e_flag);
    DFREE(V.hist); DFREE(V.lvl_off); DFREE(V.lvl_cursor); DFREE(V.counters);
    DFREE(V.minpix); DFREE(V.scan_tmp); DFREE(V.tree_id); DFREE(V.tree_size); DFREE(V.tree_rootpix);
    DFREE(V.tree_start); DFREE(V.tree_depth); DFREE(V.unit_tree);
    DFREE(V.node_pixel); DFREE(V.pixel_node); DFREE(V.parent); DFREE(V.level); DFREE(V.pw); DFREE(V.node_up); DFREE(V.node_dn); DFREE(V.leaf_bits);
    DFREE(V.lvl_start);
    DFREE(V.cost); DFREE(V.aup); V.cost_cap = V.aup_cap = 0;
    DFREE(V.disp_i); DFREE(V.best); DFREE(V.abc); DFREE(V.min_cost); DFREE(V.disp_f); DFREE(V.lr_mask);
    DFREE(V.adj_ptr); DFREE(V.adj); V.adj_ptr_cap = V.adj_cap = 0; V.n_adj = 0; V.adj_ready = false;
    V.forest_ready = V.cost_ready = V.agg_ready = V.labels_ready = false;
    V.T = 0; V.D = V.Dp = 0;
}

static int alloc_view(s3dmst_ctx* ctx, View& V, int N) {
    const size_t n = N;
    S3_CUDA(dalloc(&V.bgr, 3 * n)); S3_CUDA(dalloc(&V.raw4, n)); S3_CUDA(dalloc(&V.med, n)); S3_CUDA(dalloc(&V.gray, n));
    S3_CUDA(dalloc(&V.ew, 2 * n));
    { unsigned char* b = nullptr; S3_CUDA(dalloc(&b, 16 * n)); V.uf_comp = b; }
    S3_CUDA(dalloc(&V.uf_parent, n));
    S3_CUDA(dalloc(&V.adjw, n)); S3_CUDA(dalloc(&V.bfs_front, n));
    S3_CUDA(dalloc(&V.uf_resv, n));
    for (int i = 0; i < 2; i++) { unsigned char* b = nullptr; S3_CUDA(dalloc(&b, 32 * n + 16 * (size_t)S3_FH_MAX_CTAS * (S3_FH_SEG_SLACK + 1))); V.fh_ent[i] = b; }
    S3_CUDA(dalloc(&V.mask, 2 * n)); S3_CUDA(dalloc(&V.elist, 2 * n)); S3_CUDA(dalloc(&V.e_ra, 2 * n));
    S3_CUDA(dalloc(&V.e_rb, 2 * n)); S3_CUDA(dalloc(&V.e_flag, 2 * n));
    S3_CUDA(dalloc(&V.hist, S3_NUM_W)); S3_CUDA(dalloc(&V.lvl_off, S3_NUM_W + 1)); S3_CUDA(dalloc(&V.lvl_cursor, S3_NUM_W));
    S3_CUDA(dalloc(&V.counters, S3_MAX_ROUNDS));
    S3_CUDA(dalloc(&V.minpix, n)); S3_CUDA(dalloc(&V.scan_tmp, n + 4096)); S3_CUDA(dalloc(&V.tree_id, n));
    S3_CUDA(dalloc(&V.tree_size, n)); S3_CUDA(dalloc(&V.tree_rootpix, n));
    S3_CUDA(dalloc(&V.tree_start, n + 1)); S3_CUDA(dalloc(&V.tree_depth, n)); S3_CUDA(dalloc(&V.unit_tree, n));
    S3_CUDA(dalloc(&V.node_pixel, n)); S3_CUDA(dalloc(&V.pixel_node, n)); S3_CUDA(dalloc(&V.parent, n));
    S3_CUDA(dalloc(&V.level, n)); S3_CUDA(dalloc(&V.pw, n)); S3_CUDA(dalloc(&V.node_up, n)); S3_CUDA(dalloc(&V.node_dn, n)); S3_CUDA(dalloc(&V.leaf_bits, n / 32 + 2));
    S3_CUDA(dalloc(&V.lvl_start, 2 * n + 2));
    S3_CUDA(dalloc(&V.disp_i, n)); S3_CUDA(dalloc(&V.best, n)); S3_CUDA(dalloc(&V.abc, 3 * n)); S3_CUDA(dalloc(&V.min_cost, n));
    S3_CUDA(dalloc(&V.disp_f, n)); S3_CUDA(dalloc(&V.lr_mask, n));
    return 0;
}

static int set_size(s3dmst_ctx* ctx, int W, int H) {
    if (W < 1 || H < 1 || (long long)W * H > (1ll << 27)) return s3_fail(ctx, S3DMST_E_ARG, "image size %dx%d unsupported", W, H);
    if (ctx->W == W && ctx->H == H) return 0;
    S3_TRY(s3_comm_before_free(ctx));  // result buffers mapped into the other ranks: they let go of them first (collective)
    for (int i = 0; i < 2; i++) free_view(ctx->v[i]);
    ctx->W = ctx->H = ctx->N = 0;
    ctx->forest_pending = 0;
    for (int i = 0; i < 2; i++) {
        const int r = alloc_view(ctx, ctx->v[i], W * H);
        if (r) {  // a failed allocation leaves a context without images, not one that claims buffers it does not have
            for (int k = 0; k < 2; k++) free_view(ctx->v[k]);
            return r;
        }
    }
    ctx->W = W; ctx->H = H; ctx->N = W * H;
    return 0;
}

extern "C" {

void s3dmst_default_params(s3dmst_params* p) {
    p->fh_c = 5000.0f;
    p->min_cc_size = 200;
    p->gamma = 1.0f / 12.f;
    p->median = 3;
    p->cost_cap = 0.5f;
    p->cost_offset = 0.0f;
    p->cost_scale = 1.0f;
    p->oob_cost = 0.5f;
    p->num_iter = 100;
    p->refine_floor = 0.1f;
    p->exact = 1;
    p->keep_aggregated = 0;
    p->agg_threads = 0;
    p->agg_cache_nodes = 0;
    p->agg_ring_nodes = 0;
    p->agg_kernel = 0;
    p->fh_ctas = 0;
    p->fh_threads = 0;
    p->agg_cluster_nodes = 0;
    p->fuse_cost = 0;
    p->comm_p2p = 0;
    p->fh_cluster = 0;
    p->pms_cost_mode = 0;
    p->pm_alpha = 0.9f;
    p->pm_tau_c = 10.0f;
    p->pm_tau_g = 2.0f;
}

int s3dmst_create(s3dmst_ctx** out, int device, const s3dmst_params* params, void* stream) {
    if (!out) return S3DMST_E_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
        g_create_err = std::string("s3dmst_create: no usable CUDA device (") + cudaGetErrorString(e) + ")";
        return S3DMST_E_CUDA;  // no CPU fallback
    }
    s3dmst_ctx* ctx = new s3dmst_ctx;
    ctx->device = device;
    if (params) ctx->P = *params; else s3dmst_default_params(&ctx->P);
    memset(ctx->ev, 0, sizeof ctx->ev);
    memset(ctx->ev_set, 0, sizeof ctx->ev_set);
    memset(ctx->ev_slot, 0, sizeof ctx->ev_slot);
    memset(ctx->ev_state, 0, sizeof ctx->ev_state);
    memset(ctx->ev_acc_ms, 0, sizeof ctx->ev_acc_ms);
    memset(ctx->ev_acc_n, 0, sizeof ctx->ev_acc_n);
    memset(ctx->ev_lost, 0, sizeof ctx->ev_lost);
    memset(ctx->p2p_best, 0, sizeof ctx->p2p_best);
    memset(ctx->p2p_disp, 0, sizeof ctx->p2p_disp);
    memset(ctx->p2p_flags, 0, sizeof ctx->p2p_flags);
    bool ok = cudaSetDevice(device) == cudaSuccess;
    cudaDeviceProp prop;
    ok = ok && cudaGetDeviceProperties(&prop, device) == cudaSuccess;
    if (ok) ctx->num_sms = prop.multiProcessorCount;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else if (ok) {
        ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess;
        ctx->own_stream = ok;
    }
    {
        cudaEvent_t* evs = &ctx->ev[0][0][0][0];
        for (int i = 0; ok && i < S3DMST_T_COUNT * 2 * S3_EV_SLOTS * 2; i++) ok = cudaEventCreate(&evs[i]) == cudaSuccess;
    }
    if (ok) ok = cudaEventCreateWithFlags(&ctx->ev_xctx, cudaEventDisableTiming) == cudaSuccess;
    {
        // the launches of the large trees (clusters, 32-warp CTAs) are a launch set's critical path: their CTAs go first
        // when they compete with the small trees' launch for SMs (without it the FLIR pair's aggregation is 6.6 or 8.1 ms
        // from run to run, depending on which launch the block scheduler happens to serve first)
        int prio_lo = 0, prio_hi = 0;
        if (ok) ok = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) == cudaSuccess;
        if (ok) ok = cudaStreamCreateWithPriority(&ctx->stream_aux, cudaStreamNonBlocking, prio_hi) == cudaSuccess;
        if (ok) ok = cudaStreamCreateWithPriority(&ctx->stream_big, cudaStreamNonBlocking, prio_hi) == cudaSuccess;
    }
    if (ok) ok = cudaEventCreateWithFlags(&ctx->ev_join_big, cudaEventDisableTiming) == cudaSuccess;
    if (ok) ok = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    if (ok) ok = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) == cudaSuccess;
    if (ok) ok = cudaEventCreateWithFlags(&ctx->ev_block, cudaEventDisableTiming | cudaEventBlockingSync) == cudaSuccess;
    if (ok) {
        // weights (Stereo3DMST.cpp:444, :513): exp(-w*gamma) in double with gamma promoted from float
        std::vector<double> w(S3_NUM_W), w2(S3_NUM_W);
        std::vector<float> wf(S3_NUM_W), w2f(S3_NUM_W);
        for (int i = 0; i < S3_NUM_W; i++) {
            w[i] = exp(-(double)i * ctx->P.gamma);
            w2[i] = 1.0f - w[i] * w[i];
            wf[i] = (float)w[i];
            w2f[i] = (float)w2[i];
        }
        ok = dalloc(&ctx->lut_w, S3_NUM_W) == cudaSuccess && dalloc(&ctx->lut_w2, S3_NUM_W) == cudaSuccess &&
             dalloc(&ctx->lut_wf, S3_NUM_W) == cudaSuccess && dalloc(&ctx->lut_w2f, S3_NUM_W) == cudaSuccess;
        ok = ok && cudaMemcpy(ctx->lut_w, w.data(), sizeof(double) * S3_NUM_W, cudaMemcpyHostToDevice) == cudaSuccess;
        ok = ok && cudaMemcpy(ctx->lut_w2, w2.data(), sizeof(double) * S3_NUM_W, cudaMemcpyHostToDevice) == cudaSuccess;
        ok = ok && cudaMemcpy(ctx->lut_wf, wf.data(), sizeof(float) * S3_NUM_W, cudaMemcpyHostToDevice) == cudaSuccess;
        ok = ok && cudaMemcpy(ctx->lut_w2f, w2f.data(), sizeof(float) * S3_NUM_W, cudaMemcpyHostToDevice) == cudaSuccess;
    }
    if (!ok) {
        g_create_err = std::string("s3dmst_create: ") + cudaGetErrorString(cudaGetLastError());
        s3dmst_destroy(ctx);
        return S3DMST_E_CUDA;
    }
    *out = ctx;
    return 0;
}

void s3dmst_destroy(s3dmst_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    s3dmst_comm_destroy(ctx);
    if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
    for (int i = 0; i < 2; i++) if (ctx->ev_comm[i]) cudaEventDestroy(ctx->ev_comm[i]);
    for (int i = 0; i < 4; i++) if (ctx->ev_comm_t[i]) cudaEventDestroy(ctx->ev_comm_t[i]);
    DFREE(ctx->gmin);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 2; i++) free_view(ctx->v[i]);
    s3_rectify_free(ctx);
    DFREE(ctx->lut_w); DFREE(ctx->lut_w2); DFREE(ctx->lut_wf); DFREE(ctx->lut_w2f); DFREE(ctx->pms_scratch); DFREE(ctx->units_dev); DFREE(ctx->fh_sync); DFREE(ctx->abc_init);
    {
        cudaEvent_t* evs = &ctx->ev[0][0][0][0];
        for (int i = 0; i < S3DMST_T_COUNT * 2 * S3_EV_SLOTS * 2; i++)
            if (evs[i]) cudaEventDestroy(evs[i]);
    }
    if (ctx->ev_xctx) cudaEventDestroy(ctx->ev_xctx);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->stream_aux) { cudaStreamSynchronize(ctx->stream_aux); cudaStreamDestroy(ctx->stream_aux); }
    if (ctx->stream_big) { cudaStreamSynchronize(ctx->stream_big); cudaStreamDestroy(ctx->stream_big); }
    if (ctx->ev_join_big) cudaEventDestroy(ctx->ev_join_big);
    if (ctx->ev_block) cudaEventDestroy(ctx->ev_block);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    if (ctx->stage) cudaFreeHost(ctx->stage);
    for (int i = 0; i < 2; i++)
        if (ctx->stage_ev[i]) cudaEventDestroy(ctx->stage_ev[i]);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* s3dmst_last_error(const s3dmst_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int s3dmst_sync(s3dmst_ctx* ctx) {
    S3_CUDA(cudaSetDevice(ctx->device));
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

long long s3dmst_launch_count(const s3dmst_ctx* ctx) { return ctx->launches; }

double s3dmst_stage_ms(s3dmst_ctx* ctx, int stage) {
    if (stage < 0 || stage >= S3DMST_T_COUNT) return -1.0;
    double total = 0.0;
    for (int view = 0; view < 2; view++) {
        if (!ctx->ev_set[stage][view]) continue;
        cudaEvent_t* p = ctx->ev[stage][view][ctx->ev_slot[stage][view]];
        float ms = 0.f;
        if (cudaEventSynchronize(p[1]) != cudaSuccess) { cudaGetLastError(); return -1.0; }
        if (cudaEventElapsedTime(&ms, p[0], p[1]) != cudaSuccess) { cudaGetLastError(); return -1.0; }
        total += ms;
    }
    return total;
}

// completed pairs -> accumulators (wait: also the ones still running)
extern "C++" void s3_ev_harvest(s3dmst_ctx* ctx, int stage, int view, bool wait) {
    for (int sl = 0; sl < S3_EV_SLOTS; sl++) {
        if (!ctx->ev_state[stage][view][sl]) continue;
        cudaEvent_t* p = ctx->ev[stage][view][sl];
        const cudaError_t q = wait ? cudaEventSynchronize(p[1]) : cudaEventQuery(p[1]);
        if (q != cudaSuccess) { cudaGetLastError(); continue; }
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p[0], p[1]) == cudaSuccess) { ctx->ev_acc_ms[stage] += ms; ctx->ev_acc_n[stage]++; }
        else cudaGetLastError();
        ctx->ev_state[stage][view][sl] = 0;
    }
}

double s3dmst_stage_total_ms(s3dmst_ctx* ctx, int stage, int* samples, int reset) {
    if (!ctx || stage < 0 || stage >= S3DMST_T_COUNT) return -1.0;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return -1.0;
    for (int view = 0; view < 2; view++) s3_ev_harvest(ctx, stage, view, true);
    const double total = ctx->ev_acc_ms[stage];
    if (samples) *samples = ctx->ev_acc_n[stage];
    if (reset) { ctx->ev_acc_ms[stage] = 0.0; ctx->ev_acc_n[stage] = 0; ctx->ev_lost[stage] = 0; }
    return total;
}

static int set_images_impl(s3dmst_ctx* ctx, const uint8_t* left_bgr, const uint8_t* right_bgr, int W, int H, int stride, bool sync);
int s3dmst_set_images(s3dmst_ctx* ctx, const uint8_t* left_bgr, const uint8_t* right_bgr, int W, int H, int stride) {
    return set_images_impl(ctx, left_bgr, right_bgr, W, H, stride, true);
}
int s3dmst_set_images_async(s3dmst_ctx* ctx, const uint8_t* left_bgr, const uint8_t* right_bgr, int W, int H, int stride) {
    return set_images_impl(ctx, left_bgr, right_bgr, W, H, stride, false);
}
static int set_images_impl(s3dmst_ctx* ctx, const uint8_t* left_bgr, const uint8_t* right_bgr, int W, int H, int stride, bool sync) {
    if (!left_bgr || !right_bgr || stride < 3 * W) return s3_fail(ctx, S3DMST_E_ARG, "set_images: bad pointers/stride");
    S3_CUDA(cudaSetDevice(ctx->device));
    S3_TRY(set_size(ctx, W, H));
    const uint8_t* src[2] = {left_bgr, right_bgr};
    for (int i = 0; i < 2; i++) {
        S3_CUDA(cudaMemcpy2DAsync(ctx->v[i].bgr, 3 * (size_t)W, src[i], stride, 3 * (size_t)W, H, cudaMemcpyHostToDevice,
                                  ctx->stream));
        ctx->v[i].forest_ready = ctx->v[i].cost_ready = ctx->v[i].agg_ready = ctx->v[i].plane_ready = false;
    }
    ctx->forest_pending = 0;
    ctx->fused_D = 0;
    if (sync) S3_CUDA(cudaStreamSynchronize(ctx->stream));  // the host buffers may be pageable / reused
    return 0;
}

void s3dmst_remap_table(int16_t* tab) { s3_remap_table(tab); }

int s3dmst_set_rectify_maps(s3dmst_ctx* ctx, int view, const int16_t* map_xy, const uint16_t* map_fxy, int W, int H) {
    if (view < 0 || view > 1 || !map_xy || !map_fxy || W < 1 || H < 1) return s3_fail(ctx, S3DMST_E_ARG, "set_rectify_maps: bad arguments");
    S3_CUDA(cudaSetDevice(ctx->device));
    return s3_set_rectify_maps(ctx, view, map_xy, map_fxy, W, H);
}

int s3dmst_set_raw_images(s3dmst_ctx* ctx, const uint8_t* left_raw_bgr, const uint8_t* right_raw_bgr, int src_w, int src_h, int stride) {
    if (!left_raw_bgr || !right_raw_bgr || src_w < 1 || src_h < 1 || stride < 3 * src_w) return s3_fail(ctx, S3DMST_E_ARG, "set_raw_images: bad pointers/stride");
    if (!ctx->map_xy[0] || !ctx->map_xy[1]) return s3_fail(ctx, S3DMST_E_STATE, "set_raw_images: rectification maps of both views required");
    if (ctx->map_w[0] != ctx->map_w[1] || ctx->map_h[0] != ctx->map_h[1]) return s3_fail(ctx, S3DMST_E_ARG, "set_raw_images: the two views' maps differ in size");
    S3_CUDA(cudaSetDevice(ctx->device));
    S3_TRY(set_size(ctx, ctx->map_w[0], ctx->map_h[0]));
    for (int i = 0; i < 2; i++) ctx->v[i].forest_ready = ctx->v[i].cost_ready = ctx->v[i].agg_ready = false;
    ctx->fused_D = 0;
    S3_TRY(s3_remap_raw_pair(ctx, left_raw_bgr, right_raw_bgr, src_w, src_h, stride));
    S3_CUDA(cudaStreamSynchronize(ctx->stream));  // the host buffers may be pageable / reused
    return 0;
}

int s3dmst_get_image(s3dmst_ctx* ctx, int view, uint8_t* bgr) {
    if (view < 0 || view > 1 || !bgr || ctx->N == 0) return s3_fail(ctx, S3DMST_E_ARG, "get_image: bad view or no images");
    S3_CUDA(cudaSetDevice(ctx->device));
    S3_CUDA(cudaMemcpyAsync(bgr, ctx->v[view].bgr, 3 * (size_t)ctx->N, cudaMemcpyDeviceToHost, ctx->stream));
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int s3dmst_build_forest(s3dmst_ctx* ctx, int view) {
    if (view < 0 || view > 1 || ctx->N == 0) return s3_fail(ctx, S3DMST_E_ARG, "build_forest: bad view or no images");
    S3_CUDA(cudaSetDevice(ctx->device));
    S3_EV_BEGIN(S3DMST_T_FOREST, view);
    S3_TRY(s3_forest_stage(ctx, view));
    S3_EV_END(S3DMST_T_FOREST, view);
    return 0;
}

int s3dmst_forest_info(s3dmst_ctx* ctx, int view, int* num_trees, int* max_depth, int* adj_size);

// tree adjacency for dumps (Stereo3DMST.cpp:377-384): the device CSR (pms.cu) copied to the host
static int host_adjacency(s3dmst_ctx* ctx, int view, std::vector<int>& adj_ptr, std::vector<int>& adj) {
    View& V = ctx->v[view];
    S3_TRY(s3_tree_adjacency(ctx, view));
    adj_ptr.resize(V.T + 1);
    adj.resize(V.n_adj);
    S3_CUDA(cudaMemcpyAsync(adj_ptr.data(), V.adj_ptr, sizeof(int) * (V.T + 1), cudaMemcpyDeviceToHost, ctx->stream));
    if (V.n_adj) S3_CUDA(cudaMemcpyAsync(adj.data(), V.adj, sizeof(int) * V.n_adj, cudaMemcpyDeviceToHost, ctx->stream));
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int s3dmst_forest_info(s3dmst_ctx* ctx, int view, int* num_trees, int* max_depth, int* adj_size) {
    if (view < 0 || view > 1) return s3_fail(ctx, S3DMST_E_ARG, "bad view");
    View& V = ctx->v[view];
    if (!V.forest_ready) return s3_fail(ctx, S3DMST_E_STATE, "forest_info: no forest");
    S3_CUDA(cudaSetDevice(ctx->device));
    S3_TRY(s3_forest_finish_host(ctx));
    if (num_trees) *num_trees = V.T;
    if (max_depth) {
        S3_TRY(s3_forest_depths(ctx, view));
        *max_depth = V.max_depth;
    }
    if (adj_size) {
        S3_CUDA(cudaSetDevice(ctx->device));
        S3_TRY(s3_tree_adjacency(ctx, view));
        *adj_size = V.n_adj;
    }
    return 0;
}

#define D2H(dst, src, bytes)                                                                          \
    do {                                                                                              \
        if (dst) S3_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));       \
    } while (0)

int s3dmst_get_forest(s3dmst_ctx* ctx, int view, uint16_t* edge_weight, uint8_t* edge_mask, int32_t* tree_id,
                      int32_t* tree_start, int32_t* node_pixel, int32_t* parent, int32_t* child_begin,
                      int32_t* child_count, uint16_t* parent_weight, int32_t* level, int32_t* adj_ptr, int32_t* adj) {
    if (view < 0 || view > 1) return s3_fail(ctx, S3DMST_E_ARG, "bad view");
    View& V = ctx->v[view];
    if (!V.forest_ready) return s3_fail(ctx, S3DMST_E_STATE, "get_forest: no forest");
    S3_CUDA(cudaSetDevice(ctx->device));
    S3_TRY(s3_forest_finish_host(ctx));
    const size_t N = ctx->N;
    D2H(edge_weight, V.ew, 2 * N * sizeof(uint16_t));
    D2H(edge_mask, V.mask, 2 * N);
    D2H(tree_id, V.tree_id, N * sizeof(int));
    D2H(tree_start, V.tree_start, (V.T + 1) * sizeof(int));
    D2H(node_pixel, V.node_pixel, N * sizeof(int));
    D2H(parent, V.parent, N * sizeof(int));
    D2H(parent_weight, V.pw, N * sizeof(uint16_t));
    D2H(level, V.level, N * sizeof(int));
    if (child_begin || child_count) {
        std::vector<NodeUp> nu(N);
        S3_CUDA(cudaMemcpyAsync(nu.data(), V.node_up, N * sizeof(NodeUp), cudaMemcpyDeviceToHost, ctx->stream));
        S3_CUDA(cudaStreamSynchronize(ctx->stream));
        for (size_t i = 0; i < N; i++) {
            if (child_begin) child_begin[i] = nu[i].child_begin;
            if (child_count) child_count[i] = nu[i].child_count & 7;
        }
    }
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    if (adj_ptr || adj) {
        std::vector<int> ap, a;
        S3_TRY(host_adjacency(ctx, view, ap, a));
        if (adj_ptr) memcpy(adj_ptr, ap.data(), ap.size() * sizeof(int));
        if (adj) memcpy(adj, a.data(), a.size() * sizeof(int));
    }
    return 0;
}

int s3dmst_set_forest(s3dmst_ctx* ctx, int view, int W, int H, int T, const int32_t* tree_start, const int32_t* node_pixel,
                      const int32_t* parent, const uint16_t* parent_weight) {
    if (view < 0 || view > 1 || T < 1 || !tree_start || !node_pixel || !parent || !parent_weight)
        return s3_fail(ctx, S3DMST_E_ARG, "set_forest: bad arguments");
    S3_CUDA(cudaSetDevice(ctx->device));
    if (ctx->N == 0) S3_TRY(set_size(ctx, W, H));
    if (W != ctx->W || H != ctx->H) return s3_fail(ctx, S3DMST_E_ARG, "set_forest: size differs from the context's images");
    View& V = ctx->v[view];
    const int N = ctx->N;
    if (tree_start[T] != N) return s3_fail(ctx, S3DMST_E_ARG, "set_forest: tree_start[T] != W*H");
    // derive the per-node records the kernels read (children contiguous in BFS order)
    std::vector<NodeUp> nu(N);
    std::vector<int> level(N, 0), pixel_node(N, -1), tree_id(N, 0), lvl(2 * (size_t)N + 2, 0), depth(T, 0);
    for (int i = 0; i < N; i++) { nu[i].child_begin = 0; nu[i].child_count = 0; nu[i].cw01 = nu[i].cw23 = 0; }
    for (int t = 0; t < T; t++) {
        const int a = tree_start[t], b = tree_start[t + 1];
        if (b <= a || parent[a] != a) return s3_fail(ctx, S3DMST_E_ARG, "set_forest: tree %d malformed", t);
        for (int g = a; g < b; g++) {
            if (node_pixel[g] < 0 || node_pixel[g] >= N) return s3_fail(ctx, S3DMST_E_ARG, "set_forest: pixel out of range");
            pixel_node[node_pixel[g]] = g;
            tree_id[node_pixel[g]] = t;
            if (g == a) continue;
            const int p = parent[g];
            if (p < a || p >= g) return s3_fail(ctx, S3DMST_E_ARG, "set_forest: node %d is not in BFS order", g);
            level[g] = level[p] + 1;
            if (level[g] < level[g - 1]) return s3_fail(ctx, S3DMST_E_ARG, "set_forest: levels not monotone at %d", g);
            NodeUp& P = nu[p];
            if (P.child_count == 0) P.child_begin = g;
            if (P.child_begin + P.child_count != g || P.child_count >= 4)
                return s3_fail(ctx, S3DMST_E_ARG, "set_forest: children of %d not contiguous / more than 4", p);
            const uint32_t wv = parent_weight[g];
            if (wv >= S3_NUM_W) return s3_fail(ctx, S3DMST_E_ARG, "set_forest: weight out of range");
            if (P.child_count < 2) P.cw01 |= wv << (16 * P.child_count); else P.cw23 |= wv << (16 * (P.child_count - 2));
            P.child_count++;
        }
        int* L = lvl.data() + a + t;
        int d = 0;
        L[0] = a;
        for (int g = a + 1; g < b; g++)
            if (level[g] != level[g - 1]) L[++d] = g;
        L[++d] = b;
        depth[t] = d;
    }
    std::vector<int> order(T);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return tree_start[x + 1] - tree_start[x] > tree_start[y + 1] - tree_start[y]; });
    std::vector<int> rootpix(T);
    for (int t = 0; t < T; t++) rootpix[t] = node_pixel[tree_start[t]];
#define H2D(dst, src, bytes) S3_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream))
    H2D(V.tree_start, tree_start, sizeof(int) * (T + 1));
    H2D(V.node_pixel, node_pixel, sizeof(int) * N);
    H2D(V.parent, parent, sizeof(int) * N);
    H2D(V.pw, parent_weight, sizeof(uint16_t) * N);
    std::vector<int4> nd(N);
    for (int i = 0; i < N; i++) {  // (the records are uploaded only once they are complete)
        if (parent[i] != i && i - parent[i] >= S3_AGG_NEAR) nu[i].child_count |= S3_NU_FARPARENT;
        const bool far_child = (nu[i].child_count & 7) > 0 && nu[i].child_begin + (nu[i].child_count & 7) - 1 - i >= S3_AGG_NEAR;
        nd[i] = make_int4(parent[i], parent_weight[i], level[i] | (far_child ? S3_ND_FAR : 0) | ((nu[i].child_count & 7) == 0 ? S3_ND_LEAF : 0), node_pixel[i]);
    }
    H2D(V.node_up, nu.data(), sizeof(NodeUp) * N);
    H2D(V.node_dn, nd.data(), sizeof(int4) * N);
    std::vector<uint32_t> lbits(N / 32 + 2, 0u);
    for (int i = 0; i < N; i++)
        if (nd[i].z & S3_ND_LEAF) lbits[i >> 5] |= 1u << (i & 31);
    H2D(V.leaf_bits, lbits.data(), sizeof(uint32_t) * lbits.size());
    H2D(V.level, level.data(), sizeof(int) * N);
    H2D(V.pixel_node, pixel_node.data(), sizeof(int) * N);
    H2D(V.tree_id, tree_id.data(), sizeof(int) * N);
    H2D(V.lvl_start, lvl.data(), sizeof(int) * (N + T + 1));
    H2D(V.tree_depth, depth.data(), sizeof(int) * T);
    H2D(V.unit_tree, order.data(), sizeof(int) * T);
    H2D(V.counters + (S3_MAX_ROUNDS - 64), &T, sizeof(int));  // forest.cu CNT_T: the device-side tree count
    H2D(V.tree_rootpix, rootpix.data(), sizeof(int) * T);
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    V.T = T;
    V.h_tree_start.assign(tree_start, tree_start + T + 1);
    return s3_forest_finalize_host(ctx, view);
}

int s3dmst_build_cost_volume(s3dmst_ctx* ctx, int D, int apply_ingest) {
    S3_CUDA(cudaSetDevice(ctx->device));
    return s3_cost_adgrad(ctx, D, apply_ingest);
}

int s3dmst_set_cost_volume(s3dmst_ctx* ctx, int view, const float* vol, int D, int apply_ingest) {
    if (view < 0 || view > 1 || !vol) return s3_fail(ctx, S3DMST_E_ARG, "set_cost_volume: bad arguments");
    S3_CUDA(cudaSetDevice(ctx->device));
    ctx->fused_D = 0;
    float* stage = nullptr;
    const size_t bytes = (size_t)ctx->N * D * sizeof(float);
    S3_CUDA(cudaMalloc(&stage, bytes));
    cudaError_t e = cudaMemcpyAsync(stage, vol, bytes, cudaMemcpyHostToDevice, ctx->stream);
    int rc = e == cudaSuccess ? s3_cost_from_dmajor(ctx, view, stage, D, apply_ingest) : s3_fail(ctx, S3DMST_E_CUDA, "H2D: %s", cudaGetErrorString(e));
    cudaStreamSynchronize(ctx->stream);
    cudaFree(stage);
    return rc;
}

int s3dmst_get_cost_volume(s3dmst_ctx* ctx, int view, float* vol) {
    S3_CUDA(cudaSetDevice(ctx->device));
    if (view < 0 || view > 1 || !vol) return s3_fail(ctx, S3DMST_E_ARG, "get_cost_volume: bad arguments");
    View& V = ctx->v[view];
    S3_TRY(s3_materialize_cost(ctx));  // a dense run that computed its cost on the fly left none
    if (!V.cost_ready) return s3_fail(ctx, S3DMST_E_STATE, "get_cost_volume: no volume");
    float* stage = nullptr;
    const size_t bytes = (size_t)ctx->N * V.D * sizeof(float);
    S3_CUDA(cudaMalloc(&stage, bytes));
    int rc = s3_cost_to_dmajor(ctx, view, stage);
    if (rc == 0) {
        cudaError_t e = cudaMemcpyAsync(vol, stage, bytes, cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) rc = s3_fail(ctx, S3DMST_E_CUDA, "D2H: %s", cudaGetErrorString(e));
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(stage);
    return rc;
}

int s3dmst_aggregate_dense(s3dmst_ctx* ctx, int view, int d0, int d1, int32_t* disp, double* best_cost) {
    if (view < 0 || view > 1) return s3_fail(ctx, S3DMST_E_ARG, "bad view");
    S3_CUDA(cudaSetDevice(ctx->device));
    S3_TRY(s3_materialize_cost(ctx));
    {
        int rc = ctx->P.agg_kernel == 1 ? 1 : s3_aggregate_flow(ctx, 1 << view, d0, d1);
        if (rc == 1) rc = s3_aggregate_dense(ctx, view, d0, d1);  // simple kernel: any even d0, any depth
        if (rc) return rc;
    }
    View& V = ctx->v[view];
    D2H(disp, V.disp_i, sizeof(int32_t) * ctx->N);
    D2H(best_cost, V.best, sizeof(double) * ctx->N);
    if (disp || best_cost) S3_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int s3dmst_get_aggregated(s3dmst_ctx* ctx, int view, double* agg) {
    S3_CUDA(cudaSetDevice(ctx->device));
    if (view < 0 || view > 1 || !agg) return s3_fail(ctx, S3DMST_E_ARG, "bad arguments");
    View& V = ctx->v[view];
    if (!V.agg_ready || !ctx->P.keep_aggregated) return s3_fail(ctx, S3DMST_E_STATE, "get_aggregated: needs keep_aggregated and a dense run");
    const size_t N = ctx->N;
    std::vector<double> h(N * V.Dp);
    std::vector<int> np(N);
    const bool f32 = !ctx->P.exact && ctx->P.agg_kernel == 0;  // the fast mode keeps its running sums in fp32
    S3_CUDA(cudaMemcpyAsync(h.data(), V.aup, h.size() * (f32 ? sizeof(float) : sizeof(double)), cudaMemcpyDeviceToHost, ctx->stream));
    S3_CUDA(cudaMemcpyAsync(np.data(), V.node_pixel, N * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    const float* hf = reinterpret_cast<const float*>(h.data());
    for (size_t v = 0; v < N; v++)
        for (int d = 0; d < V.D; d++) agg[(size_t)d * N + np[v]] = f32 ? (double)hf[v * V.Dp + d] : h[v * V.Dp + d];
    return 0;
}

int s3dmst_dense_result_dev(s3dmst_ctx* ctx, int view, double** best_cost_dev, int32_t** disp_dev) {
    if (view < 0 || view > 1) return s3_fail(ctx, S3DMST_E_ARG, "bad view");
    if (best_cost_dev) *best_cost_dev = ctx->v[view].best;
    if (disp_dev) *disp_dev = ctx->v[view].disp_i;
    return 0;
}

int s3dmst_minloc_mask(s3dmst_ctx* ctx, int view, const double* global_min_dev) {
    S3_CUDA(cudaSetDevice(ctx->device));
    if (view < 0 || view > 1 || !global_min_dev) return s3_fail(ctx, S3DMST_E_ARG, "bad arguments");
    return s3_minloc_mask(ctx, view, global_min_dev);
}

int s3dmst_dense_to_disparity(s3dmst_ctx* ctx, int view) {
    S3_CUDA(cudaSetDevice(ctx->device));
    if (view < 0 || view > 1) return s3_fail(ctx, S3DMST_E_ARG, "bad view");
    return s3_dense_to_disp(ctx, view);
}

int s3dmst_prepare_plane_cost(s3dmst_ctx* ctx, int Dmax) {
    if (ctx->N == 0 || Dmax <= 0) return s3_fail(ctx, S3DMST_E_ARG, "prepare_plane_cost: images and Dmax > 0 required");
    S3_CUDA(cudaSetDevice(ctx->device));
    return s3_prepare_plane_cost(ctx, Dmax);
}
int s3dmst_get_plane_gradients(s3dmst_ctx* ctx, int view, float* grad) {
    if (view < 0 || view > 1 || !grad || !ctx->v[view].plane_ready) return s3_fail(ctx, S3DMST_E_STATE, "get_plane_gradients: s3dmst_prepare_plane_cost first");
    S3_CUDA(cudaSetDevice(ctx->device));
    S3_CUDA(cudaMemcpyAsync(grad, ctx->v[view].pgrad, sizeof(float) * 2 * (size_t)ctx->N, cudaMemcpyDeviceToHost, ctx->stream));
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int s3dmst_set_labels(s3dmst_ctx* ctx, int view, const float* abc) {
    if (view < 0 || view > 1 || !abc || ctx->N == 0) return s3_fail(ctx, S3DMST_E_ARG, "set_labels: bad arguments");
    S3_CUDA(cudaSetDevice(ctx->device));
    S3_CUDA(cudaMemcpyAsync(ctx->v[view].abc, abc, sizeof(float) * 3 * ctx->N, cudaMemcpyHostToDevice, ctx->stream));
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->v[view].labels_ready = true;
    ctx->pms_round[view] = 0;
    return 0;
}
int s3dmst_get_labels(s3dmst_ctx* ctx, int view, float* abc) {
    S3_CUDA(cudaSetDevice(ctx->device));
    if (view < 0 || view > 1 || !abc) return s3_fail(ctx, S3DMST_E_ARG, "bad arguments");
    D2H(abc, ctx->v[view].abc, sizeof(float) * 3 * ctx->N);
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int s3dmst_reset_min_cost(s3dmst_ctx* ctx, int view) {
    if (view < 0 || view > 1 || ctx->N == 0) return s3_fail(ctx, S3DMST_E_ARG, "bad arguments");
    std::vector<double> h(ctx->N, DBL_MAX);  // Stereo3DMST.cpp:820-821
    S3_CUDA(cudaSetDevice(ctx->device));
    S3_CUDA(cudaMemcpyAsync(ctx->v[view].min_cost, h.data(), sizeof(double) * ctx->N, cudaMemcpyHostToDevice, ctx->stream));
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int s3dmst_get_min_cost(s3dmst_ctx* ctx, int view, double* mc) {
    S3_CUDA(cudaSetDevice(ctx->device));
    if (view < 0 || view > 1 || !mc) return s3_fail(ctx, S3DMST_E_ARG, "bad arguments");
    D2H(mc, ctx->v[view].min_cost, sizeof(double) * ctx->N);
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int s3dmst_pms_apply(s3dmst_ctx* ctx, int view, const int32_t* tree_ids, const float* labels, size_t n) {
    if (view < 0 || view > 1 || (n && (!tree_ids || !labels))) return s3_fail(ctx, S3DMST_E_ARG, "pms_apply: bad arguments");
    S3_CUDA(cudaSetDevice(ctx->device));
    return s3_pms_apply(ctx, view, tree_ids, labels, n);
}

int s3dmst_init_labels(s3dmst_ctx* ctx, int view, int Dmax) {
    if (view < 0 || view > 1) return s3_fail(ctx, S3DMST_E_ARG, "bad view");
    S3_CUDA(cudaSetDevice(ctx->device));
    return s3_init_labels(ctx, view, Dmax);
}

int s3dmst_pms_iterate(s3dmst_ctx* ctx, int view, int n_iter, unsigned seed) {
    if (view < 0 || view > 1 || n_iter < 0) return s3_fail(ctx, S3DMST_E_ARG, "pms_iterate: bad arguments");
    S3_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->v[view].forest_ready) return s3_fail(ctx, S3DMST_E_STATE, "pms_iterate: no forest");
    if (!ctx->v[view].labels_ready) S3_TRY(s3_init_labels(ctx, view, ctx->v[view].D));  // :390-430
    return s3_pms_iterate(ctx, view, n_iter, seed);
}

int s3dmst_label_to_disp(s3dmst_ctx* ctx, int view) {
    S3_CUDA(cudaSetDevice(ctx->device));
    if (view < 0 || view > 1) return s3_fail(ctx, S3DMST_E_ARG, "bad view");
    return s3_label_to_disp(ctx, view);
}

int s3dmst_set_disparity(s3dmst_ctx* ctx, int view, const float* disp) {
    S3_CUDA(cudaSetDevice(ctx->device));
    if (view < 0 || view > 1 || !disp || ctx->N == 0) return s3_fail(ctx, S3DMST_E_ARG, "bad arguments");
    S3_CUDA(cudaMemcpyAsync(ctx->v[view].disp_f, disp, sizeof(float) * ctx->N, cudaMemcpyHostToDevice, ctx->stream));
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int s3dmst_get_disparity(s3dmst_ctx* ctx, int view, float* disp) {
    S3_CUDA(cudaSetDevice(ctx->device));
    if (view < 0 || view > 1 || !disp) return s3_fail(ctx, S3DMST_E_ARG, "bad arguments");
    D2H(disp, ctx->v[view].disp_f, sizeof(float) * ctx->N);
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int s3dmst_lr_check(s3dmst_ctx* ctx, int fill) {
    S3_CUDA(cudaSetDevice(ctx->device));
    return s3_lr_check(ctx, fill);
}

int s3dmst_get_lr_mask(s3dmst_ctx* ctx, uint8_t* mask) {
    if (!mask || ctx->N == 0) return s3_fail(ctx, S3DMST_E_ARG, "get_lr_mask: bad arguments");
    S3_CUDA(cudaSetDevice(ctx->device));
    S3_CUDA(cudaMemcpyAsync(mask, ctx->v[0].lr_mask, (size_t)ctx->N, cudaMemcpyDeviceToHost, ctx->stream));
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int s3dmst_weighted_median(s3dmst_ctx* ctx, int view, int radius, float gamma, const uint8_t* mask) {
    if (view < 0 || view > 1) return s3_fail(ctx, S3DMST_E_ARG, "bad view");
    S3_CUDA(cudaSetDevice(ctx->device));
    S3_TRY(s3_weighted_median(ctx, view, radius, gamma, mask));
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int s3dmst_norm_factor(s3dmst_ctx* ctx, int view, double* norm_factor) {
    if (view < 0 || view > 1 || !norm_factor) return s3_fail(ctx, S3DMST_E_ARG, "norm_factor: bad arguments");
    S3_CUDA(cudaSetDevice(ctx->device));
    return s3_norm_factor(ctx, view, norm_factor);
}

int s3dmst_reproject_to_3d(s3dmst_ctx* ctx, const double* Q, float disp_floor, int handle_missing, float* xyz, uint32_t* rgb) {
    if (!Q || ctx->N == 0) return s3_fail(ctx, S3DMST_E_ARG, "reproject_to_3d: Q and images required");
    S3_CUDA(cudaSetDevice(ctx->device));
    return s3_reproject(ctx, Q, disp_floor, handle_missing, xyz, rgb);
}

int s3dmst_run_dense(s3dmst_ctx* ctx, int D, int fill, float* left_disp, float* right_disp) {
    S3_CUDA(cudaSetDevice(ctx->device));
    memset(ctx->ev_set, 0, sizeof ctx->ev_set);
    S3_EV_BEGIN(S3DMST_T_FOREST, 0);
    S3_TRY(s3_forest_stage_mask(ctx, 3));  // both views share every launch of the forest stage
    S3_EV_END(S3DMST_T_FOREST, 0);
    const bool fuse = s3_want_fused_cost(ctx);
    if (fuse) S3_TRY(s3_fused_prepare(ctx, D));
    else S3_TRY(s3_cost_adgrad(ctx, D, 0));
    {
        int rc = ctx->P.agg_kernel == 1 ? 1 : s3_aggregate_flow(ctx, 3, 0, D, fuse);  // both views' trees in one launch
        if (rc == 1) {
            rc = 0;
            for (int view = 0; view < 2 && !rc; view++) rc = s3_aggregate_dense(ctx, view, 0, D);
        }
        if (rc) return rc;
    }
    for (int view = 0; view < 2; view++) S3_TRY(s3_dense_to_disp(ctx, view));
    S3_TRY(s3_lr_check(ctx, fill));
    D2H(left_disp, ctx->v[0].disp_f, sizeof(float) * ctx->N);
    D2H(right_disp, ctx->v[1].disp_f, sizeof(float) * ctx->N);
    if (left_disp || right_disp) S3_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// The reference's own pipeline (stereo3dmst, Stereo3DMST.cpp:805-904) on the current images: forests, random plane
// initialisation (:390-430), num_iter rounds of MST_PMS per view (:858-889), LabelToDisp (:189-201, :900-902) and the
// left-right check (:904; the reference passes fill = false).  Cost volumes set with s3dmst_set_cost_volume (the
// reference's mc-cnn input) are used as they are; without them the a2' volume is built and ingested.  Everything
// after the forests is queued without host synchronisation.
int s3dmst_run(s3dmst_ctx* ctx, int Dmax, unsigned seed, int fill, float* left_disp, float* right_disp) {
    if (ctx->N == 0 || Dmax <= 0) return s3_fail(ctx, S3DMST_E_ARG, "run: images and Dmax > 0 required");
    S3_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->P.exact) return s3_fail(ctx, S3DMST_E_ARG, "run: the proposal search runs in the exact mode only");
    memset(ctx->ev_set, 0, sizeof ctx->ev_set);
    if (!ctx->v[0].forest_ready || !ctx->v[1].forest_ready) {
        S3_EV_BEGIN(S3DMST_T_FOREST, 0);
        S3_TRY(s3_forest_stage_mask(ctx, 3));
        S3_EV_END(S3DMST_T_FOREST, 0);
    }
    if (ctx->P.pms_cost_mode == 1) S3_TRY(s3_prepare_plane_cost(ctx, Dmax));  // no volume: planes are scored straight from the images
    else if (!ctx->v[0].cost_ready || !ctx->v[1].cost_ready || ctx->v[0].D != Dmax || ctx->v[1].D != Dmax) S3_TRY(s3_cost_adgrad(ctx, Dmax, 1));
    for (int view = 0; view < 2; view++) {
        S3_TRY(s3_init_labels(ctx, view, Dmax));
        S3_TRY(s3_pms_iterate(ctx, view, ctx->P.num_iter, seed));
        S3_TRY(s3_label_to_disp(ctx, view));
    }
    S3_TRY(s3_lr_check(ctx, fill));
    D2H(left_disp, ctx->v[0].disp_f, sizeof(float) * ctx->N);
    D2H(right_disp, ctx->v[1].disp_f, sizeof(float) * ctx->N);
    if (left_disp || right_disp) S3_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// phases: bit 0 = front (forests + cost volumes of every frame), bit 1 = back (joint aggregation, LR check, copies)
// phases: 1 = front, 2 = back, 4 = do not wait for the copies into left_disp / right_disp
static int run_dense_batch_impl(s3dmst_ctx** ctxs, int n, int D, int fill, float** left_disp, float** right_disp, int phases) {
    if (!ctxs || n < 1) return S3DMST_E_ARG;
    s3dmst_ctx* ctx = ctxs[0];
    for (int c = 0; c < n; c++)
        if (!ctxs[c] || ctxs[c]->device != ctx->device || ctxs[c]->N != ctx->N || ctxs[c]->N == 0)
            return s3_fail(ctx, S3DMST_E_ARG, "run_dense_batch: contexts must share the device and hold images of one size");
    S3_CUDA(cudaSetDevice(ctx->device));
    const bool fuse = s3_want_fused_cost(ctx);  // a batch_back after a batch_front decides the same way (same parameters)
    // Front: forest + cost volume of every frame, queued on the frames' own streams by this one host thread.  Nothing
    // here waits for the device: the tree counts and sizes travel to the host behind the forest kernels, and the joint
    // aggregation below is the first thing that needs them.
    static const int joint_env = getenv("S3_FH_JOINT") ? atoi(getenv("S3_FH_JOINT")) : 0;   // development: one forest launch for the batch
    if ((phases & 1) && joint_env && n > 1 && 2 * n <= S3_FH_MAX_VIEWS) {
        for (int c = 0; c < n; c++) {
            memset(ctxs[c]->ev_set, 0, sizeof ctxs[c]->ev_set);
            S3_TRY(s3_forest_pre(ctxs[c], 3));
        }
        S3_TRY(s3_fh_launch_multi(ctxs, n, 3));
        for (int c = 0; c < n; c++) {
            S3_TRY(s3_forest_post(ctxs[c], 3));
            if (fuse) S3_TRY(s3_fused_prepare(ctxs[c], D));
            else S3_TRY(s3_cost_adgrad(ctxs[c], D, 0));
        }
        phases &= ~1;
    }
    for (int c = 0; (phases & 1) && c < n; c++) {
        s3dmst_ctx* cx = ctxs[c];
        const int r = [&]() -> int {
            s3dmst_ctx* ctx = cx;  // the error macros report into this frame's context
            memset(ctx->ev_set, 0, sizeof ctx->ev_set);
            S3_EV_BEGIN(S3DMST_T_FOREST, 0);
            S3_TRY(s3_forest_stage_mask(ctx, 3));
            S3_EV_END(S3DMST_T_FOREST, 0);
            return fuse ? s3_fused_prepare(ctx, D) : s3_cost_adgrad(ctx, D, 0);
        }();
        if (r) return c == 0 ? r : s3_fail(ctx, r, "run_dense_batch: frame %d: %s", c, cx->err.c_str());
    }
    if (!(phases & 2)) return 0;
    {
        int r = ctx->P.agg_kernel == 0 ? s3_aggregate_flow_multi(ctxs, n, 3, 0, D, fuse) : 1;
        if (r == 1) {
            r = 0;
            for (int c = 0; c < n && !r; c++)
                for (int view = 0; view < 2 && !r; view++) {
                    r = s3_aggregate_dense(ctxs[c], view, 0, D);
                    if (r && c) r = s3_fail(ctx, r, "run_dense_batch: frame %d: %s", c, ctxs[c]->err.c_str());
                }
        }
        if (r) return r;
    }
    for (int c = 0; c < n; c++) {
        s3dmst_ctx* cx = ctxs[c];
        int r = s3_dense_to_disp(cx, 0);
        if (!r) r = s3_dense_to_disp(cx, 1);
        if (!r) r = s3_lr_check(cx, fill);
        if (r) return c == 0 ? r : s3_fail(ctx, r, "run_dense_batch: frame %d: %s", c, cx->err.c_str());
        if (left_disp && left_disp[c]) S3_CUDA(cudaMemcpyAsync(left_disp[c], cx->v[0].disp_f, sizeof(float) * cx->N, cudaMemcpyDeviceToHost, cx->stream));
        if (right_disp && right_disp[c]) S3_CUDA(cudaMemcpyAsync(right_disp[c], cx->v[1].disp_f, sizeof(float) * cx->N, cudaMemcpyDeviceToHost, cx->stream));
    }
    if ((left_disp || right_disp) && !(phases & 4))
        for (int c = 0; c < n; c++) S3_CUDA(cudaStreamSynchronize(ctxs[c]->stream));
    return 0;
}

int s3dmst_run_dense_batch(s3dmst_ctx** ctxs, int n, int D, int fill, float** left_disp, float** right_disp) {
    return run_dense_batch_impl(ctxs, n, D, fill, left_disp, right_disp, 3);
}
int s3dmst_run_dense_batch_async(s3dmst_ctx** ctxs, int n, int D, int fill, float** left_disp, float** right_disp) {
    return run_dense_batch_impl(ctxs, n, D, fill, left_disp, right_disp, 3 | 4);
}
int s3dmst_batch_front(s3dmst_ctx** ctxs, int n, int D) { return run_dense_batch_impl(ctxs, n, D, 0, nullptr, nullptr, 1); }
int s3dmst_batch_back(s3dmst_ctx** ctxs, int n, int D, int fill, float** left_disp, float** right_disp) {
    return run_dense_batch_impl(ctxs, n, D, fill, left_disp, right_disp, 2);
}

}  // extern "C"
