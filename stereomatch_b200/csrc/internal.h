// stereomatch_b200/csrc/internal.h — context and per-view device state (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/s3dmst.h"

#define S3_NUM_W 766            // integer edge weights 0..765 (|dR|+|dG|+|dB|)
#define S3_NO_EDGE 0xFFFFu
#define S3_DEAD 0xFFFFFFFFu
#define S3_AGG_NEAR 32          // dataflow aggregation: a parent/child closer than this (in BFS index) is handed over in shared memory
#define S3_NU_FARPARENT 0x100    // NodeUp.child_count flag: the node's parent is S3_AGG_NEAR or more nodes away
#define S3_ND_FAR (1 << 30)      // node_dn.z flag: the node has a child S3_AGG_NEAR or more nodes away
#define S3_ND_LEAF (1 << 29)     // node_dn.z flag: no children (its leaf->root sum is its cost: never stored)
#define S3_ND_FLAGS (S3_ND_FAR | S3_ND_LEAF)

// Per-node record read by the leaf->root pass: children are contiguous in BFS order.
struct __align__(16) NodeUp {
    int child_begin;      // global node index of the first child
    int child_count;      // 0..4
    uint32_t cw01, cw23;  // integer edge weights of children 0..3, 16 bits each
};

#define S3_P2P_MAX 16   // ranks of a label-sharded communicator that can use the peer-memory MIN-LOC
#define S3_EV_SLOTS 4   // stage-timer samples that may be in flight per (stage, view)

struct View {
    // ---- image stage
    uint8_t* bgr = nullptr;    // [N*3] tightly packed copy of the input
    uchar4* raw4 = nullptr;    // [N] raw (b,g,r,0)
    uchar4* med = nullptr;     // [N] median-filtered (b,g,r,0)
    float* gray = nullptr;     // [N] s3_gray of the RAW image (cost kernel)
    uint2* match8 = nullptr;   // [N] {packed BGR, gray}: what the fused matching cost of the aggregation kernel reads
    float* pgrad = nullptr;    // [N][2] Sobel/8 gradients of the BGR2GRAY image (pms_cost_mode 1, pm.cpp:70-88), lazily allocated
    bool plane_ready = false;
    uint16_t* ew = nullptr;    // [2N] edge weights by canonical id
    // ---- union-find / forest construction
    void* uf_comp = nullptr;   // [N] x 16 B component records (forest.cu: FHComp)
    int* uf_parent = nullptr;  // [N] union-find parents
    void* fh_ent[2] = {nullptr, nullptr};                 // [2N] x 16 B live-edge lists of the FH rounds (ping-pong)
    unsigned long long* uf_resv = nullptr;  // [N] merge reservation key
    uint8_t* mask = nullptr;   // [2N] 0/1/2
    uint32_t* elist = nullptr; // [2N] edge ids bucketed by weight, later the merge candidate list
    int* e_ra = nullptr;       // [2N] per list position scratch
    int* e_rb = nullptr;
    uint8_t* e_flag = nullptr; // [2N]
    int* hist = nullptr;       // [S3_NUM_W] weight histogram
    int* lvl_off = nullptr;    // [S3_NUM_W+1] bucket offsets
    int* lvl_cursor = nullptr; // [S3_NUM_W]
    int* counters = nullptr;   // [S3_MAX_ROUNDS] per-round live counters + misc
    ushort4* adjw = nullptr;   // [N] forest-edge weights to the (up, left, right, down) neighbours, S3_NO_EDGE if none
    uint32_t* bfs_front = nullptr;  // [N] BFS frontier words by node
    // ---- labelling
    int* minpix = nullptr;     // [N] min pixel of the component rooted here
    int* scan_tmp = nullptr;   // [N] + block sums
    int* tree_id = nullptr;    // [N] by pixel
    int* tree_size = nullptr;  // [N] (first T used)
    int* tree_rootpix = nullptr;  // [N] (first T used)
    // ---- BFS-ordered forest
    int T = 0, max_depth = 0;
    int* tree_start = nullptr;  // [T+1] (allocated N+1)
    int* tree_depth = nullptr;  // [T]
    int* unit_tree = nullptr;   // [T] trees by decreasing size
    int* node_pixel = nullptr;  // [N]
    int* pixel_node = nullptr;  // [N]
    int* parent = nullptr;      // [N]
    int* level = nullptr;       // [N]
    uint16_t* pw = nullptr;     // [N]
    NodeUp* node_up = nullptr;  // [N]
    int4* node_dn = nullptr;    // [N] {parent, parent weight, level | flags, pixel}: the root->leaf pass record
    uint32_t* leaf_bits = nullptr;  // [N/32 + 1] bit v = node v is a leaf (prefetch target selection on the way down)
    int* lvl_start = nullptr;   // [N + T + 1]; tree t's level offsets start at tree_start[t] + t
    // tree adjacency graph (Stereo3DMST.cpp:377-384) as a device CSR, built lazily (proposal generation, parity dumps)
    int* adj_ptr = nullptr;     // [T+1]
    int* adj = nullptr;         // [n_adj] neighbours of tree t: adj[adj_ptr[t] .. adj_ptr[t+1]), ascending
    size_t adj_ptr_cap = 0, adj_cap = 0;
    int n_adj = 0;
    bool adj_ready = false;
    std::vector<int> h_tree_start, h_tree_depth, h_unit_tree;
    bool forest_ready = false;
    // ---- volumes
    int D = 0, Dp = 0;          // labels, padded row length (multiple of 4)
    float* cost = nullptr;      // [N][Dp] node-major, label-minor
    size_t cost_cap = 0;
    double* aup = nullptr;      // [N][Dp] leaf->root sums, overwritten by final values on the way down
    size_t aup_cap = 0;
    bool cost_ready = false, agg_ready = false;
    int agg_d0 = 0, agg_d1 = 0;
    // ---- dense results (pixel order)
    int32_t* disp_i = nullptr;  // [N]
    double* best = nullptr;     // [N]
    // ---- PatchMatch state (pixel order)
    float* abc = nullptr;       // [N][3]
    double* min_cost = nullptr; // [N]
    bool labels_ready = false;
    // ---- disparity maps
    float* disp_f = nullptr;    // [N]
    uint8_t* lr_mask = nullptr; // [N]
};

struct s3dmst_ctx {
    int device = 0;
    s3dmst_params P;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t stream_big = nullptr;   // third stream: the 32-warp CTAs of the large trees beside the small trees' launch
    cudaEvent_t ev_join_big = nullptr;
    cudaStream_t stream_aux = nullptr;   // second stream: the cluster launch of the giant trees runs beside the other trees' launches
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int W = 0, H = 0, N = 0;
    int num_sms = 0;
    View v[2];
    double* lut_w = nullptr;   // [S3_NUM_W] exp(-iw*gamma)
    double* lut_w2 = nullptr;  // [S3_NUM_W] 1 - w*w
    float* lut_wf = nullptr;   // fp32 copies for the fast path
    float* lut_w2f = nullptr;
    // Stage timers: a small ring of event pairs per (stage, view), so that a caller that queues call after call without
    // synchronising still gets every sample: a pair is harvested into the accumulators once its end event has completed.
    cudaEvent_t ev[S3DMST_T_COUNT][2][S3_EV_SLOTS][2];  // [stage][view][slot][begin/end]
    int ev_slot[S3DMST_T_COUNT][2];                     // slot of the latest pair
    unsigned char ev_state[S3DMST_T_COUNT][2][S3_EV_SLOTS];  // 0 = free / harvested, 1 = recorded, waiting to be harvested
    bool ev_set[S3DMST_T_COUNT][2];                     // the latest pair belongs to the current call (s3dmst_stage_ms)
    double ev_acc_ms[S3DMST_T_COUNT];                   // accumulated since the last reset (s3dmst_stage_total_ms)
    int ev_acc_n[S3DMST_T_COUNT], ev_lost[S3DMST_T_COUNT];
    long long launches = 0;
    int fused_D = 0;           // > 0: the last dense run computed its matching cost in the aggregation kernel for this D (no volume yet)
    std::string err;
    // scratch for PMS
    cudaEvent_t ev_xctx = nullptr;  // orders this context's stream against another context's in batched launches
    cudaEvent_t ev_block = nullptr; // blocking-sync event: batched contexts SLEEP through the forest kernel instead of spinning
    int forest_pending = 0;         // views whose tree count / sizes are still on their way to the host (s3_forest_finish_host)
    int* h_pin = nullptr;           // pinned landing zone of the forest stage's (T, tree sizes) copy
    size_t h_pin_cap = 0;           // ints
    char* stage = nullptr;          // pinned staging arena of s3_h2d_staged: two halves, each guarded by an event
    size_t stage_half = 0, stage_used = 0;
    int stage_cur = 0;
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    bool stage_pending[2] = {false, false};
    uint32_t* units_dev = nullptr;  // aggregation work units + view table of the current launch
    size_t units_cap = 0;
    int* fh_sync = nullptr;         // grid barrier + per-round live counters of the forest kernel
    void* pms_scratch = nullptr;
    size_t pms_scratch_cap = 0;
    // label-range sharding (comm.cu): NCCL communicator bound at run time, its stream, the global-minimum buffer
    void* comm = nullptr;
    int comm_rank = 0, comm_nranks = 0;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_comm[2] = {nullptr, nullptr}, ev_comm_t[4] = {nullptr, nullptr, nullptr, nullptr};
    bool comm_timed = false;
    // MIN-LOC over peer memory (comm.cu): every rank's (best, disparity) buffers and flag array mapped into this process
    // through CUDA IPC; rank r reduces pixel slice r reading all ranks over NVLink and writes the result to all of them
    bool p2p_tried = false, p2p_ok = false;
    int p2p_N = 0, p2p_epoch = 0;
    double* p2p_best[2][S3_P2P_MAX];
    int32_t* p2p_disp[2][S3_P2P_MAX];
    int* p2p_flags[S3_P2P_MAX];          // p2p_flags[comm_rank] = this rank's own array (cudaMalloc)
    int* p2p_counter = nullptr;          // CTAs of this rank that finished their share
    int* p2p_err_host = nullptr;         // mapped pinned: set by a kernel whose wait for the peers timed out
    int* p2p_err_dev = nullptr;
    void* p2p_xbuf = nullptr;
    double* gmin = nullptr;
    size_t gmin_cap = 0;
    float* abc_init = nullptr;      // the reference's random plane initialisation for (abc_init_w x abc_init_h, abc_init_d), kept on the device
    int abc_init_w = 0, abc_init_h = 0, abc_init_d = 0;
    uint32_t pms_round[2] = {0, 0};  // rounds of the library's generator run since the view's labels were (re)set: the RNG counter
    // rectification front-end (rectify.cu): per-view fixed-point maps, the weight table, staging for the raw pair
    int16_t* map_xy[2] = {nullptr, nullptr};    // [mh][mw][2] integer source corner (x, y)
    uint16_t* map_fxy[2] = {nullptr, nullptr};  // [mh][mw] fy * 32 + fx
    int map_w[2] = {0, 0}, map_h[2] = {0, 0};
    int16_t* remap_tab = nullptr;               // [1024][4] bilinear weights, sum 32768
    uint8_t* raw_stage = nullptr;               // 2 raw BGR images
    size_t raw_stage_cap = 0;
};

#define S3_MAX_ROUNDS 65536
#define S3_FH_MAX_VIEWS 16        // views (2 per frame) one forest-kernel launch serves
#define S3_FH_ROUNDS 8192         // round cap of the forest kernel (per-round counters)
#define S3_FH_MAX_CTAS 256        // upper bound on the cooperative grid of the forest kernel
#define S3_FH_SEG_SLACK 1024      // per-CTA slack of the live-edge list segments (one ingest event adds < 1 entry per CTA beyond its share)

// error helpers ---------------------------------------------------------------------------------
int s3_fail(s3dmst_ctx* c, int code, const char* fmt, ...);
#define S3_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return s3_fail(ctx, S3DMST_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    } while (0)
#define S3_LAUNCH_CHECK()                                                                               \
    do {                                                                                                \
        ctx->launches++;                                                                                \
        cudaError_t e__ = cudaGetLastError();                                                           \
        if (e__ != cudaSuccess)                                                                         \
            return s3_fail(ctx, S3DMST_E_CUDA, "%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
    } while (0)
void s3_ev_harvest(s3dmst_ctx* ctx, int stage, int view, bool wait);   // api.cu
#define S3_EV_BEGIN(stage, view)                                                                          \
    do {                                                                                                  \
        s3_ev_harvest(ctx, stage, view, false);                                                           \
        const int sl__ = (ctx->ev_slot[stage][view] + 1) % S3_EV_SLOTS;                                   \
        if (ctx->ev_state[stage][view][sl__]) ctx->ev_lost[stage]++; /* still running: overwritten */     \
        ctx->ev_state[stage][view][sl__] = 0;                                                             \
        ctx->ev_slot[stage][view] = sl__;                                                                 \
        S3_CUDA(cudaEventRecord(ctx->ev[stage][view][sl__][0], ctx->stream));                             \
    } while (0)
#define S3_EV_END(stage, view)                                                                            \
    do {                                                                                                  \
        S3_CUDA(cudaEventRecord(ctx->ev[stage][view][ctx->ev_slot[stage][view]][1], ctx->stream));        \
        ctx->ev_state[stage][view][ctx->ev_slot[stage][view]] = 1;                                        \
        ctx->ev_set[stage][view] = true;                                                                  \
    } while (0)
#define S3_TRY(call)            \
    do {                        \
        int r__ = (call);       \
        if (r__ != 0) return r__; \
    } while (0)

// stage entry points implemented in the .cu files ------------------------------------------------
int s3_image_stage(s3dmst_ctx* ctx, int view);                    // image.cu: median, gray, edge weights, buckets
int s3_forest_stage(s3dmst_ctx* ctx, int view);                   // forest.cu: FH + merge + labels + BFS
int s3_forest_stage_mask(s3dmst_ctx* ctx, int mask);              // both views in shared launches
int s3_fh_launch(s3dmst_ctx* ctx, int mask);
int s3_fh_launch_multi(s3dmst_ctx** ctxs, int nctx, int mask);   // one launch over several frames
int s3_forest_pre(s3dmst_ctx* ctx, int mask);                     // image stage + union-find init (async)
int s3_forest_post(s3dmst_ctx* ctx, int mask);                    // labelling + BFS (asynchronous)
int s3_forest_finish_host(s3dmst_ctx* ctx);                       // tree count / sizes -> host (waits for the copy s3_forest_post queued)
int s3_forest_finalize_host(s3dmst_ctx* ctx, int view);
int s3_forest_depths(s3dmst_ctx* ctx, int view);                  // forest.cu: lazy D2H of the tree depths           // forest.cu: unit order, depths
int s3_set_rectify_maps(s3dmst_ctx* ctx, int view, const int16_t* map_xy, const uint16_t* map_fxy, int W, int H);  // rectify.cu
int s3_remap_raw_pair(s3dmst_ctx* ctx, const uint8_t* left_raw, const uint8_t* right_raw, int sw, int sh, int stride);
void s3_rectify_free(s3dmst_ctx* ctx);
void s3_remap_table(int16_t* tab);  // [1024][4]
int s3_comm_before_free(s3dmst_ctx* ctx);                         // comm.cu: peers unmap this rank's result buffers (collective; no-op without them)
int s3_cost_adgrad(s3dmst_ctx* ctx, int D, int apply_ingest);     // cost.cu
int s3_cost_from_dmajor(s3dmst_ctx* ctx, int view, const float* dev_dmajor, int D, int apply_ingest);
int s3_cost_to_dmajor(s3dmst_ctx* ctx, int view, float* dev_dmajor);
int s3_aggregate_dense(s3dmst_ctx* ctx, int view, int d0, int d1); // aggregate.cu (v1, reference kernel)
int s3_aggregate_flow(s3dmst_ctx* ctx, int views_mask, int d0, int d1, int fuse = 0);    // aggregate3.cu (dataflow, default)
struct PmsFlowPlan {   // aggregate3.cu: unit list of a proposal-mode launch, uploaded once and reused by every round
    int n_cl, n_big, n_small;
    const void* views_dev;
    const void* units_dev;
};
int s3_pms_flow_plan(s3dmst_ctx* ctx, int view, const int* h_prop_off, double* scratch_dev, PmsFlowPlan* plan);
int s3_pms_flow_launch(s3dmst_ctx* ctx, int view, const PmsFlowPlan* plan, const int* prop_off_dev, const float* labels_dev, int gen, uint32_t seed, uint32_t round);
// Host -> device copy of small metadata through the context's pinned staging arena (api.cu): truly asynchronous, and the
// source may die as soon as the call returns.
int s3_h2d_staged(s3dmst_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int s3_tree_adjacency(s3dmst_ctx* ctx, int view);   // pms.cu: device CSR of the tree adjacency graph (lazy)
int s3_aggregate_flow_multi(s3dmst_ctx** ctxs, int nctx, int views_mask, int d0, int d1, int fuse = 0);  // one launch over several frames
int s3_pms_apply(s3dmst_ctx* ctx, int view, const int32_t* h_tree_ids, const float* h_labels, size_t n);
int s3_init_labels(s3dmst_ctx* ctx, int view, int Dmax);          // pms.cu: the reference's random plane initialisation
int s3_pms_iterate(s3dmst_ctx* ctx, int view, int n_iter, unsigned seed);
int s3_label_to_disp(s3dmst_ctx* ctx, int view);                  // post.cu
int s3_dense_to_disp(s3dmst_ctx* ctx, int view);
int s3_lr_check(s3dmst_ctx* ctx, int fill);
int s3_minloc_mask(s3dmst_ctx* ctx, int view, const double* global_min_dev);
int s3_reproject(s3dmst_ctx* ctx, const double* Q16, float disp_floor, int handle_missing, float* h_xyz, uint32_t* h_rgb);
int s3_ensure_volume(s3dmst_ctx* ctx, int view, int D, bool need_cost = true);   // need_cost = false: only the running sums (fused cost)
bool s3_want_fused_cost(const s3dmst_ctx* ctx);                 // dense runs compute the matching cost inside the aggregation kernel
int s3_fused_prepare(s3dmst_ctx* ctx, int D);                   // ... what such a run needs instead of s3_cost_adgrad
int s3_materialize_cost(s3dmst_ctx* ctx);                       // build the volume a fused run skipped, if somebody asks for it
int s3_prepare_plane_cost(s3dmst_ctx* ctx, int Dmax);              // pms.cu
int s3_weighted_median(s3dmst_ctx* ctx, int view, int radius, float gamma, const uint8_t* h_mask);  // postfilter.cu
int s3_norm_factor(s3dmst_ctx* ctx, int view, double* h_out);
