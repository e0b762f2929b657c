/* include/s3dmst.h — C ABI of the B200-native Stereo3DMST hot path.
 *
 * This is the drop-in boundary for the reference's
 *     extern "C" void stereo3dmst(std::string, std::string, cv::Mat&, cv::Mat&, cv::Mat&, cv::Mat&,
 *                                 std::string, int)            (include/Stereo3DMST.h:7)
 * and the host functions it is made of (src/Stereo3DMST.cpp).  The reference signature passes
 * C++ objects, so it is not an ABI; what a maintainer binds instead is this header (plain
 * pointers and sizes, no C++/torch types).  A header-compatible C++ shim that forwards
 * stereo3dmst(cv::Mat...) to these entry points lives in stereomatch_b200/csrc/stereo3dmst_shim.cpp;
 * INTEGRATION.md shows the call-site change.
 *
 * Conventions: every function returns 0 on success or a negative S3DMST_E_* code;
 * s3dmst_last_error() gives the message.  No exceptions cross the ABI.  All buffers are
 * caller-owned; unless a name ends in _dev they are HOST pointers and the call synchronises
 * the context's stream before returning.  One context per GPU; a context is not thread-safe.
 * "view": 0 = left, 1 = right.  Pixels are raster order p = y*W + x.  Volumes handed across the
 * ABI are the reference's layout float[D][H][W] (Stereo3DMST.cpp:117, :769-773).
 * There is NO CPU fallback: every entry point fails with S3DMST_E_CUDA if no device is usable.
 */
#ifndef S3DMST_H_
#define S3DMST_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct s3dmst_ctx s3dmst_ctx;

enum {
    S3DMST_OK = 0,
    S3DMST_E_ARG = -1,   /* bad argument / call order */
    S3DMST_E_CUDA = -2,  /* CUDA runtime error (message has the detail) */
    S3DMST_E_STATE = -3, /* required stage has not run */
    S3DMST_E_LIMIT = -4, /* internal iteration cap hit (forest kernels) */
    S3DMST_E_COMM = -5   /* NCCL missing or failed (label-range sharding) */
};

/* The literals of src/Stereo3DMST.cpp lifted into one struct (reference values are the defaults). */
typedef struct s3dmst_params {
    float fh_c;         /* 5000      Stereo3DMST.cpp:831   FH threshold constant                     */
    int min_cc_size;    /* 200       :832                  min-size merge bound (max(2, .) applied)  */
    float gamma;        /* 1/12      :830                  edge weight = exp(-w*gamma)               */
    int median;         /* 3         :214                  0 disables the 3x3 median                 */
    float cost_cap;     /* 0.5       :789-801              ingest: NaN -> cap, else min(cap, v)      */
    float cost_offset;  /* 0         :792 (commented)      ingest: v = (c + offset) * scale          */
    float cost_scale;   /* 1                                                                       */
    float oob_cost;     /* 0.5       :115                  label cost outside [0, Dmax)              */
    int num_iter;       /* 100       :854                                                         */
    float refine_floor; /* 0.1       :600                                                         */
    int exact;          /* 1: fp64 state, reference association order (bit-exact); 0: fp32 state in the dense
                           aggregation (12 instead of 20 B of HBM traffic per pixel-label; costs within ~1e-5
                           relative of the exact ones, disparities may differ at near-ties)               */
    int keep_aggregated;/* 1: s3dmst_aggregate_dense keeps the final aggregated volume for dumps    */
    int agg_threads;    /* 0 = auto; threads per CTA of the aggregation kernels (multiple of 32)    */
    int agg_cache_nodes;/* 0 = auto; nodes of a tree level cached in shared memory per CTA          */
    int agg_ring_nodes; /* unused (kept so that the layout of round 1 callers stays valid)                 */
    int agg_kernel;     /* 0 = auto (dataflow kernel), 1 = simple level-synchronous kernels (any even d0)        */
    int fh_ctas;        /* 0 = one CTA per SM (one pair alone on the GPU); > 0: CTAs of the cooperative forest kernel per
                           frame, for contexts that run beside others in a batch (measured best at C2: 36 for 8
                           frames; also selects the narrower live-edge band, forest.cu fill_fh_args)           */
    int fh_threads;     /* 0 = 1024; threads per CTA of the forest kernel */
    int fuse_cost;      /* dense runs (run_dense, run_dense_batch, aggregate_dense_sharded) on the library's AD+gradient cost:
                           0 / 1 = the aggregation kernel computes the matching cost itself and no volume is built
                           (s3dmst_get_cost_volume / s3dmst_aggregate_dense build it on demand afterwards); -1 = build the
                           volume first.  Results are bit-identical either way.                                      */
    int comm_p2p;       /* label sharding: 0 / 1 = the MIN-LOC runs as ONE kernel over peer memory (every rank's result buffers
                           mapped through CUDA IPC, NVLink loads and stores) when all ranks can map each other, else over
                           NCCL; -1 = always the two NCCL all-reduces                                                 */
    int fh_cluster;     /* 0: the forest kernel is a cooperative grid with a software barrier per view; 8 or 16: every view gets one
                           thread-block cluster of that many CTAs and the hardware cluster barrier                      */
    int pms_cost_mode;  /* data term of the 3D-label (PatchMatch) search: 0 = compute3DLabelCost on the cost volume (the
                           reference, Stereo3DMST.cpp:103-118); 1 = slanted-plane truncated colour + gradient cost straight from
                           the two images, at the sub-pixel match position (pm::PatchMatch, src/pm.cpp:97-154): no volume */
    float pm_alpha;     /* 0.9   stereo_opencv.cpp:156   weight of the gradient term                                  */
    float pm_tau_c;     /* 10    truncation of the L1 colour difference                                             */
    float pm_tau_g;     /* 2     truncation of the L1 gradient difference                                           */
    int agg_cluster_nodes; /* 0 = auto (32768): trees of at least this many nodes are walked by a thread-block cluster of
                           8 CTAs (256 warps) instead of one CTA; < 0: never */
} s3dmst_params;

void s3dmst_default_params(s3dmst_params* p);

/* stream: a cudaStream_t to run on (e.g. torch's current stream), or NULL to own one. */
int s3dmst_create(s3dmst_ctx** out, int device, const s3dmst_params* params, void* stream);
void s3dmst_destroy(s3dmst_ctx* ctx);
const char* s3dmst_last_error(const s3dmst_ctx* ctx);
int s3dmst_sync(s3dmst_ctx* ctx);

/* a1 inputs: two interleaved BGR u8 images (cv::Mat CV_8UC3), row stride in bytes.  H2D copy. */
int s3dmst_set_images(s3dmst_ctx* ctx, const uint8_t* left_bgr, const uint8_t* right_bgr, int W, int H,
                      int stride_bytes);

/* The same without the final synchronisation: the copies are queued on the context's stream (truly asynchronous from
 * pinned host memory) and the buffers must stay untouched until the next synchronising call on the context
 * (e.g. s3dmst_run_dense* with output pointers, s3dmst_sync).  Lets a batch overlap its uploads with compute. */
int s3dmst_set_images_async(s3dmst_ctx* ctx, const uint8_t* left_bgr, const uint8_t* right_bgr, int W, int H, int stride_bytes);

/* Rectification front-end: the caller's per-frame cv::remap(img, imgr, map1, map2, INTER_LINEAR) of
 * src/stereo_Yin.cpp:143-144 on the device.  map_xy / map_fxy are the CV_16SC2 / CV_16UC1 pair that
 * initUndistortRectifyMap(..., CV_16SC2, map1, map2) returns (:139-140): int16 [H][W][2] source corner (x, y) and
 * uint16 [H][W] fraction index fy*32+fx.  Uploaded once per camera and kept on the device; W x H is the rectified size. */
int s3dmst_set_rectify_maps(s3dmst_ctx* ctx, int view, const int16_t* map_xy, const uint16_t* map_fxy, int W, int H);

/* Raw (unrectified) BGR u8 pair, both src_w x src_h: H2D copy + fixed-point bilinear remap (bit-identical to cv::remap
 * with INTER_LINEAR, BORDER_CONSTANT 0) straight into the context's images; afterwards as after s3dmst_set_images. */
int s3dmst_set_raw_images(s3dmst_ctx* ctx, const uint8_t* left_raw_bgr, const uint8_t* right_raw_bgr, int src_w, int src_h,
                          int stride_bytes);

/* The image of a view as the path sees it (after s3dmst_set_images / s3dmst_set_raw_images): BGR u8, W*H*3 bytes. */
int s3dmst_get_image(s3dmst_ctx* ctx, int view, uint8_t* bgr);

/* The 1024 x 4 int16 bilinear weight table of the remap (host-side; no device needed): parity dumps. */
void s3dmst_remap_table(int16_t* tab);

/* a3,a4,a5,a7 (Stereo3DMST.cpp:226-307, :342-384, :434-522; segment-graph.h:54-89): median, edge
 * weights, FH forest (asynchronous exact rounds with a (weight, edge id) tie-break), min-size merge, tree ids, BFS
 * re-indexing — all on the device; the call queues the work and returns (the tree count reaches the host lazily). */
int s3dmst_build_forest(s3dmst_ctx* ctx, int view);
int s3dmst_forest_info(s3dmst_ctx* ctx, int view, int* num_trees, int* max_depth, int* adj_size);
/* Parity dump; any pointer may be NULL.  edge_weight/edge_mask are [2N] by canonical edge id
 * (2p = right neighbour, 2p+1 = down neighbour; mask 1 = FH edge, 2 = min-size-merge edge);
 * tree_start [T+1]; per node (tree-major, BFS order): node_pixel, parent, child_begin,
 * child_count, parent_weight (integer edge weight to the parent), level;  tree_id [N] by pixel;
 * tree adjacency CSR adj_ptr [T+1], adj [adj_size] (ascending = boost setS order). */
int s3dmst_get_forest(s3dmst_ctx* ctx, int view, uint16_t* edge_weight, uint8_t* edge_mask, int32_t* tree_id,
                      int32_t* tree_start, int32_t* node_pixel, int32_t* parent, int32_t* child_begin,
                      int32_t* child_count, uint16_t* parent_weight, int32_t* level, int32_t* adj_ptr, int32_t* adj);
/* Upload an externally built forest instead (tests: isolates the aggregation kernels). */
int s3dmst_set_forest(s3dmst_ctx* ctx, int view, int W, int H, int num_trees, const int32_t* tree_start,
                      const int32_t* node_pixel, const int32_t* parent, const uint16_t* parent_weight);

/* a2': truncated colour + gradient absolute-difference volume for both views, D labels
 * (PatchMatchStereoGPU.cu:1482-1550), written straight into the node-major device layout;
 * then the a2 ingest (Stereo3DMST.cpp:785-803) if apply_ingest != 0.  Needs both forests. */
int s3dmst_build_cost_volume(s3dmst_ctx* ctx, int D, int apply_ingest);
/* a2: external volume in the mc-cnn file layout float[D][H][W] (left.bin / right.bin), ingested. */
int s3dmst_set_cost_volume(s3dmst_ctx* ctx, int view, const float* vol_dmajor, int D, int apply_ingest);
/* (After a dense run that computed its matching cost inside the aggregation kernel — params.fuse_cost — the volume of
 * that run is built here on demand.) */
int s3dmst_get_cost_volume(s3dmst_ctx* ctx, int view, float* vol_dmajor);

/* a9,a10 + WTA, dense-label mode: labels [d0,d1), two-pass tree filter, strict '<'
 * so the lowest d wins ties.  disp [N] int32 and best_cost [N] double by pixel (NULL = leave on device; with both
 * NULL the call only queues the work on the context's stream and returns without waiting for it). */
int s3dmst_aggregate_dense(s3dmst_ctx* ctx, int view, int d0, int d1, int32_t* disp, double* best_cost);
/* Final aggregated volume of the last dense call, double[D][H][W] (needs params.keep_aggregated). */
int s3dmst_get_aggregated(s3dmst_ctx* ctx, int view, double* agg_dmajor);
/* Device pointers of the dense result (pixel order) for collectives: best cost f64 [N], disparity i32 [N].  Written in
 * the order of the context's stream: s3dmst_sync() before another stream (e.g. NCCL's) reads them. */
int s3dmst_dense_result_dev(s3dmst_ctx* ctx, int view, double** best_cost_dev, int32_t** disp_dev);
/* After an all-reduce(MIN) of best cost into global_min_dev: disp := INT32_MAX where the local cost is not
 * the global minimum, so that a second all-reduce(MIN) on disp yields the lowest d attaining the minimum. */
int s3dmst_minloc_mask(s3dmst_ctx* ctx, int view, const double* global_min_dev);
/* Label-range sharding of ONE large pair over the GPUs of a box (BASELINE config C5): one process and one
 * context per GPU; every rank builds the (deterministic) forests and its own label range of the cost volume, and the
 * only exchange is the per-pixel MIN-LOC that replaces the reference's serial `agg < min_cost` over ascending labels
 * (Stereo3DMST.cpp:173-185).  NCCL is bound at run time (dlopen libnccl.so.2); nothing else needs it.
 *   s3dmst_comm_unique_id  rank 0 fills 128 bytes (ncclUniqueId) and hands them to the other ranks by its own means
 *   s3dmst_comm_init       collective: every rank, with the same id
 *   s3dmst_comm_label_range  the labels [d0, d1) this rank aggregates (boundaries are multiples of 4)
 *   s3dmst_reduce_minloc   MIN-LOC all-reduce of a view's dense result (best cost f64, disparity i32; lowest d wins
 *                          ties), queued on the context's stream — no host synchronisation
 *   s3dmst_aggregate_dense_sharded  both views: this rank's labels, then the reduction; the left view's reduction runs
 *                          on the context's communication stream while the right view is aggregated.  Afterwards every
 *                          rank holds the global result (s3dmst_dense_to_disparity / s3dmst_lr_check as usual).
 *                          Needs images and forests; a cost volume of D labels on both views is used if there is one,
 *                          otherwise (params.fuse_cost >= 0) the AD+gradient cost of the rank's labels is computed
 *                          inside the aggregation kernel and no rank ever holds a volume.
 *   s3dmst_comm_minloc_ms  device time of the two views' reductions of the last sharded call (CUDA events), ms */
int s3dmst_comm_unique_id(void* id128);
int s3dmst_comm_init(s3dmst_ctx* ctx, const void* id128, int rank, int nranks);
int s3dmst_comm_destroy(s3dmst_ctx* ctx);
int s3dmst_comm_label_range(const s3dmst_ctx* ctx, int D, int* d0, int* d1);
int s3dmst_reduce_minloc(s3dmst_ctx* ctx, int view);
int s3dmst_aggregate_dense_sharded(s3dmst_ctx* ctx, int D);
/* How the MIN-LOC of the sharded calls travels: 1 = one kernel over peer memory (every rank's result buffers mapped through
 * CUDA IPC; decided collectively at the first s3dmst_aggregate_dense_sharded, params.comm_p2p), 0 = two NCCL all-reduces,
 * -1 = no communicator.  While the mapping is live, changing the image size (s3dmst_set_images with another W x H) and
 * s3dmst_comm_destroy / s3dmst_destroy are collective: every rank's buffers are unmapped everywhere before one is freed. */
int s3dmst_comm_transport(const s3dmst_ctx* ctx);
double s3dmst_comm_minloc_ms(s3dmst_ctx* ctx);

/* Dense disparity (int) -> float disparity map used by the LR check. */
int s3dmst_dense_to_disparity(s3dmst_ctx* ctx, int view);

/* a2' slanted variant (north-star item 1): prepares the data term of params.pms_cost_mode = 1 for Dmax labels — Sobel/8
 * gradients of both views (pm.cpp:70-88).  After it s3dmst_pms_apply / s3dmst_pms_iterate / s3dmst_run score a plane at
 * a pixel by the truncated colour + gradient difference to the other image at x -+ d (sub-pixel, pm.cpp:130-154) times
 * params.cost_scale, and params.oob_cost outside [0, Dmax] (the reference's PLANE_PENALTY is 120: set oob_cost to
 * 120 * cost_scale, well above the 2.8 * cost_scale of a mismatch); no cost volume is needed or touched. */
int s3dmst_prepare_plane_cost(s3dmst_ctx* ctx, int Dmax);
/* the gradients it computed: float [H*W][2] (parity dumps) */
int s3dmst_get_plane_gradients(s3dmst_ctx* ctx, int view, float* grad);

/* a6/a11/a12 PatchMatch state: labels abc [N][3] fp32 and min_cost [N] fp64 (init DBL_MAX). */
int s3dmst_set_labels(s3dmst_ctx* ctx, int view, const float* abc);
int s3dmst_get_labels(s3dmst_ctx* ctx, int view, float* abc);
int s3dmst_reset_min_cost(s3dmst_ctx* ctx, int view);
int s3dmst_get_min_cost(s3dmst_ctx* ctx, int view, double* min_cost);
/* Injected proposals, applied in order (MSTCostAggregationAndLabelUpdate, :160-186): proposal i tests
 * label labels[3i..3i+2] on tree tree_ids[i]. */
int s3dmst_pms_apply(s3dmst_ctx* ctx, int view, const int32_t* tree_ids, const float* labels, size_t n);
/* a6 random plane initialisation (:390-430), bit-identical to the reference (same libstdc++ engine and
 * distribution, raster order); also resets min_cost to DBL_MAX (:820-821). */
int s3dmst_init_labels(s3dmst_ctx* ctx, int view, int Dmax);
/* a12 MST_PMS (:546-629) x n_iter with the library's own proposal generator, on the device and without host
 * synchronisation between rounds: per tree one proposal per neighbouring tree (ascending id: the label of a random
 * pixel of that tree), then the refinement ladder around a random pixel of the tree itself, generated inside the
 * proposal kernel from the label the tree holds AFTER its propagation proposals (as :584-595).  Labels are initialised
 * as above if none were set.  Parity for this stage is by injection (s3dmst_pms_apply); the generator's defined
 * deviations (round-start snapshot for the propagation labels, counter-based random numbers) are stated in
 * csrc/pms.cu; tests/test_gpu_parity.py replays the generator's own proposal stream through s3dmst_pms_apply. */
int s3dmst_pms_iterate(s3dmst_ctx* ctx, int view, int n_iter, unsigned seed);
/* a13 LabelToDisp (:189-201) followed by the *(Dmax-1) of :900-902. */
int s3dmst_label_to_disp(s3dmst_ctx* ctx, int view);

int s3dmst_set_disparity(s3dmst_ctx* ctx, int view, const float* disp);
int s3dmst_get_disparity(s3dmst_ctx* ctx, int view, float* disp);
/* a14 leftRightConsistencyCheck (:632-710): invalid left pixels -> 0; fill != 0 runs the scan-line fill. */
int s3dmst_lr_check(s3dmst_ctx* ctx, int fill);

/* Post-filters the reference's author ran around the same outputs.
 * s3dmst_get_lr_mask: the left view's invalid-pixel mask of the last s3dmst_lr_check (1 = invalid), uint8 [H*W].
 * s3dmst_weighted_median: weightedMedianFilter (src/PatchMatchStereoGPU.cu:2436-2599) on a view's disparity map, applied
 *   to the pixels of `mask` (host uint8 [H*W]; NULL = the left-right check's mask, left view only): window (2*radius+1)^2
 *   (the reference: radius 10), weight exp(-sqrt(|dR|+|dG|+|dB|) * gamma) against the centre pixel of the view's image
 *   (gamma 0.1 for 0..255 intensities), stable sort by disparity, first disparity whose running weight reaches half.
 *   Every pixel reads the map as it was before the call (the reference filters in place and races).
 * s3dmst_wmf_table: the 766 weights exp(-sqrt(i) * gamma) the filter uses (host-side, no device needed): parity dumps.
 * s3dmst_norm_factor: 1 / (tree filter of the all-ones volume) per pixel, double [H*W] (ComputeMSTCostNormFactor /
 *   cost_norm_factor, PatchMatchStereoGPU.cu:5333-5429, :5898-5919): multiplying a view's aggregated costs by it is the
 *   reference's "normalised aggregation" (a per-pixel positive factor: it rescales `best`, never changes a winner). */
int s3dmst_get_lr_mask(s3dmst_ctx* ctx, uint8_t* mask);
int s3dmst_weighted_median(s3dmst_ctx* ctx, int view, int radius, float gamma, const uint8_t* mask);
void s3dmst_wmf_table(float gamma, float* tab766);
int s3dmst_norm_factor(s3dmst_ctx* ctx, int view, double* norm_factor);

/* The step after the path in its caller (src/stereo_Yin.cpp:218-243): left disparities below disp_floor are raised to it
 * (in the context's left map), then cv::reprojectImageTo3D(disp, xyz, Q, handle_missing) with the 4x4 row-major Q of
 * stereoRectify — bit-identical to OpenCV's CV_32F path — and the packed colour of the point cloud
 * (r << 16 | g << 8 | b of the left image).  xyz float[H*W][3], rgb uint32[H*W]; either may be NULL. */
int s3dmst_reproject_to_3d(s3dmst_ctx* ctx, const double* Q, float disp_floor, int handle_missing, float* xyz, uint32_t* rgb);

/* Whole dense pipeline on the current images: forests, AD+gradient matching cost, aggregation + WTA for both views,
 * LR check (+fill).  Outputs float[H][W] (NULL = leave on device).  By default (params.fuse_cost) the matching cost is
 * computed inside the aggregation kernel and no cost volume is built; the results are bit-identical either way. */
int s3dmst_run_dense(s3dmst_ctx* ctx, int D, int fill, float* left_disp, float* right_disp);

/* a1: the reference's own pipeline (stereo3dmst, Stereo3DMST.cpp:805-904) on the current images: forests (built if
 * missing), random plane initialisation (:390-430), params.num_iter rounds of MST_PMS per view with the library's
 * on-device proposal generator (:858-889; s3dmst_pms_iterate), LabelToDisp and *(Dmax-1) (:189-201, :900-902), left-right
 * check (:904; the reference passes fill = 0).  Volumes given through s3dmst_set_cost_volume (mc-cnn's left.bin /
 * right.bin) are used as they are; otherwise the a2' volume is built and ingested.  Outputs float[H][W] (NULL = leave on
 * the device; with both NULL the call returns without waiting).  This is what the header-compatible stereo3dmst() of
 * csrc/stereo3dmst_shim.cpp calls. */
int s3dmst_run(s3dmst_ctx* ctx, int Dmax, unsigned seed, int fill, float* left_disp, float* right_disp);

/* The same pipeline over a batch of frames, one context per frame (all on one device, same image size; images set
 * with s3dmst_set_images on each).  Forest and cost stages of the frames run concurrently on the contexts' streams
 * (one host thread each), the tree aggregation of ALL frames is one launch (frames are independent: the trees of a
 * batch fill the GPU where a single pair waits on its deepest tree), then LR check/fill per frame.
 * left_disp[i] / right_disp[i]: float[H][W] per frame (the arrays or any entry may be NULL). */
int s3dmst_run_dense_batch(s3dmst_ctx** ctxs, int n, int D, int fill, float** left_disp, float** right_disp);
/* The same for a stream of batches: returns once the copies into left_disp / right_disp (pinned host memory for a truly
 * asynchronous copy) are QUEUED on the frames' streams; they are complete after s3dmst_sync on the frame's context (or
 * any later synchronising call on it).  The next batch on the same contexts (s3dmst_set_images_async + this call) can be
 * queued at once: a frame's upload and forest start while the other frames' results are still on their way out. */
int s3dmst_run_dense_batch_async(s3dmst_ctx** ctxs, int n, int D, int fill, float** left_disp, float** right_disp);
/* The two halves of s3dmst_run_dense_batch, for software pipelining over consecutive batches (two sets of contexts,
 * two host threads): front = forests + cost volumes of every frame (latency-bound, many small kernels), back = the
 * joint aggregation launch (HBM-bound), LR check/fill and the copies.  front(k+1) may run while back(k) does. */
int s3dmst_batch_front(s3dmst_ctx** ctxs, int n, int D);
int s3dmst_batch_back(s3dmst_ctx** ctxs, int n, int D, int fill, float** left_disp, float** right_disp);

/* Per-stage device time of the most recent call, in ms (CUDA events on the context's stream). */
enum {
    S3DMST_T_FOREST = 0, S3DMST_T_COST = 1, S3DMST_T_AGG = 2, S3DMST_T_POST = 3, S3DMST_T_PMS = 4, S3DMST_T_COUNT = 5
};
double s3dmst_stage_ms(s3dmst_ctx* ctx, int stage);
/* The same accumulated over every call since the last reset (a caller that queues call after call without synchronising
 * still gets every sample: up to 4 per stage and view may be in flight).  Waits for the samples still running.
 * *samples = number of timed intervals in the total (may be NULL); reset != 0 clears the accumulator afterwards. */
double s3dmst_stage_total_ms(s3dmst_ctx* ctx, int stage, int* samples, int reset);
/* Number of kernels this library has launched on the context since creation. */
long long s3dmst_launch_count(const s3dmst_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* S3DMST_H_ */
