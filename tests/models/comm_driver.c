/* tests/models/comm_driver.c — TEST INFRASTRUCTURE: a plain-C caller of the label-sharded path through include/s3dmst.h
 * (no Python, no torch on the data path): communicator from a ncclUniqueId, this rank's label range, MIN-LOC reduction,
 * disparity maps.  Built by tests/test_comm_c.py with gcc and called with nranks = 1 on the single-GPU test box (NCCL
 * with one rank still runs both all-reduces); with more ranks each process calls shard_run with its rank and the id
 * rank 0 made. */
#include <stdio.h>
#include <string.h>

#include "s3dmst.h"

int shard_make_id(unsigned char* id128) { return s3dmst_comm_unique_id(id128); }

int shard_run(const unsigned char* id128, int rank, int nranks, int device, const unsigned char* left_bgr, const unsigned char* right_bgr, int W, int H,
              int D, float* left_disp, float* right_disp, int* d0, int* d1, double* minloc_ms, char* err, int errlen) {
    s3dmst_ctx* ctx = NULL;
    int rc = s3dmst_create(&ctx, device, NULL, NULL);
    if (rc) { snprintf(err, errlen, "%s", s3dmst_last_error(NULL)); return rc; }
    if (!rc) rc = s3dmst_comm_init(ctx, id128, rank, nranks);
    if (!rc) rc = s3dmst_comm_label_range(ctx, D, d0, d1);
    if (!rc) rc = s3dmst_set_images(ctx, left_bgr, right_bgr, W, H, 3 * W);
    if (!rc) rc = s3dmst_build_forest(ctx, 0);
    if (!rc) rc = s3dmst_build_forest(ctx, 1);
    if (!rc) rc = s3dmst_build_cost_volume(ctx, D, 0);
    if (!rc) rc = s3dmst_aggregate_dense_sharded(ctx, D);   /* asynchronous: aggregation + reduction queued */
    if (!rc) rc = s3dmst_dense_to_disparity(ctx, 0);
    if (!rc) rc = s3dmst_dense_to_disparity(ctx, 1);
    if (!rc) rc = s3dmst_lr_check(ctx, 1);
    if (!rc) rc = s3dmst_get_disparity(ctx, 0, left_disp);
    if (!rc) rc = s3dmst_get_disparity(ctx, 1, right_disp);
    if (!rc) *minloc_ms = s3dmst_comm_minloc_ms(ctx);
    if (rc) snprintf(err, errlen, "%s", s3dmst_last_error(ctx));
    s3dmst_destroy(ctx);
    return rc;
}
