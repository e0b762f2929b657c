// stereomatch_b200/csrc/stereo3dmst_shim.cpp — header-compatible entry point over the C ABI.
//
// Defines the three symbols the reference declares in include/Stereo3DMST.h:7-11
//     extern "C" void stereo3dmst(std::string, std::string, cv::Mat&, cv::Mat&, cv::Mat&, cv::Mat&, std::string, int);
//     void startTimer();   double getTimer();
// with the same argument meaning, ownership and error behaviour (src/Stereo3DMST.cpp:714-912), so
// src/stereo_Yin.cpp:207 links against this file + libs3dmst.so instead of src/Stereo3DMST.cpp.
// cv::Mat is only unwrapped here (rows / cols / data); everything else goes through include/s3dmst.h.
//
// Build (maintainer, with OpenCV):  g++ -std=c++11 -fPIC -shared stereo3dmst_shim.cpp -I<repo>/include
//                                       `pkg-config --cflags opencv` -L<repo>/stereomatch_b200 -ls3dmst
// In this repo's tests it is compiled against oracle/ref_shims (the container has no OpenCV C++ headers).
//
// data_cost (the reference's selector, :725-759):
//   "MCCNN_acrt" / "MCCNN_fst"  the reference shells out to mc-cnn and then mmaps mc-cnn-master/left.bin and
//                               right.bin (float32 [1][Dmax][rows][cols], :764-775).  mc-cnn is an external
//                               process and is not spawned here: the two files are read if they exist, else
//                               the call prints the reference's message and returns with the outputs
//                               allocated but unfilled, exactly as the reference does on a failed step (:727-759).
//   "ADGRAD"                    extension: truncated colour + gradient volume built on the GPU.
//   anything else               prints "wrong data cost" and returns (:756-759).
// Search: by default what the reference does — random plane initialisation, num_iter (100) rounds of MST_PMS per view,
// LabelToDisp, left-right check (s3dmst_run): sub-pixel slanted-plane disparities.  The proposal stream is the library's
// own (include/s3dmst.h: s3dmst_pms_iterate), so the maps agree with the reference's in quality, not bit for bit.
// S3DMST_MODE=dense opts into the dense integer-label WTA instead.  No GUI windows, no chdir.
// Inputs are validated before anything is allocated on their account (the reference trusts them): both images non-empty
// CV_8UC3 of one size, Dmax > 0; rows are read with each Mat's own step.
#include <sys/time.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include <opencv2/highgui/highgui.hpp>
#include <opencv2/imgproc/imgproc.hpp>

#include "s3dmst.h"

static struct timeval g_timer_start;  // process-global like the reference's (Stereo3DMST.cpp:15)

void startTimer() { gettimeofday(&g_timer_start, NULL); }

double getTimer() {  // milliseconds since startTimer(), Stereo3DMST.cpp:20-26
    struct timeval now;
    gettimeofday(&now, NULL);
    return (now.tv_sec - g_timer_start.tv_sec) * 1000.0 + (now.tv_usec - g_timer_start.tv_usec) / 1000.0;
}

static bool read_volume(const char* path, size_t count, std::vector<float>& out) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    out.resize(count);
    const size_t got = fread(out.data(), sizeof(float), count, f);
    fclose(f);
    return got == count;
}

extern "C" void stereo3dmst(std::string left_name, std::string right_name, cv::Mat& leftImg, cv::Mat& rightImg,
                            cv::Mat& leftDisp, cv::Mat& rightDisp, std::string data_cost, int Dmax) {
    (void)left_name;  // only forwarded to mc-cnn's command line by the reference (:733-748)
    (void)right_name;
    if (leftImg.empty() || rightImg.empty() || leftImg.type() != CV_8UC3 || rightImg.type() != CV_8UC3 || leftImg.rows != rightImg.rows ||
        leftImg.cols != rightImg.cols || Dmax <= 0) {
        std::cout << "stereo3dmst: need two non-empty CV_8UC3 images of one size and Dmax > 0" << std::endl;
        return;
    }
    const int rows = leftImg.rows, cols = leftImg.cols;
    leftDisp.create(rows, cols, CV_32F);  // :722-723
    rightDisp.create(rows, cols, CV_32F);
    // rows are addressed through each Mat's own step (a ROI or a padded Mat is not cols*3 wide); the C ABI takes one
    // stride for both images, so a pair with different steps is packed first
    const size_t lstep = (size_t)leftImg.step, rstep = (size_t)rightImg.step;
    std::vector<unsigned char> packed;
    const unsigned char* lptr = leftImg.data;
    const unsigned char* rptr = rightImg.data;
    size_t stride = lstep;
    if (lstep != rstep) {
        packed.resize((size_t)2 * rows * cols * 3);
        for (int y = 0; y < rows; y++) {
            memcpy(&packed[(size_t)y * cols * 3], leftImg.data + (size_t)y * lstep, (size_t)cols * 3);
            memcpy(&packed[((size_t)rows + y) * cols * 3], rightImg.data + (size_t)y * rstep, (size_t)cols * 3);
        }
        lptr = packed.data();
        rptr = packed.data() + (size_t)rows * cols * 3;
        stride = (size_t)cols * 3;
    }

    const bool mccnn = data_cost == "MCCNN_acrt" || data_cost == "MCCNN_fst";
    if (!mccnn && data_cost != "ADGRAD") {
        std::cout << "wrong data cost" << std::endl;  // :756-759
        return;
    }
    std::vector<float> lvol, rvol;
    if (mccnn) {
        const size_t count = (size_t)Dmax * rows * cols;
        if (!read_volume("mc-cnn-master/left.bin", count, lvol) || !read_volume("mc-cnn-master/right.bin", count, rvol)) {
            std::cout << "mc-cnn-master/left.bin / right.bin not readable" << std::endl;  // :727-731
            return;
        }
    }

    s3dmst_params P;
    s3dmst_default_params(&P);
    if (data_cost == "MCCNN_fst") {  // the "fast" nets score in [-1, 1]: (c + 1) / 2  (:792, PatchMatchStereoGPU.cu:4713-4745)
        P.cost_offset = 1.0f;
        P.cost_scale = 0.5f;
    }
    const char* mode = getenv("S3DMST_MODE");
    const bool pms = !(mode && !strcmp(mode, "dense"));
    if (!mccnn && pms) P.cost_scale = 1.0f / 6.0f;  // the a2' volume spans [0, 3]: scaled into the [0, 0.5] range the ingest caps at (:789-801)
    const char* dev_env = getenv("S3DMST_DEVICE");
    s3dmst_ctx* ctx = NULL;
    if (s3dmst_create(&ctx, dev_env ? atoi(dev_env) : 0, &P, NULL) != S3DMST_OK) {
        std::cout << "stereo3dmst: " << s3dmst_last_error(NULL) << std::endl;
        return;
    }
    int rc = s3dmst_set_images(ctx, lptr, rptr, cols, rows, (int)stride);
    if (!rc) rc = s3dmst_build_forest(ctx, 0);
    if (!rc) rc = s3dmst_build_forest(ctx, 1);
    if (!rc) {
        if (mccnn) {
            rc = s3dmst_set_cost_volume(ctx, 0, lvol.data(), Dmax, 1);
            if (!rc) rc = s3dmst_set_cost_volume(ctx, 1, rvol.data(), Dmax, 1);
        } else
            rc = s3dmst_build_cost_volume(ctx, Dmax, pms ? 1 : 0);
    }
    if (pms) {
        // :805-904: plane init, num_iter rounds of MST_PMS per view, LabelToDisp, LR check with fill = false
        if (!rc) rc = s3dmst_run(ctx, Dmax, 1u, 0, reinterpret_cast<float*>(leftDisp.data), reinterpret_cast<float*>(rightDisp.data));
    } else {
        for (int view = 0; view < 2 && !rc; view++) {
            rc = s3dmst_aggregate_dense(ctx, view, 0, Dmax, NULL, NULL);
            if (!rc) rc = s3dmst_dense_to_disparity(ctx, view);
        }
        if (!rc) rc = s3dmst_lr_check(ctx, 0);  // :904, fill = false
        if (!rc) rc = s3dmst_get_disparity(ctx, 0, reinterpret_cast<float*>(leftDisp.data));
        if (!rc) rc = s3dmst_get_disparity(ctx, 1, reinterpret_cast<float*>(rightDisp.data));
    }
    if (rc) std::cout << "stereo3dmst: " << s3dmst_last_error(ctx) << std::endl;
    s3dmst_destroy(ctx);
}
