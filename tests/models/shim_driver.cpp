// tests/models/shim_driver.cpp — TEST INFRASTRUCTURE: calls the header-compatible stereo3dmst() the way
// src/stereo_Yin.cpp:205-210 does (cv::Mat in, cv::Mat out, timer around it) and hands the result to ctypes.
#include <cstring>
#include <string>

#include <opencv2/highgui/highgui.hpp>

extern "C" void stereo3dmst(std::string left_name, std::string right_name, cv::Mat& leftImg, cv::Mat& rightImg,
                            cv::Mat& leftDisp, cv::Mat& rightDisp, std::string data_cost, int Dmax);
void startTimer();
double getTimer();

extern "C" int shim_call(const unsigned char* l, const unsigned char* r, int W, int H, int Dmax, const char* data_cost,
                         float* out_l, float* out_r, double* ms, int* out_rows, int* out_cols) {
    cv::Mat L(H, W, CV_8UC3), R(H, W, CV_8UC3), DL, DR;
    memcpy(L.data, l, (size_t)W * H * 3);
    memcpy(R.data, r, (size_t)W * H * 3);
    startTimer();
    stereo3dmst("img1r.png", "img2r.png", L, R, DL, DR, data_cost, Dmax);
    *ms = getTimer();
    *out_rows = DL.rows;
    *out_cols = DL.cols;
    if (DL.rows != H || DL.cols != W || DL.type() != CV_32F || DR.type() != CV_32F) return -1;
    memcpy(out_l, DL.data, (size_t)W * H * 4);
    memcpy(out_r, DR.data, (size_t)W * H * 4);
    return 0;
}

// a right image of another size and type: the shim must return before reading either image
extern "C" int shim_call_mismatched(int W, int H) {
    cv::Mat L(H, W, CV_8UC3), R(H / 2, W, CV_32F), DL, DR;
    memset(L.data, 0, (size_t)W * H * 3);
    stereo3dmst("img1r.png", "img2r.png", L, R, DL, DR, "ADGRAD", 16);
    return DL.rows;
}
