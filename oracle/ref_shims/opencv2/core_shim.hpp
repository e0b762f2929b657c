// oracle/ref_shims/opencv2/core_shim.hpp — TEST INFRASTRUCTURE ONLY.
//
// Minimal stand-in for the handful of OpenCV 3.x types/functions that
// /root/reference/src/Stereo3DMST.cpp touches, so that the reference translation unit can be
// compiled UNMODIFIED (from where it lies) into oracle/_ref/ in a container without OpenCV C++
// headers.  Nothing here implements stereo logic; it is storage + three pixel utilities:
//   cv::Mat (create / ptr / at / convertTo / =scalar / *=scalar), cv::split, cv::medianBlur(3x3, 8U,
//   replicated border — checked against cv2.medianBlur by tests/test_oracle_vs_ref.py), and
//   no-op cv::imshow / cv::waitKey (the reference opens debug windows inside the hot path).
#pragma once
#include <algorithm>
#include <cstring>
#include <functional>
#include <memory>
#include <queue>
#include <random>
#include <string>
#include <unistd.h>
#include <vector>

typedef unsigned char uchar;
#ifndef MAX
#define MAX(a, b) ((a) < (b) ? (b) : (a))
#endif
#ifndef MIN
#define MIN(a, b) ((a) > (b) ? (b) : (a))
#endif

#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)

namespace cv {

struct Vec3b {
    uchar val[3];
    uchar& operator[](int i) { return val[i]; }
};

class Mat {
public:
    int rows = 0, cols = 0;
    uchar* data = nullptr;
    Mat() {}
    Mat(int r, int c, int t) { create(r, c, t); }
    Mat(int r, int c, int t, void* ext) : rows(r), cols(c), data((uchar*)ext), type_(t) {}  // wraps, no copy
    void create(int r, int c, int t) {
        if (data && r == rows && c == cols && t == type_) return;
        rows = r;
        cols = c;
        type_ = t;
        buf_ = std::make_shared<std::vector<uchar>>((size_t)r * c * elemSize());
        data = buf_->data();
    }
    int type() const { return type_; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return true; }                                  // this stand-in never pads rows
    struct MatStep {                                                            // cv::Mat::step converts to size_t (bytes per row)
        const Mat* m;
        operator size_t() const { return (size_t)m->cols * m->elemSize(); }
    };
    MatStep step{this};
    Mat(const Mat& o) : rows(o.rows), cols(o.cols), data(o.data), type_(o.type_), buf_(o.buf_) {}
    Mat& operator=(const Mat& o) {
        rows = o.rows; cols = o.cols; data = o.data; type_ = o.type_; buf_ = o.buf_;
        return *this;
    }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize1() const { return depth() == CV_8U ? 1 : depth() == CV_32F ? 4 : 8; }
    size_t elemSize() const { return elemSize1() * channels(); }
    template <class T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * cols * elemSize()); }
    template <class T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * cols * elemSize()); }
    template <class T> T& at(int y, int x) { return ptr<T>(y)[x]; }

    // dst = saturate_cast<rtype>(src*alpha + beta); only the conversions the reference needs
    void convertTo(Mat& dst, int rtype, double alpha = 1.0, double beta = 0.0) const {
        Mat out;
        out.create(rows, cols, CV_MAKETYPE(rtype & 7, channels()));
        const size_t n = (size_t)rows * cols * channels();
        const int sd = depth(), dd = rtype & 7;
        for (size_t i = 0; i < n; i++) {
            double v = sd == CV_8U ? (double)data[i] : sd == CV_32F ? (double)((const float*)data)[i] : ((const double*)data)[i];
            if (sd == CV_32F && dd == CV_32F) {  // OpenCV scales 32F->32F in float
                ((float*)out.data)[i] = ((const float*)data)[i] * (float)alpha + (float)beta;
                continue;
            }
            v = v * alpha + beta;
            if (dd == CV_64F)
                ((double*)out.data)[i] = v;
            else if (dd == CV_32F)
                ((float*)out.data)[i] = (float)v;
            else
                out.data[i] = (uchar)std::min(255.0, std::max(0.0, std::nearbyint(v)));
        }
        dst = out;
    }
    Mat& operator=(double s) {
        const size_t n = (size_t)rows * cols * channels();
        for (size_t i = 0; i < n; i++) {
            if (depth() == CV_8U)
                data[i] = (uchar)s;
            else if (depth() == CV_32F)
                ((float*)data)[i] = (float)s;
            else
                ((double*)data)[i] = s;
        }
        return *this;
    }
    Mat& operator*=(double s) {  // Mat::operator*= == convertTo(*this, type, s): float math for 32F
        const size_t n = (size_t)rows * cols * channels();
        for (size_t i = 0; i < n; i++) {
            if (depth() == CV_32F)
                ((float*)data)[i] = ((float*)data)[i] * (float)s;
            else if (depth() == CV_64F)
                ((double*)data)[i] *= s;
        }
        return *this;
    }

private:
    int type_ = 0;
    std::shared_ptr<std::vector<uchar>> buf_;
};

inline void split(const Mat& src, std::vector<Mat>& mv) {
    const int cn = src.channels();
    mv.resize(cn);
    for (int c = 0; c < cn; c++) {
        mv[c] = Mat();
        mv[c].create(src.rows, src.cols, src.depth());
        const size_t n = (size_t)src.rows * src.cols, es = src.elemSize1();
        for (size_t i = 0; i < n; i++) memcpy(mv[c].data + i * es, src.data + (i * cn + c) * es, es);
    }
}

inline void medianBlur(const Mat& src, Mat& dst, int ksize) {
    Mat out;
    out.create(src.rows, src.cols, src.type());
    const int r = ksize / 2, W = src.cols, H = src.rows;
    std::vector<uchar> win((size_t)ksize * ksize);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int k = 0;
            for (int dy = -r; dy <= r; dy++)
                for (int dx = -r; dx <= r; dx++) {
                    int yy = y + dy < 0 ? 0 : (y + dy >= H ? H - 1 : y + dy);
                    int xx = x + dx < 0 ? 0 : (x + dx >= W ? W - 1 : x + dx);
                    win[k++] = src.data[(size_t)yy * W + xx];
                }
            std::sort(win.begin(), win.end());
            out.data[(size_t)y * W + x] = win[win.size() / 2];
        }
    dst = out;
}

inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int = 0) { return -1; }

}  // namespace cv
