// stereomatch_b200/csrc/hd_math.h
//
// Pure per-element arithmetic of the hot path, written once for device code and (for the CPU-side
// unit tests in tests/test_hd_math.py, which compile this header with g++) host code.  Every
// floating-point step is an explicitly rounded IEEE op — no FMA contraction — because parity with
// the reference's serial build is bit-exact.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define S3_HD __host__ __device__ __forceinline__
#else
#define S3_HD inline
#include <cmath>
#endif

#if defined(__CUDA_ARCH__)
#define S3_FMUL(a, b) __fmul_rn((a), (b))
#define S3_FADD(a, b) __fadd_rn((a), (b))
#define S3_FSUB(a, b) __fsub_rn((a), (b))
#define S3_DMUL(a, b) __dmul_rn((a), (b))
#define S3_DADD(a, b) __dadd_rn((a), (b))
#else  // host: translation units including this header are built with -ffp-contract=off
#define S3_FMUL(a, b) ((float)(a) * (float)(b))
#define S3_FADD(a, b) ((float)(a) + (float)(b))
#define S3_FSUB(a, b) ((float)(a) - (float)(b))
#define S3_DMUL(a, b) ((double)(a) * (double)(b))
#define S3_DADD(a, b) ((double)(a) + (double)(b))
#endif

// 3x3 median, Devillard's 19-exchange network.  cv::medianBlur(ksize=3), Stereo3DMST.cpp:226-228.
S3_HD void s3_sort2(int& a, int& b) {
    int lo = a < b ? a : b, hi = a < b ? b : a;
    a = lo;
    b = hi;
}
S3_HD int s3_median9(int p0, int p1, int p2, int p3, int p4, int p5, int p6, int p7, int p8) {
    s3_sort2(p1, p2); s3_sort2(p4, p5); s3_sort2(p7, p8);
    s3_sort2(p0, p1); s3_sort2(p3, p4); s3_sort2(p6, p7);
    s3_sort2(p1, p2); s3_sort2(p4, p5); s3_sort2(p7, p8);
    s3_sort2(p0, p3); s3_sort2(p5, p8); s3_sort2(p4, p7);
    s3_sort2(p3, p6); s3_sort2(p1, p4); s3_sort2(p2, p5);
    s3_sort2(p4, p7); s3_sort2(p4, p2); s3_sort2(p6, p4);
    s3_sort2(p4, p2);
    return p4;
}

// gray = .114 B + .587 G + .299 R  (PatchMatchStereoGPU.cu:1526-1527), left-to-right, fp32.
S3_HD float s3_gray(int b, int g, int r) {
    return S3_FADD(S3_FADD(S3_FMUL(0.114f, (float)b), S3_FMUL(0.587f, (float)g)), S3_FMUL(0.299f, (float)r));
}

// Truncated colour + gradient AD for one (ref, match) pixel pair and their right neighbours
// (PatchMatchStereoGPU.cu:1519-1536).  ref = right image at x, match = left image at x+d.
// b,g,r are 0..255; gray values are s3_gray() of the same pixels.
S3_HD float s3_adgrad(int rb, int rg, int rr, float rgray, float rgray_next, int mb, int mg, int mr, float mgray,
                      float mgray_next) {
    int l1 = (rb > mb ? rb - mb : mb - rb) + (rg > mg ? rg - mg : mg - rg) + (rr > mr ? rr - mr : mr - rr);
    // colour_l1 is a sum of three integer-valued floats < 2^24: exact in fp32, so the int sum is the same value
    float color_l1 = (float)l1;
    float g = S3_FSUB(mgray, rgray);
    g = S3_FADD(g, S3_FSUB(rgray_next, mgray_next));
    float cterm = (float)S3_DMUL((double)color_l1, 0.33333333333);
    cterm = cterm < 7.0f ? cterm : 7.0f;
    float ag = g < 0.0f ? -g : g;
    float gterm = ag < 2.0f ? ag : 2.0f;
    return S3_FADD(S3_FMUL(0.11f, cterm), S3_FMUL(0.89f, gterm));
}

// a2 ingest, Stereo3DMST.cpp:785-803.
S3_HD float s3_ingest(float v, float cap, float offset, float scale) {
    if (v != v) return cap;
    if (offset != 0.0f || scale != 1.0f) v = S3_FMUL(S3_FADD(v, offset), scale);
    return cap < v ? cap : v;  // std::min(cap, v)
}

// compute3DLabelCost, Stereo3DMST.cpp:103-118.  `row` points at this pixel's D costs (label-minor).
// x86 (int)NaN / (int)huge is INT_MIN => "floor < 0" => oob; reproduced explicitly (device casts saturate).
S3_HD float s3_label_cost(const float* row, float a, float b, float c, int x, int y, int max_disp, float oob) {
    const float disp = S3_FADD(S3_FADD(S3_FMUL((float)x, a), S3_FMUL((float)y, b)), c);
    if (!(disp == disp)) return oob;
    if (!(disp > -2147483648.0f && disp < 2147483648.0f)) return oob;
    const float dc = ceilf(disp), df = floorf(disp);
    const int dci = (int)dc, dfi = (int)df;
    if (dci >= max_disp || dfi < 0) return oob;
    return S3_FADD(S3_FMUL(S3_FSUB(dc, disp), row[dfi]), S3_FMUL(S3_FSUB(disp, df), row[dci]));
}

// LabelToDisp (:197) then *(Dmax-1) (:900-902).
S3_HD float s3_label_disp(float a, float b, float c, int x, int y, int max_disp) {
    float v = S3_FADD(S3_FADD(S3_FMUL((float)x, a), S3_FMUL((float)y, b)), c) / ((float)max_disp - 1.0f);
    v = v < 1.0f ? v : 1.0f;          // std::min(1.0f, v)
    v = 0.0f < v ? v : 0.0f;          // MAX(0.0f, v)
    return S3_FMUL(v, (float)max_disp - 1.0f);
}
