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

// ---- proposal generator of s3dmst_pms_iterate / s3dmst_run (MST_PMS, Stereo3DMST.cpp:546-629).
// The reference draws from one sequential minstd_rand0 stream (and std::rand) whose position depends on every earlier
// tree's data: unusable in parallel, and parity for this stage is by proposal injection.  The library's own
// generator is counter based: draw (seed, round, tree, slot) is a pure function, so any tree can be generated anywhere.
S3_HD uint64_t s3_mix64(uint64_t z) {  // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
S3_HD uint32_t s3_rng(uint32_t seed, uint32_t round, uint32_t tree, uint32_t slot) {
    const uint64_t a = s3_mix64(((uint64_t)seed << 32 | round) + 0x9E3779B97F4A7C15ull);
    return (uint32_t)(s3_mix64(a ^ ((uint64_t)tree << 32 | slot)) >> 32);
}
S3_HD float s3_rng_unit(uint32_t r) { return (float)(r >> 8) * 5.9604644775390625e-08f; }           // U[0,1), 24 bits, exact
S3_HD float s3_rng_sym(uint32_t r) { return S3_FSUB(S3_FMUL(2.0f, s3_rng_unit(r)), 1.0f); }         // U[-1,1), exact
#define S3_SLOT_REFINE_PIXEL 0x80000000u   // slot of the refinement pixel draw; ladder step k draws slots 0x40000000 + 4k .. + 3
#define S3_SLOT_LADDER 0x40000000u
#define S3_MAX_LADDER 64

// index of the sampled pixel inside a tree of `size` pixels: (int)((dice+1)*0.5*size) of :569 with the Q11 clamp
S3_HD int s3_sample_index(uint32_t r, int size) {
    const int i = (int)S3_FMUL(s3_rng_unit(r), (float)size);
    return i < size - 1 ? i : size - 1;
}

// RANDOM REFINEMENT ladder (:584-625) around label (a, b, c) of pixel (px, py): max_d = Dmax/2, /2, ... > floor with
// max_n = 1, 1/2, ...; step k is skipped when its disparity leaves [0, Dmax] (:602).  Writes up to S3_MAX_LADDER
// labels to out[3*i..], returns their number.  fp32 throughout, every operation individually rounded.
S3_HD int s3_refine_ladder(float a, float b, float c, float px, float py, int max_disp, float floor_d, uint32_t seed, uint32_t round,
                           uint32_t tree, float* out) {
    const float nz = 1.0f / sqrtf(S3_FADD(S3_FADD(S3_FMUL(a, a), S3_FMUL(b, b)), 1.0f));
    const float nx = S3_FMUL(-a, nz), ny = S3_FMUL(-b, nz);
    const float d = S3_FADD(S3_FADD(S3_FMUL(px, a), S3_FMUL(py, b)), c);
    float max_n = 1.0f, max_d = S3_FMUL(0.5f, (float)max_disp);
    int n = 0;
    for (uint32_t k = 0; max_d > floor_d && k < S3_MAX_LADDER; k++, max_d = S3_FMUL(max_d, 0.5f), max_n = S3_FMUL(max_n, 0.5f)) {
        const float rd = S3_FADD(d, S3_FMUL(s3_rng_sym(s3_rng(seed, round, tree, S3_SLOT_LADDER + 4 * k)), max_d));
        if (rd < 0.0f || rd > (float)max_disp) continue;
        float rx = S3_FADD(nx, S3_FMUL(s3_rng_sym(s3_rng(seed, round, tree, S3_SLOT_LADDER + 4 * k + 1)), max_n));
        float ry = S3_FADD(ny, S3_FMUL(s3_rng_sym(s3_rng(seed, round, tree, S3_SLOT_LADDER + 4 * k + 2)), max_n));
        float rz = S3_FADD(nz, S3_FMUL(s3_rng_sym(s3_rng(seed, round, tree, S3_SLOT_LADDER + 4 * k + 3)), max_n));
        const float inv = 1.0f / sqrtf(S3_FADD(S3_FADD(S3_FMUL(rx, rx), S3_FMUL(ry, ry)), S3_FMUL(rz, rz)));
        rx = S3_FMUL(rx, inv);
        ry = S3_FMUL(ry, inv);
        rz = fabsf(S3_FMUL(rz, inv));
        out[3 * n] = -rx / rz;
        out[3 * n + 1] = -ry / rz;
        out[3 * n + 2] = S3_FADD(S3_FADD(S3_FMUL(rx, px), S3_FMUL(ry, py)), S3_FMUL(rz, rd)) / rz;
        n++;
    }
    return n;
}

// ---- slanted-plane matching cost straight from the two images (north-star item 1; pm::PatchMatch, src/pm.cpp).
// Gradients (pm.cpp:70-88): cv::cvtColor(BGR2GRAY) on u8 — OpenCV's fixed-point weights — then cv::Sobel(CV_32F, ksize 3,
// BORDER_DEFAULT = reflect-101) / 8.  The gray weights are those of the OpenCV this repo can check against (cv2 4.13:
// (B*3735 + G*19235 + R*9798 + 2^14) >> 15); the reference's OpenCV 3.4.3 used the 14-bit set (1868, 9617, 4899), which
// differs by one gray level on ~0.3 % of pixels.
S3_HD int s3_cv_gray(int b, int g, int r) { return (b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15; }
S3_HD int s3_reflect101(int i, int n) {  // cv::borderInterpolate(BORDER_REFLECT_101)
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}
// cv::saturate_cast<uchar>(float): cvRound (nearest, ties to even) then clamped to 0..255
S3_HD int s3_sat_u8(float v) {
    const int i = (int)rintf(v);
    return i < 0 ? 0 : (i > 255 ? 255 : i);
}
// Cost of plane (a, b, c) at pixel (x, y) of view `view` (0 = left: the match lies at x - d in the other image; 1 = right:
// x + d), pm.cpp:130-154 for ONE pixel (the reference sums it over a 35x35 window with adaptive weights; here the tree
// filter is the aggregation) and dissimilarity :97-104:
//   d = a*x + b*y + c (pm.h:153-156); outside [0, max_disp] -> oob (the reference's PLANE_PENALTY);
//   match = x -+ d; xm = (int)match; wm = 1 - (match - xm); xm clamped to [0, W-2];
//   colour and gradient of the match = vecAverage(pixel xm, pixel xm+1, wm) (pm.h:166-169) — both assigned to BYTE vectors
//   in the reference (cv::Vec3b mcolo, cv::Vec2b mgrad, :150-151), so every product wm*v and (1-wm)*v and their sum is
//   saturate_cast to u8 (negative gradients clamp to 0: reproduced);
//   cost = (1-alpha) * min(L1 colour, tau_c) + alpha * min(L1 gradient, tau_g), then * scale (the path's cost range).
// self_bgr / self_grad: this pixel; orow_bgr / orow_grad: row y of the other view (BGR u8 triples, gradient float pairs).
S3_HD float s3_plane_cost(const uint8_t* self_bgr, const float* self_grad, const uint8_t* orow_bgr, const float* orow_grad, int x, int y, int W,
                          int view, float a, float b, float c, int max_disp, float alpha, float tau_c, float tau_g, float scale, float oob) {
    const float d = S3_FADD(S3_FADD(S3_FMUL(a, (float)x), S3_FMUL(b, (float)y)), c);
    if (!(d >= 0.0f) || d > (float)max_disp || W < 2) return oob;
    const float match = view ? S3_FADD((float)x, d) : S3_FSUB((float)x, d);
    int xm = (int)match;
    const float wm = S3_FSUB(1.0f, S3_FSUB(match, (float)xm));
    if (xm > W - 2) xm = W - 2;
    if (xm < 0) xm = 0;
    const float w1 = S3_FSUB(1.0f, wm);
    float cc = 0.0f;
    for (int k = 0; k < 3; k++) {  // Vec3b = sat(wm * p0) + sat((1 - wm) * p1), saturated again
        int m = s3_sat_u8(S3_FMUL(wm, (float)orow_bgr[3 * xm + k])) + s3_sat_u8(S3_FMUL(w1, (float)orow_bgr[3 * (xm + 1) + k]));
        m = m > 255 ? 255 : m;
        const int df = (int)self_bgr[k] - m;
        cc = S3_FADD(cc, (float)(df < 0 ? -df : df));
    }
    float cg = 0.0f;
    for (int k = 0; k < 2; k++) {  // Vec2f average, then converted to Vec2b
        const float gm = (float)s3_sat_u8(S3_FADD(S3_FMUL(wm, orow_grad[2 * xm + k]), S3_FMUL(w1, orow_grad[2 * (xm + 1) + k])));
        const float df = S3_FSUB(self_grad[k], gm);
        cg = S3_FADD(cg, df < 0.0f ? -df : df);
    }
    cc = cc < tau_c ? cc : tau_c;
    cg = cg < tau_g ? cg : tau_g;
    return S3_FMUL(S3_FADD(S3_FMUL(S3_FSUB(1.0f, alpha), cc), S3_FMUL(alpha, cg)), scale);
}
