// tests/models/hd_math_host.cpp — compiles stereomatch_b200/csrc/hd_math.h for the HOST so the per-element
// arithmetic the CUDA kernels use can be checked against the oracle without a GPU (tests/test_hd_math.py).
#include <cmath>
#include <cstddef>
#include <cstdint>
#include "../../stereomatch_b200/csrc/hd_math.h"

extern "C" {
void hd_median3(const uint8_t* in, int W, int H, uint8_t* out) {
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int v[9], k = 0;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    int yy = y + dy < 0 ? 0 : (y + dy >= H ? H - 1 : y + dy);
                    int xx = x + dx < 0 ? 0 : (x + dx >= W ? W - 1 : x + dx);
                    v[k++] = in[yy * W + xx];
                }
            out[y * W + x] = (uint8_t)s3_median9(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8]);
        }
}
// same per-pixel logic as k_cost_adgrad, [d][p] output for direct comparison with orc_cost_adgrad
void hd_cost_adgrad(const uint8_t* L, const uint8_t* R, int W, int H, int D, float* lv, float* rv) {
    const size_t N = (size_t)W * H;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const size_t p = (size_t)y * W + x, rb = (size_t)y * W;
            auto g = [&](const uint8_t* I, size_t q) { return s3_gray(I[3 * q], I[3 * q + 1], I[3 * q + 2]); };
            const bool has_next = x + 1 < W;
            for (int d = 0; d < D; d++) {
                float cl = 3.0f, cr = 3.0f;
                const int xr = x - d;
                if (xr >= 0 && has_next)
                    cl = s3_adgrad(R[3 * (rb + xr)], R[3 * (rb + xr) + 1], R[3 * (rb + xr) + 2], g(R, rb + xr), g(R, rb + xr + 1),
                                   L[3 * p], L[3 * p + 1], L[3 * p + 2], g(L, p), g(L, p + 1));
                const int xl = x + d;
                if (xl + 1 < W)
                    cr = s3_adgrad(R[3 * p], R[3 * p + 1], R[3 * p + 2], g(R, p), g(R, p + 1), L[3 * (rb + xl)],
                                   L[3 * (rb + xl) + 1], L[3 * (rb + xl) + 2], g(L, rb + xl), g(L, rb + xl + 1));
                lv[(size_t)d * N + p] = cl;
                rv[(size_t)d * N + p] = cr;
            }
        }
}
float hd_label_cost(const float* row, float a, float b, float c, int x, int y, int D, float oob) {
    return s3_label_cost(row, a, b, c, x, y, D, oob);
}
float hd_label_disp(float a, float b, float c, int x, int y, int D) { return s3_label_disp(a, b, c, x, y, D); }
float hd_ingest(float v, float cap, float off, float sc) { return s3_ingest(v, cap, off, sc); }
}

// ---- host mirror of the on-device proposal generator (s3dmst_pms_iterate): the same hd_math.h functions, called in
// the order the kernels call them
extern "C" {
// propagation proposals of one round: for tree t and its j-th neighbour (CSR, ascending), the label of a sampled pixel
void hd_gen_propagation(int T, const int* adj_ptr, const int* adj, const int* tree_start, const int* node_pixel, const float* abc,
                        unsigned seed, unsigned round, int* out_tree, float* out_labels) {
    for (int t = 0; t < T; t++)
        for (int j = 0; j < adj_ptr[t + 1] - adj_ptr[t]; j++) {
            const int e = adj_ptr[t] + j, nb = adj[e];
            const int b = tree_start[nb], sz = tree_start[nb + 1] - b;
            const int pix = node_pixel[b + s3_sample_index(s3_rng(seed, round, (uint32_t)t, (uint32_t)j), sz)];
            out_tree[e] = t;
            for (int k = 0; k < 3; k++) out_labels[3 * (size_t)e + k] = abc[3 * (size_t)pix + k];
        }
}
// refinement ladders of one round from the labels as the propagation proposals left them; returns the proposal count
int hd_gen_refinement(int T, int W, const int* tree_start, const int* node_pixel, const float* abc, int Dmax, float floor_d,
                      unsigned seed, unsigned round, int* out_tree, float* out_labels) {
    int n = 0;
    for (int t = 0; t < T; t++) {
        const int b = tree_start[t], sz = tree_start[t + 1] - b;
        const int pix = node_pixel[b + (int)(s3_rng(seed, round, (uint32_t)t, S3_SLOT_REFINE_PIXEL) % (uint32_t)sz)];
        float lab[3 * S3_MAX_LADDER];
        const int k = s3_refine_ladder(abc[3 * (size_t)pix], abc[3 * (size_t)pix + 1], abc[3 * (size_t)pix + 2], (float)(pix % W), (float)(pix / W), Dmax,
                                       floor_d, seed, round, (uint32_t)t, lab);
        for (int i = 0; i < k; i++, n++) {
            out_tree[n] = t;
            for (int c = 0; c < 3; c++) out_labels[3 * (size_t)n + c] = lab[3 * i + c];
        }
    }
    return n;
}
}

// ---- host build of the plane cost (pms_cost_mode 1) for the CPU parity test against the oracle
extern "C" void hd_plane_cost_map(int view, const uint8_t* left_bgr, const uint8_t* right_bgr, const float* left_grad, const float* right_grad, int W, int H,
                                  float a, float b, float c, int max_disp, float alpha, float tau_c, float tau_g, float scale, float oob, float* out) {
    const uint8_t* simg = view ? right_bgr : left_bgr; const uint8_t* oimg = view ? left_bgr : right_bgr;
    const float* sgrad = view ? right_grad : left_grad; const float* ograd = view ? left_grad : right_grad;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const size_t p = (size_t)y * W + x;
            out[p] = s3_plane_cost(simg + 3 * p, sgrad + 2 * p, oimg + 3 * (size_t)y * W, ograd + 2 * (size_t)y * W, x, y, W, view, a, b, c, max_disp, alpha, tau_c,
                                   tau_g, scale, oob);
        }
}
extern "C" void hd_pm_gradients(const uint8_t* bgr, int W, int H, float* grad) {
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int g[3][3];
            for (int j = 0; j < 3; j++)
                for (int i = 0; i < 3; i++) {
                    const uint8_t* c = bgr + 3 * ((size_t)s3_reflect101(y + j - 1, H) * W + s3_reflect101(x + i - 1, W));
                    g[j][i] = s3_cv_gray(c[0], c[1], c[2]);
                }
            grad[2 * ((size_t)y * W + x)] = (float)((g[0][2] + 2 * g[1][2] + g[2][2]) - (g[0][0] + 2 * g[1][0] + g[2][0])) / 8.f;
            grad[2 * ((size_t)y * W + x) + 1] = (float)((g[2][0] + 2 * g[2][1] + g[2][2]) - (g[0][0] + 2 * g[0][1] + g[0][2])) / 8.f;
        }
}
