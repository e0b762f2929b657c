// oracle/s3dmst_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's Stereo3DMST hot path
// (lr-xiang/StereoMatch: src/Stereo3DMST.cpp + include/segment-graph.h + include/disjoint-set.h)
// and of the dense cost-volume / WTA formulas in src/PatchMatchStereoGPU.cu.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library, and only as the checker or the timed CPU baseline.  Nothing under
// stereomatch_b200/ links, imports or calls it.
//
// Parity status: PINNED.  Every stage below is checked bit-for-bit against the reference's own
// translation unit (src/Stereo3DMST.cpp compiled unmodified into oracle/_ref/ against minimal
// OpenCV/Boost container shims, see oracle/ref_shims/ and oracle/ref_driver.cpp) by
// tests/test_oracle_vs_ref.py, and against golden vectors generated from that build
// (tests/golden/, made by tests/golden/make_golden.py).  The dense cost volume (a2') has no
// reference counterpart that runs on the 3DMST path (the reference shells out to mc-cnn), so
// that one function is "defined here" from PatchMatchStereoGPU.cu:1482-1550 with the two
// unspecified spots stated in the comments.
//
// All citations are relative to /root/reference.  Build: see oracle/Makefile
// (-O2 -ffp-contract=off for parity; a second -O3 -march=x86-64-v3 -ffp-contract=off build for timing).

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <queue>
#include <random>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// glibc random()/rand() TYPE_3 additive-feedback generator, restated so the oracle does not
// depend on (or disturb) the process-global libc state.  The reference draws std::rand() at
// src/Stereo3DMST.cpp:584 and random() at :75-77 from that shared state.
// ---------------------------------------------------------------------------------------------
struct GlibcRand {
    std::vector<uint32_t> hist;  // r[i] = r[i-31] + r[i-3]; output k is r[k+344] >> 1
    explicit GlibcRand(uint32_t seed = 1) { reseed(seed); }
    void reseed(uint32_t seed) {
        if (seed == 0) seed = 1;
        std::vector<int32_t> s(34);
        s[0] = (int32_t)seed;
        for (int i = 1; i < 31; i++) {
            int64_t hi = s[i - 1] / 127773, lo = s[i - 1] % 127773;
            int64_t word = 16807 * lo - 2836 * hi;
            if (word < 0) word += 2147483647;
            s[i] = (int32_t)word;
        }
        for (int i = 31; i < 34; i++) s[i] = s[i - 31];
        hist.assign(s.begin(), s.end());
        for (int i = 34; i < 344; i++) hist.push_back(hist[i - 31] + hist[i - 3]);
    }
    uint32_t next() {
        size_t i = hist.size();
        uint32_t v = hist[i - 31] + hist[i - 3];
        hist.push_back(v);
        if (hist.size() > 4096) hist.erase(hist.begin(), hist.begin() + 2048);
        return v >> 1;
    }
};

struct UnionFind {
    std::vector<int> p, sz;
    int num;
    explicit UnionFind(int n) : p(n), sz(n, 1), num(n) {
        for (int i = 0; i < n; i++) p[i] = i;
    }
    int find(int x) {
        while (p[x] != x) {
            p[x] = p[p[x]];
            x = p[x];
        }
        return x;
    }
    // include/disjoint-set.h:66-86 — only membership and sizes are observable to callers
    void join(int a, int b) {
        a = find(a);
        b = find(b);
        if (a == b) return;
        if (sz[a] > sz[b]) std::swap(a, b);
        p[a] = b;
        sz[b] += sz[a];
        num--;
    }
    int size(int x) { return sz[find(x)]; }
};

struct Edge {
    double w;
    int a, b;
    int eid;  // canonical id: 2*p (right neighbour), 2*p+1 (down neighbour)
};

// include/segment-graph.h:34-42
inline bool edge_less(const Edge& x, const Edge& y) {
    if (x.w != y.w) return x.w < y.w;
    if (x.a != y.a) return x.a < y.a;
    return x.b < y.b;
}

}  // namespace

struct OrcForest {
    int W = 0, H = 0, N = 0, T = 0;
    float gamma = 0.f;
    std::vector<uint16_t> ew;      // [2N] integer edge weights by canonical id, 0xFFFF where no edge
    std::vector<uint8_t> mask;     // [2N] 0 = not in forest, 1 = FH edge, 2 = min-size-merge edge
    std::vector<int> fh_comp;      // [N] FH component label (min pixel of the component) before the merge
    std::vector<int> tree_id;      // [N] pixel -> tree (first-seen raster order, Stereo3DMST.cpp:352-367)
    std::vector<int> tree_start;   // [T+1] node offsets, trees in id order
    std::vector<int> node_pixel;   // [N] node (tree-major, BFS order inside the tree) -> pixel
    std::vector<int> pixel_node;   // [N] inverse
    std::vector<int> parent;       // [N] global node index of the parent; root -> itself
    std::vector<int> child_begin;  // [N] global node index of first child (children are contiguous)
    std::vector<int> child_count;  // [N]
    std::vector<uint16_t> pw;      // [N] integer weight of the edge to the parent (root: 0)
    std::vector<int> level;        // [N] BFS depth
    std::vector<double> wlut, w2lut;  // exp(-iw*gamma), 1 - w*w for iw in 0..765
    std::vector<int> adj_ptr, adj;    // tree adjacency, CSR, ascending neighbour id (setS order)
};

extern "C" {

// -------------------------------------------------------------------------------------------
// a3: 3x3 median per channel (cv::medianBlur(ksize=3) on CV_8U replicates the border),
// Stereo3DMST.cpp:226-228.
// -------------------------------------------------------------------------------------------
void orc_median3(const uint8_t* in, int W, int H, uint8_t* out) {
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            uint8_t v[9];
            int k = 0;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    int yy = std::min(H - 1, std::max(0, y + dy));
                    int xx = std::min(W - 1, std::max(0, x + dx));
                    v[k++] = in[yy * W + xx];
                }
            std::nth_element(v, v + 4, v + 9);
            out[y * W + x] = v[4];
        }
}

// Edge weights on the median-filtered planes: |dR|+|dG|+|dB| (Stereo3DMST.cpp:83-91, :244-262).
// bgr is interleaved, 3 bytes per pixel, already median filtered.  ew[2p] = right, ew[2p+1] = down.
void orc_edge_weights(const uint8_t* bgr_med, int W, int H, uint16_t* ew) {
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int p = y * W + x;
            ew[2 * p] = ew[2 * p + 1] = 0xFFFF;
            if (x < W - 1) {
                int q = p + 1, s = 0;
                for (int c = 0; c < 3; c++) s += std::abs((int)bgr_med[3 * p + c] - (int)bgr_med[3 * q + c]);
                ew[2 * p] = (uint16_t)s;
            }
            if (y < H - 1) {
                int q = p + W, s = 0;
                for (int c = 0; c < 3; c++) s += std::abs((int)bgr_med[3 * p + c] - (int)bgr_med[3 * q + c]);
                ew[2 * p + 1] = (uint16_t)s;
            }
        }
}

// -------------------------------------------------------------------------------------------
// a3..a7: forest construction.  bgr = raw interleaved BGR u8 (median is applied here when
// do_median != 0, as the reference does at :226-228).
// -------------------------------------------------------------------------------------------
OrcForest* orc_forest_build(const uint8_t* bgr, int W, int H, float c, int min_size, float gamma, int do_median) {
    OrcForest* F = new OrcForest;
    const int N = W * H;
    F->W = W;
    F->H = H;
    F->N = N;
    F->gamma = gamma;

    std::vector<uint8_t> med(3 * (size_t)N);
    {
        std::vector<uint8_t> pl(N), pm(N);
        for (int ch = 0; ch < 3; ch++) {
            for (int i = 0; i < N; i++) pl[i] = bgr[3 * i + ch];
            if (do_median)
                orc_median3(pl.data(), W, H, pm.data());
            else
                pm = pl;
            for (int i = 0; i < N; i++) med[3 * i + ch] = pm[i];
        }
    }
    F->ew.resize(2 * (size_t)N);
    orc_edge_weights(med.data(), W, H, F->ew.data());

    // edge list in emission order (:244-262): right then down, raster
    std::vector<Edge> edges;
    edges.reserve(2 * (size_t)N);
    for (int p = 0; p < N; p++) {
        int x = p % W, y = p / W;
        if (x < W - 1) edges.push_back({(double)F->ew[2 * p], p, p + 1, 2 * p});
        if (y < H - 1) edges.push_back({(double)F->ew[2 * p + 1], p, p + W, 2 * p + 1});
    }
    const int num = (int)edges.size();

    // segment-graph.h:54-89
    std::sort(edges.begin(), edges.end(), edge_less);
    UnionFind u(N);
    std::vector<double> thr(N);
    for (int i = 0; i < N; i++) thr[i] = c / 1;  // THRESHOLD(1,c): float / int -> float -> double
    F->mask.assign(2 * (size_t)N, 0);
    for (int i = 0; i < num; i++) {
        const Edge& e = edges[i];
        int a = u.find(e.a), b = u.find(e.b);
        if (a != b) {
            if (e.w <= thr[a] && e.w <= thr[b]) {
                u.join(a, b);
                a = u.find(a);
                thr[a] = e.w + (c / u.size(a));  // Q2: float c / int size is a float division
                F->mask[e.eid] = 1;
            }
        }
    }
    F->fh_comp.assign(N, -1);
    {
        std::vector<int> minpix(N, -1);
        for (int i = 0; i < N; i++) {
            int r = u.find(i);
            if (minpix[r] < 0) minpix[r] = i;
            F->fh_comp[i] = minpix[r];
        }
    }

    // min-size merge, Stereo3DMST.cpp:293-307
    min_size = std::max(2, min_size);
    for (int i = 0; i < num; i++) {
        const Edge& e = edges[i];
        int a = u.find(e.a), b = u.find(e.b);
        if (a != b && (u.size(a) < min_size || u.size(b) < min_size)) {
            u.join(a, b);
            F->mask[e.eid] = 2;
        }
    }

    // tree ids, first-seen raster order (:352-367)
    const int T = u.num;
    F->T = T;
    F->tree_id.assign(N, -1);
    std::vector<std::vector<int>> tree_pixels(T);
    std::vector<int> id_in_tree(N);
    {
        std::vector<int> rep_to_id(N, -1);
        int next = 0;
        for (int i = 0; i < N; i++) {
            int r = u.find(i);
            if (rep_to_id[r] < 0) rep_to_id[r] = next++;
            int t = rep_to_id[r];
            F->tree_id[i] = t;
            id_in_tree[i] = (int)tree_pixels[t].size();
            tree_pixels[t].push_back(i);
        }
    }

    // tree adjacency graph (:377-384); setS => unique, ascending iteration
    {
        std::vector<std::vector<int>> nb(T);
        for (int i = 0; i < num; i++) {
            int ta = F->tree_id[edges[i].a], tb = F->tree_id[edges[i].b];
            if (ta != tb) {
                nb[ta].push_back(tb);
                nb[tb].push_back(ta);
            }
        }
        F->adj_ptr.assign(T + 1, 0);
        for (int t = 0; t < T; t++) {
            std::sort(nb[t].begin(), nb[t].end());
            nb[t].erase(std::unique(nb[t].begin(), nb[t].end()), nb[t].end());
            F->adj_ptr[t + 1] = F->adj_ptr[t] + (int)nb[t].size();
        }
        F->adj.reserve(F->adj_ptr[T]);
        for (int t = 0; t < T; t++) F->adj.insert(F->adj.end(), nb[t].begin(), nb[t].end());
    }

    // weights (:444, :513): exp(-w*gamma) with w double, gamma float promoted
    F->wlut.resize(766);
    F->w2lut.resize(766);
    for (int iw = 0; iw < 766; iw++) {
        double w = std::exp(-(double)iw * gamma);
        F->wlut[iw] = w;
        F->w2lut[iw] = 1.0f - w * w;
    }

    // per-tree adjacency in sorted-edge insertion order (:436-446)
    struct Nb {
        int v;        // local id of the neighbour
        uint16_t iw;  // integer edge weight
    };
    std::vector<std::vector<Nb>> adjl(N);  // indexed by pixel; neighbours as pixels
    for (int i = 0; i < num; i++) {
        const Edge& e = edges[i];
        if (F->mask[e.eid]) {
            adjl[e.a].push_back({e.b, (uint16_t)e.w});
            adjl[e.b].push_back({e.a, (uint16_t)e.w});
        }
    }

    // BFS re-index (:450-522)
    F->tree_start.assign(T + 1, 0);
    for (int t = 0; t < T; t++) F->tree_start[t + 1] = F->tree_start[t] + (int)tree_pixels[t].size();
    F->node_pixel.assign(N, -1);
    F->pixel_node.assign(N, -1);
    F->parent.assign(N, -1);
    F->child_begin.assign(N, 0);
    F->child_count.assign(N, 0);
    F->pw.assign(N, 0);
    F->level.assign(N, 0);
    for (int t = 0; t < T; t++) {
        const int base = F->tree_start[t];
        int next = base;
        int root_pix = tree_pixels[t][0];
        F->node_pixel[next] = root_pix;
        F->pixel_node[root_pix] = next;
        F->parent[next] = next;
        next++;
        for (int g = base; g < next; g++) {  // FIFO == increasing node id
            int pix = F->node_pixel[g];
            F->child_begin[g] = next;
            for (const Nb& nbh : adjl[pix]) {
                if (F->pixel_node[nbh.v] >= 0) continue;  // already coloured
                F->node_pixel[next] = nbh.v;
                F->pixel_node[nbh.v] = next;
                F->parent[next] = g;
                F->pw[next] = nbh.iw;
                F->level[next] = F->level[g] + 1;
                F->child_count[g]++;
                next++;
            }
        }
        // next == tree_start[t+1] iff the masked edges span the component (always true)
    }
    (void)id_in_tree;
    return F;
}

void orc_forest_free(OrcForest* F) { delete F; }
int orc_forest_num_trees(const OrcForest* F) { return F->T; }
int orc_forest_adj_size(const OrcForest* F) { return (int)F->adj.size(); }
int orc_forest_max_depth(const OrcForest* F) {
    int m = 0;
    for (int v : F->level) m = std::max(m, v);
    return m + 1;
}
// Copies out every array the parity tests compare; any pointer may be null.
void orc_forest_get(const OrcForest* F, uint16_t* ew, uint8_t* mask, int* fh_comp, int* tree_id, int* tree_start,
                    int* node_pixel, int* parent, int* child_begin, int* child_count, uint16_t* pw, int* level,
                    int* adj_ptr, int* adj) {
    const size_t N = F->N;
    if (ew) memcpy(ew, F->ew.data(), 2 * N * sizeof(uint16_t));
    if (mask) memcpy(mask, F->mask.data(), 2 * N);
    if (fh_comp) memcpy(fh_comp, F->fh_comp.data(), N * sizeof(int));
    if (tree_id) memcpy(tree_id, F->tree_id.data(), N * sizeof(int));
    if (tree_start) memcpy(tree_start, F->tree_start.data(), (F->T + 1) * sizeof(int));
    if (node_pixel) memcpy(node_pixel, F->node_pixel.data(), N * sizeof(int));
    if (parent) memcpy(parent, F->parent.data(), N * sizeof(int));
    if (child_begin) memcpy(child_begin, F->child_begin.data(), N * sizeof(int));
    if (child_count) memcpy(child_count, F->child_count.data(), N * sizeof(int));
    if (pw) memcpy(pw, F->pw.data(), N * sizeof(uint16_t));
    if (level) memcpy(level, F->level.data(), N * sizeof(int));
    if (adj_ptr) memcpy(adj_ptr, F->adj_ptr.data(), (F->T + 1) * sizeof(int));
    if (adj) memcpy(adj, F->adj.data(), F->adj.size() * sizeof(int));
}
void orc_forest_get_luts(const OrcForest* F, double* w, double* w2) {
    memcpy(w, F->wlut.data(), 766 * sizeof(double));
    memcpy(w2, F->w2lut.data(), 766 * sizeof(double));
}

// -------------------------------------------------------------------------------------------
// a6: random plane init (Stereo3DMST.cpp:390-430).  Uses the same libstdc++ objects.
// abc is [N][3] float.
// -------------------------------------------------------------------------------------------
void orc_plane_init(int W, int H, int max_disp, float* abc) {
    std::default_random_engine generator;
    std::uniform_real_distribution<float> distribution(0.0f, 1.0f);
    auto dice = std::bind(distribution, generator);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const int idx = y * W + x;
            const float d = dice() * max_disp;
            float x1, x2, nx, ny, nz;
            while (true) {
                x1 = dice();
                x2 = dice();
                if (x1 * x1 + x2 * x2 < 1.0f) break;
            }
            nx = 2.0f * x1 * std::sqrt(1.0f - x1 * x1 - x2 * x2);
            ny = 2.0f * x2 * std::sqrt(1.0f - x1 * x1 - x2 * x2);
            nz = std::sqrt(1.0f - nx * nx - ny * ny);
            abc[3 * idx + 0] = -nx / nz;
            abc[3 * idx + 1] = -ny / nz;
            abc[3 * idx + 2] = (nx * x + ny * y + nz * d) / nz;
        }
}

// -------------------------------------------------------------------------------------------
// a2: cost-volume ingest (Stereo3DMST.cpp:785-803): NaN -> cap, else min(cap, (c+offset)*scale).
// Reference: offset 0, scale 1, cap 0.5 ("accurate"); the commented "fast" variant is offset 1, scale 0.5.
// -------------------------------------------------------------------------------------------
void orc_ingest(float* vol, size_t n, float cap, float offset, float scale) {
    for (size_t i = 0; i < n; i++) {
        float v = vol[i];
        if (std::isnan(v))
            vol[i] = cap;
        else {
            if (offset != 0.0f || scale != 1.0f) v = (v + offset) * scale;
            vol[i] = std::min(cap, v);
        }
    }
}

// -------------------------------------------------------------------------------------------
// a2': dense truncated colour + gradient cost, from PatchMatchStereoGPU.cu:1482-1550
// (buildCostVolumeSharedMemoryBGR), d in [0, D).  Layout [d][y][x] fp32 for both views.
// DEFINED HERE (Q19): (i) every arithmetic step is a separately rounded fp32 op in the written
// order, except colour_l1*0.33333333333 which the reference evaluates in double; (ii) left-volume
// entries the reference kernel never writes (column x = W-1) are 3.0.
// -------------------------------------------------------------------------------------------
}  // extern "C"
static inline float gray_of(const uint8_t* p) {
    float b = (float)p[0], g = (float)p[1], r = (float)p[2];
    float t = 0.114f * b;
    float u = 0.587f * g;
    float v = 0.299f * r;
    float s = t + u;
    return s + v;
}
static inline float adgrad_pair(const uint8_t* ref0, const uint8_t* ref1, const uint8_t* mat0, const uint8_t* mat1) {
    float color_l1 = 0.0f;
    for (int i = 0; i < 3; i++) color_l1 += std::fabs((float)ref0[i] - (float)mat0[i]);
    float g = gray_of(mat0) - gray_of(ref0);
    float g2 = gray_of(ref1) - gray_of(mat1);
    g = g + g2;
    float cterm = std::fmin((float)((double)color_l1 * 0.33333333333), 7.0f);
    float gterm = std::fmin(std::fabs(g), 2.0f);
    float a = 0.11f * cterm;
    float b = 0.89f * gterm;
    return a + b;
}
extern "C" {
void orc_cost_adgrad(const uint8_t* left_bgr, const uint8_t* right_bgr, int W, int H, int D, float* left_vol,
                     float* right_vol) {
    const size_t N = (size_t)W * H;
    for (size_t i = 0; i < N * D; i++) left_vol[i] = 3.0f;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
            for (int d = 0; d < D; d++) {
                size_t idx = (size_t)d * N + (size_t)y * W + x;
                if (d + x + 1 < W) {
                    const uint8_t* ref0 = right_bgr + 3 * ((size_t)y * W + x);
                    const uint8_t* mat0 = left_bgr + 3 * ((size_t)y * W + x + d);
                    float cost = adgrad_pair(ref0, ref0 + 3, mat0, mat0 + 3);
                    right_vol[idx] = cost;
                    left_vol[idx + d] = cost;
                } else
                    right_vol[idx] = 3.0f;
                if (x - d < 0) left_vol[idx] = 3.0f;
            }
}

// the labels [d0, d1) of the same two volumes (rows 0 .. d1-d0-1 of the outputs; either output may be null): what one
// shard of the CPU arm of bench.py builds; every element is computed exactly as above
void orc_cost_adgrad_range(const uint8_t* left_bgr, const uint8_t* right_bgr, int W, int H, int d0, int d1, float* left_vol,
                           float* right_vol) {
    const size_t N = (size_t)W * H;
    for (int d = d0; d < d1; d++)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                const size_t p = (size_t)y * W + x, idx = (size_t)(d - d0) * N + p;
                if (right_vol) {
                    if (d + x + 1 < W) {  // right reference at x, left match at x + d
                        const uint8_t* ref0 = right_bgr + 3 * p;
                        const uint8_t* mat0 = left_bgr + 3 * (p + d);
                        right_vol[idx] = adgrad_pair(ref0, ref0 + 3, mat0, mat0 + 3);
                    } else
                        right_vol[idx] = 3.0f;
                }
                if (left_vol) {
                    if (x - d >= 0 && x + 1 < W) {  // the left entry at x was mirrored from the right reference at x - d
                        const uint8_t* ref0 = right_bgr + 3 * (p - d);
                        const uint8_t* mat0 = left_bgr + 3 * p;
                        left_vol[idx] = adgrad_pair(ref0, ref0 + 3, mat0, mat0 + 3);
                    } else
                        left_vol[idx] = 3.0f;
                }
            }
}

// -------------------------------------------------------------------------------------------
// a8: compute3DLabelCost (Stereo3DMST.cpp:103-118)
// -------------------------------------------------------------------------------------------
}  // extern "C"
static inline float label_cost(const float* vol, float a, float b, float c, int pixel, int max_disp, int W, size_t N) {
    const float disp = (pixel % W) * a + (pixel / W) * b + c;
    const float dc = std::ceil(disp), df = std::floor(disp);
    const int dci = (int)dc, dfi = (int)df;
    if (dci >= max_disp || dfi < 0) return 0.5f;
    return (dc - disp) * vol[(size_t)dfi * N + pixel] + (disp - df) * vol[(size_t)dci * N + pixel];
}

// Tree filter on one tree, given per-node costs already placed in agg[pixel] order.
// Up pass: Stereo3DMST.cpp:120-138; down pass :141-158.  `cost_of(pixel)` supplies the fp32 cost.
template <class CostFn>
static inline void tree_filter(const OrcForest* F, int t, double* agg, CostFn cost_of) {
    const int a = F->tree_start[t], b = F->tree_start[t + 1];
    for (int g = a; g < b; g++) agg[F->node_pixel[g]] = 0.0;  // :165
    for (int g = b - 1; g > a; g--) {
        const int pix = F->node_pixel[g];
        const int ppix = F->node_pixel[F->parent[g]];
        agg[pix] += cost_of(pix);
        agg[ppix] += F->wlut[F->pw[g]] * agg[pix];
    }
    agg[F->node_pixel[a]] += cost_of(F->node_pixel[a]);
    for (int g = a; g < b; g++) {
        const int pix = F->node_pixel[g];
        const int cb = F->child_begin[g], cc = F->child_count[g];
        for (int k = 0; k < cc; k++) {
            const int ch = cb + k, cpix = F->node_pixel[ch];
            agg[cpix] = F->wlut[F->pw[ch]] * agg[pix] + F->w2lut[F->pw[ch]] * agg[cpix];
        }
    }
}

extern "C" {
// a11: MSTCostAggregationAndLabelUpdate (Stereo3DMST.cpp:160-186) for one proposal on one tree.
// vol is [D][N] (already ingested).  Returns number of pixels updated.
int orc_eval_proposal(const OrcForest* F, const float* vol, int max_disp, int tree, float a, float b, float c,
                      double* min_cost, float* abc, double* agg) {
    const int W = F->W;
    const size_t N = F->N;
    tree_filter(F, tree, agg, [&](int pix) { return label_cost(vol, a, b, c, pix, max_disp, W, N); });
    int upd = 0;
    for (int g = F->tree_start[tree]; g < F->tree_start[tree + 1]; g++) {
        const int pix = F->node_pixel[g];
        if (agg[pix] < min_cost[pix]) {
            min_cost[pix] = agg[pix];
            abc[3 * pix + 0] = a;
            abc[3 * pix + 1] = b;
            abc[3 * pix + 2] = c;
            upd++;
        }
    }
    return upd;
}

// Apply a list of injected proposals in order: props is [n][4] float rows {tree_id, a, b, c}
// with tree_id stored as a float-encoded integer is lossy for big T, so tree ids come separately.
void orc_pms_apply(const OrcForest* F, const float* vol, int max_disp, const int* tree_ids, const float* labels, int n,
                   double* min_cost, float* abc) {
    std::vector<double> agg(F->N, 0.0);
    for (int i = 0; i < n; i++)
        orc_eval_proposal(F, vol, max_disp, tree_ids[i], labels[3 * i], labels[3 * i + 1], labels[3 * i + 2], min_cost,
                          abc, agg.data());
}

// Only the aggregated cost of one proposal, in pixel order over the tree's pixels (others untouched).
void orc_aggregate_label(const OrcForest* F, const float* vol, int max_disp, int tree, float a, float b, float c,
                         double* agg) {
    const int W = F->W;
    const size_t N = F->N;
    tree_filter(F, tree, agg, [&](int pix) { return label_cost(vol, a, b, c, pix, max_disp, W, N); });
}

// -------------------------------------------------------------------------------------------
// Slanted-plane matching cost straight from the images (north-star item 1; formula source pm::PatchMatch, src/pm.cpp):
// gradients pm.cpp:70-88 (cvtColor BGR2GRAY u8 with OpenCV's fixed-point weights, Sobel 3x3 CV_32F reflect-101, / 8),
// per-pixel cost pm.cpp:130-154 + dissimilarity :97-104 with the reference's byte-vector quirks (cv::Vec3b mcolo,
// cv::Vec2b mgrad: every scaled term and the sums are saturate_cast<uchar>).  Checked against cv2 for the gradients
// (tests/test_oracle_golden.py).  Floating-point order: separately rounded fp32 operations, left to right.
// -------------------------------------------------------------------------------------------
static inline int pm_reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}
static inline int pm_sat_u8(float v) {
    const long i = lrintf(v);  // cvRound: nearest, ties to even
    return i < 0 ? 0 : (i > 255 ? 255 : (int)i);
}
void orc_pm_gradients(const uint8_t* bgr, int W, int H, float* grad) {
    std::vector<int> gray((size_t)W * H);
    for (size_t p = 0; p < (size_t)W * H; p++) gray[p] = (bgr[3 * p] * 3735 + bgr[3 * p + 1] * 19235 + bgr[3 * p + 2] * 9798 + (1 << 14)) >> 15;  // cv2 4.x (3.4.3: 1868/9617/4899 >> 14)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const int xm = pm_reflect101(x - 1, W), xp = pm_reflect101(x + 1, W), ym = pm_reflect101(y - 1, H), yp = pm_reflect101(y + 1, H);
            auto g = [&](int yy, int xx) { return gray[(size_t)yy * W + xx]; };
            const int gx = (g(ym, xp) + 2 * g(y, xp) + g(yp, xp)) - (g(ym, xm) + 2 * g(y, xm) + g(yp, xm));
            const int gy = (g(yp, xm) + 2 * g(yp, x) + g(yp, xp)) - (g(ym, xm) + 2 * g(ym, x) + g(ym, xp));
            grad[2 * ((size_t)y * W + x)] = (float)gx / 8.f;
            grad[2 * ((size_t)y * W + x) + 1] = (float)gy / 8.f;
        }
}
static inline float pm_plane_cost(const uint8_t* simg, const float* sgrad, const uint8_t* oimg, const float* ograd, int pix, int W, int view, float a,
                                  float b, float c, int max_disp, float alpha, float tau_c, float tau_g, float scale, float oob) {
    const int x = pix % W, y = pix / W;
    const float d = a * x + b * y + c;                      // pm.h:153-156
    if (!(d >= 0.0f) || d > (float)max_disp || W < 2) return oob;    // :132-135 (PLANE_PENALTY)
    const float match = view ? x + d : x - d;               // :139 sign = -1 + 2*cpv
    int xm = (int)match;                                    // :140
    const float wm = 1.0f - (match - xm);                   // :142
    if (xm > W - 2) xm = W - 2;                             // :144-147
    if (xm < 0) xm = 0;
    const uint8_t* o0 = oimg + 3 * ((size_t)y * W + xm);
    const float* g0 = ograd + 2 * ((size_t)y * W + xm);
    float cc = 0.0f, cg = 0.0f;
    for (int k = 0; k < 3; k++) {                           // :150 Vec3b = wm * p0 + (1 - wm) * p1 in byte vectors
        int m = pm_sat_u8(wm * o0[k]) + pm_sat_u8((1.0f - wm) * o0[3 + k]);
        if (m > 255) m = 255;
        cc += std::fabs((float)((int)simg[3 * (size_t)pix + k] - m));
    }
    for (int k = 0; k < 2; k++) {                           // :151 Vec2b <- Vec2f average
        const float gm = (float)pm_sat_u8(wm * g0[k] + (1.0f - wm) * g0[2 + k]);
        cg += std::fabs(sgrad[2 * (size_t)pix + k] - gm);
    }
    cc = std::min(cc, tau_c);                               // :99-103
    cg = std::min(cg, tau_g);
    return ((1.0f - alpha) * cc + alpha * cg) * scale;
}
// the plane cost of one label at every pixel of a view (parity dumps)
void orc_pm_plane_cost_map(int view, const uint8_t* left_bgr, const uint8_t* right_bgr, const float* left_grad, const float* right_grad, int W, int H,
                           float a, float b, float c, int max_disp, float alpha, float tau_c, float tau_g, float scale, float oob, float* out) {
    const uint8_t* simg = view ? right_bgr : left_bgr; const uint8_t* oimg = view ? left_bgr : right_bgr;
    const float* sgrad = view ? right_grad : left_grad; const float* ograd = view ? left_grad : right_grad;
    for (int p = 0; p < W * H; p++) out[p] = pm_plane_cost(simg, sgrad, oimg, ograd, p, W, view, a, b, c, max_disp, alpha, tau_c, tau_g, scale, oob);
}
// injected proposals evaluated with the plane cost (MSTCostAggregationAndLabelUpdate with the data term above)
void orc_pms_apply_plane(const OrcForest* F, int view, const uint8_t* left_bgr, const uint8_t* right_bgr, const float* left_grad, const float* right_grad,
                         int max_disp, float alpha, float tau_c, float tau_g, float scale, float oob, const int* tree_ids, const float* labels, int n,
                         double* min_cost, float* abc) {
    std::vector<double> agg(F->N, 0.0);
    const uint8_t* simg = view ? right_bgr : left_bgr; const uint8_t* oimg = view ? left_bgr : right_bgr;
    const float* sgrad = view ? right_grad : left_grad; const float* ograd = view ? left_grad : right_grad;
    for (int i = 0; i < n; i++) {
        const float a = labels[3 * i], b = labels[3 * i + 1], c = labels[3 * i + 2];
        const int tree = tree_ids[i];
        tree_filter(F, tree, agg.data(), [&](int pix) { return pm_plane_cost(simg, sgrad, oimg, ograd, pix, F->W, view, a, b, c, max_disp, alpha, tau_c, tau_g, scale, oob); });
        for (int g = F->tree_start[tree]; g < F->tree_start[tree + 1]; g++) {
            const int pix = F->node_pixel[g];
            if (agg[pix] < min_cost[pix]) {
                min_cost[pix] = agg[pix];
                abc[3 * pix] = a; abc[3 * pix + 1] = b; abc[3 * pix + 2] = c;
            }
        }
    }
}

// -------------------------------------------------------------------------------------------
// A13 dense-label mode: for d ascending, cost = C[d][p] (selectDisparity, PatchMatchStereoGPU.cu:1706-1717),
// tree filter with the reference arithmetic, strict '<' so the lowest d wins ties.
// agg_out (optional) is [D][N] double; disp [N] int32; best [N] double.  Labels d0..d1-1 only.
// -------------------------------------------------------------------------------------------
void orc_aggregate_dense(const OrcForest* F, const float* vol, int D, int d0, int d1, double* agg_out, int32_t* disp,
                         double* best) {
    const size_t N = F->N;
    std::vector<double> agg(N, 0.0);
    for (size_t i = 0; i < N; i++) {
        best[i] = DBL_MAX;
        disp[i] = -1;
    }
    (void)D;
    for (int d = d0; d < d1; d++) {
        const float* slice = vol + (size_t)d * N;
        for (int t = 0; t < F->T; t++) tree_filter(F, t, agg.data(), [&](int pix) { return slice[pix]; });
        if (agg_out) memcpy(agg_out + (size_t)d * N, agg.data(), N * sizeof(double));
        for (size_t i = 0; i < N; i++)
            if (agg[i] < best[i]) {
                best[i] = agg[i];
                disp[i] = d;
            }
    }
}

// -------------------------------------------------------------------------------------------
// a12: one MST_PMS call (Stereo3DMST.cpp:546-629).  Records every tested proposal when rec_* are
// non-null (capacity rec_cap); returns the number of proposals tested.  `grand` carries the
// glibc rand() stream across calls (opaque handle from orc_rand_new).
// -------------------------------------------------------------------------------------------
void* orc_rand_new(unsigned seed) { return new GlibcRand(seed); }
void orc_rand_free(void* g) { delete (GlibcRand*)g; }
void orc_rand_burn(void* g, size_t n) {
    GlibcRand* r = (GlibcRand*)g;
    for (size_t i = 0; i < n; i++) r->next();
}
unsigned orc_rand_next(void* g) { return ((GlibcRand*)g)->next(); }

int orc_mst_pms(const OrcForest* F, const float* vol, int max_disp, double* min_cost, float* abc, void* grand,
                int* rec_tree, float* rec_label, int rec_cap) {
    const int W = F->W;
    GlibcRand* gr = (GlibcRand*)grand;
    std::default_random_engine generator;  // Q8: bind copies => every call replays the same stream
    std::uniform_real_distribution<float> distribution(-1.0f, 1.0f);
    auto dice = std::bind(distribution, generator);
    std::vector<double> agg(F->N, 0.0);
    int nrec = 0;
    auto test = [&](int tree, float a, float b, float c) {
        if (rec_tree && nrec < rec_cap) {
            rec_tree[nrec] = tree;
            rec_label[3 * nrec] = a;
            rec_label[3 * nrec + 1] = b;
            rec_label[3 * nrec + 2] = c;
        }
        nrec++;
        orc_eval_proposal(F, vol, max_disp, tree, a, b, c, min_cost, abc, agg.data());
    };
    for (int t = 0; t < F->T; t++) {
        // spatial propagation (:563-580)
        for (int k = F->adj_ptr[t]; k < F->adj_ptr[t + 1]; k++) {
            const int nb = F->adj[k];
            const int nsz = F->tree_start[nb + 1] - F->tree_start[nb];
            int idx = (int)((dice() + 1.0f) * 0.5f * nsz);
            if (idx >= nsz) idx = nsz - 1;  // Q11
            const int pix = F->node_pixel[F->tree_start[nb] + idx];
            test(t, abc[3 * pix], abc[3 * pix + 1], abc[3 * pix + 2]);
        }
        // random refinement (:584-625)
        const int sz = F->tree_start[t + 1] - F->tree_start[t];
        const int pix = F->node_pixel[F->tree_start[t] + (int)(gr->next() % (unsigned)sz)];
        const float px = (float)(pix % W), py = (float)(pix / W);
        const float la = abc[3 * pix], lb = abc[3 * pix + 1], lc = abc[3 * pix + 2];
        const float nz = 1.0f / std::sqrt(la * la + lb * lb + 1.0f);
        const float nx = -la * nz, ny = -lb * nz;
        const float d = px * la + py * lb + lc;
        float max_n = 1.0f, max_d = 0.5f * max_disp;
        for (; max_d > 0.1f; max_d *= 0.5f, max_n *= 0.5f) {
            float rand_d = d + dice() * max_d;
            if (rand_d < 0.0f || rand_d > (float)max_disp) continue;
            float rnx = nx + dice() * max_n;
            float rny = ny + dice() * max_n;
            float rnz = nz + dice() * max_n;
            const float norm_inv = 1.0f / std::sqrt(rnx * rnx + rny * rny + rnz * rnz);
            rnx *= norm_inv;
            rny *= norm_inv;
            rnz = std::fabs(rnz * norm_inv);
            test(t, -rnx / rnz, -rny / rnz, (rnx * px + rny * py + rnz * rand_d) / rnz);
        }
    }
    return nrec;
}

// a13: LabelToDisp (:189-201), normalised to [0,1]
void orc_label_to_disp(const float* abc, int W, int H, int max_disp, float* disp) {
    for (int i = 0; i < W * H; i++) {
        float v = ((i % W) * abc[3 * i] + (i / W) * abc[3 * i + 1] + abc[3 * i + 2]) / (max_disp - 1.0f);
        v = std::min(1.0f, v);
        disp[i] = (0.0f < v) ? v : 0.0f;  // MAX(0.0f, v) with OpenCV's MAX(a,b) ((a) < (b) ? (b) : (a))
    }
}

// a14: leftRightConsistencyCheck (:632-710).  mask_out optional [N] u8 (1 = invalid after pass 1).
void orc_lr_check(float* left, const float* right, int W, int H, int max_disp, int fill, uint8_t* mask_out) {
    std::vector<uint8_t> mask((size_t)W * H, 0);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const int idx = y * W + x;
            const float d_f = left[idx];
            int d = (int)std::round(d_f);
            if (x - d >= 0 && d >= 0 && d < max_disp) {
                if (std::fabs(d_f - right[idx - d]) > 1.0f) {
                    mask[idx] = 1;
                    left[idx] = 0.0f;
                }
            } else {
                mask[idx] = 1;
                left[idx] = 0.0f;
            }
        }
    if (mask_out) memcpy(mask_out, mask.data(), mask.size());
    if (!fill) return;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const int idx = y * W + x;
            if (mask[idx] == 0) continue;
            int i = 1;
            while (true) {
                if (x - i < 0) break;
                if (mask[idx - i] == 0) {
                    left[idx] = left[idx - i];
                    mask[idx] = 0;
                    break;
                }
                i++;
            }
            i = 1;
            while (true) {
                if (x + i >= W) break;
                if (mask[idx + i] == 0) {
                    if (left[idx + i] < left[idx] || mask[idx] == 1) left[idx] = left[idx + i];
                    break;
                }
                i++;
            }
        }
}

// -------------------------------------------------------------------------------------------
// a1: the whole stereo3dmst() (:714-912) minus the mc-cnn subprocess: volumes are inputs
// ([Dmax][N] fp32, raw; ingested here as :785-803 does).  emulate_gui_rand != 0 burns the
// 3*N random() draws per view that the debug colouring makes (Q6) so that the std::rand()
// stream position matches the reference binary.  Proposals can be recorded per view.
// -------------------------------------------------------------------------------------------
void orc_stereo3dmst(const uint8_t* left_bgr, const uint8_t* right_bgr, int W, int H, const float* left_vol_raw,
                     const float* right_vol_raw, int Dmax, int num_iter, int emulate_gui_rand, float* left_disp,
                     float* right_disp, float* left_abc_out, float* right_abc_out, double* left_min_cost_out,
                     double* right_min_cost_out) {
    const size_t N = (size_t)W * H;
    std::vector<float> lv(left_vol_raw, left_vol_raw + N * Dmax), rv(right_vol_raw, right_vol_raw + N * Dmax);
    orc_ingest(lv.data(), N * Dmax, 0.5f, 0.0f, 1.0f);
    orc_ingest(rv.data(), N * Dmax, 0.5f, 0.0f, 1.0f);
    const float gamma = 1.0f / 12.f, c = 5000.0f;
    const int min_cc = 200;
    GlibcRand* gr = new GlibcRand(1);
    std::vector<float> labc(3 * N), rabc(3 * N);
    OrcForest* FL = orc_forest_build(left_bgr, W, H, c, min_cc, gamma, 1);
    if (emulate_gui_rand) orc_rand_burn(gr, 3 * N);
    orc_plane_init(W, H, Dmax, labc.data());
    OrcForest* FR = orc_forest_build(right_bgr, W, H, c, min_cc, gamma, 1);
    if (emulate_gui_rand) orc_rand_burn(gr, 3 * N);
    orc_plane_init(W, H, Dmax, rabc.data());
    std::vector<double> lmin(N, DBL_MAX), rmin(N, DBL_MAX);
    for (int i = 0; i < num_iter; i++) orc_mst_pms(FL, lv.data(), Dmax, lmin.data(), labc.data(), gr, nullptr, nullptr, 0);
    for (int i = 0; i < num_iter; i++) orc_mst_pms(FR, rv.data(), Dmax, rmin.data(), rabc.data(), gr, nullptr, nullptr, 0);
    orc_label_to_disp(labc.data(), W, H, Dmax, left_disp);
    orc_label_to_disp(rabc.data(), W, H, Dmax, right_disp);
    for (size_t i = 0; i < N; i++) {
        left_disp[i] = left_disp[i] * (Dmax - 1.f);  // :900-902
        right_disp[i] = right_disp[i] * (Dmax - 1.f);
    }
    orc_lr_check(left_disp, right_disp, W, H, Dmax, 0, nullptr);  // :904
    if (left_abc_out) memcpy(left_abc_out, labc.data(), 3 * N * sizeof(float));
    if (right_abc_out) memcpy(right_abc_out, rabc.data(), 3 * N * sizeof(float));
    if (left_min_cost_out) memcpy(left_min_cost_out, lmin.data(), N * sizeof(double));
    if (right_min_cost_out) memcpy(right_min_cost_out, rmin.data(), N * sizeof(double));
    orc_forest_free(FL);
    orc_forest_free(FR);
    delete gr;
}

}  // extern "C"
