// oracle/ref_driver.cpp — TEST INFRASTRUCTURE ONLY.
//
// C-ABI handles around the reference's OWN code, compiled from where it lies:
//   #include REF_TU   ==  /root/reference/src/Stereo3DMST.cpp   (unmodified; includes
//   /root/reference/include/{Stereo3DMST.h,segment-graph.h,disjoint-set.h})
// built against the container shims in oracle/ref_shims/ (the image has no OpenCV/Boost C++
// headers).  Output goes to oracle/_ref/libref3dmst.so only (git-ignored).  Used by
// tests/test_oracle_vs_ref.py and tests/golden/make_golden.py to pin oracle/s3dmst_oracle.cpp.
// No reference source is copied into this repo.

#ifndef REF_TU
#error "build with -DREF_TU='\"/root/reference/src/Stereo3DMST.cpp\"'"
#endif
#include REF_TU

#include <cstdio>

struct RefView {
    int W, H, Dmax;
    std::vector<mst_graph_t> mst_vec;
    std::vector<std::vector<int>> mst_vertices_vec;
    tree_graph_t tree_g;
    std::vector<abc> abc_map;
};

extern "C" {

// include/segment-graph.h:54-89 on a caller-supplied edge list (arrays are sorted in place,
// as the reference does).  mask is indexed by sorted position; comp[i] = find(i).
int ref_segment_graph(int num_vertices, int num_edges, double* w, int* a, int* b, int* mask, float c, int* comp,
                      int* comp_size) {
    std::vector<edge> ed(num_edges);
    for (int i = 0; i < num_edges; i++) {
        ed[i].w = w[i];
        ed[i].a = a[i];
        ed[i].b = b[i];
    }
    memset(mask, 0, sizeof(int) * num_edges);
    universe* u = segment_graph(num_vertices, num_edges, ed.data(), mask, c);
    for (int i = 0; i < num_edges; i++) {
        w[i] = ed[i].w;
        a[i] = ed[i].a;
        b[i] = ed[i].b;
    }
    for (int i = 0; i < num_vertices; i++) {
        comp[i] = u->find(i);
        if (comp_size) comp_size[i] = u->size(comp[i]);
    }
    int n = u->num_sets();
    delete u;
    return n;
}

// src/Stereo3DMST.cpp:213-543 (segment_image_other_init) on an interleaved BGR u8 image.
RefView* ref_view_build(const uint8_t* bgr, int W, int H, float c, int min_size, int max_disp, float gamma) {
    RefView* v = new RefView;
    v->W = W;
    v->H = H;
    v->Dmax = max_disp;
    cv::Mat img(H, W, CV_8UC3, (void*)bgr);
    std::vector<cv::Mat> ch;
    cv::split(img, ch);
    v->abc_map.resize((size_t)W * H);
    segment_image_other_init(ch[2], ch[1], ch[0], v->mst_vec, v->mst_vertices_vec, v->tree_g, v->abc_map.data(), c,
                             min_size, max_disp, gamma);
    return v;
}
void ref_view_free(RefView* v) { delete v; }
int ref_view_num_trees(RefView* v) { return (int)v->mst_vec.size(); }
int ref_view_adj_size(RefView* v) {
    int n = 0;
    for (size_t t = 0; t < v->mst_vec.size(); t++) {
        auto pr = boost::adjacent_vertices((int)t, v->tree_g);
        for (auto it = pr.first; it != pr.second; ++it) n++;
    }
    return n;
}
// Flattened dump, tree-major / BFS order: node_pixel, parent (global node), child_begin (global
// node of children_indices[0], or -1), child_count, weight, weight2; tree_start [T+1];
// adjacency CSR; abc [N][3].
void ref_view_get(RefView* v, int* tree_start, int* node_pixel, int* parent, int* child_begin, int* child_count,
                  double* weight, double* weight2, int* children4, int* adj_ptr, int* adj, float* abc_out) {
    int T = (int)v->mst_vec.size();
    int off = 0;
    int k = 0;
    for (int t = 0; t < T; t++) {
        mst_graph_t& g = v->mst_vec[t];
        int n = (int)boost::num_vertices(g);
        tree_start[t] = off;
        for (int i = 0; i < n; i++) {
            node_pixel[off + i] = v->mst_vertices_vec[t][i];
            parent[off + i] = off + g[i].parent_idx;
            child_count[off + i] = g[i].num_children;
            child_begin[off + i] = g[i].num_children > 0 ? off + g[i].children_indices[0] : -1;
            for (int j = 0; j < 4; j++) children4[4 * (off + i) + j] = j < g[i].num_children ? off + g[i].children_indices[j] : -1;
            weight[off + i] = i == 0 ? 0.0 : g[i].weight;
            weight2[off + i] = i == 0 ? 0.0 : g[i].weight2;
        }
        off += n;
        adj_ptr[t] = k;
        auto pr = boost::adjacent_vertices(t, v->tree_g);
        for (auto it = pr.first; it != pr.second; ++it) adj[k++] = *it;
    }
    tree_start[T] = off;
    adj_ptr[T] = k;
    if (abc_out) memcpy(abc_out, v->abc_map.data(), sizeof(abc) * v->abc_map.size());
}
void ref_view_set_abc(RefView* v, const float* abc_in) { memcpy(v->abc_map.data(), abc_in, sizeof(abc) * v->abc_map.size()); }

// src/Stereo3DMST.cpp:160-186 for one label on one tree.  vol = [Dmax][N] fp32.
void ref_eval_proposal(RefView* v, float* vol, int tree, float a, float b, float c, double* min_cost, double* agg) {
    abc lab;
    lab.a = a;
    lab.b = b;
    lab.c = c;
    MSTCostAggregationAndLabelUpdate(min_cost, agg, v->mst_vec[tree], v->abc_map.data(), lab, v->mst_vertices_vec[tree],
                                     vol, nullptr, v->Dmax, v->W, v->H, v->W * v->H);
}

// src/Stereo3DMST.cpp:546-629, called the way stereo3dmst() calls it (:851-852, :862-864).
void ref_mst_pms(RefView* v, float* vol, double* min_cost, double* agg) {
    std::default_random_engine generator;
    std::uniform_real_distribution<float> distribution(-1.0f, 1.0f);
    MST_PMS(v->mst_vec, v->mst_vertices_vec, v->tree_g, v->abc_map.data(), min_cost, agg, vol, nullptr, v->Dmax, v->W,
            v->H, v->W * v->H, generator, distribution);
}

// :189-201
void ref_label_to_disp(RefView* v, float* disp) {
    cv::Mat d(v->H, v->W, CV_32F, disp);
    LabelToDisp(v->abc_map.data(), v->mst_vec, v->mst_vertices_vec, d, v->H, v->W, v->Dmax);
}

// :632-710
void ref_lr_check(float* left, float* right, int W, int H, int max_disp, int fill) {
    leftRightConsistencyCheck(left, right, W, H, max_disp, fill != 0);
}

// :714-912, the real entry point.  The caller must have prepared <workdir>/mc-cnn-master/{left,right}.bin
// ([1][Dmax][H][W] fp32) — the reference's system("./main.lua ...") then fails harmlessly (or runs a
// no-op script) and the function mmaps those files exactly as it would mc-cnn's output.
int ref_stereo3dmst(const char* workdir, const uint8_t* left_bgr, const uint8_t* right_bgr, int W, int H, int Dmax,
                    float* left_disp, float* right_disp) {
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd)) return -1;
    if (chdir(workdir) < 0) return -2;
    cv::Mat L(H, W, CV_8UC3, (void*)left_bgr), R(H, W, CV_8UC3, (void*)right_bgr);
    cv::Mat dl, dr;
    stereo3dmst("img1r.png", "img2r.png", L, R, dl, dr, "MCCNN_acrt", Dmax);
    int rc = chdir(cwd);
    memcpy(left_disp, dl.data, sizeof(float) * W * H);
    memcpy(right_disp, dr.data, sizeof(float) * W * H);
    return rc;
}

}  // extern "C"
