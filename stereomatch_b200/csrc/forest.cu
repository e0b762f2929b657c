// stereomatch_b200/csrc/forest.cu — forest construction on the GPU (north-star items 2 and 3).
//
// Reference semantics (sequential): sort the 4-connected grid edges by (w, a, b), run the
// Felzenszwalb-Huttenlocher union rule  w <= thr[a] && w <= thr[b], thr = w + c/size
// (include/segment-graph.h:54-89), then merge every component smaller than max(2, min_size) along the
// same sorted order (src/Stereo3DMST.cpp:293-307), number the trees in first-seen raster order
// (:352-367) and re-index every tree in BFS order from its minimum pixel (:450-522).
//
// Parallel formulation (validated edge-for-edge against the sequential code by
// tests/models/forest_model.py and tests/test_forest_model.py):
//  * FH decomposes exactly by integer weight level.  A component closed at level w (w > thr) stays
//    closed for ever; a component that merges at level w is open for the rest of the level
//    (thr = w + c/size >= w).  Hence inside one level the accepted edges are the minimum spanning
//    forest, under the key "edge id" (== the reference's (a,b) tie-break: right edge 2p before down
//    edge 2p+1, ascending p), of the level's edges between open components: Boruvka rounds with
//    atomicMin picks.  thr is never stored: thr(r) = lastw[r] + f32(c)/f32(size[r]).
//  * The min-size merge is order dependent; it is replayed with deterministic reservations: every
//    pending edge atomicMin's its key (w<<32 | id) onto both endpoint components; an edge commits
//    when each endpoint component is either already >= m (its size can no longer matter) or holds
//    this edge as its reservation (no earlier pending edge touches it).  Edges between two big
//    components can never fire and are dropped.
//  Both phases run in ONE persistent cooperative kernel (grid = all co-resident CTAs, grid.sync()
//  between phases of a round); a round is: reserve | commit+hook | size/cleanup.
//  * BFS: one CTA per tree, level-synchronous inside the CTA; children of a node are its forest
//    neighbours except the parent, ordered by (w, edge id) (= the reference's adjacency insertion
//    order), numbered by a block-wide exclusive scan so BFS numbering equals the reference's queue order.
#include <cooperative_groups.h>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "internal.h"

namespace cg = cooperative_groups;

#define CNT_ERR (S3_MAX_ROUNDS - 1)
#define CNT_LIST (S3_MAX_ROUNDS - 2)
#define CNT_ROUNDS (S3_MAX_ROUNDS - 3)
#define CNT_FIRSTS (S3_MAX_ROUNDS - 4)
#define ROUND_CAP (S3_MAX_ROUNDS - 32)

struct FHArgs {
    int W, N;
    float c;
    int m;
    const uint16_t* ew;
    uint32_t* elist;
    const int* lvl_off;
    int* parent;
    int* size;
    int* lastw;
    uint32_t* best;
    unsigned long long* resv;
    uint8_t* mask;
    int* e_ra;
    int* e_rb;
    uint8_t* e_flag;
    int* counters;
};

__device__ __forceinline__ int uf_find(int* parent, int x) {
    while (true) {
        const int p = __ldcg(parent + x);
        if (p == x) return x;
        const int gp = __ldcg(parent + p);
        if (gp == p) return p;
        __stcg(parent + x, gp);  // path halving; racing writers only ever store ancestors
        x = gp;
    }
}

// segment-graph.h:27,80 — THRESHOLD(size,c) is a float division (Q2), added to a double w
__device__ __forceinline__ bool uf_open(const FHArgs& A, int r, int w) {
    const double thr = (double)__ldcg(A.lastw + r) + (double)__fdiv_rn(A.c, (float)__ldcg(A.size + r));
    return (double)w <= thr;
}

__device__ __forceinline__ int block_sum_to_counter(int v, int* counter) {
    // warp reduce then one atomic per warp
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(counter, v);
    return v;
}

struct FHArgs2 {
    FHArgs v[2];
};

// One thread-block CLUSTER per view (16 CTAs x 1024 threads on one GPC): the phases of a Boruvka / reservation
// round are separated by cluster barriers (~0.3 us) instead of grid-wide barriers (~3-4 us), and the ~1500
// barriers of a forest build are what bounds this kernel.  Clusters (views) never synchronise with each other.
__global__ void __launch_bounds__(1024) k_fh_merge(FHArgs2 AA) {
    cg::cluster_group grid = cg::this_cluster();
    const FHArgs& A = AA.v[blockIdx.x / grid.num_blocks()];
    const int gtid = grid.block_rank() * blockDim.x + threadIdx.x;
    const int gstride = grid.num_blocks() * blockDim.x;
    const int cbase = grid.block_rank() * blockDim.x;  // this CTA's first thread in the cluster-wide numbering
    int round = 0;
    long long tA = 0, tB = 0, tC = 0, tS = 0, t0 = clock64(), t1;
    int nlev = 0;
#define FH_T(acc) do { t1 = clock64(); acc += t1 - t0; t0 = t1; } while (0)

    // cluster-wide sum of a per-thread count through distributed shared memory (one cluster barrier, no
    // global-memory round trip): every CTA deposits its partial sum in every CTA's slot array
    __shared__ int s_part[2];
    __shared__ int s_cnt[2][16];
    const int nb = (int)grid.num_blocks(), myrank = (int)grid.block_rank();
    int sync_no = 0;
    auto cluster_sum = [&](int v) -> int {
        const int par = sync_no & 1;
        sync_no++;
        if (threadIdx.x == 0) s_part[par] = 0;
        __syncthreads();
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_part[par], v);
        __syncthreads();
        if ((int)threadIdx.x < nb) grid.map_shared_rank(&s_cnt[par][0], threadIdx.x)[myrank] = s_part[par];
        grid.sync();
        int tot = 0;
        for (int r = 0; r < nb; r++) tot += s_cnt[par][r];
        return tot;
    };
    // the three phases of a Boruvka round on one edge (e becomes S3_DEAD when the edge can never join)
    auto phase_a = [&](uint32_t& e, int& ra, int& rb, int w) -> int {
        const int a = (int)(e >> 1), b = a + ((e & 1u) ? A.W : 1);
        ra = uf_find(A.parent, a);
        rb = uf_find(A.parent, b);
        if (ra == rb || !uf_open(A, ra, w) || !uf_open(A, rb, w)) {
            e = S3_DEAD;  // same component / closed component: permanent
            return 0;
        }
        atomicMin(A.best + ra, e);
        atomicMin(A.best + rb, e);
        return 1;
    };
    auto phase_b = [&](uint32_t e, int ra, int rb, uint8_t& fl) -> int {  // returns 1 if the edge stays pending
        const bool pa = __ldcg(A.best + ra) == e, pb = __ldcg(A.best + rb) == e;
        fl = 0;
        if (!(pa || pb)) return 1;
        A.mask[e] = 1;
        const bool mutual = pa && pb;
        if (pa && !(mutual && ra < rb)) { __stcg(A.parent + ra, rb); fl |= 1; }
        if (pb && !(mutual && rb < ra)) { __stcg(A.parent + rb, ra); fl |= 2; }
        fl |= 4;
        return 0;
    };
    auto phase_c = [&](int ra, int rb, uint8_t fl, int w) {
        __stcg(A.best + ra, S3_DEAD);
        __stcg(A.best + rb, S3_DEAD);
        if (fl & 1) {
            const int R = uf_find(A.parent, ra);
            atomicAdd(A.size + R, __ldcg(A.size + ra));
            __stcg(A.lastw + R, w);
        }
        if (fl & 2) {
            const int R = uf_find(A.parent, rb);
            atomicAdd(A.size + R, __ldcg(A.size + rb));
            __stcg(A.lastw + R, w);
        }
    };

    // ------------------------------------------------------------------ FH, level by level
    for (int w = 0; w < S3_NUM_W; ++w) {
        const int lo = A.lvl_off[w], hi = A.lvl_off[w + 1];
        if (lo == hi) continue;
        nlev++;
        // a level with at most one edge per thread (the common case) keeps its edge state in registers
        const bool single = hi - lo <= gstride;
        uint32_t e1 = S3_DEAD;
        int ra1 = 0, rb1 = 0;
        uint8_t fl1 = 0;
        if (single && lo + gtid < hi) e1 = A.elist[lo + gtid];
        while (true) {
            round++;
            // phase A: each live edge picks itself as the minimum edge of both endpoint components
            int live = 0;
            if (single) {
                if (e1 != S3_DEAD) live = phase_a(e1, ra1, rb1, w);
            } else {
                for (int pos = lo + gtid; pos < hi; pos += gstride) {
                    uint32_t e = A.elist[pos];
                    if (e == S3_DEAD) continue;
                    int ra, rb;
                    if (phase_a(e, ra, rb, w)) {
                        A.e_ra[pos] = ra;
                        A.e_rb[pos] = rb;
                        live++;
                    } else
                        A.elist[pos] = S3_DEAD;
                }
            }
            FH_T(tA);
            const int nlive = cluster_sum(live);
            FH_T(tS);
            if (nlive == 0) break;
            // phase B: picked edges join the forest; the picking component hooks under the other one
            int remain = 0;
            if (single) {
                if (e1 != S3_DEAD) remain = phase_b(e1, ra1, rb1, fl1);
            } else {
                for (int pos = lo + gtid; pos < hi; pos += gstride) {
                    const uint32_t e = A.elist[pos];
                    if (e == S3_DEAD) continue;
                    uint8_t fl;
                    remain += phase_b(e, A.e_ra[pos], A.e_rb[pos], fl);
                    A.e_flag[pos] = fl;
                }
            }
            FH_T(tB);
            const int nrem = cluster_sum(remain);
            FH_T(tS);
            // phase C: sizes flow to the new roots, picks are cleared
            if (single) {
                if (e1 != S3_DEAD) {
                    phase_c(ra1, rb1, fl1, w);
                    if (fl1 & 4) e1 = S3_DEAD;
                }
            } else {
                for (int pos = lo + gtid; pos < hi; pos += gstride) {
                    const uint32_t e = A.elist[pos];
                    if (e == S3_DEAD) continue;
                    const uint8_t fl = A.e_flag[pos];
                    phase_c(A.e_ra[pos], A.e_rb[pos], fl, w);
                    if (fl & 4) A.elist[pos] = S3_DEAD;
                }
            }
            FH_T(tC);
            grid.sync();
            FH_T(tS);
            if (nrem == 0) break;  // every live edge of the level joined: nothing left to pick
        }
    }
    if (gtid == 0) {
        A.counters[S3_MAX_ROUNDS - 8] = (int)(tA >> 10); A.counters[S3_MAX_ROUNDS - 7] = (int)(tB >> 10);
        A.counters[S3_MAX_ROUNDS - 6] = (int)(tC >> 10); A.counters[S3_MAX_ROUNDS - 5] = (int)(tS >> 10);
        A.counters[S3_MAX_ROUNDS - 9] = nlev; A.counters[S3_MAX_ROUNDS - 10] = round;
        tA = 0;
    }
    t0 = clock64();

    // ------------------------------------------------------------------ min-size merge
    grid.sync();
    const int E2 = 2 * A.N;
    for (int base = cbase; base < E2; base += gstride) {  // warp-uniform trip count
        const int e = base + threadIdx.x;
        bool cand = false;
        if (e < E2 && A.ew[e] != S3_NO_EDGE) {
            const int a = e >> 1, b = a + ((e & 1) ? A.W : 1);
            const int ra = uf_find(A.parent, a), rb = uf_find(A.parent, b);
            cand = ra != rb && (__ldcg(A.size + ra) < A.m || __ldcg(A.size + rb) < A.m);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, cand);
        if (bal) {
            const int lane = threadIdx.x & 31;
            int off = 0;
            if (lane == 0) off = atomicAdd(A.counters + CNT_LIST, __popc(bal));
            off = __shfl_sync(0xffffffffu, off, 0);
            if (cand) A.elist[off + __popc(bal & ((1u << lane) - 1))] = (uint32_t)e;
        }
    }
    grid.sync();
    const int nlist = __ldcg(A.counters + CNT_LIST);
    while (true) {
        if (round >= ROUND_CAP) {
            if (gtid == 0) A.counters[CNT_ERR] = 1;
            return;
        }
        int live = 0;
        for (int pos = gtid; pos < nlist; pos += gstride) {
            const uint32_t e = A.elist[pos];
            if (e == S3_DEAD) continue;
            const int a = (int)(e >> 1), b = a + ((e & 1u) ? A.W : 1);
            const int ra = uf_find(A.parent, a), rb = uf_find(A.parent, b);
            if (ra == rb) {
                A.elist[pos] = S3_DEAD;
                continue;
            }
            const bool fa = __ldcg(A.size + ra) < A.m, fb = __ldcg(A.size + rb) < A.m;
            if (!fa && !fb) {
                A.elist[pos] = S3_DEAD;  // big-big: can never fire (sizes only grow)
                continue;
            }
            const unsigned long long key = ((unsigned long long)A.ew[e] << 32) | e;
            atomicMin(A.resv + ra, key);
            atomicMin(A.resv + rb, key);
            A.e_ra[pos] = ra;
            A.e_rb[pos] = rb;
            A.e_flag[pos] = (uint8_t)((fa ? 1 : 0) | (fb ? 2 : 0));
            live++;
        }
        block_sum_to_counter(live, A.counters + round);
        grid.sync();
        const int nlive = __ldcg(A.counters + round);
        round++;
        if (nlive == 0) break;
        for (int pos = gtid; pos < nlist; pos += gstride) {
            const uint32_t e = A.elist[pos];
            if (e == S3_DEAD) continue;
            const int ra = A.e_ra[pos], rb = A.e_rb[pos];
            const uint8_t fl = A.e_flag[pos];
            const unsigned long long key = ((unsigned long long)A.ew[e] << 32) | e;
            const bool fa = fl & 1, fb = fl & 2;
            const bool oka = !fa || __ldcg(A.resv + ra) == key;
            const bool okb = !fb || __ldcg(A.resv + rb) == key;
            if (oka && okb) {
                A.mask[e] = 2;
                const bool a_hooks = (fa && !fb) || (fa && fb && ra > rb);
                const int frm = a_hooks ? ra : rb, to = a_hooks ? rb : ra;
                __stcg(A.parent + frm, to);
                atomicAdd(A.size + to, __ldcg(A.size + frm));
                A.e_flag[pos] = fl | 4;
            }
        }
        grid.sync();
        for (int pos = gtid; pos < nlist; pos += gstride) {
            const uint32_t e = A.elist[pos];
            if (e == S3_DEAD) continue;
            __stcg(A.resv + A.e_ra[pos], ~0ull);
            __stcg(A.resv + A.e_rb[pos], ~0ull);
            if (A.e_flag[pos] & 4) A.elist[pos] = S3_DEAD;
        }
        grid.sync();
    }
    FH_T(tA);
    if (gtid == 0) { A.counters[CNT_ROUNDS] = round; A.counters[S3_MAX_ROUNDS - 11] = (int)(tA >> 10); }
}

__global__ void k_uf_init(int N, int* parent, int* size, int* lastw, uint32_t* best, unsigned long long* resv,
                          int* minpix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    parent[i] = i;
    size[i] = 1;
    lastw[i] = 0;
    best[i] = S3_DEAD;
    resv[i] = ~0ull;
    minpix[i] = 0x7fffffff;
}

// ---- labelling: trees numbered by their minimum pixel (first-seen raster order, :352-367)
__global__ void k_label_roots(int N, int* parent, int* root_of, int* minpix) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const int r = uf_find(parent, p);
    root_of[p] = r;
    atomicMin(minpix + r, p);
}
__global__ void k_label_firsts(int N, const int* __restrict__ root_of, const int* __restrict__ minpix, int* firsts,
                               int* counter) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    if (minpix[root_of[p]] == p) firsts[atomicAdd(counter, 1)] = p;
}
__global__ void k_label_mark(int T, const int* __restrict__ rootpix, int* tid_at, int* tree_size) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    tid_at[rootpix[t]] = t;
    tree_size[t] = 0;
}
__global__ void k_label_ids(int N, const int* __restrict__ root_of, const int* __restrict__ minpix,
                            const int* __restrict__ tid_at, int* tree_id, int* tree_size) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const int t = tid_at[minpix[root_of[p]]];
    tree_id[p] = t;
    atomicAdd(tree_size + t, 1);
}

// ---- BFS re-indexing, one CTA per tree
#define BFS_THREADS 256
struct BfsArgs {
    int T;
    const int* unit_tree;
    const int* tree_start;
    const int* tree_rootpix;
    const uint16_t* ew;
    const uint8_t* mask;
    int* node_pixel;
    int* pixel_node;
    int* parent;
    int* level;
    uint16_t* pw;
    NodeUp* node_up;
    int4* node_dn;
    int* lvl_start;
    int* tree_depth;
    int4* tile_desc;
    int* tree_ntiles;
};
struct BfsArgs2 {
    BfsArgs v[2];
    int W, H, NN;
    int grid0;  // CTAs [0, grid0) serve v[0], the rest v[1]: both views' trees re-indexed by one launch
};
__global__ void __launch_bounds__(BFS_THREADS) k_bfs(BfsArgs2 AA) {
    const int vi = (int)blockIdx.x >= AA.grid0;
    const BfsArgs& B = AA.v[vi];
    const int bid = vi ? blockIdx.x - AA.grid0 : blockIdx.x, nb = vi ? gridDim.x - AA.grid0 : AA.grid0;
    const int T = B.T, W = AA.W, H = AA.H, NN = AA.NN;
    const int* __restrict__ unit_tree = B.unit_tree;
    const int* __restrict__ tree_start = B.tree_start;
    const int* __restrict__ tree_rootpix = B.tree_rootpix;
    const uint16_t* __restrict__ ew = B.ew;
    const uint8_t* __restrict__ mask = B.mask;
    int* node_pixel = B.node_pixel; int* pixel_node = B.pixel_node; int* parent = B.parent; int* level = B.level;
    uint16_t* pw = B.pw; NodeUp* node_up = B.node_up; int4* node_dn = B.node_dn; int* lvl_start = B.lvl_start;
    int* tree_depth = B.tree_depth; int4* tile_desc = B.tile_desc; int* tree_ntiles = B.tree_ntiles;
    __shared__ int s_warp[BFS_THREADS / 32];
    __shared__ int s_total;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int u = bid; u < T; u += nb) {
        const int t = unit_tree[u];
        const int base = tree_start[t];
        int* lvl = lvl_start + base + t;
        if (tid == 0) {
            const int rp = tree_rootpix[t];
            node_pixel[base] = rp;
            pixel_node[rp] = base;
            parent[base] = base;
            level[base] = 0;
            pw[base] = 0;
            node_dn[base] = make_int4(base, 0, 0, rp);
            lvl[0] = base;
        }
        __syncthreads();
        int a = base, b = base + 1, L = 0;
        while (a < b) {
            int run = 0;  // children emitted so far for this level (uniform)
            for (int chunk = a; chunk < b; chunk += BFS_THREADS) {
                const int g = chunk + tid;
                int cc = 0;
                int q[4];
                uint32_t wq[4];
                unsigned long long key[4];
                if (g < b) {
                    const int pix = node_pixel[g];
                    const int ppix = node_pixel[parent[g]];
                    const int x = pix % W, y = pix / W;
                    // candidate neighbours: left, right, up, down — forest edges only
                    int nq[4];
                    int ne[4];
                    nq[0] = pix - 1; ne[0] = x > 0 ? 2 * (pix - 1) : -1;
                    nq[1] = pix + 1; ne[1] = x < W - 1 ? 2 * pix : -1;
                    nq[2] = pix - W; ne[2] = y > 0 ? 2 * (pix - W) + 1 : -1;
                    nq[3] = pix + W; ne[3] = y < H - 1 ? 2 * pix + 1 : -1;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        if (ne[k] < 0 || nq[k] == ppix) continue;
                        if (!mask[ne[k]]) continue;
                        const uint32_t wv = ew[ne[k]];
                        const unsigned long long kk = ((unsigned long long)wv << 32) | (uint32_t)ne[k];
                        int j = cc++;  // insertion sort by (w, edge id)
                        while (j > 0 && key[j - 1] > kk) {
                            key[j] = key[j - 1]; q[j] = q[j - 1]; wq[j] = wq[j - 1];
                            j--;
                        }
                        key[j] = kk; q[j] = nq[k]; wq[j] = wv;
                    }
                }
                // block exclusive scan of cc
                int incl = cc;
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                if (lane == 31) s_warp[wid] = incl;
                __syncthreads();
                if (wid == 0) {
                    int v = lane < BFS_THREADS / 32 ? s_warp[lane] : 0;
                    int iv = v;
                    for (int o = 1; o < 32; o <<= 1) {
                        const int u2 = __shfl_up_sync(0xffffffffu, iv, o);
                        if (lane >= o) iv += u2;
                    }
                    if (lane < BFS_THREADS / 32) s_warp[lane] = iv - v;
                    if (lane == 31) s_total = iv;
                }
                __syncthreads();
                const int excl = incl - cc + s_warp[wid];
                const int chunk_total = s_total;
                if (g < b) {
                    const int cb = b + run + excl;
                    NodeUp nu;
                    nu.child_begin = cb;
                    nu.child_count = cc;
                    nu.cw01 = (cc > 0 ? wq[0] : 0u) | ((cc > 1 ? wq[1] : 0u) << 16);
                    nu.cw23 = (cc > 2 ? wq[2] : 0u) | ((cc > 3 ? wq[3] : 0u) << 16);
                    node_up[g] = nu;
                    for (int k = 0; k < cc; k++) {
                        const int h = cb + k;
                        node_pixel[h] = q[k];
                        pixel_node[q[k]] = h;
                        parent[h] = g;
                        level[h] = L + 1;
                        pw[h] = (uint16_t)wq[k];
                        node_dn[h] = make_int4(g, (int)wq[k], L + 1, q[k]);
                    }
                }
                run += chunk_total;
                __syncthreads();  // s_warp / s_total reuse + global writes visible to the block
            }
            a = b;
            b = b + run;
            L++;
            if (tid == 0) lvl[L] = a;
            __syncthreads();
        }
        if (tid == 0) tree_depth[t] = L;
        // aggregation tiles (<= S3_TILE_NODES consecutive nodes of one level): one thread per level, block scan of the
        // per-level tile counts; written twice: root->leaf order (levels ascending) at [0,2N) and leaf->root order
        // (levels descending) at [2N,4N)
        for (int dir = 0; dir < 2; dir++) {
            int run_t = 0;
            for (int l0 = 0; l0 < L; l0 += BFS_THREADS) {
                const int li = l0 + tid;                       // position in processing order
                const int l = dir == 0 ? li : L - 1 - li;      // level
                int ls = 0, le = 0, cnt = 0;
                if (li < L) {
                    ls = lvl[l];
                    le = lvl[l + 1];
                    cnt = (le - ls + S3_TILE_NODES - 1) / S3_TILE_NODES;
                }
                int incl = cnt;
                for (int o = 1; o < 32; o <<= 1) {
                    const int vv = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += vv;
                }
                if (lane == 31) s_warp[wid] = incl;
                __syncthreads();
                if (wid == 0) {
                    int vv = lane < BFS_THREADS / 32 ? s_warp[lane] : 0;
                    int iv = vv;
                    for (int o = 1; o < 32; o <<= 1) {
                        const int u2 = __shfl_up_sync(0xffffffffu, iv, o);
                        if (lane >= o) iv += u2;
                    }
                    if (lane < BFS_THREADS / 32) s_warp[lane] = iv - vv;
                    if (lane == 31) s_total = iv;
                }
                __syncthreads();
                size_t ti = (size_t)(dir ? NN : 0) + (size_t)base + run_t + incl - cnt + s_warp[wid];  // tile index (2 int4 each)
                if (li < L) {
                    const int ps = l > 0 ? lvl[l - 1] : 0;
                    for (int s0 = ls; s0 < le; s0 += S3_TILE_NODES, ti++) {
                        const int n = min(S3_TILE_NODES, le - s0);
                        tile_desc[2 * ti] = make_int4(s0, n, s0 - ls, (s0 == ls ? S3_TF_FIRST : 0) | (s0 + n >= le ? S3_TF_LAST : 0));
                        tile_desc[2 * ti + 1] = make_int4(le, ps, 0, 0);
                    }
                }
                run_t += s_total;
                __syncthreads();
            }
            if (tid == 0) tree_ntiles[t] = run_t;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
static void fill_fh_args(s3dmst_ctx* ctx, View& V, FHArgs& A) {
    A.W = ctx->W; A.N = ctx->N; A.c = ctx->P.fh_c; A.m = std::max(2, ctx->P.min_cc_size);
    A.ew = V.ew; A.elist = V.elist; A.lvl_off = V.lvl_off;
    A.parent = V.uf_parent; A.size = V.uf_size; A.lastw = V.uf_lastw; A.best = V.uf_best; A.resv = V.uf_resv;
    A.mask = V.mask; A.e_ra = V.e_ra; A.e_rb = V.e_rb; A.e_flag = V.e_flag; A.counters = V.counters;
}

// FH + min-size merge for the views in `mask`, one cluster per view, one launch
int s3_fh_launch(s3dmst_ctx* ctx, int mask) {
    FHArgs2 AA;
    int nv = 0;
    for (int view = 0; view < 2; view++)
        if (mask & (1 << view)) fill_fh_args(ctx, ctx->v[view], AA.v[nv++]);
    if (!nv) return 0;
    static int cluster_size = 0;  // largest cluster this device schedules: 16 (non-portable) or 8
    if (!cluster_size) {
        cudaFuncSetAttribute(k_fh_merge, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaGetLastError();
        for (int cs : {16, 8, 4, 2, 1}) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cs * 2); cfg.blockDim = dim3(1024);
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int ncl = 0;
            if (cudaOccupancyMaxActiveClusters(&ncl, k_fh_merge, &cfg) == cudaSuccess && ncl >= 2) { cluster_size = cs; break; }
            cudaGetLastError();
        }
        if (!cluster_size) return s3_fail(ctx, S3DMST_E_CUDA, "k_fh_merge: no cluster configuration is schedulable");
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster_size * nv); cfg.blockDim = dim3(1024); cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster_size; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    S3_CUDA(cudaLaunchKernelEx(&cfg, k_fh_merge, AA));
    ctx->launches++;
    return 0;
}

// tree sizes straight from the union-find (exact: every union added the hooked size)
__global__ void k_label_sizes(int T, const int* __restrict__ rootpix, const int* __restrict__ root_of,
                              const int* __restrict__ uf_size, int* tid_at, int* tree_size) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    tid_at[rootpix[t]] = t;
    tree_size[t] = uf_size[root_of[rootpix[t]]];
}
__global__ void k_label_ids2(int N, const int* __restrict__ root_of, const int* __restrict__ minpix,
                             const int* __restrict__ tid_at, int* tree_id) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < N) tree_id[p] = tid_at[minpix[root_of[p]]];
}

// Forest construction for the views in `mask` (bit 0 left, bit 1 right).  The views are independent, so each
// device stage is ONE launch covering both (FH+merge: one cluster per view; BFS: one CTA per tree of either
// view) and the two host round trips (tree count, tree sizes -> offsets and work order) are shared.
int s3_forest_stage_mask(s3dmst_ctx* ctx, int mask) {
    const int N = ctx->N, W = ctx->W, H = ctx->H;
    const int TB = 256;
    for (int view = 0; view < 2; view++) {
        if (!(mask & (1 << view))) continue;
        View& V = ctx->v[view];
        V.forest_ready = false;
        S3_TRY(s3_image_stage(ctx, view));
        k_uf_init<<<(N + TB - 1) / TB, TB, 0, ctx->stream>>>(N, V.uf_parent, V.uf_size, V.uf_lastw, V.uf_best, V.uf_resv, V.minpix);
        S3_LAUNCH_CHECK();
        S3_CUDA(cudaMemsetAsync(V.counters, 0, sizeof(int) * S3_MAX_ROUNDS, ctx->stream));
        S3_CUDA(cudaMemsetAsync(V.mask, 0, 2 * (size_t)N, ctx->stream));
    }
    S3_TRY(s3_fh_launch(ctx, mask));
    int hc[2][16];
    for (int view = 0; view < 2; view++) {
        if (!(mask & (1 << view))) continue;
        View& V = ctx->v[view];
        k_label_roots<<<(N + TB - 1) / TB, TB, 0, ctx->stream>>>(N, V.uf_parent, V.scan_tmp, V.minpix);
        S3_LAUNCH_CHECK();
        k_label_firsts<<<(N + TB - 1) / TB, TB, 0, ctx->stream>>>(N, V.scan_tmp, V.minpix, V.tree_rootpix, V.counters + CNT_FIRSTS);
        S3_LAUNCH_CHECK();
        S3_CUDA(cudaMemcpyAsync(hc[view], V.counters + S3_MAX_ROUNDS - 16, sizeof(int) * 16, cudaMemcpyDeviceToHost, ctx->stream));
    }
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<int> rootpix[2];
    for (int view = 0; view < 2; view++) {
        if (!(mask & (1 << view))) continue;
        View& V = ctx->v[view];
        if (hc[view][15]) return s3_fail(ctx, S3DMST_E_LIMIT, "forest kernel hit the round cap");
        if (getenv("S3_DEBUG_FH"))
            fprintf(stderr, "[fh view %d] levels %d fh-rounds %d total-rounds %d | kcycles A %d B %d C %d sync %d merge %d\n", view, hc[view][7], hc[view][6],
                    hc[view][13], hc[view][8], hc[view][9], hc[view][10], hc[view][11], hc[view][5]);
        const int T = hc[view][16 - (S3_MAX_ROUNDS - CNT_FIRSTS)];
        if (T <= 0 || T > N) return s3_fail(ctx, S3DMST_E_CUDA, "labelling produced T=%d", T);
        V.T = T;
        rootpix[view].resize(T);
        S3_CUDA(cudaMemcpyAsync(rootpix[view].data(), V.tree_rootpix, sizeof(int) * T, cudaMemcpyDeviceToHost, ctx->stream));
    }
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<int> tsize[2];
    for (int view = 0; view < 2; view++) {
        if (!(mask & (1 << view))) continue;
        View& V = ctx->v[view];
        const int T = V.T;
        std::sort(rootpix[view].begin(), rootpix[view].end());  // tree ids = first-seen raster order = by minimum pixel
        S3_CUDA(cudaMemcpyAsync(V.tree_rootpix, rootpix[view].data(), sizeof(int) * T, cudaMemcpyHostToDevice, ctx->stream));
        int* tid_at = V.pixel_node;  // scratch until BFS fills it: [N]
        k_label_sizes<<<(T + TB - 1) / TB, TB, 0, ctx->stream>>>(T, V.tree_rootpix, V.scan_tmp, V.uf_size, tid_at, V.tree_size);
        S3_LAUNCH_CHECK();
        k_label_ids2<<<(N + TB - 1) / TB, TB, 0, ctx->stream>>>(N, V.scan_tmp, V.minpix, tid_at, V.tree_id);
        S3_LAUNCH_CHECK();
        tsize[view].resize(T);
        S3_CUDA(cudaMemcpyAsync(tsize[view].data(), V.tree_size, sizeof(int) * T, cudaMemcpyDeviceToHost, ctx->stream));
    }
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    BfsArgs2 BA;
    memset(&BA, 0, sizeof BA);
    BA.W = W; BA.H = H; BA.NN = N;
    int nv = 0, grid = 0;
    for (int view = 0; view < 2; view++) {
        if (!(mask & (1 << view))) continue;
        View& V = ctx->v[view];
        const int T = V.T;
        V.h_tree_start.assign(T + 1, 0);
        for (int t = 0; t < T; t++) V.h_tree_start[t + 1] = V.h_tree_start[t] + tsize[view][t];
        if (V.h_tree_start[T] != N) return s3_fail(ctx, S3DMST_E_CUDA, "tree sizes sum to %d, expected %d", V.h_tree_start[T], N);
        V.h_unit_tree.resize(T);
        std::iota(V.h_unit_tree.begin(), V.h_unit_tree.end(), 0);
        std::stable_sort(V.h_unit_tree.begin(), V.h_unit_tree.end(), [&](int x, int y) { return tsize[view][x] > tsize[view][y]; });
        S3_CUDA(cudaMemcpyAsync(V.tree_start, V.h_tree_start.data(), sizeof(int) * (T + 1), cudaMemcpyHostToDevice, ctx->stream));
        S3_CUDA(cudaMemcpyAsync(V.unit_tree, V.h_unit_tree.data(), sizeof(int) * T, cudaMemcpyHostToDevice, ctx->stream));
        BfsArgs& B = BA.v[nv];
        B.T = T; B.unit_tree = V.unit_tree; B.tree_start = V.tree_start; B.tree_rootpix = V.tree_rootpix; B.ew = V.ew; B.mask = V.mask;
        B.node_pixel = V.node_pixel; B.pixel_node = V.pixel_node; B.parent = V.parent; B.level = V.level; B.pw = V.pw;
        B.node_up = V.node_up; B.node_dn = V.node_dn; B.lvl_start = V.lvl_start; B.tree_depth = V.tree_depth;
        B.tile_desc = V.tile_desc; B.tree_ntiles = V.tree_ntiles;
        const int g = std::min(T, ctx->num_sms * 8);
        if (nv == 0) BA.grid0 = g;
        grid += g;
        nv++;
    }
    if (nv == 1) BA.v[1] = BA.v[0];
    k_bfs<<<grid, BFS_THREADS, 0, ctx->stream>>>(BA);
    S3_LAUNCH_CHECK();
    for (int view = 0; view < 2; view++)
        if (mask & (1 << view)) {
            View& V = ctx->v[view];
            V.h_tree_depth.resize(V.T);
            S3_CUDA(cudaMemcpyAsync(V.h_tree_depth.data(), V.tree_depth, sizeof(int) * V.T, cudaMemcpyDeviceToHost, ctx->stream));
        }
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int view = 0; view < 2; view++)
        if (mask & (1 << view)) {
            View& V = ctx->v[view];
            V.max_depth = 0;
            for (int d : V.h_tree_depth) V.max_depth = std::max(V.max_depth, d);
            V.forest_ready = true;
            V.cost_ready = false;
            V.agg_ready = false;
        }
    return 0;
}

int s3_forest_stage(s3dmst_ctx* ctx, int view) { return s3_forest_stage_mask(ctx, 1 << view); }

int s3_forest_finalize_host(s3dmst_ctx* ctx, int view) {
    View& V = ctx->v[view];
    V.h_tree_depth.resize(V.T);
    S3_CUDA(cudaMemcpyAsync(V.h_tree_depth.data(), V.tree_depth, sizeof(int) * V.T, cudaMemcpyDeviceToHost, ctx->stream));
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    V.max_depth = 0;
    for (int d : V.h_tree_depth) V.max_depth = std::max(V.max_depth, d);
    V.forest_ready = true;
    V.cost_ready = false;
    V.agg_ready = false;
    return 0;
}
