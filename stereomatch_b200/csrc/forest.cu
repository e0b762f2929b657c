// stereomatch_b200/csrc/forest.cu — forest construction on the GPU (north-star items 2 and 3).
//
// Reference semantics (sequential): sort the 4-connected grid edges by (w, a, b), run the
// Felzenszwalb-Huttenlocher union rule  w <= thr[a] && w <= thr[b], thr = w + c/size
// (include/segment-graph.h:54-89), then merge every component smaller than max(2, min_size) along the
// same sorted order (src/Stereo3DMST.cpp:293-307), number the trees in first-seen raster order
// (:352-367) and re-index every tree in BFS order from its minimum pixel (:450-522).
//
// Parallel formulation (validated edge-for-edge against the sequential code by
// tests/models/forest_model.py + tests/test_forest_model.py on the CPU and tests/test_gpu_parity.py on the GPU):
//  * FH: asynchronous exact rounds over a weight-ordered live prefix of the edges — see k_fh_merge.  thr is never
//    stored: thr(r) = lastw[r] + f32(c)/f32(size[r]).  Edge key (w, edge id) == the reference's (w, a, b) order
//    (right edge 2p before down edge 2p+1, ascending p).
//  * The min-size merge is order dependent; it is replayed with deterministic reservations: every
//    pending edge atomicMin's its key (w<<32 | id) onto both endpoint components; an edge commits
//    when each endpoint component is either already >= m (its size can no longer matter) or holds
//    this edge as its reservation (no earlier pending edge touches it).  Edges between two big
//    components can never fire and are dropped.
//  Both phases run in ONE persistent cooperative kernel (software grid barrier between the phases of a round).
//  * BFS: one CTA per tree, level-synchronous inside the CTA (narrow levels: one warp, no block barrier); children
//    of a node are its forest neighbours except the parent, ordered by (w, edge id) (= the reference's adjacency
//    insertion order), numbered by an exclusive scan so BFS numbering equals the reference's queue order.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "internal.h"

#define CNT_T (S3_MAX_ROUNDS - 64)          // number of trees (device side)
#define CNT_UNIT_HIST (S3_MAX_ROUNDS - 256) // [64] size-class histogram + cursors of the work order
#define CNT_ERR (S3_MAX_ROUNDS - 1)
#define CNT_LIST (S3_MAX_ROUNDS - 2)
#define CNT_ROUNDS (S3_MAX_ROUNDS - 3)
#define CNT_ACT (S3_MAX_ROUNDS - 20)  // [2] live-list lengths of the FH rounds
#define ROUND_CAP (S3_FH_ROUNDS - 32)

// Component record: everything a round reads about a component (a root) is ONE 16-byte load (the forest kernel is
// bound by random sector requests: L2 for one pair, DRAM for a batch).  The pick word cleans itself: a posted key carries
// (S3_FH_ROUNDS - round) above the weight, so a value left by an earlier round is larger than anything the current
// round posts and simply loses the atomicMin — no second buffer, no clearing stores.  The union-find parents stay in a
// compact int array: the find chains walk ordinary pixels, whose neighbours share sectors (packing them too was measured
// 2x slower).  The kernels address the fields through strided pointers: int fields stride FHC_I ints, 64-bit fields FHC_L.
struct __align__(16) FHComp {
    unsigned long long pick;  // minimum live edge key of the component in the round named by its prefix
    int size;                 // pixels of the component (valid at the root)
    int lastw;                // weight of the last accepted edge (valid at the root)
};
#define FHC_I 4
#define FHC_L 2
#define FH_RSHIFT 42          // pick = (S3_FH_ROUNDS - round) << 42 | w << 32 | edge id   (w < 1024)

struct __align__(16) FHEntry {
    unsigned long long key;  // (w << 32) | edge id; bit 63 = accepted last round (ra = the root that hooked), bit 62 = rejected
    int ra, rb;              // endpoint components when the entry was last looked at
};
#define FH_HOOKED (1ull << 63)
#define FH_REJECT (1ull << 62)
#define FH_NOKEY (~0ull)

struct FHArgs {
    int W, N;
    float c;
    int m;
    int band_low, band_high;  // live-list control: ingest further weight levels when fewer than band_low entries are live
    const uint16_t* ew;
    uint32_t* elist;
    const int* lvl_off;
    int* parent;
    int* size;
    int* lastw;
    unsigned long long* pick;     // [N] minimum live key per component (round-prefixed, see FHComp)
    FHEntry* ent[2];              // [2N] live-edge lists, ping-pong
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

// two finds with their dependent loads interleaved (the chains are independent; each hop is an L2 round trip).
// Full path compression for the first few nodes of each path (kept in registers), halving beyond.
__device__ __forceinline__ void uf_find2(int* parent, int& x, int& y) {
    int ax = -1, bx = -1, cx = -1, dx4 = -1, ay = -1, by = -1, cy = -1, dy4 = -1;  // the last 4 non-root nodes of each path
    bool dx = false, dy = false;
    while (!(dx && dy)) {
        int px = x, py = y;
        if (!dx) px = __ldcg(parent + x);
        if (!dy) py = __ldcg(parent + y);
        if (px == py) {  // the chains have met: one component, whatever its root is (most edges by the time their level comes)
            x = y = px;
            break;
        }
        if (!dx) {
            if (px == x) dx = true;
            else { dx4 = cx; cx = bx; bx = ax; ax = x; x = px; }
        }
        if (!dy) {
            if (py == y) dy = true;
            else { dy4 = cy; cy = by; by = ay; ay = y; y = py; }
        }
    }
    // ax / ay already point at the root; the three before them are re-pointed (racing writers only store ancestors)
    if (bx >= 0) __stcg(parent + bx, x);
    if (cx >= 0) __stcg(parent + cx, x);
    if (dx4 >= 0) __stcg(parent + dx4, x);
    if (by >= 0) __stcg(parent + by, y);
    if (cy >= 0) __stcg(parent + cy, y);
    if (dy4 >= 0) __stcg(parent + dy4, y);
}

// segment-graph.h:27,80 — THRESHOLD(size,c) is a float division (Q2), added to a double w
__device__ __forceinline__ bool uf_open(float c, int size, int lastw, int w) {
    const double thr = (double)lastw + (double)__fdiv_rn(c, (float)size);
    return (double)w <= thr;
}

__device__ __forceinline__ int block_sum_to_counter(int v, int* counter) {
    // warp reduce then one atomic per warp
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(counter, v);
    return v;
}

struct FHArgs2 {
    FHArgs v[S3_FH_MAX_VIEWS];
    int nviews;
    int seg_cap;         // capacity of one CTA's segment of the live-edge lists (entries)
    int cluster;         // > 0: the launch gives every view ONE thread-block cluster of this many CTAs (hardware barrier)
    unsigned* bar;       // [S3_FH_MAX_VIEWS][32] one barrier counter per view (zeroed by the host)
    int* gcnt;           // [S3_FH_ROUNDS][S3_FH_MAX_VIEWS] per-round, per-view live counts (zeroed by the host)
};

// Grid-wide barrier for a co-resident grid (cooperative launch): one arrive per CTA, thread 0 spins on the counter.
// Mutable data is accessed with __ldcg/__stcg (L2), so the fence + barrier pair orders it across CTAs.
__device__ __forceinline__ void fh_grid_bar(unsigned* bar, unsigned& target, unsigned nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += nblocks;
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

// The same for a view that owns one thread-block cluster: the hardware cluster barrier (release / acquire at cluster
// scope orders the L2 accesses of the view's CTAs) costs a few hundred cycles instead of a round trip of atomics
// through L2 per CTA — the ~300 barriers of a forest were most of its time when one pair is alone on the GPU.
__device__ __forceinline__ void fh_cluster_bar() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// FH phase — asynchronous exact formulation.  Sequential FH visits edges in key order (w, id) and accepts an edge iff
// its endpoint components differ and are both open (w <= thr) at that moment.  Facts used (each validated
// edge-for-edge against the sequential code on the FLIR and synthetic images, tests/models/forest_model.py):
//   (1) if e is the minimum-key undecided edge of component C, C is unchanged when e's turn comes, so "C open at
//       w(e)" can be evaluated now; if it fails, C can never merge again (every other edge of C is heavier): C is
//       dead and all its edges are rejected;
//   (2) if e is the minimum of both its components, the decision is final now;
//   (3) if e = (C, D) is C's minimum, D is alive and D's minimum has the SAME weight, e is accepted: whatever D
//       merges with before e at that level stays open for the rest of the level (thr = w + c/size >= w).
// A round = every live edge posts its key on both components (atomicMin), then every edge that is somebody's
// minimum is decided by (1)-(3); chains of (3) resolve in one round like Boruvka hooks.  Only a weight-ordered
// PREFIX of the edges is live at any time (any prefix of the key order is self-contained for these rules); more
// levels are ingested when the live list runs low.  ~110-145 rounds at the C2 workload instead of ~900 with
// one-level-at-a-time Boruvka.
//
// Launch: cooperative, one CTA per SM; the first half of the grid serves view 0, the second half view 1 (random
// 4-8 byte accesses: the kernel is bound by L2 sector throughput and by the two barriers per round, so it wants
// every GPC's L2 ports, not one cluster's).  Every CTA keeps its OWN segment of the live list (survivors stay with
// the CTA that looked at them, new levels are dealt out evenly), so compaction needs no global cursor.
__global__ void __launch_bounds__(1024, 2) k_fh_merge(FHArgs2 AA) {
    const int nblk_all = gridDim.x;
    const bool cl = AA.cluster > 0;
    const int per_view = cl ? AA.cluster : nblk_all / AA.nviews;
    const int vi = min((int)blockIdx.x / per_view, AA.nviews - 1);
    const FHArgs& A = AA.v[vi];
    const int crank = blockIdx.x - vi * per_view;                                   // CTA rank inside its view (= its rank in the cluster)
    const int nblk = cl ? per_view : (vi == AA.nviews - 1 ? nblk_all - vi * per_view : per_view);     // CTAs of this view
    const int gtid = crank * blockDim.x + threadIdx.x;
    const int gstride = nblk * blockDim.x;
    const int cbase = crank * blockDim.x;
    const int lane = threadIdx.x & 31;
    unsigned bar_target = 0;
    int round = 0;
    __shared__ int s_cnt;
#define FH_BAR() do { if (cl) fh_cluster_bar(); else fh_grid_bar(AA.bar + 32 * vi, bar_target, (unsigned)nblk); } while (0)  // the views never wait for each other

    // ------------------------------------------------------------------ FH
    {
        int n_live = 0;    // live entries of this view after the previous round (all its CTAs)
        int lev = 0;       // weight levels ingested so far
        int band_pos = 0;  // == lvl_off[lev]
        int my_src = 0;    // entries in this CTA's source segment (live + flagged ones of the previous round)
        int par = 0;
        const size_t seg = (size_t)crank * AA.seg_cap;
        while (true) {
            const int band_lo_vi = band_pos;
            if (n_live < A.band_low)
                while (lev < S3_NUM_W && n_live + (band_pos - band_lo_vi) < A.band_high) band_pos = A.lvl_off[++lev];
            const int work = n_live + (band_pos - band_lo_vi);
            if (work == 0) break;  // nothing live and nothing left to ingest
            if (++round >= ROUND_CAP) {
                if (gtid == 0) A.counters[CNT_ERR] = 1;
                return;
            }
            const FHEntry* src = A.ent[par] + seg;
            FHEntry* dst = A.ent[par ^ 1] + seg;
            unsigned long long* pick = A.pick;
            const unsigned long long rtag = (unsigned long long)(S3_FH_ROUNDS - round) << FH_RSHIFT;
            // ---- phase 1: settle last round's decisions, re-root the survivors, post keys, compact
            const int n_new = band_pos - band_lo_vi;
            const int share = (n_new + nblk - 1) / nblk;
            const int new_lo = min(n_new, crank * share), new_hi = min(n_new, (crank + 1) * share);
            const int total = my_src + (new_hi - new_lo);
            if (threadIdx.x == 0) s_cnt = 0;
            __syncthreads();
            for (int base = 0; base < total; base += blockDim.x) {  // warp-uniform trip count
                const int pos = base + threadIdx.x;
                bool live = false;
                FHEntry en;
                en.key = FH_NOKEY; en.ra = 0; en.rb = 0;
                if (pos < my_src) {
                    en = src[pos];
                    if (en.key & FH_HOOKED) {
                        // accepted last round, en.ra hooked: its size flows to the root it ended under
                        const int R = uf_find(A.parent, en.ra);
                        atomicAdd(A.size + FHC_I * R, __ldcg(A.size + FHC_I * en.ra));
                        __stcg(A.lastw + FHC_I * R, (int)((en.key >> 32) & 0x3FFu));
                    } else if (!(en.key & FH_REJECT)) {
                        uf_find2(A.parent, en.ra, en.rb);
                        live = en.ra != en.rb;
                    }
                } else if (pos < total) {
                    const uint32_t e = A.elist[band_lo_vi + new_lo + (pos - my_src)];
                    en.key = ((unsigned long long)A.ew[e] << 32) | e;
                    en.ra = (int)(e >> 1);
                    en.rb = en.ra + ((e & 1u) ? A.W : 1);
                    uf_find2(A.parent, en.ra, en.rb);
                    live = en.ra != en.rb;
                }
                if (live) {
                    atomicMin(pick + FHC_L * en.ra, rtag | en.key);
                    atomicMin(pick + FHC_L * en.rb, rtag | en.key);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, live);
                if (bal) {
                    int off = 0;
                    if (lane == 0) off = atomicAdd(&s_cnt, __popc(bal));
                    off = __shfl_sync(0xffffffffu, off, 0);
                    if (live) dst[off + __popc(bal & ((1u << lane) - 1))] = en;
                }
            }
            __syncthreads();
            const int my_cnt = s_cnt;
            if (threadIdx.x == 0 && my_cnt) atomicAdd(AA.gcnt + S3_FH_MAX_VIEWS * round + vi, my_cnt);
            FH_BAR();
            n_live = __ldcg(AA.gcnt + S3_FH_MAX_VIEWS * round + vi);
            // ---- phase 2: decide every edge that is the minimum of one of its components
            for (int pos = threadIdx.x; pos < my_cnt; pos += blockDim.x) {
                FHEntry en = dst[pos];
                // one 16-byte load per endpoint: {pick, size, lastw}
                const int4 ca = __ldcg(reinterpret_cast<const int4*>(pick + FHC_L * en.ra)), cb = __ldcg(reinterpret_cast<const int4*>(pick + FHC_L * en.rb));
                const unsigned long long ka = ((unsigned long long)(unsigned)ca.y << 32 | (unsigned)ca.x) ^ rtag, kb = ((unsigned long long)(unsigned)cb.y << 32 | (unsigned)cb.x) ^ rtag;
                const int wa = (int)(ka >> 32), wb = (int)(kb >> 32), w = (int)(en.key >> 32);
                if (!uf_open(A.c, ca.z, ca.w, wa) || !uf_open(A.c, cb.z, cb.w, wb)) {  // (1): a dead endpoint
                    dst[pos].key = en.key | FH_REJECT;
                    continue;
                }
                const bool pa = ka == en.key, pb = kb == en.key;
                if (!(pa || pb)) continue;
                if ((pa && pb) || (pa && wb == w) || (pb && wa == w)) {  // (2), (3)
                    A.mask[(uint32_t)en.key] = 1;
                    int frm, to;
                    if (pa && pb) {  // only one side of a mutual pick hooks: the smaller component goes under the larger (shorter paths)
                        const int sa = ca.z, sb = cb.z;
                        const bool a_hooks = sa < sb || (sa == sb && en.ra > en.rb);
                        frm = a_hooks ? en.ra : en.rb; to = a_hooks ? en.rb : en.ra;
                    }
                    else if (pa) { frm = en.ra; to = en.rb; }
                    else { frm = en.rb; to = en.ra; }
                    __stcg(A.parent + frm, to);
                    en.key |= FH_HOOKED;
                    en.ra = frm; en.rb = to;
                    dst[pos] = en;
                }
            }
            FH_BAR();
            my_src = my_cnt;
            par ^= 1;
        }
    }

    // ------------------------------------------------------------------ min-size merge
    const int E2 = 2 * A.N;
    for (int base = cbase; base < E2; base += gstride) {  // warp-uniform trip count
        const int e = base + threadIdx.x;
        bool cand = false;
        if (e < E2 && A.ew[e] != S3_NO_EDGE) {
            int ra = e >> 1, rb = ra + ((e & 1) ? A.W : 1);
            uf_find2(A.parent, ra, rb);
            cand = ra != rb && (__ldcg(A.size + FHC_I * ra) < A.m || __ldcg(A.size + FHC_I * rb) < A.m);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, cand);
        if (bal) {
            int off = 0;
            if (lane == 0) off = atomicAdd(A.counters + CNT_LIST, __popc(bal));
            off = __shfl_sync(0xffffffffu, off, 0);
            if (cand) A.elist[off + __popc(bal & ((1u << lane) - 1))] = (uint32_t)e;
        }
    }
    FH_BAR();
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
            int ra = (int)(e >> 1), rb = ra + ((e & 1u) ? A.W : 1);
            uf_find2(A.parent, ra, rb);
            if (ra == rb) {
                A.elist[pos] = S3_DEAD;
                continue;
            }
            const bool fa = __ldcg(A.size + FHC_I * ra) < A.m, fb = __ldcg(A.size + FHC_I * rb) < A.m;
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
        round++;
        block_sum_to_counter(live, AA.gcnt + S3_FH_MAX_VIEWS * round + vi);
        FH_BAR();
        const int nlive = __ldcg(AA.gcnt + S3_FH_MAX_VIEWS * round + vi);
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
                atomicAdd(A.size + FHC_I * to, __ldcg(A.size + FHC_I * frm));
                A.e_flag[pos] = fl | 4;
            }
        }
        FH_BAR();
        for (int pos = gtid; pos < nlist; pos += gstride) {
            const uint32_t e = A.elist[pos];
            if (e == S3_DEAD) continue;
            __stcg(A.resv + A.e_ra[pos], ~0ull);
            __stcg(A.resv + A.e_rb[pos], ~0ull);
            if (A.e_flag[pos] & 4) A.elist[pos] = S3_DEAD;
        }
        FH_BAR();
    }
    if (gtid == 0) A.counters[CNT_ROUNDS] = round;
}

__global__ void k_uf_init(int N, FHComp* comp, int* parent, unsigned long long* resv, int* minpix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    FHComp c;
    c.pick = FH_NOKEY;
    c.size = 1;
    c.lastw = 0;
    comp[i] = c;
    parent[i] = i;
    resv[i] = ~0ull;
    minpix[i] = 0x7fffffff;
}

// ---- labelling: trees numbered by their minimum pixel (first-seen raster order, :352-367).  Entirely on the device:
// a tree's number is the rank of its minimum pixel among all minimum pixels = an exclusive scan of the "I am my
// tree's minimum pixel" flags over the image; tree_start is a scan of the tree sizes (exact: every union added the
// hooked size).  The host learns T and the sizes from ONE asynchronous copy it only waits for when it builds the
// aggregation work list — by then the BFS and the cost kernels are already queued behind these.
__global__ void k_label_roots(int N, int* parent, int* root_of, int* minpix) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const int r = uf_find(parent, p);
    root_of[p] = r;
    atomicMin(minpix + r, p);
}
__global__ void k_label_flags(int N, const int* __restrict__ root_of, const int* __restrict__ minpix, int* __restrict__ flag) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < N) flag[p] = minpix[root_of[p]] == p;
}
// exclusive scan of n ints (n on the host, or *n_dev when n_dev != nullptr): 1024 elements per CTA, block totals scanned
// by one CTA, then added back.  out[n] = total.
__global__ void __launch_bounds__(1024) k_scan_local(int n_host, const int* n_dev, const int* __restrict__ in, int* __restrict__ out, int* __restrict__ bsum) {
    const int n = n_dev ? *n_dev : n_host;
    __shared__ int s_w[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int i = blockIdx.x * 1024 + tid;
    if (blockIdx.x * 1024 >= n) { if (tid == 0) bsum[blockIdx.x] = 0; return; }
    const int v = i < n ? in[i] : 0;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const int w = s_w[lane];
        int winc = w;
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += u;
        }
        s_w[lane] = winc - w;
        if (lane == 31) bsum[blockIdx.x] = winc;
    }
    __syncthreads();
    if (i < n) out[i] = s_w[wid] + inc - v;
}
__global__ void __launch_bounds__(1024) k_scan_sums(int nb, int* bsum) {  // in-place exclusive scan of nb block totals, total -> bsum[nb]
    __shared__ int s_w[32];
    __shared__ int s_run;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_run = 0;
    __syncthreads();
    for (int b = 0; b < nb; b += 1024) {
        const int i = b + tid;
        const int v = i < nb ? bsum[i] : 0;
        int inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) s_w[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            const int w = s_w[lane];
            int winc = w;
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += u;
            }
            s_w[lane] = winc - w;
        }
        __syncthreads();
        const int run = s_run;
        if (i < nb) bsum[i] = run + s_w[wid] + inc - v;
        __syncthreads();
        if (tid == 1023) s_run = run + s_w[31] + inc;
        __syncthreads();
    }
    if (tid == 0) bsum[nb] = s_run;
}
__global__ void k_scan_apply(int n_host, const int* n_dev, int* __restrict__ out, const int* __restrict__ bsum, int nb, int* total_out) {
    const int n = n_dev ? *n_dev : n_host;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += bsum[i >> 10];
    if (i == 0) {
        out[n] = bsum[nb];
        if (total_out) *total_out = bsum[nb];
    }
}
// the minimum pixel of tree t: its rank, its pixel, its size
__global__ void k_label_assign(int N, const int* __restrict__ flag, const int* __restrict__ rank, const int* __restrict__ root_of,
                               const int* __restrict__ uf_size, int* __restrict__ rootpix, int* __restrict__ tree_size, int* __restrict__ tid_at) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N || !flag[p]) return;
    const int t = rank[p];
    rootpix[t] = p;
    tree_size[t] = uf_size[FHC_I * root_of[p]];
    tid_at[p] = t;
}
__global__ void k_label_ids2(int N, const int* __restrict__ root_of, const int* __restrict__ minpix,
                             const int* __restrict__ tid_at, int* tree_id) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < N) tree_id[p] = tid_at[minpix[root_of[p]]];
}
// work order of the per-tree kernels (BFS): trees by decreasing size class (log2 of the node count); the order inside
// a class is arbitrary (it only balances the load)
__global__ void k_unit_hist(const int* T_dev, const int* __restrict__ tree_size, int* hist) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < *T_dev) atomicAdd(hist + (31 - __clz(max(1, tree_size[t]))), 1);
}
__global__ void k_unit_scatter(const int* T_dev, const int* __restrict__ tree_size, int* hist, int* __restrict__ unit_tree) {
    __shared__ int s_off[32];
    if (threadIdx.x == 0) {  // offsets of the classes, largest first (hist[32..63] = cursors)
        int run = 0;
        for (int b = 31; b >= 0; b--) { s_off[b] = run; run += hist[b]; }
    }
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *T_dev) return;
    const int b = 31 - __clz(max(1, tree_size[t]));
    unit_tree[s_off[b] + atomicAdd(hist + 32 + b, 1)] = t;
}

// ---- per-pixel forest adjacency, pre-sorted: four 16-bit entries (w << 3) | direction (0 up, 1 left, 2 right, 3 down) in
// ascending order = the order the reference's per-vertex adjacency lists hold them (edges are inserted in sorted-edge
// order (w, a, b), Stereo3DMST.cpp:436-445; for equal w the edge ids order the directions up < left < right < down);
// 0xFFFF = no forest edge (its direction field, 7, matches no parent direction).  One 8-byte record per BFS node instead of 4 mask + 4 weight gathers and a sort.
__global__ void k_pix_adj(int W, int H, const uint16_t* __restrict__ ew, const uint8_t* __restrict__ mask,
                          unsigned long long* __restrict__ adjw) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= W * H) return;
    const int x = p % W, y = p / W;
    uint32_t e[4] = {0xFFFFu, 0xFFFFu, 0xFFFFu, 0xFFFFu};
    if (y > 0 && mask[2 * (p - W) + 1]) e[0] = ((uint32_t)ew[2 * (p - W) + 1] << 3) | 0u;
    if (x > 0 && mask[2 * (p - 1)]) e[1] = ((uint32_t)ew[2 * (p - 1)] << 3) | 1u;
    if (x < W - 1 && mask[2 * p]) e[2] = ((uint32_t)ew[2 * p] << 3) | 2u;
    if (y < H - 1 && mask[2 * p + 1]) e[3] = ((uint32_t)ew[2 * p + 1] << 3) | 3u;
#define S3_CSWAP(i, j) { const uint32_t lo = min(e[i], e[j]), hi = max(e[i], e[j]); e[i] = lo; e[j] = hi; }
    S3_CSWAP(0, 1) S3_CSWAP(2, 3) S3_CSWAP(0, 2) S3_CSWAP(1, 3) S3_CSWAP(1, 2)
#undef S3_CSWAP
    adjw[p] = (unsigned long long)e[0] | ((unsigned long long)e[1] << 16) | ((unsigned long long)e[2] << 32) | ((unsigned long long)e[3] << 48);
}

// ---- BFS re-indexing, one CTA per tree
#define BFS_THREADS 64
#ifndef BFS_INSTR
#define BFS_INSTR 0
#endif
#if BFS_INSTR
#define BFS_CLK(acc) do { const long long c__ = clock64(); acc += c__ - tq; tq = c__; } while (0)
#else
#define BFS_CLK(acc) do { } while (0)
#endif
#define BFS_FRONT 512   // frontier entries kept in shared memory per level (wider levels go through global memory)
struct BfsArgs {
    const int* T;     // device: number of trees
    const int* unit_tree;
    const int* tree_start;
    const int* tree_rootpix;
    const unsigned long long* adjw;
    uint32_t* front;  // [N] frontier words by node (pixel | direction of the parent << 28), the wide-level fallback
    int* node_pixel;
    int* pixel_node;
    int* parent;
    int* level;
    uint16_t* pw;
    NodeUp* node_up;
    int4* node_dn;
    int* lvl_start;
    int* tree_depth;
};
struct BfsArgs2 {
    BfsArgs v[2];
    int W, H, NN;
    int grid0;  // CTAs [0, grid0) serve v[0], the rest v[1]: both views' trees re-indexed by one launch
};

#if BFS_INSTR
__device__ long long g_dummy;
#endif
// One node of one BFS level: decode the frontier word, fetch the adjacency record, drop the parent's entry.
// Returns the child count; k0..k3 = the child entries in order, 0xFFFF-padded (kept in four registers: every use below
// is straight-line code — the per-level instruction count of ONE warp is what bounds this kernel).
__device__ __forceinline__ int bfs_expand(uint32_t fw, unsigned long long rec, int& pix, uint32_t& k0, uint32_t& k1, uint32_t& k2,
                                          uint32_t& k3) {
    pix = (int)(fw & 0x0FFFFFFFu);
    const uint32_t pdir = (fw >> 28) & 7u;  // 4 = no parent (root); bit 31: the parent is S3_AGG_NEAR or more nodes back
    const uint32_t lo = (uint32_t)rec, hi = (uint32_t)(rec >> 32);
    const uint32_t e0 = lo & 0xFFFFu, e1 = lo >> 16, e2 = hi & 0xFFFFu, e3 = hi >> 16;
    // f_k: the parent's entry sits at a position <= k; the list without it is e_k below it and e_{k+1} from there on
    const bool f0 = (e0 & 7u) == pdir, f1 = f0 || (e1 & 7u) == pdir, f2 = f1 || (e2 & 7u) == pdir, f3 = f2 || (e3 & 7u) == pdir;
    k0 = f0 ? e1 : e0;
    k1 = f1 ? e2 : e1;
    k2 = f2 ? e3 : e2;
    k3 = f3 ? 0xFFFFu : e3;
    return (k0 != 0xFFFFu) + (k1 != 0xFFFFu) + (k2 != 0xFFFFu) + (k3 != 0xFFFFu);
}

// Level-synchronous BFS from the tree's minimum pixel.  The reference's queue order (Stereo3DMST.cpp:477-518) is:
// nodes of a level in order, each appending its unvisited neighbours in adjacency order.  An exclusive scan of the
// child counts over the level reproduces that numbering.
// A tree level here is a handful of nodes (median 10 at C2) and the deepest tree has > 1000 levels, so what bounds
// the kernel is the number of dependent instructions ONE warp executes per level.  Levels of <= 32 nodes are run by
// warp 0 alone (no block barrier, ballot-based scan); the adjacency records are pre-sorted; the lines holding a
// node's possible grandchildren are pulled into L1 when the node is discovered, two levels before they are needed.
__global__ void __launch_bounds__(BFS_THREADS) k_bfs(BfsArgs2 AA) {
    const int vi = (int)blockIdx.x >= AA.grid0;
    const BfsArgs& B = AA.v[vi];
    const int bid = vi ? blockIdx.x - AA.grid0 : blockIdx.x, nb = vi ? gridDim.x - AA.grid0 : AA.grid0;
    const int T = *B.T, W = AA.W, NN = AA.NN;
    const int* __restrict__ unit_tree = B.unit_tree;
    const int* __restrict__ tree_start = B.tree_start;
    const int* __restrict__ tree_rootpix = B.tree_rootpix;
    const unsigned long long* __restrict__ adjw = B.adjw;
    uint32_t* front = B.front;
    int* pixel_node = B.pixel_node;
    NodeUp* node_up = B.node_up; int4* node_dn = B.node_dn; int* lvl_start = B.lvl_start;
    int* tree_depth = B.tree_depth;
    __shared__ int s_warp[BFS_THREADS / 32];
    __shared__ int s_total;
    __shared__ int s_state[4];
    __shared__ uint32_t s_front[2][BFS_FRONT];
    // the adjacency records of the frontier, fetched by the lane that DISCOVERS a node: the load (an L2 round trip, ~700
    // cycles — an L1 prefetch did not shorten it) flies while that lane writes the node's records, and the level that
    // expands the node finds its record in shared memory
    __shared__ unsigned long long s_rec[2][BFS_FRONT];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (int u = bid; u < T; u += nb) {
        const int t = unit_tree[u];
        const int base = tree_start[t];
        int* lvl = lvl_start + base + t;
        if (tid == 0) {
            const int rp = tree_rootpix[t];
            pixel_node[rp] = base;
            node_dn[base] = make_int4(base, 0, 0, rp);
            lvl[0] = base;
            s_front[0][0] = (uint32_t)rp | (4u << 28);  // direction 4 = no parent
            s_rec[0][0] = __ldg(adjw + rp);
        }
        __syncthreads();
        int a = base, b = base + 1, L = 0, cur = 0;
#if BFS_INSTR
        long long tq = clock64(), q_load = 0, q_scan = 0, q_store = 0, q_sync = 0;
#endif
        // emits the children of node g (child slots cb..cb+cc-1 of level L+1) and g's leaf->root record
        // one child of node g: frontier word, pixel -> node, root->leaf record, warm-up of its own children's lines
        auto emit_child = [&](int g, int pix, uint32_t en, int h, int bnext, int Lc, int curc) -> unsigned long long {
            const int dir = (int)(en & 7u);
            const int q = pix + ((dir & 2) ? 1 : -1) * ((dir == 0 || dir == 3) ? W : 1);  // 0 up, 1 left, 2 right, 3 down
            const uint32_t fw = (uint32_t)q | ((uint32_t)(3 - dir) << 28) | (h - g >= S3_AGG_NEAR ? 0x80000000u : 0u);  // the parent lies in the opposite direction
            const bool in_smem = h - bnext < BFS_FRONT;
            unsigned long long rec = 0ull;
            if (in_smem) rec = __ldg(adjw + q);  // in flight during the stores below
            if (in_smem) s_front[curc ^ 1][h - bnext] = fw;
            else front[h] = fw;
            pixel_node[q] = h;
            node_dn[h] = make_int4(g, (int)(en >> 3), Lc + 1, q);
            {  // q's possible children: rows above and below (q +- 1 share q's line)
                if (q >= W) asm volatile("prefetch.global.L2 [%0];" ::"l"(adjw + q - W));
                if (q + W < NN) asm volatile("prefetch.global.L2 [%0];" ::"l"(adjw + q + W));
            }
            return rec;
        };
        // emits the children of node g (child slots cb..cb+cc-1 of level L+1) and g's leaf->root record; `active` is false
        // for lanes without a node (they only take part in the warp-uniform votes)
        auto emit = [&](bool active, int g, int pix, int cc, uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3, int cb, int bnext, int Lc, int curc, uint32_t fwg) {
            if (active) {
                NodeUp nu;
                nu.child_begin = cb;
                nu.child_count = cc | ((fwg >> 31) ? S3_NU_FARPARENT : 0);
                nu.cw01 = (cc > 0 ? k0 >> 3 : 0u) | ((cc > 1 ? k1 >> 3 : 0u) << 16);
                nu.cw23 = (cc > 2 ? k2 >> 3 : 0u) | ((cc > 3 ? k3 >> 3 : 0u) << 16);
                node_up[g] = nu;
                if (cc == 0) reinterpret_cast<int*>(node_dn + g)[2] = Lc | S3_ND_LEAF;
                else if (cb + cc - 1 - g >= S3_AGG_NEAR) reinterpret_cast<int*>(node_dn + g)[2] = Lc | S3_ND_FAR;
            }
            unsigned long long r0 = 0ull, r1 = 0ull, r2 = 0ull, r3 = 0ull;
            if (__any_sync(0xffffffffu, active && cc > 0)) {
                if (active && cc > 0) r0 = emit_child(g, pix, k0, cb, bnext, Lc, curc);
                if (__any_sync(0xffffffffu, active && cc > 1)) {
                    if (active && cc > 1) r1 = emit_child(g, pix, k1, cb + 1, bnext, Lc, curc);
                    if (__any_sync(0xffffffffu, active && cc > 2)) {
                        if (active && cc > 2) r2 = emit_child(g, pix, k2, cb + 2, bnext, Lc, curc);
                        if (active && cc > 3) r3 = emit_child(g, pix, k3, cb + 3, bnext, Lc, curc);
                    }
                }
            }
            // the children's records, once they have arrived (every other store of the level is already on its way)
            if (active) {
                const int s0 = cb - bnext;
                if (cc > 0 && s0 < BFS_FRONT) s_rec[curc ^ 1][s0] = r0;
                if (cc > 1 && s0 + 1 < BFS_FRONT) s_rec[curc ^ 1][s0 + 1] = r1;
                if (cc > 2 && s0 + 2 < BFS_FRONT) s_rec[curc ^ 1][s0 + 2] = r2;
                if (cc > 3 && s0 + 3 < BFS_FRONT) s_rec[curc ^ 1][s0 + 3] = r3;
            }
        };
        while (true) {
            // ---- runs of narrow levels: warp 0 alone, no block barrier
            if (wid == 0) {
                while (a < b && b - a <= 32) {
                    int cc = 0, pix = 0;
                    uint32_t k0 = 0xFFFFu, k1 = 0xFFFFu, k2 = 0xFFFFu, k3 = 0xFFFFu;
                    const int g = a + lane;
                    uint32_t fwg = 0;
                    if (g < b) { fwg = s_front[cur][lane]; cc = bfs_expand(fwg, s_rec[cur][lane], pix, k0, k1, k2, k3); }
                    BFS_CLK(q_load);
                    // exclusive prefix of cc (0..4) from three independent ballots
                    const uint32_t b0 = __ballot_sync(0xffffffffu, cc & 1), b1 = __ballot_sync(0xffffffffu, cc & 2), b2 = __ballot_sync(0xffffffffu, cc & 4);
                    const int excl = __popc(b0 & lt_mask) + 2 * __popc(b1 & lt_mask) + 4 * __popc(b2 & lt_mask);
                    const int total = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
                    BFS_CLK(q_scan);
                    emit(g < b, g, pix, cc, k0, k1, k2, k3, b + excl, b, L, cur, fwg);
                    a = b;
                    b += total;
                    L++;
                    cur ^= 1;
                    if (lane == 0) lvl[L] = a;
                    BFS_CLK(q_store);
                    __syncwarp();
                    BFS_CLK(q_sync);
                }
                if (lane == 0) { s_state[0] = a; s_state[1] = b; s_state[2] = L; s_state[3] = cur; }
            }
            __syncthreads();
            a = s_state[0]; b = s_state[1]; L = s_state[2]; cur = s_state[3];
            if (a >= b) break;
            // ---- one wide level: the whole block, chunks of BFS_THREADS nodes
            int run = 0;  // children emitted so far for this level (uniform)
            for (int chunk = a; chunk < b; chunk += BFS_THREADS) {
                const int g = chunk + tid;
                int cc = 0, pix = 0;
                uint32_t k0 = 0xFFFFu, k1 = 0xFFFFu, k2 = 0xFFFFu, k3 = 0xFFFFu;
                uint32_t fwg = 0;
                if (g < b) {
                    const bool in_smem = g - a < BFS_FRONT;
                    fwg = in_smem ? s_front[cur][g - a] : front[g];
                    cc = bfs_expand(fwg, in_smem ? s_rec[cur][g - a] : __ldg(adjw + (fwg & 0x0FFFFFFFu)), pix, k0, k1, k2, k3);
                }
                int incl = cc;
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                if (lane == 31) s_warp[wid] = incl;
                __syncthreads();
                int wbase = 0, chunk_total = 0;
#pragma unroll
                for (int i = 0; i < BFS_THREADS / 32; i++) {
                    const int v = s_warp[i];
                    if (i < wid) wbase += v;
                    chunk_total += v;
                }
                emit(g < b, g, pix, cc, k0, k1, k2, k3, b + run + incl - cc + wbase, b, L, cur, fwg);
                run += chunk_total;
                __syncthreads();  // s_warp reuse; frontier words visible to the block
            }
            a = b;
            b = b + run;
            L++;
            cur ^= 1;
            if (tid == 0) { lvl[L] = a; s_state[0] = a; s_state[1] = b; s_state[2] = L; s_state[3] = cur; }
            __syncthreads();
        }
        __syncthreads();
        if (tid == 0) tree_depth[t] = L;
        __syncthreads();
#if BFS_INSTR
        if (u == 0 && tid == 0) printf("bfs cycles/level: load %lld scan %lld store %lld sync %lld\n", q_load / L, q_scan / L, q_store / L, q_sync / L);
#endif
    }
}

// node_dn -> the flat per-node arrays the other stages and the parity dumps read (off the BFS critical path)
__global__ void k_bfs_unpack(int N, const int4* __restrict__ node_dn, int* __restrict__ node_pixel, int* __restrict__ parent,
                             int* __restrict__ level, uint16_t* __restrict__ pw, uint32_t* __restrict__ leaf_bits) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    const int4 nd = h < N ? node_dn[h] : make_int4(0, 0, 0, 0);
    const uint32_t lb = __ballot_sync(0xffffffffu, h < N && (nd.z & S3_ND_LEAF));
    if ((threadIdx.x & 31) == 0 && h < N) leaf_bits[h >> 5] = lb;
    if (h >= N) return;
    parent[h] = nd.x;
    pw[h] = (uint16_t)nd.y;
    level[h] = nd.z & ~S3_ND_FLAGS;
    node_pixel[h] = nd.w;
}

// ------------------------------------------------------------------------------------------------
static void fill_fh_args(s3dmst_ctx* ctx, View& V, FHArgs& A) {
    A.W = ctx->W; A.N = ctx->N; A.c = ctx->P.fh_c; A.m = std::max(2, ctx->P.min_cc_size);
    A.ew = V.ew; A.elist = V.elist; A.lvl_off = V.lvl_off;
    FHComp* comp = reinterpret_cast<FHComp*>(V.uf_comp);
    A.parent = V.uf_parent; A.size = &comp->size; A.lastw = &comp->lastw; A.resv = V.uf_resv;
    A.pick = &comp->pick;
    A.ent[0] = reinterpret_cast<FHEntry*>(V.fh_ent[0]); A.ent[1] = reinterpret_cast<FHEntry*>(V.fh_ent[1]);
    // Live-list band.  A wide band means fewer rounds (latency: one pair alone on the GPU, 16384/65536), a narrow one
    // fewer futile re-visits of edges whose turn has not come (throughput: contexts set up for batching share the GPU and
    // are bound by random DRAM accesses, 8192/24576).  Measured at C2: batch of 8 20.65 -> 19.77 ms, single 5.00 -> 5.20.
    const bool batching = ctx->P.fh_ctas > 0;
    A.band_low = batching ? 8192 : 16384;
    A.band_high = batching ? 24576 : 65536;
    A.mask = V.mask; A.e_ra = V.e_ra; A.e_rb = V.e_rb; A.e_flag = V.e_flag; A.counters = V.counters;
}

// FH + min-size merge for the views in `mask` of every context in `ctxs` (one device): ONE cooperative launch on
// ctxs[0]'s stream, the grid split evenly between the views.  A batch of frames shares the ~150 barrier rounds.
int s3_fh_launch_multi(s3dmst_ctx** ctxs, int nctx, int mask) {
    s3dmst_ctx* ctx = ctxs[0];
    FHArgs2 AA;
    memset(&AA, 0, sizeof AA);
    int nv = 0;
    for (int c = 0; c < nctx; c++)
        for (int view = 0; view < 2; view++)
            if (mask & (1 << view)) {
                if (nv >= S3_FH_MAX_VIEWS) return s3_fail(ctx, S3DMST_E_ARG, "forest kernel: at most %d views per launch", S3_FH_MAX_VIEWS);
                fill_fh_args(ctxs[c], ctxs[c]->v[view], AA.v[nv++]);
            }
    if (!nv) return 0;
    int threads = ctx->P.fh_threads > 0 ? ctx->P.fh_threads : 1024;
    threads = std::max(64, std::min(1024, threads / 32 * 32));
    int ctas_per_sm = 0;
    S3_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_fh_merge, threads, 0));
    if (ctas_per_sm < 1) return s3_fail(ctx, S3DMST_E_CUDA, "k_fh_merge does not fit on an SM");
    const int want = (nctx > 1 || ctx->P.fh_ctas <= 0) ? ctx->num_sms : ctx->P.fh_ctas;  // a joint launch takes every SM
    const int cluster = ctx->P.fh_cluster > 0 ? std::min(16, ctx->P.fh_cluster) : 0;
    const int per_view = cluster ? cluster : std::max(1, std::min(std::min(want, ctx->num_sms * ctas_per_sm), S3_FH_MAX_CTAS) / nv);
    const int grid = per_view * nv;  // every view gets the same number of CTAs (and list segments of one size)
    if (!cluster && grid > ctx->num_sms * ctas_per_sm) return s3_fail(ctx, S3DMST_E_ARG, "forest kernel: %d views do not fit the GPU in one cooperative launch", nv);
    AA.nviews = nv;
    AA.seg_cap = (2 * ctx->N + per_view - 1) / per_view + S3_FH_SEG_SLACK;
    const size_t sync_ints = 32 * S3_FH_MAX_VIEWS + (size_t)S3_FH_ROUNDS * S3_FH_MAX_VIEWS;  // one barrier word (own 128-byte line) per view + counters
    if (!ctx->fh_sync) S3_CUDA(cudaMalloc(&ctx->fh_sync, sizeof(int) * sync_ints));
    S3_CUDA(cudaMemsetAsync(ctx->fh_sync, 0, sizeof(int) * sync_ints, ctx->stream));
    AA.bar = reinterpret_cast<unsigned*>(ctx->fh_sync);
    AA.gcnt = ctx->fh_sync + 32 * S3_FH_MAX_VIEWS;
    for (int c = 1; c < nctx; c++) {  // the other frames' image stages come first
        S3_CUDA(cudaEventRecord(ctxs[c]->ev_xctx, ctxs[c]->stream));
        S3_CUDA(cudaStreamWaitEvent(ctx->stream, ctxs[c]->ev_xctx, 0));
    }
    AA.cluster = cluster;
    if (cluster) {  // one cluster per view: its CTAs are co-scheduled by the hardware, the views are independent
        if (cluster > 8) S3_CUDA(cudaFuncSetAttribute(k_fh_merge, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(threads);
        cfg.stream = ctx->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        S3_CUDA(cudaLaunchKernelEx(&cfg, k_fh_merge, AA));
    } else {
        void* args[] = {&AA};
        S3_CUDA(cudaLaunchCooperativeKernel((const void*)k_fh_merge, dim3(grid), dim3(threads), args, 0, ctx->stream));
    }
    ctx->launches++;
    if (nctx > 1) {
        S3_CUDA(cudaEventRecord(ctx->ev_xctx, ctx->stream));
        for (int c = 1; c < nctx; c++) S3_CUDA(cudaStreamWaitEvent(ctxs[c]->stream, ctx->ev_xctx, 0));
    }
    return 0;
}

int s3_fh_launch(s3dmst_ctx* ctx, int mask) { return s3_fh_launch_multi(&ctx, 1, mask); }

// Forest construction for the views in `mask` (bit 0 left, bit 1 right).  The views are independent, so each
// device stage is ONE launch covering both (FH+merge: one cluster per view; BFS: one CTA per tree of either
// view) and the two host round trips (tree count, tree sizes -> offsets and work order) are shared.
// image stage + union-find initialisation of the views in `mask` (asynchronous on the context's stream)
int s3_forest_pre(s3dmst_ctx* ctx, int mask) {
    const int N = ctx->N;
    const int TB = 256;
    for (int view = 0; view < 2; view++) {
        if (!(mask & (1 << view))) continue;
        View& V = ctx->v[view];
        V.forest_ready = false;
        S3_TRY(s3_image_stage(ctx, view));
        k_uf_init<<<(N + TB - 1) / TB, TB, 0, ctx->stream>>>(N, reinterpret_cast<FHComp*>(V.uf_comp), V.uf_parent, V.uf_resv, V.minpix);
        S3_LAUNCH_CHECK();
        S3_CUDA(cudaMemsetAsync(V.counters, 0, sizeof(int) * S3_MAX_ROUNDS, ctx->stream));
        S3_CUDA(cudaMemsetAsync(V.mask, 0, 2 * (size_t)N, ctx->stream));
    }
    return 0;
}

static int scan_exclusive(s3dmst_ctx* ctx, int n_host, const int* n_dev, int n_max, const int* in, int* out, int* bsum, int* total_out) {
    const int nb = (n_max + 1023) / 1024;
    k_scan_local<<<nb, 1024, 0, ctx->stream>>>(n_host, n_dev, in, out, bsum);
    S3_LAUNCH_CHECK();
    k_scan_sums<<<1, 1024, 0, ctx->stream>>>(nb, bsum);
    S3_LAUNCH_CHECK();
    k_scan_apply<<<(n_max + 255) / 256, 256, 0, ctx->stream>>>(n_host, n_dev, out, bsum, nb, total_out);
    S3_LAUNCH_CHECK();
    return 0;
}

static int forest_tmax(const s3dmst_ctx* ctx) {  // a tree has at least max(2, min_cc_size) pixels unless the whole image is one small component
    return std::min(ctx->N, ctx->N / std::max(2, ctx->P.min_cc_size) + 2);
}

// Everything after the forest kernel, queued without any host synchronisation: tree numbering, tree_start, work order,
// adjacency records, BFS re-indexing, flat copies — and one copy of (error flag, T, tree sizes) into pinned memory,
// marked by an event.  s3_forest_finish_host() waits for that event when the host needs the sizes.
int s3_forest_post(s3dmst_ctx* ctx, int mask) {
    const int N = ctx->N, W = ctx->W, H = ctx->H;
    const int TB = 256;
    const int Tmax = forest_tmax(ctx);
    const size_t pin_view = 16 + (size_t)Tmax;
    if (ctx->h_pin_cap < 2 * pin_view) {
        if (ctx->h_pin) {
            S3_CUDA(cudaStreamSynchronize(ctx->stream));
            S3_CUDA(cudaFreeHost(ctx->h_pin));
        }
        ctx->h_pin = nullptr; ctx->h_pin_cap = 0;
        S3_CUDA(cudaHostAlloc(&ctx->h_pin, 2 * pin_view * sizeof(int), cudaHostAllocDefault));
        ctx->h_pin_cap = 2 * pin_view;
    }
    BfsArgs2 BA;
    memset(&BA, 0, sizeof BA);
    BA.W = W; BA.H = H; BA.NN = N;
    int nv = 0, grid = 0;
    for (int view = 0; view < 2; view++) {
        if (!(mask & (1 << view))) continue;
        View& V = ctx->v[view];
        int* root_of = V.scan_tmp;            // [N]
        int* flag = V.e_rb;                   // [N] (forest-kernel scratch, idle now)
        int* rank = V.e_ra;                   // [N + 1], block totals behind it
        int* bsum = V.e_ra + N + 8;
        int* tid_at = V.pixel_node;           // scratch until the BFS fills it
        int* T_dev = V.counters + CNT_T;
        int* hist = V.counters + CNT_UNIT_HIST;
        k_label_roots<<<(N + TB - 1) / TB, TB, 0, ctx->stream>>>(N, V.uf_parent, root_of, V.minpix);
        S3_LAUNCH_CHECK();
        k_label_flags<<<(N + TB - 1) / TB, TB, 0, ctx->stream>>>(N, root_of, V.minpix, flag);
        S3_LAUNCH_CHECK();
        S3_TRY(scan_exclusive(ctx, N, nullptr, N, flag, rank, bsum, T_dev));
        k_label_assign<<<(N + TB - 1) / TB, TB, 0, ctx->stream>>>(N, flag, rank, root_of, &reinterpret_cast<FHComp*>(V.uf_comp)->size, V.tree_rootpix,
                                                                  V.tree_size, tid_at);
        S3_LAUNCH_CHECK();
        k_label_ids2<<<(N + TB - 1) / TB, TB, 0, ctx->stream>>>(N, root_of, V.minpix, tid_at, V.tree_id);
        S3_LAUNCH_CHECK();
        S3_TRY(scan_exclusive(ctx, 0, T_dev, Tmax, V.tree_size, V.tree_start, bsum, nullptr));
        S3_CUDA(cudaMemsetAsync(hist, 0, sizeof(int) * 64, ctx->stream));
        k_unit_hist<<<(Tmax + TB - 1) / TB, TB, 0, ctx->stream>>>(T_dev, V.tree_size, hist);
        S3_LAUNCH_CHECK();
        k_unit_scatter<<<(Tmax + TB - 1) / TB, TB, 0, ctx->stream>>>(T_dev, V.tree_size, hist, V.unit_tree);
        S3_LAUNCH_CHECK();
        // the host's copy: [0] round-cap flag of the forest kernel, [1] T, [16 ..] tree sizes
        int* pin = ctx->h_pin + view * pin_view;
        S3_CUDA(cudaMemcpyAsync(pin, V.counters + CNT_ERR, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        S3_CUDA(cudaMemcpyAsync(pin + 1, T_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        S3_CUDA(cudaMemcpyAsync(pin + 16, V.tree_size, sizeof(int) * Tmax, cudaMemcpyDeviceToHost, ctx->stream));
        BfsArgs& B = BA.v[nv];
        B.T = T_dev; B.unit_tree = V.unit_tree; B.tree_start = V.tree_start; B.tree_rootpix = V.tree_rootpix; B.adjw = reinterpret_cast<const unsigned long long*>(V.adjw); B.front = V.bfs_front;
        B.node_pixel = V.node_pixel; B.pixel_node = V.pixel_node; B.parent = V.parent; B.level = V.level; B.pw = V.pw;
        B.node_up = V.node_up; B.node_dn = V.node_dn; B.lvl_start = V.lvl_start; B.tree_depth = V.tree_depth;
        const int g = std::min(Tmax, ctx->num_sms * 16);
        if (nv == 0) BA.grid0 = g;
        grid += g;
        nv++;
    }
    S3_CUDA(cudaEventRecord(ctx->ev_block, ctx->stream));
    ctx->forest_pending |= mask;
    if (nv == 1) BA.v[1] = BA.v[0];
    for (int view = 0; view < 2; view++)
        if (mask & (1 << view)) {
            View& V = ctx->v[view];
            k_pix_adj<<<(N + TB - 1) / TB, TB, 0, ctx->stream>>>(W, H, V.ew, V.mask, reinterpret_cast<unsigned long long*>(V.adjw));
            S3_LAUNCH_CHECK();
        }
    k_bfs<<<grid, BFS_THREADS, 0, ctx->stream>>>(BA);
    S3_LAUNCH_CHECK();
    for (int view = 0; view < 2; view++)
        if (mask & (1 << view)) {
            View& V = ctx->v[view];
            k_bfs_unpack<<<(N + TB - 1) / TB, TB, 0, ctx->stream>>>(N, V.node_dn, V.node_pixel, V.parent, V.level, V.pw, V.leaf_bits);
            S3_LAUNCH_CHECK();
            V.max_depth = -1;  // tree depths stay on the device until somebody asks (s3_forest_depths)
            V.adj_ready = false;
            V.forest_ready = true;   // the device side is complete in stream order; T and the sizes reach the host lazily
            V.cost_ready = false;
            V.agg_ready = false;
            V.T = -1;
        }
    return 0;
}

// The host's view of the forests queued by s3_forest_post: tree count and tree_start.  Waits (sleeping, not spinning:
// a batch has many frames and a box many ranks) for the copy s3_forest_post queued; no-op when nothing is pending.
int s3_forest_finish_host(s3dmst_ctx* ctx) {
    if (!ctx->forest_pending) return 0;
    const int N = ctx->N;
    const int Tmax = forest_tmax(ctx);
    const size_t pin_view = 16 + (size_t)Tmax;
    S3_CUDA(cudaEventSynchronize(ctx->ev_block));
    const int mask = ctx->forest_pending;
    ctx->forest_pending = 0;
    for (int view = 0; view < 2; view++) {
        if (!(mask & (1 << view))) continue;
        View& V = ctx->v[view];
        const int* pin = ctx->h_pin + view * pin_view;
        V.forest_ready = false;
        if (pin[0]) return s3_fail(ctx, S3DMST_E_LIMIT, "forest kernel hit the round cap");
        const int T = pin[1];
        if (T <= 0 || T > Tmax) return s3_fail(ctx, S3DMST_E_CUDA, "labelling produced T=%d (bound %d)", T, Tmax);
        V.T = T;
        V.h_tree_start.assign(T + 1, 0);
        for (int t = 0; t < T; t++) V.h_tree_start[t + 1] = V.h_tree_start[t] + pin[16 + t];
        if (V.h_tree_start[T] != N) return s3_fail(ctx, S3DMST_E_CUDA, "tree sizes sum to %d, expected %d", V.h_tree_start[T], N);
        V.forest_ready = true;
    }
    return 0;
}

int s3_forest_stage_mask(s3dmst_ctx* ctx, int mask) {
    S3_TRY(s3_forest_pre(ctx, mask));
    S3_TRY(s3_fh_launch(ctx, mask));
    return s3_forest_post(ctx, mask);
}

int s3_forest_stage(s3dmst_ctx* ctx, int view) { return s3_forest_stage_mask(ctx, 1 << view); }

// tree depths -> host (forest_info / parity dumps); the pipeline itself never needs them on the host
int s3_forest_depths(s3dmst_ctx* ctx, int view) {
    View& V = ctx->v[view];
    S3_TRY(s3_forest_finish_host(ctx));
    if (V.max_depth >= 0) return 0;
    V.h_tree_depth.resize(V.T);
    S3_CUDA(cudaMemcpyAsync(V.h_tree_depth.data(), V.tree_depth, sizeof(int) * V.T, cudaMemcpyDeviceToHost, ctx->stream));
    S3_CUDA(cudaStreamSynchronize(ctx->stream));
    V.max_depth = 0;
    for (int d : V.h_tree_depth) V.max_depth = std::max(V.max_depth, d);
    return 0;
}

// an uploaded forest (s3dmst_set_forest): the host already knows T and tree_start
int s3_forest_finalize_host(s3dmst_ctx* ctx, int view) {
    View& V = ctx->v[view];
    ctx->forest_pending &= ~(1 << view);
    V.max_depth = -1;
    V.adj_ready = false;
    V.forest_ready = true;
    V.cost_ready = false;
    V.agg_ready = false;
    return s3_forest_depths(ctx, view);
}
