// stereomatch_b200/csrc/aggregate3.cu — tree-filter aggregation + WTA, dataflow kernel (default).
//
// Same arithmetic and HBM layout as aggregate.cu (see the header there): FP64 in the reference's association
// order (src/Stereo3DMST.cpp:120-158, :173-185), node-major label-minor volumes, a lane owns the label pairs
// label(h, e) = l0 + h*64 + 2*lane + e for the whole kernel.  What changes is the schedule.
//
// Measured on B200 (profiles/r01_agg2_full.txt): the level-synchronous kernels are bound by what every warp has
// to execute between two block barriers — a tree level is ~10 nodes, the deepest tree of the C2 workload has 1401
// levels, and a lone warp retires a dependent instruction only every 6-15 cycles, so a barrier interval costs
// ~1100 cycles however the bytes arrive; 53 % of all warp samples sit in the barrier and the SM issues 0.9
// instructions per cycle.  Here there is NO block barrier inside a pass:
//   * The nodes of a tree are dealt round-robin to the warps in traversal order (descending BFS index on the way
//     up, ascending on the way down).  A warp owns a node from start to finish.
//   * A node waits only for what it really depends on: on the way up for its children, on the way down for its
//     parent.  Completion is published as one word per warp ("last node finished", release/acquire in shared
//     memory); warps finish their nodes in order, so node c is done iff its owner's word has passed c.
//   * Everything that does not depend on other nodes (node record, edge weights, the node's own cost / running-sum
//     row — loaded one iteration ahead into registers, with the rows pulled into L2 further ahead by one bulk
//     prefetch per round) happens before the wait; the dependent part is
//     poll -> shared-memory load -> DMUL/DADD chain -> shared-memory store -> publish, ~200 cycles per tree level.
//   * Values are handed over through a shared-memory ring indexed by node index (R = 64 rows).  A child/parent
//     further than NEAR = 32 nodes away (levels wider than ~16 nodes, where the chain has slack) is read from the
//     copy in L2.  Ring rows are reused safely without per-row flags: a row of node x is only ever read by nodes
//     closer than NEAR, so writing row x -/+ R waits until every warp's progress word has passed x -/+ (R - NEAR)
//     (a warp max/min over the progress words, cached while it holds).
// Several CTAs (trees) share an SM, so the issue slots one tree leaves empty while it waits are used by another.
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "hd_math.h"
#include "internal.h"

#define A3_PF 6      // rounds ahead of the bulk L2 prefetch
#ifndef A3_SLEEP
#define A3_SLEEP 20  // ns a waiting warp yields the issue slot for between two polls
#endif
#ifndef A3_INSTR
#define A3_INSTR 0
#endif
#if A3_INSTR
#define A3_CLK(acc) do { const long long c__ = clock64(); acc += c__ - tq; tq = c__; } while (0)
#else
#define A3_CLK(acc) do { } while (0)
#endif

struct Agg3View {
    const int* tree_start;
    const NodeUp* node_up;
    const int4* node_dn;
    const int* node_pixel;
    const float* cost;
    double* aup;
    int32_t* disp;    // pixel order: final results when the launch has one slice
    double* best;
    int32_t* pdisp;   // [slice][node] partial results otherwise
    double* pbest;
    // proposal mode with the plane cost (params.pms_cost_mode = 1): this view's and the other view's image and gradients
    const uint8_t* img_self;
    const uint8_t* img_other;
    const float* grad_self;
    const float* grad_other;
    // dense mode with the matching cost computed in the kernel (FUSE): {packed BGR, gray} of this view's and the other view's image
    const uint2* m_self;
    const uint2* m_other;
};

struct Agg3Args {
    const Agg3View* views;  // device table: [frame][view] of every context in the launch
    const int4* units;  // {view, tree, first label, slice index}, longest tree first
    int Dp, d1, N, n_slices;
    int unit0;          // first unit of this launch
    const void* lut_w;   // exp(-w*gamma) and 1 - w*w tables in the state type
    const void* lut_w2;
    int keep;
    int sleep_ns;
    // proposal mode (PMS = true): labels are 3D planes tested on whole trees (MSTCostAggregationAndLabelUpdate, :160-186)
    const int* prop_off;   // [T+1] proposals of tree t: [prop_off[t], prop_off[t+1]) of `labels`, in list order
    const float* labels;   // [n][3] (a, b, c)
    double* min_cost;      // [N] pixel order
    float* abc;            // [N][3] pixel order
    int img_w, D;          // image width, labels of the cost rows
    unsigned long long w_magic;  // ceil(2^40 / img_w): pix / img_w == (pix * w_magic) >> 40 for pix < 2^28 (FUSE)
    float oob;             // label cost outside [0, D)
    // proposal generation inside the kernel (s3dmst_pms_iterate): after a tree's listed proposals (its neighbours' labels),
    // the refinement ladder around a random pixel of the tree itself
    int cost_mode, view, img_h;            // proposal mode: 1 = plane cost from the images (hd_math.h: s3_plane_cost)
    float pm_alpha, pm_tau_c, pm_tau_g, pm_scale;
    int gen;
    uint32_t seed, round;
    float refine_floor;
};

__device__ __forceinline__ uint32_t a3_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Truncated colour + gradient matching cost of one (pixel, label) — the expression of k_cost_adgrad (cost.cu;
// PatchMatchStereoGPU.cu:1482-1550), operation for operation, so that a dense run that never materialises the volume
// aggregates bit-identical numbers.  ct = the 22-entry colour-term table (L1 distance saturates at 21).
__device__ __forceinline__ float a3_adgrad(const float* ct, int view, uint32_t me, float me_g, float me_gn, uint32_t o, float og, float ogn) {
    const int l1 = (int)__dp4a(__vabsdiffu4(o, me), 0x00010101u, 0u);
    const float ctv = ct[min(l1, 21)];
    const float g = view == 0 ? S3_FADD(S3_FSUB(me_g, og), S3_FSUB(ogn, me_gn)) : S3_FADD(S3_FSUB(og, me_g), S3_FSUB(me_gn, ogn));
    const float ag = fabsf(g);
    const float gterm = ag < 2.0f ? ag : 2.0f;
    return S3_FADD(ctv, S3_FMUL(0.89f, gterm));
}
__global__ void k_pack_match(int N, const uchar4* __restrict__ raw4, const float* __restrict__ gray, uint2* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const uchar4 o = raw4[p];
    out[p] = make_uint2((uint32_t)o.x | ((uint32_t)o.y << 8) | ((uint32_t)o.z << 16), __float_as_uint(gray[p]));
}
__device__ __forceinline__ int a3_ld_acquire(uint32_t a) {
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void a3_st_release(uint32_t a, int v) {
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ int a3_ld_relaxed(uint32_t a) {
    int v;
    asm volatile("ld.relaxed.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
// ---- thread-block cluster (one giant tree walked by CL CTAs): progress words and ring rows of the other CTAs are read
// through distributed shared memory
__device__ __forceinline__ uint32_t a3_crank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t a3_mapa(uint32_t a, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(cta));
    return r;
}
// Cluster-scope acquire / release compile to MEMBAR.ALL.GPU and an L1 invalidation (CCTL.IVALL) around EVERY access: far
// too heavy for a poll loop.  The hand-over does not need them: a row and its progress word are written by one warp
// into its own shared memory in program order (STS row, bar.warp, STS word), remote reads of that memory are serviced
// in arrival order, and a reader issues its row loads only after the poll loop's branch has consumed the word (no
// speculation).  What does travel through global memory (far rows) is fenced explicitly by the writer (a3_fence_cl
// before the word is stored) and read with ld.global.cg (L2), never from L1.
__device__ __forceinline__ int a3_ld_acquire_cl(uint32_t a) {
    int v;
    asm volatile("ld.relaxed.cluster.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void a3_st_release_cl(uint32_t a, int v) {  // a: this CTA's own shared window
    asm volatile("st.volatile.shared::cta.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");  // an STS behind the row's STSs, same pipe
}
__device__ __forceinline__ void a3_fence_cl() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void a3_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double2 a3_ldcg_d2(const double* p) {
    double2 v;
    asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double a3_lds_d(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double2 a3_lds_d2(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void a3_sts_d2(uint32_t a, double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void a3_prefetch_l2(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// total order on doubles as unsigned 64-bit keys (handles negative costs of the mc-cnn "fast" volumes)
__device__ __forceinline__ unsigned long long a3_dkey(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double a3_dkey_inv(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// ---- the two state types: double (exact: the reference's arithmetic, bit for bit) and float (fast: 12 instead of 20
// bytes of HBM traffic per pixel-label; results within rounding of the exact ones, no bit-exactness promise)
template <typename T> struct A3T;
template <> struct A3T<double> {
    typedef double2 T2;
    static __device__ __forceinline__ T2 zero2() { return make_double2(0.0, 0.0); }
    static __device__ __forceinline__ double maxv() { return DBL_MAX; }
    static __device__ __forceinline__ double add(double a, double b) { return S3_DADD(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return S3_DMUL(a, b); }
    static __device__ __forceinline__ double ldsw(uint32_t a) { return a3_lds_d(a); }
    static __device__ __forceinline__ T2 lds2(uint32_t a) { return a3_lds_d2(a); }
    static __device__ __forceinline__ void sts2(uint32_t a, T2 v) { a3_sts_d2(a, v); }
    static __device__ __forceinline__ T2 ldc2(uint32_t a) {  // distributed shared memory (mapa address)
        double2 v;
        asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
        return v;
    }
    static __device__ __forceinline__ T2 ldcg2(const char* p) { return a3_ldcg_d2(reinterpret_cast<const double*>(p)); }
    // warp arg-min with three 32-bit REDUX steps: (cost hi, cost lo, label); ties -> lowest label
    static __device__ __forceinline__ unsigned warp_argmin(double bc, int bd, double& mc) {
        const unsigned long long key = a3_dkey(bc);
        const unsigned khi = (unsigned)(key >> 32), klo = (unsigned)key;
        const unsigned mhi = __reduce_min_sync(0xffffffffu, khi);
        const unsigned mlo = __reduce_min_sync(0xffffffffu, khi == mhi ? klo : 0xffffffffu);
        const unsigned md = __reduce_min_sync(0xffffffffu, (khi == mhi && klo == mlo) ? (unsigned)bd : 0x7fffffffu);
        mc = a3_dkey_inv(((unsigned long long)mhi << 32) | mlo);
        return md;
    }
};
template <> struct A3T<float> {
    typedef float2 T2;
    static __device__ __forceinline__ T2 zero2() { return make_float2(0.f, 0.f); }
    static __device__ __forceinline__ float maxv() { return FLT_MAX; }
    static __device__ __forceinline__ float add(float a, float b) { return S3_FADD(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return S3_FMUL(a, b); }
    static __device__ __forceinline__ float ldsw(uint32_t a) {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
        return v;
    }
    static __device__ __forceinline__ T2 lds2(uint32_t a) {
        float2 v;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
        return v;
    }
    static __device__ __forceinline__ void sts2(uint32_t a, T2 v) {
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
    }
    static __device__ __forceinline__ T2 ldc2(uint32_t a) {
        float2 v;
        asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
        return v;
    }
    static __device__ __forceinline__ T2 ldcg2(const char* p) {
        float2 v;
        asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
        return v;
    }
    static __device__ __forceinline__ unsigned warp_argmin(float bc, int bd, double& mc) {
        const unsigned b = __float_as_uint(bc);
        const unsigned key = (b >> 31) ? ~b : (b | 0x80000000u);
        const unsigned mk = __reduce_min_sync(0xffffffffu, key);
        const unsigned md = __reduce_min_sync(0xffffffffu, key == mk ? (unsigned)bd : 0x7fffffffu);
        mc = (double)__uint_as_float((mk >> 31) ? (mk & 0x7fffffffu) : ~mk);
        return md;
    }
};

// FULL: every lane's label pairs are real labels in every half (no per-lane predication in the loops)
// BIG: 32 warps per tree, one CTA per SM (the biggest trees: enough warps that a level's nodes do not need every warp)
// PMS: the "labels" of a pass are up to 64 injected proposals of the tree (two per lane); the cost of (node, proposal)
// is compute3DLabelCost on the node's cost row; the epilogue is the label update instead of the WTA; a tree with more
// than 64 proposals runs its passes once per batch of 64, in list order.
// CL: thread-block cluster size.  CL > 1: ONE tree is walked by the CL x 32 warps of a cluster (a tree's time on one CTA is
// proportional to its node count: the 204 k-node tree of the FLIR pair bounded the whole launch).  Processing position k
// (k = top - v on the way up, v - base on the way down) is owned by CTA k % CL, warp (k / CL) % 32; its ring row is row
// (k / CL) % R of the owner's ring, so the cluster's rings together hold the last R x CL rows and hand over values between
// nodes closer than NEAR x CL.  Progress words and ring rows of other CTAs are read through distributed shared memory
// (mapa + ld.acquire.cluster / ld.shared::cluster), published with st.release.cluster; the far path through L2 is the
// same (the release is cluster scope, so global stores before it are visible to the other SMs of the cluster).
template <typename T, int NH, bool FULL, bool BIG, int A3_R, int A3_NEAR, bool PMS, int CL, bool FUSE = false>
__global__ void __launch_bounds__(BIG ? 1024 : 512, BIG ? 1 : 2) k_agg_flow(Agg3Args A) {
    static_assert(CL == 1 || BIG, "a cluster walks a tree with 32 warps per CTA");
    static_assert(!(FUSE && PMS), "the fused matching cost belongs to the dense mode");
    static_assert(CL == 1 || CL == 2 || CL == 4 || CL == 8, "portable cluster sizes");
    constexpr int LOGCL = CL == 1 ? 0 : CL == 2 ? 1 : CL == 4 ? 2 : 3;
    constexpr int NEAR_E = A3_NEAR * CL;   // hand-over distance and ring rows of the whole cluster
    constexpr int R_E = A3_R * CL;
    extern __shared__ __align__(16) unsigned char s_raw[];
    using TT = A3T<T>;
    using T2 = typename TT::T2;
    constexpr uint32_t HB = 64 * sizeof(T);                         // bytes of 64 labels (one half) of a running-sum row
    T* s_w = reinterpret_cast<T*>(s_raw);                           // [S3_NUM_W] exp(-w*gamma)
    T* s_w2 = s_w + S3_NUM_W;                                       // [S3_NUM_W] 1 - w*w
    T2* s_ring = reinterpret_cast<T2*>(s_raw + ((2 * S3_NUM_W * sizeof(T) + 15) / 16) * 16);  // [A3_R][NH][32]
    int* s_prog = reinterpret_cast<int*>(s_ring + A3_R * NH * 32);  // [32] progress words
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, W = blockDim.x >> 5, WM = W - 1;
    const int WT = W * CL;                                          // warps walking the tree
    const uint32_t crank = CL > 1 ? a3_crank() : 0u;
    const int gw = w * CL + (int)crank;                             // this warp's position in the deal

    const int4 unit = A.units[A.unit0 + blockIdx.x / CL];
    const Agg3View V = A.views[unit.x];
    const int t = unit.y, slice = unit.w;
    const int base = V.tree_start[t], end = V.tree_start[t + 1], top = end - 1;
    const size_t Dp = (size_t)A.Dp;                 // cost row length
    const size_t DA = PMS ? (size_t)64 : Dp;        // running-sum row length (proposal mode: a [node][64] scratch)
    __shared__ float s_lab[PMS ? 64 * 3 : 1];
    __shared__ int s_nlad;
    __shared__ float s_ct[FUSE ? 22 : 1];  // colour term of L1 = i: 0.11f * min((float)((double)i * 0.33333333333), 7.0f)
    if constexpr (FUSE) {
        if (tid < 22) {
            float ctv = (float)S3_DMUL((double)(float)tid, 0.33333333333);
            ctv = ctv < 7.0f ? ctv : 7.0f;
            s_ct[tid] = S3_FMUL(0.11f, ctv);
        }
    }
    const int p_lo = PMS ? A.prop_off[t] : 0, p_hi = PMS ? A.prop_off[t + 1] : 1;
    // passes over the tree: dense mode one; proposal mode one per batch of 64 listed proposals, then (A.gen) one for the
    // refinement ladder this kernel generates itself from the label the tree holds after those (MST_PMS, :584-625)
    const int n_listed = PMS ? (p_hi - p_lo + 63) / 64 : 1;
    const int n_pass = n_listed + ((PMS && A.gen) ? 1 : 0);

    for (int i = tid; i < S3_NUM_W; i += blockDim.x) {
        s_w[i] = reinterpret_cast<const T*>(A.lut_w)[i];
        s_w2[i] = reinterpret_cast<const T*>(A.lut_w2)[i];
    }
    if (tid < 32) s_prog[tid] = end;  // up pass: node c is done iff its owner's word is <= c
    if constexpr (CL > 1) a3_cluster_sync(); else __syncthreads();
    const uint32_t prog_a = a3_smem(s_prog);
    const unsigned sleep_ns = (unsigned)A.sleep_ns;
    const uint32_t ring_a = a3_smem(s_ring) + (uint32_t)sizeof(T2) * lane;  // this lane's column of the ring
    const uint32_t w_a = a3_smem(s_w);
    constexpr uint32_t ROWB = NH * HB;                      // bytes of one ring row
    const long long strideC = (long long)WT * (long long)Dp * 4, strideA = (long long)WT * (long long)(PMS ? 64 : A.Dp) * (long long)sizeof(T);  // bytes between a warp's consecutive rows
    T* const aupT = reinterpret_cast<T*>(V.aup);  // running sums in the state type (the buffer is sized for doubles)
    // progress word / ring row of processing position k (see the header: owner CTA k % CL, warp (k / CL) % W)
    auto prog_of = [&](int k) -> uint32_t {
        if constexpr (CL == 1) return prog_a + 4u * (uint32_t)(k & WM);
        else return a3_mapa(prog_a + 4u * (uint32_t)((k >> LOGCL) & WM), (uint32_t)(k & (CL - 1)));
    };
    auto poll = [&](uint32_t pa) -> int {
        if constexpr (CL == 1) return a3_ld_acquire(pa);
        else return a3_ld_acquire_cl(pa);
    };
    auto row_of = [&](int k) -> uint32_t {  // this lane's column of the row, in the owner's window
        const uint32_t ra = ring_a + (uint32_t)((k >> LOGCL) & (A3_R - 1)) * ROWB;
        if constexpr (CL == 1) return ra;
        else return a3_mapa(ra, (uint32_t)(k & (CL - 1)));
    };
    auto ld_row = [&](uint32_t ra) -> T2 {
        if constexpr (CL == 1) return TT::lds2(ra);
        else return TT::ldc2(ra);
    };
    auto publish = [&](int v) {
        if constexpr (CL == 1) a3_st_release(prog_a + 4u * w, v);
        else a3_st_release_cl(prog_a + 4u * w, v);
    };
    // maximum / minimum over the progress words of every warp walking the tree (ring-reuse guard)
    auto prog_max = [&]() -> int {
        int m = INT_MIN;
        if (lane < W) {
            if constexpr (CL == 1) m = a3_ld_acquire(prog_a + 4u * lane);
            else {
#pragma unroll
                for (int j = 0; j < CL; j++) m = max(m, a3_ld_acquire_cl(a3_mapa(prog_a + 4u * lane, (uint32_t)j)));
            }
        }
        return __reduce_max_sync(0xffffffffu, m);
    };
    auto prog_min = [&]() -> int {
        int m = INT_MAX;
        if (lane < W) {
            if constexpr (CL == 1) m = a3_ld_acquire(prog_a + 4u * lane);
            else {
#pragma unroll
                for (int j = 0; j < CL; j++) m = min(m, a3_ld_acquire_cl(a3_mapa(prog_a + 4u * lane, (uint32_t)j)));
            }
        }
        return __reduce_min_sync(0xffffffffu, m);
    };

    for (int pass = 0; pass < n_pass; pass++) {
    const bool gen = PMS && pass >= n_listed;
    const int b0 = PMS ? (gen ? 0 : p_lo + 64 * pass) : 0;   // proposal mode: index of the batch's first proposal (s_lab[0])
    int b_end = gen ? 0 : p_hi;                              // ... and of the first one that does not exist
    if (PMS) {
        if constexpr (CL > 1) a3_cluster_sync(); else __syncthreads();  // the previous batch is done with s_lab, the ring and the progress words
        if (gen) {
            // every CTA walking the tree derives the same ladder: the label of a random node of the tree as the listed
            // proposals left it (written by this cluster before the barrier above; read from L2)
            if (tid == 0) {
                const int pix = V.node_pixel[base + (int)(s3_rng(A.seed, A.round, (uint32_t)t, S3_SLOT_REFINE_PIXEL) % (uint32_t)(end - base))];
                const float* l = A.abc + 3 * (size_t)pix;
                s_nlad = s3_refine_ladder(__ldcg(l), __ldcg(l + 1), __ldcg(l + 2), (float)(pix % A.img_w), (float)(pix / A.img_w), A.D, A.refine_floor,
                                          A.seed, A.round, (uint32_t)t, s_lab);
            }
        } else
            for (int i = tid; i < 3 * min(64, p_hi - b0); i += blockDim.x) s_lab[i] = A.labels[3 * (size_t)b0 + i];
        if (tid < 32) s_prog[tid] = end;
        if constexpr (CL > 1) a3_cluster_sync(); else __syncthreads();
        if (gen) b_end = s_nlad;
        if (b_end <= b0) continue;  // every step of the ladder left [0, Dmax]
    }
    const int lim = PMS ? b_end : A.d1; // first label / proposal index that does not exist
    const int l0 = PMS ? b0 : unit.z;   // first label (dense) / proposal (PMS) of this pass
    const int aoff = PMS ? 0 : l0;      // column of l0 in the running-sum rows
    bool act[NH];
#pragma unroll
    for (int h = 0; h < NH; h++) act[h] = FULL || l0 + h * 64 + 2 * lane < lim;  // the lane's pair holds at least one real label
    // proposal mode: cost of this lane's two proposals at node vv (compute3DLabelCost, :103-118)
    auto pms_cost = [&](int vv) -> float2 {
        const int pix = V.node_pixel[vv];
        const int x = pix % A.img_w, y = pix / A.img_w;
        const int k = 2 * lane;
        float2 c = make_float2(0.f, 0.f);
        if (A.cost_mode) {  // slanted-plane colour + gradient cost at the sub-pixel match position (pm.cpp:97-154)
            const uint8_t* sb = V.img_self + 3 * (size_t)pix;
            const float* sg = V.grad_self + 2 * (size_t)pix;
            const uint8_t* ob = V.img_other + 3 * (size_t)y * A.img_w;
            const float* og = V.grad_other + 2 * (size_t)y * A.img_w;
            if (b0 + k < b_end) c.x = s3_plane_cost(sb, sg, ob, og, x, y, A.img_w, A.view, s_lab[3 * k], s_lab[3 * k + 1], s_lab[3 * k + 2], A.D, A.pm_alpha, A.pm_tau_c, A.pm_tau_g, A.pm_scale, A.oob);
            if (b0 + k + 1 < b_end) c.y = s3_plane_cost(sb, sg, ob, og, x, y, A.img_w, A.view, s_lab[3 * k + 3], s_lab[3 * k + 4], s_lab[3 * k + 5], A.D, A.pm_alpha, A.pm_tau_c, A.pm_tau_g, A.pm_scale, A.oob);
            return c;
        }
        const float* row = V.cost + (size_t)vv * Dp;
        if (b0 + k < b_end) c.x = s3_label_cost(row, s_lab[3 * k], s_lab[3 * k + 1], s_lab[3 * k + 2], x, y, A.D, A.oob);
        if (b0 + k + 1 < b_end) c.y = s3_label_cost(row, s_lab[3 * k + 3], s_lab[3 * k + 4], s_lab[3 * k + 5], x, y, A.D, A.oob);
        return c;
    };

    // ================================================================== leaf -> root
    {
        int v = top - gw;
        const char* nup_p = reinterpret_cast<const char*>(V.node_up + v);
        // FUSE: the matching cost of this lane's labels at node vv, straight from the two images (cost.cu: k_cost_adgrad; the
        // left volume reads the right image at x - d, the right volume the left image at x + d; labels without a
        // counterpart cost 3.0, row padding 0).  Issued where the cost row would be loaded: after the publish, in the
        // shadow of the wait for the next node's children.
        auto fused_cost = [&](int vv, float2* cf) {
            const int pix = __ldg(V.node_pixel + vv);
            const int x = pix - (int)(((unsigned long long)(unsigned)pix * A.w_magic) >> 40) * A.img_w;
            const int view = unit.x & 1;
            const uint2 me = __ldg(V.m_self + pix);
            const bool has_next = x + 1 < A.img_w;
            const float me_g = __uint_as_float(me.y);
            const float me_gn = has_next ? __uint_as_float(__ldg(V.m_self + pix + 1).y) : 0.0f;
            const int dvalid = view == 0 ? (has_next ? x + 1 : 0) : A.img_w - 1 - x;
            // nodes whose whole label slice has a counterpart in the other image (all but the columns next to one border):
            // no per-label validity
            if (FULL && (view == 0 ? (has_next && x >= l0 + 64 * NH - 1) : (x + l0 + 64 * NH < A.img_w))) {
#pragma unroll
                for (int h = 0; h < NH; h++) {
                    const int dA = l0 + h * 64 + 2 * lane;
                    // the three pixels q0 q1 q2 from two aligned 16-byte loads (lanes 16 bytes apart: dense wavefronts; the
                    // parity of the first pixel is the same for every lane of the node)
                    const int lo = view == 0 ? pix - dA - 1 : pix + dA;
                    const uint4* qq = reinterpret_cast<const uint4*>(V.m_other + (lo & ~1));
                    const uint4 u0 = __ldg(qq), u1 = __ldg(qq + 1);
                    const bool odd = lo & 1;
                    const uint2 q0 = odd ? make_uint2(u0.z, u0.w) : make_uint2(u0.x, u0.y);
                    const uint2 q1 = odd ? make_uint2(u1.x, u1.y) : make_uint2(u0.z, u0.w);
                    const uint2 q2 = odd ? make_uint2(u1.z, u1.w) : make_uint2(u1.x, u1.y);
                    const uint2 a = view == 0 ? q1 : q0, b = view == 0 ? q0 : q1;
                    const float an = __uint_as_float(view == 0 ? q2.y : q1.y), bn = __uint_as_float(view == 0 ? q1.y : q2.y);
                    cf[h].x = a3_adgrad(s_ct, view, me.x, me_g, me_gn, a.x, __uint_as_float(a.y), an);
                    cf[h].y = a3_adgrad(s_ct, view, me.x, me_g, me_gn, b.x, __uint_as_float(b.y), bn);
                }
                return;
            }
#pragma unroll
            for (int h = 0; h < NH; h++) {
                cf[h] = make_float2(0.f, 0.f);
                if (!act[h]) continue;
                const int dA = l0 + h * 64 + 2 * lane;
                const bool vA = dA < A.D && dA < dvalid, vB = dA + 1 < A.D && dA + 1 < dvalid;
                float cA = 3.0f, cB = 3.0f;  // bad_cost
                if (vA) {
                    // three consecutive pixels of the other view cover both labels and their right neighbours
                    const uint2* q = V.m_other + (view == 0 ? pix - dA - 1 : pix + dA);
                    const uint2 q1 = __ldg(q + 1);
                    uint2 q0 = make_uint2(0u, 0u), q2 = make_uint2(0u, 0u);
                    if (view == 0 ? vB : true) q0 = __ldg(q);
                    if (view == 0 ? true : vB) q2 = __ldg(q + 2);
                    const uint2 a = view == 0 ? q1 : q0, b = view == 0 ? q0 : q1;
                    const float an = __uint_as_float(view == 0 ? q2.y : q1.y), bn = __uint_as_float(view == 0 ? q1.y : q2.y);
                    cA = a3_adgrad(s_ct, view, me.x, me_g, me_gn, a.x, __uint_as_float(a.y), an);
                    if (vB) cB = a3_adgrad(s_ct, view, me.x, me_g, me_gn, b.x, __uint_as_float(b.y), bn);
                }
                if (dA >= A.D) cA = 0.0f;
                if (dA + 1 >= A.D) cB = 0.0f;
                cf[h] = make_float2(cA, cB);
            }
        };
        const char* cost_p = reinterpret_cast<const char*>(V.cost + (size_t)v * Dp + (PMS ? 0 : l0 + 2 * lane));
        char* aup_p = reinterpret_cast<char*>(aupT + (size_t)v * DA + aoff + 2 * lane);
        const char* aup_lane0 = reinterpret_cast<const char*>(aupT + aoff + 2 * lane);  // + c * DA * sizeof(T) for a far child
        // the node's record and cost row live in these registers from the end of the previous iteration (the loads are
        // issued right after the previous publish, so they are in flight during this warp's wait for the children)
        int4 nu = make_int4(0, 0, 0, 0);  // {child_begin, child_count, cw01, cw23}
        int par = 0;                      // CL > 1: the node's parent (how far the row has to travel)
        float2 cf[NH];
#pragma unroll
        for (int h = 0; h < NH; h++) cf[h] = make_float2(0.f, 0.f);
        if (v >= base) {
            nu = *reinterpret_cast<const int4*>(nup_p);
            if constexpr (CL > 1) par = V.node_dn[v].x;
            if constexpr (PMS) cf[0] = pms_cost(v);
            else if constexpr (FUSE) fused_cost(v, cf);
            else {
#pragma unroll
                for (int h = 0; h < NH; h++)
                    if (act[h]) cf[h] = *reinterpret_cast<const float2*>(cost_p + h * 256);
            }
        } else if (lane == 0)
            publish(base);  // a warp without nodes never holds anybody back
        int guard_ok = top + 1;  // writing ring row v is known to be safe for every v >= guard_ok
#if A3_INSTR
        long long tq = clock64(), q_pre = 0, q_poll = 0, q_dep = 0, q_guard = 0, q_pub = 0, q_post = 0; const long long t_up0 = tq; int n_nodes = 0, n_guard = 0;
#endif
        // ring row of v last held node v + R_E, which only nodes > v + R_E - NEAR_E may still read
        auto guard_up = [&]() {
            if (v < guard_ok && v + R_E <= top) {
#if A3_INSTR
                n_guard++;
#endif
                int m;
                while (true) {
                    m = prog_max();
                    if (m < v + R_E - NEAR_E + 1 + WT) break;
                    __nanosleep(sleep_ns);
                }
                guard_ok = m - (R_E - NEAR_E + WT);
            }
        };
        while (v >= base) {
#if A3_INSTR
            n_nodes++;
#endif
            const int vn = v - WT;
            const int cc = nu.y & 7, cb = nu.x;
            if constexpr (CL > 1) guard_up();  // a round trip through the cluster: taken before the wait for the children, not after it
            T2 acc[NH];
#pragma unroll
            for (int h = 0; h < NH; h++) acc[h] = TT::zero2();
            // (((0 + w3 A3) + w2 A2) + w1 A1) + w0 A0) + cost — children in reverse BFS order (Stereo3DMST.cpp:125-137)
#define A3_CHILD(K, IW)                                                                                              \
    if (cc > K) {                                                                                                    \
        const int c = cb + K;                                                                                        \
        const T wk = TT::ldsw(w_a + (uint32_t)sizeof(T) * (IW));                                                                 \
        const uint32_t pa = prog_of(top - c);                                                                        \
        A3_CLK(q_pre);                                                                                               \
        T2 cv[NH];                                                                                              \
        if (CL > 1 && c - v < NEAR_E) {                                                                              \
            /* another SM: the progress word and, right behind it, the row — ONE round trip through the cluster when */ \
            /* the child is done (requests of a warp to one CTA are serviced in order; a row fetched too early is    */ \
            /* simply fetched again) */                                                                              \
            const uint32_t ra = row_of(top - c);                                                                     \
            while (true) {                                                                                           \
                const int f = poll(pa);                                                                              \
                _Pragma("unroll") for (int h = 0; h < NH; h++) cv[h] = ld_row(ra + h * HB);                        \
                if (f <= c) break;                                                                                   \
            }                                                                                                        \
            A3_CLK(q_poll);                                                                                          \
        } else {                                                                                                     \
            while (poll(pa) > c) __nanosleep(sleep_ns);                                                              \
            A3_CLK(q_poll);                                                                                          \
            if (c - v < NEAR_E) {                                                                                    \
                const uint32_t ra = row_of(top - c);                                                                 \
                _Pragma("unroll") for (int h = 0; h < NH; h++) cv[h] = ld_row(ra + h * HB);                        \
            } else {                                                                                                 \
                const char* gp = aup_lane0 + (size_t)c * DA * sizeof(T);                                             \
                _Pragma("unroll") for (int h = 0; h < NH; h++)                                                       \
                    cv[h] = act[h] ? TT::ldcg2(gp + h * HB) : TT::zero2();                                           \
            }                                                                                                        \
        }                                                                                                            \
        _Pragma("unroll") for (int h = 0; h < NH; h++) {                                                             \
            acc[h].x = TT::add(acc[h].x, TT::mul(wk, cv[h].x));                                                      \
            acc[h].y = TT::add(acc[h].y, TT::mul(wk, cv[h].y));                                                      \
        }                                                                                                            \
    }
            A3_CHILD(3, (uint32_t)nu.w >> 16)
            A3_CHILD(2, (uint32_t)nu.w & 0xFFFFu)
            A3_CHILD(1, (uint32_t)nu.z >> 16)
            A3_CHILD(0, (uint32_t)nu.z & 0xFFFFu)
#undef A3_CHILD
#pragma unroll
            for (int h = 0; h < NH; h++) {
                acc[h].x = TT::add(acc[h].x, (T)cf[h].x);
                acc[h].y = TT::add(acc[h].y, (T)cf[h].y);
            }
            // a parent beyond the rings reads this row from L2: it has to be out before the publish
            const bool far_parent = CL > 1 ? (v - par >= NEAR_E) : (bool)(nu.y & S3_NU_FARPARENT);
            if (far_parent) {
#pragma unroll
                for (int h = 0; h < NH; h++)
                    if (act[h]) *reinterpret_cast<T2*>(aup_p + h * HB) = acc[h];
                if constexpr (CL > 1) a3_fence_cl();  // every lane's stores, before lane 0 publishes
            }
            A3_CLK(q_dep);
            if constexpr (CL == 1) guard_up();  // (measured: at the top it makes no difference for one CTA)
            A3_CLK(q_guard);
            {
                const uint32_t ra = ring_a + (uint32_t)(((top - v) >> LOGCL) & (A3_R - 1)) * ROWB;
#pragma unroll
                for (int h = 0; h < NH; h++) TT::sts2(ra + h * HB, acc[h]);
            }
            __syncwarp();
            if (lane == 0) publish(v);
            A3_CLK(q_pub);
            // Global accesses are issued only AFTER the publish: the release fence waits for every memory operation the warp
            // has in flight, and a load still on its way from HBM would put its latency on every level of the tree.
            const bool store_late = !far_parent;
            if (vn >= base) {  // next node of this warp: record and cost row
                nu = *reinterpret_cast<const int4*>(nup_p - (long long)WT * 16);
                if constexpr (CL > 1) par = V.node_dn[vn].x;
                if constexpr (PMS) cf[0] = pms_cost(vn);
                else if constexpr (FUSE) fused_cost(vn, cf);
                else {
#pragma unroll
                    for (int h = 0; h < NH; h++)
                        if (act[h]) cf[h] = *reinterpret_cast<const float2*>(cost_p - strideC + h * 256);
                }
            }
            // pull this warp's row of A3_PF rounds from now into L2 (one 128-byte line per lane)
            if (!FUSE && lane < (PMS ? (int)((Dp * 4 + 127) / 128) : NH * 2) && v - A3_PF * WT >= base && !(PMS && A.cost_mode))
                asm volatile("prefetch.global.L2 [%0];" ::"l"(cost_p - (PMS ? 0 : 2 * lane * 4) - A3_PF * strideC + lane * 128));
            if (store_late) {  // read back on the way down
#pragma unroll
                for (int h = 0; h < NH; h++)
                    if (act[h]) *reinterpret_cast<T2*>(aup_p + h * HB) = acc[h];
            }
            nup_p -= (long long)WT * 16;
            cost_p -= strideC;
            aup_p -= strideA;
            v = vn;
            A3_CLK(q_post);
        }
#if A3_INSTR
        if (blockIdx.x == 0 && lane == 0 && (w == 0 || w == 9)) printf("up W=%d warp %d: %d nodes (%d guards), cycles/node total %lld | pre %lld poll %lld dep %lld guard %lld publish %lld post %lld\n", W, w, n_nodes, n_guard, (clock64() - t_up0) / n_nodes, q_pre / n_nodes, q_poll / n_nodes, q_dep / n_nodes, q_guard / n_nodes, q_pub / n_nodes, q_post / n_nodes);
#endif
    }
    // every row of the way up is in L2 / HBM and every warp of the tree is done with the rings before the way down starts
    if constexpr (CL > 1) a3_cluster_sync(); else __syncthreads();
    if (tid < 32) s_prog[tid] = base - 1;  // down pass: node p is done iff its owner's word is >= p
    if constexpr (CL > 1) a3_cluster_sync(); else __syncthreads();

    // ================================================================== root -> leaf, WTA folded in
    {
        int v = base + gw;
        const char* ndn_p = reinterpret_cast<const char*>(V.node_dn + v);
        char* aup_p = reinterpret_cast<char*>(aupT + (size_t)v * DA + aoff + 2 * lane);
        const char* aup_lane0 = reinterpret_cast<const char*>(aupT + aoff + 2 * lane);
        int4 nd = make_int4(0, 0, 0, 0);  // {parent, parent weight, level | flags, pixel}
        int2 ch = make_int2(0, 0);        // CL > 1: {child_begin, child_count} (how far the final row has to travel)
        T2 au[NH];
#pragma unroll
        for (int h = 0; h < NH; h++) au[h] = TT::zero2();
        if (v < end) {
            nd = *reinterpret_cast<const int4*>(ndn_p);
            if constexpr (CL > 1) ch = *reinterpret_cast<const int2*>(V.node_up + v);
#pragma unroll
            for (int h = 0; h < NH; h++)  // (cluster: the row was written by another SM — read it from L2, never from this SM's L1)
                if (act[h]) au[h] = CL > 1 ? TT::ldcg2(aup_p + h * HB) : *reinterpret_cast<const T2*>(aup_p + h * HB);
        } else if (lane == 0)
            publish(end);
        int guard_ok = base - 1;  // writing ring row v is known to be safe for every v <= guard_ok
        // WTA over a node's labels held by this warp: strict '<', lowest label wins ties
        auto wta = [&](const T2* f, int vv, int pix) {
            T bc = TT::maxv();  // the oracle's initial best is DBL_MAX (a cost below it is required to win)
            int bd = 0x7fffffff;
#pragma unroll
            for (int h = 0; h < NH; h++) {
                const int lab = l0 + h * 64 + 2 * lane;
                if ((FULL || lab < lim) && f[h].x < bc) { bc = f[h].x; bd = lab; }
                if ((FULL || lab + 1 < lim) && f[h].y < bc) { bc = f[h].y; bd = lab + 1; }
            }
            double mc;
            const unsigned md = TT::warp_argmin(bc, bd, mc);
            if (PMS) {
                // label update (:173-185): proposals in list order, strict '<' == the first proposal attaining the minimum, if
                // it beats what the pixel holds (batches run in list order too)
                if (lane == 0 && md != 0x7fffffffu && mc < A.min_cost[pix]) {
                    A.min_cost[pix] = mc;
                    const float* l = s_lab + 3 * ((int)md - b0);
                    A.abc[3 * (size_t)pix] = l[0];
                    A.abc[3 * (size_t)pix + 1] = l[1];
                    A.abc[3 * (size_t)pix + 2] = l[2];
                }
            } else if (lane == 0) {
                if (A.n_slices == 1) {
                    V.disp[pix] = (int)md;
                    V.best[pix] = mc;
                } else {
                    V.pdisp[(size_t)slice * A.N + vv] = (int)md;
                    V.pbest[(size_t)slice * A.N + vv] = mc;
                }
            }
        };
        auto guard_dn = [&]() {
            if (v > guard_ok && v - R_E >= base) {
                int m;
                while (true) {
                    m = prog_min();
                    if (m > v - R_E + NEAR_E - 1 - WT) break;
                    __nanosleep(sleep_ns);
                }
                guard_ok = m + (R_E - NEAR_E + WT);
            }
        };
        while (v < end) {
            const int vn = v + WT;
            const int p = nd.x;
            if constexpr (CL > 1) guard_dn();
            T2 fin[NH];
            if (p != v) {
                // A[c] = w * A[parent] + (1 - w*w) * A_up[c]   (Stereo3DMST.cpp:155)
                const T wp = TT::ldsw(w_a + (uint32_t)sizeof(T) * (uint32_t)nd.y), wq = TT::ldsw(w_a + (uint32_t)sizeof(T) * (uint32_t)(S3_NUM_W + nd.y));
#pragma unroll
                for (int h = 0; h < NH; h++) {  // the half that does not depend on the parent, before the wait
                    au[h].x = TT::mul(wq, au[h].x);
                    au[h].y = TT::mul(wq, au[h].y);
                }
                const uint32_t pa = prog_of(p - base);
                T2 pv[NH];
                if (CL > 1 && v - p < NEAR_E) {  // word and row in one round trip through the cluster (see the way up)
                    const uint32_t ra = row_of(p - base);
                    while (true) {
                        const int f = poll(pa);
#pragma unroll
                        for (int h = 0; h < NH; h++) pv[h] = ld_row(ra + h * HB);
                        if (f >= p) break;
                    }
                } else {
                    while (poll(pa) < p) __nanosleep(sleep_ns);
                    if (v - p < NEAR_E) {
                        const uint32_t ra = row_of(p - base);
#pragma unroll
                        for (int h = 0; h < NH; h++) pv[h] = ld_row(ra + h * HB);
                    } else {
                        const char* gp = aup_lane0 + (size_t)p * DA * sizeof(T);
#pragma unroll
                        for (int h = 0; h < NH; h++) pv[h] = act[h] ? TT::ldcg2(gp + h * HB) : TT::zero2();
                    }
                }
#pragma unroll
                for (int h = 0; h < NH; h++) {
                    fin[h].x = TT::add(TT::mul(wp, pv[h].x), au[h].x);
                    fin[h].y = TT::add(TT::mul(wp, pv[h].y), au[h].y);
                }
            } else {
#pragma unroll
                for (int h = 0; h < NH; h++) fin[h] = au[h];  // the root keeps its leaf->root sum
            }
            // children further than the rings reach read the final value from L2
            const bool far_child = CL > 1 ? ((ch.y & 7) > 0 && ch.x + (ch.y & 7) - 1 - v >= NEAR_E) : (bool)(nd.z & S3_ND_FAR);
            if (far_child || (!PMS && A.keep)) {
#pragma unroll
                for (int h = 0; h < NH; h++)
                    if (act[h]) *reinterpret_cast<T2*>(aup_p + h * HB) = fin[h];
                if constexpr (CL > 1) a3_fence_cl();
            }
            if constexpr (CL == 1) guard_dn();
            {
                const uint32_t ra = ring_a + (uint32_t)(((v - base) >> LOGCL) & (A3_R - 1)) * ROWB;
#pragma unroll
                for (int h = 0; h < NH; h++) TT::sts2(ra + h * HB, fin[h]);
            }
            __syncwarp();
            if (lane == 0) publish(v);
            // next node's loads: after the publish (see the leaf->root pass), into the registers this node is done with
            const int pix = nd.w;
            if (vn < end) {
                nd = *reinterpret_cast<const int4*>(ndn_p + (long long)WT * 16);
                if constexpr (CL > 1) ch = *reinterpret_cast<const int2*>(V.node_up + vn);
#pragma unroll
                for (int h = 0; h < NH; h++)
                    if (act[h]) au[h] = CL > 1 ? TT::ldcg2(aup_p + strideA + h * HB) : *reinterpret_cast<const T2*>(aup_p + strideA + h * HB);
            }
            if (lane < NH * (int)sizeof(T) / 2 && v + A3_PF * WT < end)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(aup_p - 2 * lane * (int)sizeof(T) + A3_PF * strideA + lane * 128));
            // the WTA of this node runs while those loads are in flight and the next parent is still being computed
            wta(fin, v, pix);
            ndn_p += (long long)WT * 16;
            aup_p += strideA;
            v = vn;
        }
    }
    }  // passes
    // a CTA's shared memory must stay alive until no other CTA of the cluster can read it any more
    if constexpr (CL > 1) a3_cluster_sync();
}

static size_t agg3_smem_bytes(int NH, int R, size_t tsz) { return (2 * S3_NUM_W * tsz + 15) / 16 * 16 + (size_t)R * NH * 32 * 2 * tsz + 32 * sizeof(int); }

__global__ void k_wta_finish3(int N, int n_slices, const int4* __restrict__ node_dn, const int32_t* __restrict__ pdisp,
                              const double* __restrict__ pbest, int32_t* __restrict__ disp, double* __restrict__ best) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N) return;
    double bc = pbest[v];
    int bd = pdisp[v];
    for (int s = 1; s < n_slices; s++) {
        const double c = pbest[(size_t)s * N + v];
        const int d = pdisp[(size_t)s * N + v];
        if (c < bc || (c == bc && d < bd)) { bc = c; bd = d; }
    }
    const int pix = node_dn[v].w;
    disp[pix] = bd;
    best[pix] = bc;
}

#define A3_CLUSTER 8   // CTAs walking one giant tree
#ifndef A3_CL_NEAR
#define A3_CL_NEAR 32   // hand-over distance per CTA of a cluster (x A3_CLUSTER nodes for the cluster)
#endif

// trees of at least this many nodes are walked by a cluster / get a CTA of 32 warps (development overrides: S3_AGG_CL, S3_AGG_BIG)
static int s3_agg_cluster_nodes(const s3dmst_ctx* ctx) {
    static const int env = getenv("S3_AGG_CL") ? atoi(getenv("S3_AGG_CL")) : 0;
    const int p = ctx->P.agg_cluster_nodes;
    return p < 0 ? 0 : p > 0 ? p : env ? env : 32768;
}
// Trees of at least this many nodes get a CTA of 32 warps and an SM of their own; the others 16 warps, two trees per SM.
// Two small CTAs per SM hide each other's hand-over latencies — measured better whenever the launch has enough trees to
// fill the GPU (C4 batch of 8: 40.4 -> 36.5 ms with every tree on 16 warps) — but a tree that would be the launch's critical
// path on 16 warps needs the 32 (one C2 pair, whose 19 k-node tree is 1401 levels deep: 1.83 vs 2.93 ms; the FLIR pair: 6.7
// vs 8.1).  So: the big CTA for trees whose nodes exceed alpha x what one of the 2 x SMs small-CTA slots would get if the
// launch were spread perfectly.  alpha = 0.25 measured (0.25 / 0.5 / 1: C2 pair 1.85 / 1.88 / 2.02 ms, FLIR pair 6.66 / 7.55 /
// 8.17, C4 pair 5.02 / 5.79 / 4.64, C4 batch 36.8 / 36.5 / 36.5); development overrides: S3_AGG_BIG_ALPHA, S3_AGG_BIG (the
// threshold itself).  The three size classes are launched on three streams and run side by side.
static int s3_agg_big_nodes(const s3dmst_ctx* ctx, double unit_nodes_total) {
    static const int v = getenv("S3_AGG_BIG") ? atoi(getenv("S3_AGG_BIG")) : 0;
    static const double alpha = getenv("S3_AGG_BIG_ALPHA") ? atof(getenv("S3_AGG_BIG_ALPHA")) : 0.25;
    if (v > 0) return v;
    const double per_slot = unit_nodes_total / (2.0 * ctx->num_sms);
    return (int)std::max(256.0, std::min(1.0e9, alpha * per_slot));
}

template <typename K>
static int agg3_launch(s3dmst_ctx* ctx, K kernel, int grid, int threads, size_t smem, int cluster, cudaStream_t stream, const Agg3Args& A) {
    S3_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (cluster == 1) {
        kernel<<<grid, threads, smem, stream>>>(A);
    } else {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(threads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cluster;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        S3_CUDA(cudaLaunchKernelEx(&cfg, kernel, A));
    }
    S3_LAUNCH_CHECK();
    return 0;
}

// One launch over the trees of every view in `views_mask` (bit 0 = left, bit 1 = right) of every context in `ctxs`
// (all on one device, same image size and D): the unit list spans frames, so a batch of stereo pairs fills the GPU with
// independent trees instead of waiting on the deepest tree of a single pair.  Runs on ctxs[0]'s stream; the other
// contexts' streams are ordered before and after it with events.
// Returns 1 (and does nothing) if this kernel cannot serve the request, so the caller falls back to the simple one.
// fuse != 0: the matching cost is computed inside the up pass from the images (no cost volume is read; the views need
// images, forests and V.D / V.aup from s3_ensure_volume).
int s3_aggregate_flow_multi(s3dmst_ctx** ctxs, int nctx, int views_mask, int d0, int d1, int fuse) {
    s3dmst_ctx* ctx = ctxs[0];
    if (fuse && !ctx->P.exact) return s3_fail(ctx, S3DMST_E_ARG, "aggregate_dense: the fused matching cost needs the exact mode");
    int first = -1;
    for (int view = 0; view < 2 && first < 0; view++)
        if (views_mask & (1 << view)) first = view;
    if (first < 0) return s3_fail(ctx, S3DMST_E_ARG, "aggregate_dense: empty view mask");
    const int Dv = ctx->v[first].D, Dp = ctx->v[first].Dp;
    for (int c = 0; c < nctx; c++) {  // tree counts and sizes of every frame: the only thing the host waits for
        const int r = s3_forest_finish_host(ctxs[c]);
        if (r) return c == 0 ? r : s3_fail(ctx, r, "aggregate_dense: frame %d: %s", c, ctxs[c]->err.c_str());
    }
    for (int c = 0; c < nctx; c++) {
        if (ctxs[c]->device != ctx->device || ctxs[c]->N != ctx->N)
            return s3_fail(ctx, S3DMST_E_ARG, "aggregate_dense: batched contexts must share the device and the image size");
        for (int view = 0; view < 2; view++) {
            if (!(views_mask & (1 << view))) continue;
            View& V = ctxs[c]->v[view];
            if (!V.forest_ready || !(fuse ? V.aup != nullptr && V.D > 0 : V.cost_ready)) return s3_fail(ctx, S3DMST_E_STATE, "aggregate_dense: forest and cost volume required");
            if (V.D != Dv) return s3_fail(ctx, S3DMST_E_ARG, "aggregate_dense: views hold different D");
        }
    }
    if (d0 < 0 || d1 > Dv || d0 >= d1) return s3_fail(ctx, S3DMST_E_ARG, "aggregate_dense: need 0 <= d0 < d1 <= D");
    if (d0 & 3) return 1;  // 16-byte row alignment of the lane's label pairs
    const int nl = d1 - d0;
    static const int force_nh = getenv("S3_AGG_NH") ? atoi(getenv("S3_AGG_NH")) : 0;
    // (A tree is walked by one CTA per label slice and its time is proportional to its NODE count, so narrower slices do
    // not shorten the largest tree — measured on the FLIR pair: 17.4 -> 14.8 ms — they only add instructions.)
    const int NH = force_nh ? force_nh : (nl > 64 ? 2 : 1);
    const int SW = 64 * NH;
    const int n_slices = (nl + SW - 1) / SW;

    // unit list: (frame/view, tree, slice), longest (most nodes) tree first
    std::vector<std::pair<int, int4>> u;
    for (int c = 0; c < nctx; c++)
        for (int view = 0; view < 2; view++) {
            if (!(views_mask & (1 << view))) continue;
            View& V = ctxs[c]->v[view];
            for (int t = 0; t < V.T; t++)
                for (int s = 0; s < n_slices; s++)
                    u.push_back({V.h_tree_start[t + 1] - V.h_tree_start[t], make_int4(2 * c + view, t, d0 + s * SW, s)});
        }
    std::stable_sort(u.begin(), u.end(), [](const auto& a, const auto& b) { return a.first > b.first; });
    std::vector<int4> units(u.size());
    for (size_t i = 0; i < u.size(); i++) units[i] = u[i].second;

    // per-view tables and (for several slices) partial WTA results
    std::vector<Agg3View> table(2 * (size_t)nctx);
    memset(table.data(), 0, table.size() * sizeof(Agg3View));
    std::vector<int32_t*> pdisp(2 * (size_t)nctx, nullptr);
    std::vector<double*> pbest(2 * (size_t)nctx, nullptr);
    for (int c = 0; c < nctx; c++) {
        s3dmst_ctx* cx = ctxs[c];
        if (n_slices > 1) {
            const size_t per_view = (size_t)n_slices * cx->N * (sizeof(double) + sizeof(int32_t));
            const size_t need = 2 * per_view;
            if (cx->pms_scratch_cap < need) {
                if (cx->pms_scratch) S3_CUDA(cudaFree(cx->pms_scratch));
                cx->pms_scratch = nullptr; cx->pms_scratch_cap = 0;
                S3_CUDA(cudaMalloc(&cx->pms_scratch, need));
                cx->pms_scratch_cap = need;
            }
            for (int view = 0; view < 2; view++) {
                pbest[2 * c + view] = (double*)((char*)cx->pms_scratch + view * per_view);
                pdisp[2 * c + view] = (int32_t*)(pbest[2 * c + view] + (size_t)n_slices * cx->N);
            }
        }
        for (int view = 0; view < 2; view++) {
            View& V = cx->v[view];
            Agg3View& G = table[2 * c + view];
            G.tree_start = V.tree_start; G.node_up = V.node_up; G.node_dn = V.node_dn; G.node_pixel = V.node_pixel;
            G.cost = V.cost; G.aup = V.aup;
            G.disp = V.disp_i; G.best = V.best; G.pdisp = pdisp[2 * c + view]; G.pbest = pbest[2 * c + view];
        }
        if (fuse) {  // {BGR, gray} of both images, 8 bytes per pixel (on the frame's own stream, before the hand-over below)
            for (int view = 0; view < 2; view++) {
                View& V = cx->v[view];
                if (!V.match8) S3_CUDA(cudaMalloc(&V.match8, sizeof(uint2) * ((size_t)cx->N + 2)));
                k_pack_match<<<(cx->N + 255) / 256, 256, 0, cx->stream>>>(cx->N, V.raw4, V.gray, V.match8);
                S3_LAUNCH_CHECK();
            }
            for (int view = 0; view < 2; view++) {
                table[2 * c + view].m_self = cx->v[view].match8;
                table[2 * c + view].m_other = cx->v[view ^ 1].match8;
            }
        }
    }
    const size_t ubytes = units.size() * sizeof(int4), tbytes = (table.size() * sizeof(Agg3View) + 15) / 16 * 16;
    if (ctx->units_cap < ubytes + tbytes) {
        if (ctx->units_dev) S3_CUDA(cudaFree(ctx->units_dev));
        ctx->units_dev = nullptr; ctx->units_cap = 0;
        S3_CUDA(cudaMalloc(&ctx->units_dev, ubytes + tbytes));
        ctx->units_cap = ubytes + tbytes;
    }
    char* ubase = reinterpret_cast<char*>(ctx->units_dev);
    S3_TRY(s3_h2d_staged(ctx, ubase, table.data(), table.size() * sizeof(Agg3View)));
    S3_TRY(s3_h2d_staged(ctx, ubase + tbytes, units.data(), ubytes));
    // everything the other contexts have queued (their cost volumes) comes first
    for (int c = 1; c < nctx; c++) {
        S3_CUDA(cudaEventRecord(ctxs[c]->ev_xctx, ctxs[c]->stream));
        S3_CUDA(cudaStreamWaitEvent(ctx->stream, ctxs[c]->ev_xctx, 0));
    }

    Agg3Args A;
    memset(&A, 0, sizeof A);
    A.views = reinterpret_cast<const Agg3View*>(ubase);
    A.units = reinterpret_cast<const int4*>(ubase + tbytes);
    A.Dp = Dp; A.d1 = d1; A.N = ctx->N; A.n_slices = n_slices;
    const bool exact = ctx->P.exact != 0;
    A.lut_w = exact ? (const void*)ctx->lut_w : (const void*)ctx->lut_wf;
    A.lut_w2 = exact ? (const void*)ctx->lut_w2 : (const void*)ctx->lut_w2f;
    A.keep = ctx->P.keep_aggregated;
    A.img_w = ctx->W; A.D = Dv;
    A.w_magic = ((1ull << 40) + (unsigned long long)ctx->W - 1) / (unsigned long long)ctx->W;
    static const int sleep_env = getenv("S3_AGG_SLEEP") ? atoi(getenv("S3_AGG_SLEEP")) : 0;
    A.sleep_ns = sleep_env < 0 ? 0 : sleep_env ? sleep_env : 20;

    // The giant trees get a thread-block cluster (A3_CLUSTER CTAs = 256 warps on one tree), on the context's second
    // stream so that they run beside the rest; all but the smallest of the others get 32 warps and an SM of their own;
    // the rest 16 warps, two trees per SM.
    double unit_nodes_total = 0.0;
    for (const auto& e : u) unit_nodes_total += e.first;
    const int cl_nodes = s3_agg_cluster_nodes(ctx), big_nodes = s3_agg_big_nodes(ctx, unit_nodes_total);
    int n_cl = 0;
    while (cl_nodes > 0 && n_cl < (int)u.size() && u[n_cl].first >= cl_nodes) n_cl++;
    int n_big = n_cl;
    while (n_big < (int)u.size() && u[n_big].first >= big_nodes) n_big++;
    n_big -= n_cl;
    const int n_small = (int)u.size() - n_cl - n_big;
    const bool full = nl % SW == 0;  // every slice covers SW real labels
    S3_EV_BEGIN(S3DMST_T_AGG, first);
    cudaStream_t launch_stream = ctx->stream;
#define A3_LAUNCH_T(T_, NH_, FULL_, BIG_, R_, NEAR_, CL_, GRID_, THREADS_)                                                      \
    do {                                                                                                                       \
        const size_t smem = agg3_smem_bytes(NH_, R_, sizeof(T_));                                                              \
        S3_TRY(agg3_launch(ctx, k_agg_flow<T_, NH_, FULL_, BIG_, R_, NEAR_, false, CL_>, GRID_, THREADS_, smem, CL_, launch_stream, A)); \
    } while (0)
#define A3_LAUNCH_F(NH_, FULL_, BIG_, R_, NEAR_, CL_, GRID_, THREADS_)                                                          \
    do {                                                                                                                       \
        const size_t smem = agg3_smem_bytes(NH_, R_, sizeof(double));                                                          \
        S3_TRY(agg3_launch(ctx, k_agg_flow<double, NH_, FULL_, BIG_, R_, NEAR_, false, CL_, true>, GRID_, THREADS_, smem, CL_, launch_stream, A)); \
    } while (0)
#define A3_LAUNCH(NH_, FULL_, BIG_, R_, NEAR_, CL_, GRID_, THREADS_)                                                            \
    do {                                                                                                                       \
        if (fuse) A3_LAUNCH_F(NH_, FULL_, BIG_, R_, NEAR_, CL_, GRID_, THREADS_);                                              \
        else if (exact) A3_LAUNCH_T(double, NH_, FULL_, BIG_, R_, NEAR_, CL_, GRID_, THREADS_);                                \
        else A3_LAUNCH_T(float, NH_, FULL_, BIG_, (R_) * 2, NEAR_, CL_, GRID_, THREADS_);                                      \
    } while (0)
    // ring geometry: rows R and hand-over distance NEAR (>= S3_AGG_NEAR, the distance the forest stage flags nodes by).
    // A warp may not run more than (R - NEAR) / W rounds ahead of the slowest one, so R - NEAR >= ~2 W.
#define A3_DISPATCH(BIG_, RB_, NEARB_, CL_, GRID_, THREADS_)                                            \
    do {                                                                                               \
        if (NH == 2) {                                                                                 \
            if (full) A3_LAUNCH(2, true, BIG_, RB_ / 2, NEARB_, CL_, GRID_, THREADS_); else A3_LAUNCH(2, false, BIG_, RB_ / 2, NEARB_, CL_, GRID_, THREADS_); \
        } else {                                                                                       \
            if (full) A3_LAUNCH(1, true, BIG_, RB_, NEARB_, CL_, GRID_, THREADS_); else A3_LAUNCH(1, false, BIG_, RB_, NEARB_, CL_, GRID_, THREADS_); \
        }                                                                                              \
    } while (0)
    if (n_cl) {
        S3_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
        S3_CUDA(cudaStreamWaitEvent(ctx->stream_aux, ctx->ev_fork, 0));
        launch_stream = ctx->stream_aux;
        A.unit0 = 0;
        A3_DISPATCH(true, 256, A3_CL_NEAR, A3_CLUSTER, n_cl * A3_CLUSTER, 1024);
        S3_CUDA(cudaEventRecord(ctx->ev_join, ctx->stream_aux));
        launch_stream = ctx->stream;
    }
    // the three size classes run side by side (streams in one launch order would make the small trees wait for the last
    // big one while most SMs idle)
    if (n_big && n_small) {
        if (!n_cl) S3_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
        S3_CUDA(cudaStreamWaitEvent(ctx->stream_big, ctx->ev_fork, 0));
        launch_stream = ctx->stream_big;
    }
    if (n_big) {
        A.unit0 = n_cl;
        A3_DISPATCH(true, 256, 64, 1, n_big, 1024);   // 128 KB ring, one tree per SM
    }
    if (n_big && n_small) {
        S3_CUDA(cudaEventRecord(ctx->ev_join_big, ctx->stream_big));
        launch_stream = ctx->stream;
    }
    if (n_small) {
        A.unit0 = n_cl + n_big;
        A3_DISPATCH(false, 128, 32, 1, n_small, 512);  // 64 KB ring: two trees per SM
    }
    if (n_cl) S3_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    if (n_big && n_small) S3_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join_big, 0));
#undef A3_DISPATCH
#undef A3_LAUNCH
#undef A3_LAUNCH_F
#undef A3_LAUNCH_T
    if (n_slices > 1) {
        for (int c = 0; c < nctx; c++)
            for (int view = 0; view < 2; view++) {
                if (!(views_mask & (1 << view))) continue;
                View& V = ctxs[c]->v[view];
                k_wta_finish3<<<(ctx->N + 255) / 256, 256, 0, ctx->stream>>>(ctx->N, n_slices, V.node_dn, pdisp[2 * c + view], pbest[2 * c + view], V.disp_i, V.best);
                S3_LAUNCH_CHECK();
            }
    }
    S3_EV_END(S3DMST_T_AGG, first);
    // the other contexts continue (WTA results -> disparity maps) only after the joint launch
    if (nctx > 1) {
        S3_CUDA(cudaEventRecord(ctx->ev_xctx, ctx->stream));
        for (int c = 1; c < nctx; c++) S3_CUDA(cudaStreamWaitEvent(ctxs[c]->stream, ctx->ev_xctx, 0));
    }
    // No host synchronisation: the unit list went through the context's pinned staging; the callers queue the
    // post-processing behind this launch while it runs.
    for (int c = 0; c < nctx; c++)
        for (int view = 0; view < 2; view++)
            if (views_mask & (1 << view)) {
                ctxs[c]->v[view].agg_ready = true;
                ctxs[c]->v[view].agg_d0 = d0;
                ctxs[c]->v[view].agg_d1 = d1;
            }
    return 0;
}

int s3_aggregate_flow(s3dmst_ctx* ctx, int views_mask, int d0, int d1, int fuse) { return s3_aggregate_flow_multi(&ctx, 1, views_mask, d0, d1, fuse); }

// Proposal mode of the dataflow kernel.  Plan: the unit list (trees with work, longest first) and the view table, uploaded
// once; launch: one evaluation of a tree-grouped proposal list (labels_dev [n][3] grouped by tree, order inside a tree
// preserved; prop_off_dev [T+1]) on one view, optionally followed by the in-kernel refinement ladder (gen).
// h_prop_off = the offsets on the host, or nullptr for "every tree" (the generator: every tree refines).
// scratch_dev: N * 64 doubles.  Returns 1 if this kernel cannot serve the request (the caller then uses the simple one).
int s3_pms_flow_plan(s3dmst_ctx* ctx, int view, const int* h_prop_off, double* scratch_dev, PmsFlowPlan* plan) {
    View& V = ctx->v[view];
    if (!ctx->P.exact) return 1;  // proposals are always evaluated in the reference's arithmetic
    S3_TRY(s3_forest_finish_host(ctx));
    std::vector<std::pair<int, int4>> u;
    for (int t = 0; t < V.T; t++)
        if (!h_prop_off || h_prop_off[t + 1] > h_prop_off[t]) u.push_back({V.h_tree_start[t + 1] - V.h_tree_start[t], make_int4(0, t, 0, 0)});
    memset(plan, 0, sizeof *plan);
    if (u.empty()) return 0;
    std::stable_sort(u.begin(), u.end(), [](const auto& a, const auto& b) { return a.first > b.first; });
    std::vector<int4> units(u.size());
    for (size_t i = 0; i < u.size(); i++) units[i] = u[i].second;
    Agg3View G;
    memset(&G, 0, sizeof G);
    G.tree_start = V.tree_start; G.node_up = V.node_up; G.node_dn = V.node_dn; G.node_pixel = V.node_pixel;
    G.cost = V.cost; G.aup = scratch_dev;
    G.img_self = V.bgr; G.img_other = ctx->v[view ^ 1].bgr; G.grad_self = V.pgrad; G.grad_other = ctx->v[view ^ 1].pgrad;
    const size_t ubytes = units.size() * sizeof(int4), tbytes = (sizeof(Agg3View) + 15) / 16 * 16;
    if (ctx->units_cap < ubytes + tbytes) {
        if (ctx->units_dev) S3_CUDA(cudaFree(ctx->units_dev));
        ctx->units_dev = nullptr; ctx->units_cap = 0;
        S3_CUDA(cudaMalloc(&ctx->units_dev, ubytes + tbytes));
        ctx->units_cap = ubytes + tbytes;
    }
    char* ubase = reinterpret_cast<char*>(ctx->units_dev);
    S3_TRY(s3_h2d_staged(ctx, ubase, &G, sizeof G));
    S3_TRY(s3_h2d_staged(ctx, ubase + tbytes, units.data(), ubytes));
    double unit_nodes_total = 0.0;
    for (const auto& e : u) unit_nodes_total += e.first;
    const int cl_nodes = s3_agg_cluster_nodes(ctx), big_nodes = s3_agg_big_nodes(ctx, unit_nodes_total);
    int n_cl = 0;
    while (cl_nodes > 0 && n_cl < (int)u.size() && u[n_cl].first >= cl_nodes) n_cl++;
    int n_big = n_cl;
    while (n_big < (int)u.size() && u[n_big].first >= big_nodes) n_big++;
    plan->n_cl = n_cl; plan->n_big = n_big - n_cl; plan->n_small = (int)u.size() - n_big;
    plan->views_dev = ubase; plan->units_dev = ubase + tbytes;
    return 0;
}

int s3_pms_flow_launch(s3dmst_ctx* ctx, int view, const PmsFlowPlan* plan, const int* prop_off_dev, const float* labels_dev, int gen, uint32_t seed, uint32_t round) {
    View& V = ctx->v[view];
    const int n_cl = plan->n_cl, n_big = plan->n_big, n_small = plan->n_small;
    if (n_cl + n_big + n_small == 0) return 0;
    Agg3Args A;
    memset(&A, 0, sizeof A);
    A.views = reinterpret_cast<const Agg3View*>(plan->views_dev);
    A.units = reinterpret_cast<const int4*>(plan->units_dev);
    A.Dp = V.Dp; A.d1 = 0; A.N = ctx->N; A.n_slices = 1;
    A.lut_w = ctx->lut_w; A.lut_w2 = ctx->lut_w2;
    A.keep = 0;
    A.sleep_ns = 20;
    A.prop_off = prop_off_dev; A.labels = labels_dev; A.min_cost = V.min_cost; A.abc = V.abc;
    A.img_w = ctx->W; A.D = V.D; A.oob = ctx->P.oob_cost;
    A.gen = gen; A.seed = seed; A.round = round; A.refine_floor = ctx->P.refine_floor;
    A.cost_mode = ctx->P.pms_cost_mode == 1; A.view = view; A.img_h = ctx->H;
    A.pm_alpha = ctx->P.pm_alpha; A.pm_tau_c = ctx->P.pm_tau_c; A.pm_tau_g = ctx->P.pm_tau_g; A.pm_scale = ctx->P.cost_scale;
    S3_EV_BEGIN(S3DMST_T_PMS, view);
    if (n_cl) {
        S3_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
        S3_CUDA(cudaStreamWaitEvent(ctx->stream_aux, ctx->ev_fork, 0));
        A.unit0 = 0;
        S3_TRY(agg3_launch(ctx, k_agg_flow<double, 1, false, true, 256, A3_CL_NEAR, true, A3_CLUSTER>, n_cl * A3_CLUSTER, 1024, agg3_smem_bytes(1, 256, sizeof(double)), A3_CLUSTER, ctx->stream_aux, A));
        S3_CUDA(cudaEventRecord(ctx->ev_join, ctx->stream_aux));
    }
    const bool split = n_big && n_small;  // the size classes run side by side (see s3_aggregate_flow_multi)
    if (split) {
        if (!n_cl) S3_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
        S3_CUDA(cudaStreamWaitEvent(ctx->stream_big, ctx->ev_fork, 0));
    }
    if (n_big) {
        A.unit0 = n_cl;
        S3_TRY(agg3_launch(ctx, k_agg_flow<double, 1, false, true, 256, 64, true, 1>, n_big, 1024, agg3_smem_bytes(1, 256, sizeof(double)), 1, split ? ctx->stream_big : ctx->stream, A));
        if (split) S3_CUDA(cudaEventRecord(ctx->ev_join_big, ctx->stream_big));
    }
    if (n_small) {
        A.unit0 = n_cl + n_big;
        S3_TRY(agg3_launch(ctx, k_agg_flow<double, 1, false, false, 128, 32, true, 1>, n_small, 512, agg3_smem_bytes(1, 128, sizeof(double)), 1, ctx->stream, A));
    }
    if (n_cl) S3_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    if (split) S3_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join_big, 0));
    S3_EV_END(S3DMST_T_PMS, view);
    return 0;
}
