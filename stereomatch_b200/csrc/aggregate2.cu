// stereomatch_b200/csrc/aggregate2.cu — tree-filter aggregation + WTA, bulk-copy (TMA) pipelined kernel (default).
//
// Same arithmetic, layout and work decomposition as aggregate.cu (see the header there): one CTA per
// (tree, label slice), level-synchronous walk, FP64 in the reference's association order
// (src/Stereo3DMST.cpp:120-158, :173-185).  What changes is how bytes reach the math.
//
// The simple kernel is bound by its critical path: per tree level a chain of dependent global loads
// (level offsets -> node record -> child rows), ~3.8 us per level on B200, i.e. 11 ms for a 1400-level tree
// while the HBM time of the whole volume is 0.4 ms.  Measured facts that shaped this version (ncu source
// view + clock64 counters, profiles/r01_*): a single warp retires a dependent instruction every ~5-10
// cycles, so what matters is the NUMBER of instructions every warp executes between two level barriers.
//   * The tile sequence (<= 16 consecutive nodes of one level per tile) is precomputed at forest-build time
//     as 32-byte descriptors in both traversal orders; a 64-entry descriptor ring in shared memory is kept
//     filled by bulk copies, so no warp walks level tables or touches global memory to find its work.
//   * One "DMA" warp stages, per tile, the node records and the cost rows (up pass) / running-sum rows (down
//     pass) into a shared-memory ring with ONE cp.async.bulk (TMA, UBLKCP) each — BFS order makes a tile a
//     contiguous byte range — and arrives at the tile barrier only when the NEXT tile has landed
//     (mbarrier complete_tx).
//   * 16 math warps, one node per warp per tile, software pipelined: right after barrier k they issue the
//     shared-memory loads of the children/parent values of tile k, then pre-load everything of tile k+1
//     that does not depend on tile k (descriptor, node record, weights, cost, w2*A_up), then run the short
//     FP64 chain of tile k.  A warp without a node in the tile does ~10 instructions.
//   * Both views' trees are scheduled in one launch, longest first.
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "hd_math.h"
#include "internal.h"

#ifndef A2_INSTRUMENT
#define A2_INSTRUMENT 0  // 1: clock64 counters of CTA 0 (S3_DEBUG_AGG=1 prints them); costs registers
#endif
#if A2_INSTRUMENT
#define A2_CLK() clock64()
#else
#define A2_CLK() 0ll
#endif
#define A2_MAXT 16  // tiles in flight (mbarrier slots)
#define A2_DR 64    // descriptor ring entries: two halves of 32
#define A2_NP 2     // prep warps (tile k is prepared by prep warp k % A2_NP)
#define A2_ND 4     // DMA warps  (tile k is staged by DMA warp k % A2_ND)

struct Agg2View {
    const int* tree_start;
    const int* tree_ntiles;
    const int4* tile_desc;
    const NodeUp* node_up;
    const int4* node_dn;
    const float* cost;
    double* aup;
    int32_t* disp;  // pixel order when n_slices == 1, else [slice][node] partials
    double* best;
};

struct Agg2Args {
    Agg2View v[2];
    const uint32_t* units;  // (view << 31) | tree, longest first
    int n_slices, Dp, d0, d1, N;
    const double* lut_w;
    const double* lut_w2;
    int cap, rn, keep, use_tma;
    long long* dbg;  // optional cycle counters of CTA 0 (development aid)
};

// ---- PTX helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.b32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > (1u << 26)) __trap();  // never hang the GPU: a lost copy becomes an error
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src_gmem), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ double2 lds_d2(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_d2(uint32_t a, double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ int4 lds_i4(uint32_t a) {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds_d(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}

// total order on doubles as unsigned 64-bit keys (handles negative costs of the mc-cnn "fast" volumes)
__device__ __forceinline__ unsigned long long dkey(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dkey_inv(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// descriptor ring slot of tile k of a pass whose first chunk has global chunk number gbase
__device__ __forceinline__ uint32_t desc_addr(uint32_t s_desc, int gbase, int k) {
    return s_desc + (uint32_t)(((((gbase + (k >> 5)) & 1) << 5) | (k & 31)) * 32);
}

template <int HV>
__global__ void __launch_bounds__(32 * (S3_TILE_NODES + A2_NP + A2_ND), 1) k_agg_dense2(Agg2Args A) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    constexpr int NW = S3_TILE_NODES;  // chain warps == nodes per tile; then A2_NP prep warps, then A2_ND DMA warps
    constexpr int NALL = 32 * (NW + A2_NP + A2_ND);
    constexpr int PREC = 80;           // bytes of one prepared node record
    constexpr uint32_t NOSLOT = 0x7fffffffu;
    constexpr int SW = 64 * HV;
    constexpr int ROWL = HV * 512;  // bytes of one node in the level hand-over buffers
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cap = A.cap, RN = A.rn;

    const uint32_t unit_code = A.units[blockIdx.x / A.n_slices];
    const int slice = blockIdx.x % A.n_slices;
    const Agg2View& V = A.v[unit_code >> 31];
    const int t = (int)(unit_code & 0x7fffffffu);
    const int tbase = V.tree_start[t];   // node offset == tile index offset of the tree
    const int NT = V.tree_ntiles[t];
    const int nch = (NT + 31) >> 5;      // descriptor chunks per pass
    const int lab0 = A.d0 + slice * SW;  // first label of this slice (multiple of 4)
    const int seg = min(SW, A.Dp - lab0);  // labels staged per row (multiple of 4)
    const size_t Dp = A.Dp;
    const bool contig = seg == A.Dp;  // slice covers whole rows: a tile's rows are one byte range
    const uint32_t rowb_up = (uint32_t)seg * 4u, rowb_dn = (uint32_t)seg * 8u;

    // ---- shared memory carve-up (mirrored by agg2_smem_bytes on the host)
    const uint32_t s_base = smem_u32(s_raw);
    const uint32_t s_full = s_base;                                     // [A2_MAXT] tile mbarriers
    const uint32_t s_dfull = s_base + 128;                              // [2] descriptor-half mbarriers
    const uint32_t s_desc = s_base + 256;                               // [A2_DR][32]
    const uint32_t s_meta = s_desc + A2_DR * 32;                        // [RN][16]
    const uint32_t s_rows = s_meta + (uint32_t)RN * 16u;                // [RN][SW*8]
    const uint32_t s_lvl = s_rows + (uint32_t)RN * (uint32_t)(SW * 8);  // [2][cap][ROWL]
    const uint32_t s_w = s_lvl + 2u * (uint32_t)cap * ROWL;             // [S3_NUM_W] exp(-w*gamma)
    const uint32_t s_w2 = s_w + S3_NUM_W * 8u;                          // [S3_NUM_W] 1 - w*w
    const uint32_t s_prep = s_w2 + S3_NUM_W * 8u;                       // [2][NW][PREC] prepared node records
    {
        double* gw = reinterpret_cast<double*>(s_raw + (s_w - s_base));
        for (int i = tid; i < S3_NUM_W; i += NALL) {
            gw[i] = A.lut_w[i];
            gw[S3_NUM_W + i] = A.lut_w2[i];
        }
    }
    if (tid == 0) {
        for (int i = 0; i < A2_MAXT; i++) mbar_init(s_full + 8 * i, A.use_tma ? (contig ? 1 : 33) : 32);
        mbar_init(s_dfull, 1);
        mbar_init(s_dfull + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int tend = tbase + V.tree_start[t + 1] - tbase;  // one past the tree's last node
    const int RNM = RN - 1;                                 // RN is a power of two
    // ring entry of the first node of a tile: tiles take consecutive entries in traversal order, so it is the
    // number of nodes traversed before the tile, modulo the ring size — a function of the descriptor alone
    auto tile_ent0 = [&](int pass, const int4& dA, const int4& dB) -> int {
        return (pass == 0 ? (tend - dB.x) + dA.z : dA.x - tbase) & RNM;
    };
    if (warp >= NW + A2_NP) {
        // =========================================================== DMA warps (stateless schedule)
        // Tile x is staged by DMA warp x % A2_ND as soon as the ring has room for it: tiles take consecutive ring
        // entries in traversal order, so "room" is  end(x) - start(k) <= RN  with k the tile the chain warps are on
        // (both are functions of the descriptors alone — no shared bookkeeping between the DMA warps).  The same
        // warp makes sure tile x has landed before barrier x - 1.
        const int me = warp - (NW + A2_NP);
        for (int pass = 0; pass < 2; pass++) {
            const int gbase = pass * nch;  // global chunk number of this pass's first descriptor chunk
            const int qbase = pass * NT;   // global tile number of this pass's first tile
            const int4* dsrc = V.tile_desc + 2 * ((size_t)(pass == 0 ? A.N : 0) + (size_t)tbase);
            const uint32_t rowbytes = pass == 0 ? rowb_up : rowb_dn;
            const size_t gstride = Dp * (pass == 0 ? 4 : 8);
            const unsigned char* mbase = pass == 0 ? reinterpret_cast<const unsigned char*>(V.node_up)
                                                   : reinterpret_cast<const unsigned char*>(V.node_dn);
            const unsigned char* rbase = pass == 0 ? reinterpret_cast<const unsigned char*>(V.cost + lab0)
                                                   : reinterpret_cast<const unsigned char*>(V.aup + lab0);
            auto load_chunk = [&](int c) {
                if (lane == 0) {
                    const int g = gbase + c;
                    const uint32_t bar = s_dfull + 8u * (uint32_t)(g & 1);
                    const uint32_t bytes = (uint32_t)min(32, NT - 32 * c) * 32u;
                    mbar_expect_tx(bar, bytes);
                    bulk_g2s(s_desc + (uint32_t)(g & 1) * 1024u, dsrc + 64 * (size_t)c, bytes, bar);
                }
            };
            int chunk_ok = -1;  // last descriptor chunk this warp has seen land
            auto tile_pref = [&](const int4& dA, const int4& dB) -> int {  // nodes traversed before the tile
                return pass == 0 ? (tend - dB.x) + dA.z : dA.x - tbase;
            };
            auto issue = [&](int x, const int4& dA, const int4& dB) {  // stage tile x (x % A2_ND == me)
                const int n = dA.y, t0 = dA.x;
                const int ent0 = tile_ent0(pass, dA, dB);
                const int qg = qbase + x;
                const uint32_t bar = s_full + 8u * (uint32_t)(qg % A2_MAXT);
                const unsigned char* msrc = mbase + (size_t)t0 * 16;
                const unsigned char* rsrc = rbase + (size_t)t0 * gstride;
                const int n1 = min(n, RN - ent0);  // entries before the ring wraps
                if (lane == 0) {
                    mbar_expect_tx(bar, (uint32_t)n * 16u + (contig ? (uint32_t)n * rowbytes : 0u));
                    bulk_g2s(s_meta + (uint32_t)ent0 * 16u, msrc, (uint32_t)n1 * 16u, bar);
                    if (contig) bulk_g2s(s_rows + (uint32_t)ent0 * rowbytes, rsrc, (uint32_t)n1 * rowbytes, bar);
                    if (n1 < n) {
                        bulk_g2s(s_meta, msrc + (size_t)n1 * 16, (uint32_t)(n - n1) * 16u, bar);
                        if (contig) bulk_g2s(s_rows, rsrc + (size_t)n1 * gstride, (uint32_t)(n - n1) * rowbytes, bar);
                    }
                }
                if (!contig) {
                    // label-sliced rows are strided in HBM: 16-byte cp.async chunks, one node per iteration
                    const int cpr = (int)(rowbytes >> 4);
                    for (int i = 0; i < n; i++) {
                        const int e = (ent0 + i) & RNM;
                        for (int c = lane; c < cpr; c += 32)
                            cp_async16(s_rows + (uint32_t)e * rowbytes + (uint32_t)c * 16u, rsrc + (size_t)i * gstride + (size_t)c * 16);
                    }
                    cp_async_arrive_noinc(bar);
                }
            };
            auto wait_tile = [&](int x) {
                const int qg = qbase + x;
                mbar_wait(s_full + 8u * (uint32_t)(qg % A2_MAXT), (uint32_t)((qg / A2_MAXT) & 1));
            };
            if (me == 0) {
                load_chunk(0);
                if (nch > 1) load_chunk(1);
            }
            int nx = me;  // my next tile to stage
            auto try_issue = [&](int k) {  // k: tile the chain warps are on (its entries and everything after are in use)
                while (nx < NT && nx - k < A2_MAXT) {
                    if ((nx >> 5) != chunk_ok) {  // first touch of a descriptor chunk: wait until it has landed
                        const int g = gbase + (nx >> 5);
                        mbar_wait(s_dfull + 8u * (uint32_t)(g & 1), (uint32_t)((g >> 1) & 1));
                        chunk_ok = nx >> 5;
                    }
                    const uint32_t dx = desc_addr(s_desc, gbase, nx), dk = desc_addr(s_desc, gbase, k);
                    const int4 xA = lds_i4(dx), xB = lds_i4(dx + 16);
                    const int4 kA = lds_i4(dk), kB = lds_i4(dk + 16);
                    if (tile_pref(xA, xB) + xA.y - tile_pref(kA, kB) > RN) break;
                    issue(nx, xA, xB);
                    nx += A2_ND;
                }
            };
            try_issue(0);
            if (me == 0) wait_tile(0);
            named_bar_sync(1, NALL);  // B_init: tile 0 of the pass has landed
            named_bar_sync(1, NALL);  // B_init2: tile 0 is prepared
            for (int k = 0; k < NT; k++) {
                if (k + 1 < NT && (k + 1) % A2_ND == me) wait_tile(k + 1);  // the prep warps read tile k+1 after B_k
                named_bar_sync(1, NALL);  // B_k: the chain warps have finished tile k-1
                if (me == 0 && k > 0 && (k & 31) == 0 && (k >> 5) + 1 < nch) load_chunk((k >> 5) + 1);
                try_issue(k);
            }
            named_bar_sync(1, NALL);  // pass boundary: all chain warps done (and, after pass 0, fenced)
        }
        return;
    }

    const uint32_t lane16 = (uint32_t)lane * 16u;
    if (warp >= NW) {
        // =========================================================== prep warps: lane i prepares node i of the NEXT tile
        // Everything about a node that is scalar (the same for all of its labels) is decoded here, one tile
        // ahead of the chain warps and one node per LANE: child/parent slots in the hand-over buffer, edge
        // weights as doubles, ring offsets, global row addresses.  The chain warps then read 48-80 ready bytes.
        //   up:   R0 {code, own slot, cost row, child0 slot}  R1 {w0, w1}  R2 {row addr lo, hi, child1 slot, child2 slot}
        //         R3 {w2, w3}  R4 {child3 slot, child_begin, -, -}
        //   down: R0 {code, own slot, A_up row, parent slot}  R1 {w, 1-w*w}  R2 {row addr lo, hi, pixel, node}
        //         R3 {parent row addr lo, hi, -, -}
        //   code: bits 0-2 child count (up) / bit 0 root (down); bit 8 = some operand is outside the shared-memory
        //         window (slow path); bits 16-17 tile flags; bit 31 = no node for this warp in the tile
        const int pme = warp - NW;
        for (int pass = 0; pass < 2; pass++) {
            const int gbase = pass * nch;
            auto prep = [&](int k) {
                if (k % A2_NP != pme) return;
                const uint32_t da = desc_addr(s_desc, gbase, k);
                const int4 dA = lds_i4(da);       // {t0, n, loff, flags}
                const int4 dB = lds_i4(da + 16);  // {level end, parent level start, 0, 0}
                const int n = dA.y;
                const int e = (tile_ent0(pass, dA, dB) + lane) & RNM;
                const uint32_t out = s_prep + (uint32_t)((k & 1) * NW + lane) * PREC;
                if (lane < NW) {
                    const int v = dA.x + lane, li = dA.z + lane;
                    const int tf = dA.w << 16;
                    int4 r0, r2, r4 = make_int4(0, 0, 0, 0);
                    double w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0;
                    const uint32_t li_off = li < cap ? (uint32_t)li * ROWL : NOSLOT;
                    const unsigned long long ga = (unsigned long long)(V.aup + (size_t)v * Dp + lab0);
                    if (lane >= n) {
                        r0 = make_int4((int)0x80000000u | tf, 0, 0, 0);
                        r2 = make_int4(0, 0, 0, 0);
                    } else if (pass == 0) {
                        int4 nu = lds_i4(s_meta + (uint32_t)e * 16u);  // {child_begin, child_count, cw01, cw23}
                        nu.y &= 7;  // bit 8 is the dataflow kernel's far-parent flag
                        const int j0 = nu.x - dB.x;                          // children live in the next level
                        w0 = lds_d(s_w + 8u * ((uint32_t)nu.z & 0xFFFFu));
                        w1 = lds_d(s_w + 8u * ((uint32_t)nu.z >> 16));
                        w2 = lds_d(s_w + 8u * ((uint32_t)nu.w & 0xFFFFu));
                        w3 = lds_d(s_w + 8u * ((uint32_t)nu.w >> 16));
                        const int slow = (nu.y > 0 && j0 + nu.y > cap) ? 0x100 : 0;
                        r0 = make_int4(nu.y | slow | tf, (int)li_off, (int)((uint32_t)e * rowb_up), j0 * ROWL);
                        r2 = make_int4((int)(unsigned)ga, (int)(unsigned)(ga >> 32), (j0 + 1) * ROWL, (j0 + 2) * ROWL);
                        r4 = make_int4((j0 + 3) * ROWL, nu.x, j0, 0);
                    } else {
                        const int4 nd = lds_i4(s_meta + (uint32_t)e * 16u);  // {parent, pw, level, pixel}
                        const int j = nd.x - dB.y;
                        w0 = lds_d(s_w + 8u * (uint32_t)nd.y);
                        w1 = lds_d(s_w2 + 8u * (uint32_t)nd.y);
                        const bool root = nd.x == v;
                        const int slow = (!root && j >= cap) ? 0x100 : 0;
                        r0 = make_int4((root ? 1 : 0) | slow | tf, (int)li_off, (int)((uint32_t)e * rowb_dn), j * ROWL);
                        r2 = make_int4((int)(unsigned)ga, (int)(unsigned)(ga >> 32), nd.w, v);
                        const unsigned long long gp = (unsigned long long)(V.aup + (size_t)nd.x * Dp + lab0);
                        r4 = make_int4((int)(unsigned)gp, (int)(unsigned)(gp >> 32), 0, 0);
                    }
                    asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(out), "r"(r0.x), "r"(r0.y), "r"(r0.z), "r"(r0.w) : "memory");
                    sts_d2(out + 16, make_double2(w0, w1));
                    asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(out + 32), "r"(r2.x), "r"(r2.y), "r"(r2.z), "r"(r2.w) : "memory");
                    if (pass == 0) {
                        sts_d2(out + 48, make_double2(w2, w3));
                        asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(out + 64), "r"(r4.x), "r"(r4.y), "r"(r4.z), "r"(r4.w) : "memory");
                    } else
                        asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(out + 48), "r"(r4.x), "r"(r4.y), "r"(r4.z), "r"(r4.w) : "memory");
                }
            };
            named_bar_sync(1, NALL);  // B_init
            prep(0);
            named_bar_sync(1, NALL);  // B_init2
            for (int k = 0; k < NT; k++) {
                named_bar_sync(1, NALL);  // B_k: tile k+1 has landed
                if (k + 1 < NT) prep(k + 1);
            }
            named_bar_sync(1, NALL);  // pass boundary
        }
        return;
    }

    // =============================================================== chain warps: one node per warp per tile
    const int w_ = warp;
    const uint32_t lane8 = (uint32_t)lane * 8u;
    const bool full = lab0 + SW <= A.d1;  // every lane owns 2*HV real labels: no per-lane predication needed
    bool act[HV];
#pragma unroll
    for (int h = 0; h < HV; h++) act[h] = lab0 + h * 64 + 2 * lane < A.d1;
    uint32_t cur = s_lvl, prev = s_lvl + (uint32_t)cap * ROWL;
    long long m_bar = 0, m_work = 0, mc0 = A2_CLK(), mc1;
    long long seg_a = 0, seg_b = 0, seg_c = 0, seg_n = 0;

    // ------------------------------------------------------------------ leaf -> root
    named_bar_sync(1, NALL);  // B_init
    named_bar_sync(1, NALL);  // B_init2
    for (int k = 0; k < NT; k++) {
        mc1 = A2_CLK(); m_work += mc1 - mc0;
        named_bar_sync(1, NALL);  // B_k: previous tile's sums are visible, tile k is prepared
        mc0 = A2_CLK(); m_bar += mc0 - mc1;
        const uint32_t pr = s_prep + (uint32_t)((k & 1) * NW + w_) * PREC;
        const int4 r0 = lds_i4(pr);  // {code, own slot, cost row, child0 slot}
        if (r0.x >= 0) {
            const double2 w01 = lds_d2(pr + 16);
            const int4 r2 = lds_i4(pr + 32);  // {row addr lo, hi, child1 slot, child2 slot}
            const int cc = r0.x & 7;
            double2 acc[HV];
#pragma unroll
            for (int h = 0; h < HV; h++) acc[h] = make_double2(0.0, 0.0);
            float2 cf[HV];
#pragma unroll
            for (int h = 0; h < HV; h++) cf[h] = lds_f2(s_rows + (uint32_t)r0.z + (uint32_t)h * 256u + lane8);
            // (((0 + w3 A3) + w2 A2) + w1 A1) + w0 A0) + cost — children in reverse BFS order (Stereo3DMST.cpp:125-137)
            if (!(r0.x & 0x100)) {
                // fast path: all children in the shared-memory window; straight-line code per child count
                const uint32_t pb = prev + lane16;
                if (cc == 1) {
                    double2 c0[HV];
#pragma unroll
                    for (int h = 0; h < HV; h++) c0[h] = lds_d2(pb + (uint32_t)r0.w + (uint32_t)h * 512u);
#pragma unroll
                    for (int h = 0; h < HV; h++) {
                        acc[h].x = S3_DADD(0.0, S3_DMUL(w01.x, c0[h].x));
                        acc[h].y = S3_DADD(0.0, S3_DMUL(w01.x, c0[h].y));
                    }
                } else if (cc == 2) {
                    double2 c0[HV], c1[HV];
#pragma unroll
                    for (int h = 0; h < HV; h++) {
                        c0[h] = lds_d2(pb + (uint32_t)r0.w + (uint32_t)h * 512u);
                        c1[h] = lds_d2(pb + (uint32_t)r2.z + (uint32_t)h * 512u);
                    }
#pragma unroll
                    for (int h = 0; h < HV; h++) {
                        acc[h].x = S3_DADD(S3_DADD(0.0, S3_DMUL(w01.y, c1[h].x)), S3_DMUL(w01.x, c0[h].x));
                        acc[h].y = S3_DADD(S3_DADD(0.0, S3_DMUL(w01.y, c1[h].y)), S3_DMUL(w01.x, c0[h].y));
                    }
                } else if (cc >= 3) {
                    const double2 w23 = lds_d2(pr + 48);
                    const int off3 = lds_i4(pr + 64).x;
                    double2 c0[HV], c1[HV], c2[HV], c3[HV];
#pragma unroll
                    for (int h = 0; h < HV; h++) {
                        c0[h] = lds_d2(pb + (uint32_t)r0.w + (uint32_t)h * 512u);
                        c1[h] = lds_d2(pb + (uint32_t)r2.z + (uint32_t)h * 512u);
                        c2[h] = lds_d2(pb + (uint32_t)r2.w + (uint32_t)h * 512u);
                        c3[h] = cc == 4 ? lds_d2(pb + (uint32_t)off3 + (uint32_t)h * 512u) : make_double2(0.0, 0.0);
                    }
#pragma unroll
                    for (int h = 0; h < HV; h++) {
                        double ax = 0.0, ay = 0.0;
                        if (cc == 4) {
                            ax = S3_DADD(ax, S3_DMUL(w23.y, c3[h].x));
                            ay = S3_DADD(ay, S3_DMUL(w23.y, c3[h].y));
                        }
                        ax = S3_DADD(ax, S3_DMUL(w23.x, c2[h].x));
                        ay = S3_DADD(ay, S3_DMUL(w23.x, c2[h].y));
                        ax = S3_DADD(ax, S3_DMUL(w01.y, c1[h].x));
                        ay = S3_DADD(ay, S3_DMUL(w01.y, c1[h].y));
                        acc[h].x = S3_DADD(ax, S3_DMUL(w01.x, c0[h].x));
                        acc[h].y = S3_DADD(ay, S3_DMUL(w01.x, c0[h].y));
                    }
                }
            } else {
                // slow path (levels wider than the window): children beyond it come from HBM/L2
                const double2 w23 = lds_d2(pr + 48);
                const int4 r4 = lds_i4(pr + 64);  // {child3 slot, child_begin, first child's index in its level, -}
                for (int c = cc - 1; c >= 0; --c) {
                    const double w = c == 0 ? w01.x : c == 1 ? w01.y : c == 2 ? w23.x : w23.y;
#pragma unroll
                    for (int h = 0; h < HV; h++) {
                        double2 cv = make_double2(0.0, 0.0);
                        if (act[h])
                            cv = r4.z + c < cap ? lds_d2(prev + (uint32_t)(r4.z + c) * ROWL + (uint32_t)h * 512u + lane16)
                                                : *reinterpret_cast<const double2*>(V.aup + (size_t)(r4.y + c) * Dp + lab0 + h * 64 + 2 * lane);
                        acc[h].x = S3_DADD(acc[h].x, S3_DMUL(w, cv.x));
                        acc[h].y = S3_DADD(acc[h].y, S3_DMUL(w, cv.y));
                    }
                }
            }
            double* grow = reinterpret_cast<double*>(((unsigned long long)(unsigned)r2.y << 32) | (unsigned)r2.x) + 2 * lane;
#pragma unroll
            for (int h = 0; h < HV; h++) {
                acc[h].x = S3_DADD(acc[h].x, (double)cf[h].x);
                acc[h].y = S3_DADD(acc[h].y, (double)cf[h].y);
                if (full || act[h]) {
                    if ((uint32_t)r0.y != NOSLOT) sts_d2(cur + (uint32_t)r0.y + (uint32_t)h * 512u + lane16, acc[h]);
                    *reinterpret_cast<double2*>(grow + h * 64) = acc[h];
                }
            }
        }
        if (r0.x & (S3_TF_LAST << 16)) { const uint32_t tmp = cur; cur = prev; prev = tmp; }
    }
    fence_proxy_async();      // the running sums written above are read back by bulk copies in pass 2
    named_bar_sync(1, NALL);  // pass boundary
    const long long m_bar_up = m_bar, m_work_up = m_work;

    // ------------------------------------------------------------------ root -> leaf, WTA folded in
    named_bar_sync(1, NALL);  // B_init
    named_bar_sync(1, NALL);  // B_init2
    for (int k = 0; k < NT; k++) {
        mc1 = A2_CLK(); m_work += mc1 - mc0;
        named_bar_sync(1, NALL);  // B_k
        mc0 = A2_CLK(); m_bar += mc0 - mc1;
        const uint32_t pr = s_prep + (uint32_t)((k & 1) * NW + w_) * PREC;
        const int4 r0 = lds_i4(pr);  // {code, own slot, A_up row, parent slot}
        if (r0.x >= 0) {
            const double2 ww = lds_d2(pr + 16);  // {w, 1 - w*w}
            const int4 r2 = lds_i4(pr + 32);     // {row addr lo, hi, pixel, node}
            const bool root = r0.x & 1;
            double2 pv[HV], au[HV], fin[HV];
#pragma unroll
            for (int h = 0; h < HV; h++) au[h] = lds_d2(s_rows + (uint32_t)r0.z + (uint32_t)h * 512u + lane16);
            if (!(r0.x & 0x100)) {
#pragma unroll
                for (int h = 0; h < HV; h++) pv[h] = lds_d2(prev + (root ? 0u : (uint32_t)r0.w) + (uint32_t)h * 512u + lane16);
            } else {
                const int4 r3 = lds_i4(pr + 48);  // {parent row addr lo, hi, -, -}
                const double* gpar = reinterpret_cast<const double*>(((unsigned long long)(unsigned)r3.y << 32) | (unsigned)r3.x) + 2 * lane;
#pragma unroll
                for (int h = 0; h < HV; h++) pv[h] = act[h] ? *reinterpret_cast<const double2*>(gpar + h * 64) : make_double2(0.0, 0.0);
            }
            double* grow = reinterpret_cast<double*>(((unsigned long long)(unsigned)r2.y << 32) | (unsigned)r2.x) + 2 * lane;
            double bc = DBL_MAX;  // the oracle's initial best (cost < DBL_MAX is required to win)
            int bd = 0x7fffffff;
#pragma unroll
            for (int h = 0; h < HV; h++) {
                // A[c] = w * A[parent] + (1 - w*w) * A_up[c]   (Stereo3DMST.cpp:155); the root keeps its leaf->root sum
                const double fx = S3_DADD(S3_DMUL(ww.x, pv[h].x), S3_DMUL(ww.y, au[h].x));
                const double fy = S3_DADD(S3_DMUL(ww.x, pv[h].y), S3_DMUL(ww.y, au[h].y));
                fin[h] = root ? au[h] : make_double2(fx, fy);
                if (full || act[h]) {
                    if ((uint32_t)r0.y != NOSLOT) sts_d2(cur + (uint32_t)r0.y + (uint32_t)h * 512u + lane16, fin[h]);
                    if ((uint32_t)r0.y == NOSLOT || A.keep) *reinterpret_cast<double2*>(grow + h * 64) = fin[h];
                    const int l0 = lab0 + h * 64 + 2 * lane;
                    if (fin[h].x < bc) { bc = fin[h].x; bd = l0; }
                    if ((full || l0 + 1 < A.d1) && fin[h].y < bc) { bc = fin[h].y; bd = l0 + 1; }
                }
            }
            // warp arg-min with three 32-bit REDUX steps: (cost hi, cost lo, label); ties -> lowest label
            const unsigned long long key = dkey(bc);
            const unsigned khi = (unsigned)(key >> 32), klo = (unsigned)key;
            const unsigned mhi = __reduce_min_sync(0xffffffffu, khi);
            const unsigned mlo = __reduce_min_sync(0xffffffffu, khi == mhi ? klo : 0xffffffffu);
            const unsigned md = __reduce_min_sync(0xffffffffu, (khi == mhi && klo == mlo) ? (unsigned)bd : 0x7fffffffu);
            if (lane == 0) {
                const double mc = dkey_inv(((unsigned long long)mhi << 32) | mlo);
                if (A.n_slices == 1) {
                    V.disp[r2.z] = (int)md;
                    V.best[r2.z] = mc;
                } else {
                    V.disp[(size_t)slice * A.N + r2.w] = (int)md;
                    V.best[(size_t)slice * A.N + r2.w] = mc;
                }
            }
        }
        if (r0.x & (S3_TF_LAST << 16)) { const uint32_t tmp = cur; cur = prev; prev = tmp; }
    }
    named_bar_sync(1, NALL);  // pass boundary (matches the DMA warp's)
    if (A.dbg && blockIdx.x == 0 && tid == 0) {
        A.dbg[4] = m_bar_up; A.dbg[5] = m_work_up; A.dbg[6] = m_bar - m_bar_up; A.dbg[7] = m_work - m_work_up;
        A.dbg[8] = seg_a; A.dbg[9] = seg_b; A.dbg[10] = seg_c; A.dbg[11] = seg_n;
    }
}

static size_t agg2_smem_bytes(int HV, int rn, int cap) {
    const size_t SW = 64 * HV;
    return 256 + A2_DR * 32 + (size_t)rn * 16 + (size_t)rn * SW * 8 + 2 * (size_t)cap * HV * 512 + 2 * S3_NUM_W * sizeof(double) +
           2 * S3_TILE_NODES * 80 + 16;
}

__global__ void k_wta_finish2(int N, int n_slices, const int* __restrict__ node_pixel, const int32_t* __restrict__ pdisp,
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
    const int pix = node_pixel[v];
    disp[pix] = bd;
    best[pix] = bc;
}

// views_mask: bit 0 = left, bit 1 = right.  Both views must hold volumes of the same D.
// Returns 1 (and does nothing) if this kernel cannot serve the request, so the caller falls back to the simple one.
int s3_aggregate_dense2(s3dmst_ctx* ctx, int views_mask, int d0, int d1) {
    int nviews = 0, first = -1;
    for (int view = 0; view < 2; view++) {
        if (!(views_mask & (1 << view))) continue;
        View& V = ctx->v[view];
        if (!V.forest_ready || !V.cost_ready) return s3_fail(ctx, S3DMST_E_STATE, "aggregate_dense: forest and cost volume required");
        if (first < 0) first = view;
        if (V.D != ctx->v[first].D) return s3_fail(ctx, S3DMST_E_ARG, "aggregate_dense: views hold different D");
        nviews++;
    }
    if (!nviews) return s3_fail(ctx, S3DMST_E_ARG, "aggregate_dense: empty view mask");
    const int Dv = ctx->v[first].D, Dp = ctx->v[first].Dp;
    if (d0 < 0 || d1 > Dv || d0 >= d1) return s3_fail(ctx, S3DMST_E_ARG, "aggregate_dense: need 0 <= d0 < d1 <= D");
    if (d0 & 3) return 1;
    const int nl = d1 - d0;
    const int HV = nl > 64 ? 2 : 1;
    const int SW = 64 * HV;
    const int n_slices = (nl + SW - 1) / SW;
    const int NW = S3_TILE_NODES;
    const int cap = std::max(NW, ctx->P.agg_cache_nodes > 0 ? ctx->P.agg_cache_nodes : 16);
    int rn = 64;  // ring entries: power of two, >= 3 tiles
    while (rn < 3 * NW || rn < ctx->P.agg_ring_nodes) rn *= 2;
    const size_t smem = agg2_smem_bytes(HV, rn, cap);
    if (smem > 227 * 1024) return 1;

    // unit list: trees of the requested views, longest (most nodes) first
    std::vector<std::pair<int, uint32_t>> u;
    for (int view = 0; view < 2; view++) {
        if (!(views_mask & (1 << view))) continue;
        View& V = ctx->v[view];
        for (int t = 0; t < V.T; t++) u.push_back({V.h_tree_start[t + 1] - V.h_tree_start[t], ((uint32_t)view << 31) | (uint32_t)t});
    }
    std::stable_sort(u.begin(), u.end(), [](const auto& a, const auto& b) { return a.first > b.first; });
    std::vector<uint32_t> units(u.size());
    for (size_t i = 0; i < u.size(); i++) units[i] = u[i].second;
    if (ctx->units_cap < units.size()) {
        if (ctx->units_dev) S3_CUDA(cudaFree(ctx->units_dev));
        ctx->units_dev = nullptr; ctx->units_cap = 0;
        S3_CUDA(cudaMalloc(&ctx->units_dev, units.size() * sizeof(uint32_t)));
        ctx->units_cap = units.size();
    }
    S3_CUDA(cudaMemcpyAsync(ctx->units_dev, units.data(), units.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));

    Agg2Args A;
    memset(&A, 0, sizeof A);
    int32_t* pdisp[2] = {nullptr, nullptr};
    double* pbest[2] = {nullptr, nullptr};
    if (n_slices > 1) {
        const size_t per_view = (size_t)n_slices * ctx->N * (sizeof(double) + sizeof(int32_t));
        const size_t need = 2 * per_view;
        if (ctx->pms_scratch_cap < need) {
            if (ctx->pms_scratch) S3_CUDA(cudaFree(ctx->pms_scratch));
            ctx->pms_scratch = nullptr; ctx->pms_scratch_cap = 0;
            S3_CUDA(cudaMalloc(&ctx->pms_scratch, need));
            ctx->pms_scratch_cap = need;
        }
        for (int view = 0; view < 2; view++) {
            pbest[view] = (double*)((char*)ctx->pms_scratch + view * per_view);
            pdisp[view] = (int32_t*)(pbest[view] + (size_t)n_slices * ctx->N);
        }
    }
    for (int view = 0; view < 2; view++) {
        View& V = ctx->v[view];
        Agg2View& G = A.v[view];
        G.tree_start = V.tree_start; G.tree_ntiles = V.tree_ntiles; G.tile_desc = V.tile_desc;
        G.node_up = V.node_up; G.node_dn = V.node_dn;
        G.cost = V.cost; G.aup = V.aup;
        G.disp = n_slices == 1 ? V.disp_i : pdisp[view];
        G.best = n_slices == 1 ? V.best : pbest[view];
    }
    A.units = ctx->units_dev;
    A.n_slices = n_slices; A.Dp = Dp; A.d0 = d0; A.d1 = d1; A.N = ctx->N;
    A.lut_w = ctx->lut_w; A.lut_w2 = ctx->lut_w2;
    A.cap = cap; A.rn = rn; A.keep = ctx->P.keep_aggregated;
    A.use_tma = getenv("S3_AGG_TMA") ? atoi(getenv("S3_AGG_TMA")) : 1;
    const bool dbg = getenv("S3_DEBUG_AGG") != nullptr;
    if (dbg) S3_CUDA(cudaMalloc(&A.dbg, 16 * sizeof(long long)));

    const int grid = (int)units.size() * n_slices;
    const int threads = 32 * (NW + A2_NP + A2_ND);
    S3_EV_BEGIN(S3DMST_T_AGG, first);
    if (HV == 2) {
        S3_CUDA(cudaFuncSetAttribute(k_agg_dense2<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_agg_dense2<2><<<grid, threads, smem, ctx->stream>>>(A);
    } else {
        S3_CUDA(cudaFuncSetAttribute(k_agg_dense2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_agg_dense2<1><<<grid, threads, smem, ctx->stream>>>(A);
    }
    S3_LAUNCH_CHECK();
    if (n_slices > 1) {
        for (int view = 0; view < 2; view++) {
            if (!(views_mask & (1 << view))) continue;
            View& V = ctx->v[view];
            k_wta_finish2<<<(ctx->N + 255) / 256, 256, 0, ctx->stream>>>(ctx->N, n_slices, V.node_pixel, pdisp[view], pbest[view],
                                                                         V.disp_i, V.best);
            S3_LAUNCH_CHECK();
        }
    }
    S3_EV_END(S3DMST_T_AGG, first);
    S3_CUDA(cudaStreamSynchronize(ctx->stream));  // `units` (host vector) is read by the async copy above
    if (dbg) {
        long long h[16];
        S3_CUDA(cudaMemcpy(h, A.dbg, sizeof h, cudaMemcpyDeviceToHost));
        fprintf(stderr, "[agg2 cta0] dma: issue %lld wait %lld bar %lld tiles %lld | math w0: up bar %lld work %lld, down bar %lld work %lld (cycles)\n",
                h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
        fprintf(stderr, "[agg2 cta0] chain w0 up: loads %lld math+sts %lld stg %lld nodes %lld | dma lane0: expect_tx %lld bulk(meta) %lld bulk(rows) %lld\n", h[8], h[9], h[10], h[11], h[12], h[13], h[14]);
        cudaFree(A.dbg);
    }
    for (int view = 0; view < 2; view++)
        if (views_mask & (1 << view)) {
            ctx->v[view].agg_ready = true;
            ctx->v[view].agg_d0 = d0;
            ctx->v[view].agg_d1 = d1;
        }
    return 0;
}
