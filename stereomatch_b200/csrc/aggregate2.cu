// stereomatch_b200/csrc/aggregate2.cu — tree-filter aggregation + WTA, TMA-pipelined version (the default).
//
// Same arithmetic, layout and work decomposition as aggregate.cu (see the header there): one CTA per
// (tree, label slice), level-synchronous walk, FP64 in the reference's association order
// (src/Stereo3DMST.cpp:120-158, :173-185).  What changes is how bytes reach the math.  The v1 kernel is
// bound by its critical path: every level pays a chain of dependent global loads (level offsets -> node
// record -> child rows), ~3.8 us per level on B200, i.e. 11 ms for a 1400-level tree while the HBM time
// of the whole volume is 0.4 ms.  Here
//   * one extra "DMA" warp per CTA walks the tree AHEAD of the math warps and stages, per node, the node
//     record (16 B) and its cost row (up pass) / running-sum row (down pass) into a shared-memory ring with
//     cp.async.bulk (TMA bulk copies, mbarrier complete_tx); BFS order makes every tile a contiguous node
//     range, so the addresses are known without touching the data;
//   * the DMA warp joins the per-tile named barrier only after the next tile's bytes have landed, so the
//     math warps never wait on an mbarrier or on global memory on the critical path: per level it is
//     LDS(node record) -> LDS(children / parent values) -> FP64 chain -> STS -> BAR;
//   * both views' trees are scheduled in ONE launch (longest trees first), so the two deepest trees'
//     critical paths overlap.
// Ring entries are per node, tiles take consecutive entries, so narrow levels let the DMA warp run many
// levels ahead (prefetch distance is bounded by bytes, not by level count).
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "hd_math.h"
#include "internal.h"

#define A2_MAXT 16           // tiles in flight (header + mbarrier slots)
#define A2_FLAG_LEVEL_END 1
#define A2_FLAG_LAST 2

struct Agg2View {
    const int* tree_start;
    const int* tree_depth;
    const int* lvl_start;
    const NodeUp* node_up;
    const int4* node_dn;
    const int* node_pixel;
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
    int cap, rn, tn, keep;
    long long* dbg;  // optional cycle counters of CTA 0 (development aid)
};

// ---- PTX helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.b32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > (1u << 26)) __trap();  // never hang the GPU: a lost copy becomes an error
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// total order on doubles as unsigned 64-bit keys (handles negative costs of the mc-cnn "fast" volumes)
__device__ __forceinline__ unsigned long long dkey(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dkey_inv(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

struct TileHdr {  // 32 bytes
    int t0, n, ent0, loff;
    int aux, flags, pad0, pad1;
};

template <int HV>
__global__ void __launch_bounds__(1024, 1) k_agg_dense2(Agg2Args A) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NW = (blockDim.x >> 5) - 1;  // math warps; warp NW is the DMA warp
    const int nall = blockDim.x;
    const int cap = A.cap, RN = A.rn, TN = A.tn;
    constexpr int SW = 64 * HV;
    constexpr int ROWB = SW * 8;  // ring row stride in bytes (down-pass rows are doubles)

    // ---- shared memory carve-up (mirrored by agg2_smem_bytes on the host)
    uint64_t* s_full = reinterpret_cast<uint64_t*>(s_raw);                       // [A2_MAXT]
    TileHdr* s_hdr = reinterpret_cast<TileHdr*>(s_raw + 128);                    // [A2_MAXT]
    unsigned char* s_meta = s_raw + 128 + A2_MAXT * 32;                          // [RN][16]
    unsigned char* s_rows = s_meta + (size_t)RN * 16;                            // [RN][ROWB]
    double2* s_lvl = reinterpret_cast<double2*>(s_rows + (size_t)RN * ROWB);     // [2][cap][HV][32]
    double* s_w = reinterpret_cast<double*>(s_lvl + 2 * (size_t)cap * HV * 32);  // [S3_NUM_W] exp(-w*gamma)
    double* s_w2 = s_w + S3_NUM_W;                                               // [S3_NUM_W] 1 - w*w

    const uint32_t unit_code = A.units[blockIdx.x / A.n_slices];
    const int slice = blockIdx.x % A.n_slices;
    const Agg2View& V = A.v[unit_code >> 31];
    const int t = (int)(unit_code & 0x7fffffffu);
    const int base = V.tree_start[t];
    const int* lvl = V.lvl_start + base + t;
    const int depth = V.tree_depth[t];
    const int lab0 = A.d0 + slice * SW;                     // first label of this slice (multiple of 4)
    const int seg = min(SW, A.Dp - lab0);                   // labels copied per row (multiple of 4)
    const size_t Dp = A.Dp;
    const bool contig = seg == A.Dp;  // the slice covers whole rows: a tile's rows are one contiguous range

    for (int i = tid; i < S3_NUM_W; i += blockDim.x) {
        s_w[i] = A.lut_w[i];
        s_w2[i] = A.lut_w2[i];
    }
    if (tid == 0) {
        for (int i = 0; i < A2_MAXT; i++) mbar_init(s_full + i, contig ? 1 : 33);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == NW) {
        // =========================================================== DMA warp
        int q = 0, iq = 0;  // consumed / issued tile counters (continue across the two passes: mbarrier parity)
        long long t_issue = 0, t_wait = 0, t_bar = 0, n_tiles = 0;
        for (int pass = 0; pass < 2; pass++) {
            int inflight = 0, ent_next = 0, n_prev = 0;
            const int qbase = q;
            // issue-side iterator
            int L = pass == 0 ? depth - 1 : 0, off = 0;
            int ls = __ldg(lvl + L), le = __ldg(lvl + L + 1);
            int ps = 0;  // parent level start (down pass)
            bool it_done = false;
            const uint32_t rowbytes = (uint32_t)seg * (pass == 0 ? 4u : 8u);
            auto try_issue = [&]() {
                while (!it_done && iq - q < A2_MAXT) {
                    const int t0 = ls + off;
                    const int n = min(TN, le - t0);
                    // a tile takes consecutive ring entries and never wraps (so it is ONE contiguous copy);
                    // entries skipped at the end of the ring stay accounted to the tile until it is released
                    int ent0 = ent_next, skip = 0;
                    if (ent0 + n > RN) { skip = RN - ent0; ent0 = 0; }
                    if (inflight + skip + n > RN) break;
                    const bool level_end = t0 + n >= le;
                    const bool last = level_end && (pass == 0 ? L == 0 : L == depth - 1);
                    const int slot = iq % A2_MAXT;
                    const unsigned char* msrc = pass == 0 ? reinterpret_cast<const unsigned char*>(V.node_up + t0)
                                                          : reinterpret_cast<const unsigned char*>(V.node_dn + t0);
                    const unsigned char* rsrc = pass == 0
                        ? reinterpret_cast<const unsigned char*>(V.cost + (size_t)t0 * Dp + lab0)
                        : reinterpret_cast<const unsigned char*>(V.aup + (size_t)t0 * Dp + lab0);
                    if (lane == 0) {
                        TileHdr h;
                        h.t0 = t0; h.n = n; h.ent0 = ent0; h.loff = off;
                        h.aux = pass == 0 ? le : ps;
                        h.flags = (level_end ? A2_FLAG_LEVEL_END : 0) | (last ? A2_FLAG_LAST : 0);
                        h.pad0 = skip + n; h.pad1 = 0;
                        s_hdr[slot] = h;
                        // TMA bulk copies: the tile's node records, and (rows contiguous) all its rows at once
                        mbar_expect_tx(s_full + slot, (uint32_t)n * 16u + (contig ? (uint32_t)n * rowbytes : 0u));
                        bulk_g2s(s_meta + (size_t)ent0 * 16, msrc, (uint32_t)n * 16u, s_full + slot);
                        if (contig) bulk_g2s(s_rows + (size_t)ent0 * rowbytes, rsrc, (uint32_t)n * rowbytes, s_full + slot);
                    }
                    if (!contig) {
                        // label-sliced rows are strided in HBM: 16-byte cp.async chunks, one node per iteration
                        const int cpr = (int)(rowbytes >> 4);
                        const size_t rstride = Dp * (pass == 0 ? 4 : 8);
                        for (int i = 0; i < n; i++)
                            for (int k = lane; k < cpr; k += 32)
                                cp_async16(s_rows + (size_t)(ent0 + i) * rowbytes + (size_t)k * 16, rsrc + (size_t)i * rstride + (size_t)k * 16);
                        cp_async_arrive_noinc(s_full + slot);
                    }
                    __syncwarp();
                    ent_next = ent0 + n;
                    if (ent_next >= RN) ent_next = 0;
                    inflight += skip + n;
                    iq++;
                    // advance
                    if (level_end) {
                        if (last) {
                            it_done = true;
                        } else if (pass == 0) {
                            L--; le = ls; ls = __ldg(lvl + L); off = 0;
                        } else {
                            L++; ps = ls; ls = le; le = __ldg(lvl + L + 1); off = 0;
                        }
                    } else
                        off += n;
                }
            };
            long long c0 = clock64();
            try_issue();
            long long c1 = clock64();
            t_issue += c1 - c0;
            while (true) {
                c0 = clock64();
                mbar_wait(s_full + (q % A2_MAXT), (uint32_t)((q / A2_MAXT) & 1));
                c1 = clock64();
                t_wait += c1 - c0;
                named_bar_sync(1, nall);  // B_q: tile q landed; math warps finished tile q-1
                c0 = clock64();
                t_bar += c0 - c1;
                n_tiles++;
                const TileHdr& h = s_hdr[q % A2_MAXT];
                const bool last = h.flags & A2_FLAG_LAST;
                const int n_cur = h.pad0;  // entries held by the tile (incl. skipped ring tail)
                if (q > qbase) inflight -= n_prev;
                n_prev = n_cur;
                q++;
                if (last) break;
                c0 = clock64();
                try_issue();
                c1 = clock64();
                t_issue += c1 - c0;
            }
            named_bar_sync(1, nall);  // pass boundary: all math warps done (and, after pass 0, fenced)
        }
        if (A.dbg && blockIdx.x == 0 && lane == 0) {
            A.dbg[0] = t_issue; A.dbg[1] = t_wait; A.dbg[2] = t_bar; A.dbg[3] = n_tiles;
        }
        return;
    }

    // =============================================================== math warps
    int lab[HV];
    bool act[HV];
#pragma unroll
    for (int h = 0; h < HV; h++) {
        lab[h] = lab0 + h * 64 + 2 * lane;
        act[h] = lab[h] < A.d1;
    }
    double2* cur = s_lvl;
    double2* prev = s_lvl + (size_t)cap * HV * 32;
    int q = 0;
    long long m_bar = 0, m_work = 0, mc0 = clock64(), mc1;

    // ------------------------------------------------------------------ leaf -> root
    while (true) {
        mc1 = clock64(); m_work += mc1 - mc0;
        named_bar_sync(1, nall);
        mc0 = clock64(); m_bar += mc0 - mc1;
        const TileHdr h = s_hdr[q % A2_MAXT];
        const int le = h.aux;
        for (int i = warp; i < h.n; i += NW) {
            const int ent = h.ent0 + i;
            const NodeUp nu = *reinterpret_cast<const NodeUp*>(s_meta + (size_t)ent * 16);
            const float2* crow = reinterpret_cast<const float2*>(s_rows + (size_t)ent * (seg * 4));
            const int v = h.t0 + i, li = h.loff + i;
            double2 acc[HV];
#pragma unroll
            for (int hh = 0; hh < HV; hh++) acc[hh] = make_double2(0.0, 0.0);
            for (int k = nu.child_count - 1; k >= 0; --k) {
                const int ch = nu.child_begin + k;
                const int j = ch - le;
                const uint32_t iw = ((k & 2) ? nu.cw23 : nu.cw01) >> ((k & 1) * 16) & 0xFFFFu;
                const double w = s_w[iw];
#pragma unroll
                for (int hh = 0; hh < HV; hh++) {
                    if (!act[hh]) continue;
                    double2 cv;
                    if (j < cap)
                        cv = prev[((size_t)j * HV + hh) * 32 + lane];
                    else
                        cv = *reinterpret_cast<const double2*>(V.aup + (size_t)ch * Dp + lab[hh]);
                    acc[hh].x = S3_DADD(acc[hh].x, S3_DMUL(w, cv.x));
                    acc[hh].y = S3_DADD(acc[hh].y, S3_DMUL(w, cv.y));
                }
            }
#pragma unroll
            for (int hh = 0; hh < HV; hh++) {
                if (!act[hh]) continue;
                const float2 c = crow[hh * 32 + lane];
                acc[hh].x = S3_DADD(acc[hh].x, (double)c.x);
                acc[hh].y = S3_DADD(acc[hh].y, (double)c.y);
                if (li < cap) cur[((size_t)li * HV + hh) * 32 + lane] = acc[hh];
                *reinterpret_cast<double2*>(V.aup + (size_t)v * Dp + lab[hh]) = acc[hh];
            }
        }
        q++;
        if (h.flags & A2_FLAG_LEVEL_END) { double2* tmp = cur; cur = prev; prev = tmp; }
        if (h.flags & A2_FLAG_LAST) break;
    }
    fence_proxy_async();      // the running sums written above are read back by bulk copies in pass 2
    named_bar_sync(1, nall);  // pass boundary

    // ------------------------------------------------------------------ root -> leaf, WTA folded in
    long long m_bar_up = m_bar, m_work_up = m_work;
    while (true) {
        mc1 = clock64(); m_work += mc1 - mc0;
        named_bar_sync(1, nall);
        mc0 = clock64(); m_bar += mc0 - mc1;
        const TileHdr h = s_hdr[q % A2_MAXT];
        const int ps = h.aux;
        for (int i = warp; i < h.n; i += NW) {
            const int ent = h.ent0 + i;
            const int4 nd = *reinterpret_cast<const int4*>(s_meta + (size_t)ent * 16);  // {parent, pw, level, pixel}
            const double2* arow = reinterpret_cast<const double2*>(s_rows + (size_t)ent * (seg * 8));
            const int v = h.t0 + i, li = h.loff + i;
            double2 fin[HV];
            if (nd.x == v) {  // root
#pragma unroll
                for (int hh = 0; hh < HV; hh++)
                    if (act[hh]) fin[hh] = arow[hh * 32 + lane];
            } else {
                const int p = nd.x, j = p - ps;
                const double w = s_w[nd.y], w2 = s_w2[nd.y];
#pragma unroll
                for (int hh = 0; hh < HV; hh++) {
                    if (!act[hh]) continue;
                    double2 pv;
                    if (j < cap)
                        pv = prev[((size_t)j * HV + hh) * 32 + lane];
                    else
                        pv = *reinterpret_cast<const double2*>(V.aup + (size_t)p * Dp + lab[hh]);
                    const double2 au = arow[hh * 32 + lane];
                    fin[hh].x = S3_DADD(S3_DMUL(w, pv.x), S3_DMUL(w2, au.x));
                    fin[hh].y = S3_DADD(S3_DMUL(w, pv.y), S3_DMUL(w2, au.y));
                }
            }
            double bc = DBL_MAX;
            int bd = 0x7fffffff;
#pragma unroll
            for (int hh = 0; hh < HV; hh++) {
                if (!act[hh]) continue;
                if (li < cap) cur[((size_t)li * HV + hh) * 32 + lane] = fin[hh];
                if (li >= cap || A.keep) *reinterpret_cast<double2*>(V.aup + (size_t)v * Dp + lab[hh]) = fin[hh];
                if (fin[hh].x < bc) { bc = fin[hh].x; bd = lab[hh]; }
                if (lab[hh] + 1 < A.d1 && fin[hh].y < bc) { bc = fin[hh].y; bd = lab[hh] + 1; }
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
                    const int pix = nd.w;
                    V.disp[pix] = (int)md;
                    V.best[pix] = mc;
                } else {
                    V.disp[(size_t)slice * A.N + v] = (int)md;
                    V.best[(size_t)slice * A.N + v] = mc;
                }
            }
        }
        q++;
        if (h.flags & A2_FLAG_LEVEL_END) { double2* tmp = cur; cur = prev; prev = tmp; }
        if (h.flags & A2_FLAG_LAST) break;
    }
    named_bar_sync(1, nall);  // pass boundary (matches the DMA warp's)
    if (A.dbg && blockIdx.x == 0 && tid == 0) {
        A.dbg[4] = m_bar_up; A.dbg[5] = m_work_up; A.dbg[6] = m_bar - m_bar_up; A.dbg[7] = m_work - m_work_up;
    }
}

static size_t agg2_smem_bytes(int HV, int rn, int cap) {
    const size_t SW = 64 * HV;
    return 128 + A2_MAXT * 32 + (size_t)rn * 16 + (size_t)rn * SW * 8 + 2 * (size_t)cap * HV * 32 * 16 + 2 * S3_NUM_W * sizeof(double);
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
    if (d0 < 0 || d1 > Dv || d0 >= d1 || (d0 & 3)) return s3_fail(ctx, S3DMST_E_ARG, "aggregate_dense: need 0 <= d0 < d1 <= D and d0 % 4 == 0");
    const int nl = d1 - d0;
    const int HV = nl > 64 ? 2 : 1;
    const int SW = 64 * HV;
    const int n_slices = (nl + SW - 1) / SW;
    int threads = ctx->P.agg_threads > 0 ? ctx->P.agg_threads : 256;
    threads = std::max(32, std::min(992, threads / 32 * 32));
    const int NW = threads / 32;
    const int tn = std::min(32, NW);
    const int cap = ctx->P.agg_cache_nodes > 0 ? ctx->P.agg_cache_nodes : 16;
    const int rn = std::max(4 * tn, ctx->P.agg_ring_nodes > 0 ? ctx->P.agg_ring_nodes : 48);
    const size_t smem = agg2_smem_bytes(HV, rn, cap);
    if (smem > 227 * 1024) return s3_fail(ctx, S3DMST_E_ARG, "aggregate_dense: ring/cache sizes need %zu B of shared memory", smem);

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
        G.tree_start = V.tree_start; G.tree_depth = V.tree_depth; G.lvl_start = V.lvl_start;
        G.node_up = V.node_up; G.node_dn = V.node_dn; G.node_pixel = V.node_pixel;
        G.cost = V.cost; G.aup = V.aup;
        G.disp = n_slices == 1 ? V.disp_i : pdisp[view];
        G.best = n_slices == 1 ? V.best : pbest[view];
    }
    A.units = ctx->units_dev;
    A.n_slices = n_slices; A.Dp = Dp; A.d0 = d0; A.d1 = d1; A.N = ctx->N;
    A.lut_w = ctx->lut_w; A.lut_w2 = ctx->lut_w2;
    A.cap = cap; A.rn = rn; A.tn = tn; A.keep = ctx->P.keep_aggregated;
    const bool dbg = getenv("S3_DEBUG_AGG") != nullptr;
    if (dbg) S3_CUDA(cudaMalloc(&A.dbg, 8 * sizeof(long long)));

    const int grid = (int)units.size() * n_slices;
    S3_EV_BEGIN(S3DMST_T_AGG, first);
    if (HV == 2) {
        S3_CUDA(cudaFuncSetAttribute(k_agg_dense2<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_agg_dense2<2><<<grid, threads + 32, smem, ctx->stream>>>(A);
    } else {
        S3_CUDA(cudaFuncSetAttribute(k_agg_dense2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_agg_dense2<1><<<grid, threads + 32, smem, ctx->stream>>>(A);
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
        long long h[8];
        S3_CUDA(cudaMemcpy(h, A.dbg, sizeof h, cudaMemcpyDeviceToHost));
        fprintf(stderr, "[agg2 cta0] dma: issue %lld wait %lld bar %lld tiles %lld | math w0: up bar %lld work %lld, down bar %lld work %lld (cycles)\n",
                h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
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
