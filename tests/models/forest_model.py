"""CPU model (numpy) of the GPU forest-construction algorithm (stereomatch_b200/csrc/forest.cu), used to validate the
parallel formulation against the sequential reference semantics (tests/test_forest_model.py compares it edge for
edge with the oracle, which is pinned to the reference's own segment_graph).

FH phase  : asynchronous exact rounds.  Only a weight-ordered prefix of the edges is live; every live edge posts its
            key (w, edge id) on both endpoint components (atomicMin).  An edge that is the minimum of a component is
            decided: (1) a component whose minimum edge is heavier than its threshold thr = lastw + f32(c)/f32(size)
            is dead for ever and all its edges are rejected; (2) an edge that is the minimum of both components is
            accepted; (3) an edge that is the minimum of one component is accepted if the other component's minimum
            has the same weight.  More weight levels are ingested when fewer than `band_low` edges are live.
Merge phase: reservation rounds.  Every pending edge (endpoints in different components, at
            least one smaller than m) reserves both endpoint components with atomicMin(key),
            key = (w, edge id).  An edge commits when, for each endpoint component, that
            component is big (>= m) or holds the edge's reservation; big-big edges are dead.
"""
import numpy as np


def find_all(parent):
    """Full path compression for every vertex (pointer jumping)."""
    p = parent.copy()
    while True:
        pp = p[p]
        if np.array_equal(pp, p):
            return p
        p = pp


def build_forest_model(ew, W, H, c, min_size, band_low=64, band_high=256):
    N = W * H
    c = np.float32(c)
    eid_all = np.nonzero(ew != 0xFFFF)[0]
    a_all = eid_all >> 1
    b_all = np.where(eid_all & 1, a_all + W, a_all + 1)
    w_all = ew[eid_all].astype(np.int64)
    parent = np.arange(N)
    size = np.ones(N, np.int64)
    lastw = np.zeros(N, np.int64)
    mask = np.zeros(2 * N, np.uint8)
    rounds_fh = 0
    order = np.lexsort((eid_all, w_all))           # weight buckets (the order inside a bucket is irrelevant)
    levels = np.unique(w_all)
    lvl_end = np.searchsorted(w_all[order], levels, side="right")
    INF = np.iinfo(np.int64).max
    lk = np.zeros(0, np.int64); la = np.zeros(0, np.int64); lb = np.zeros(0, np.int64)   # live list: key, endpoints
    band_pos, li = 0, 0
    while True:
        if len(lk) < band_low:
            new_pos = band_pos
            while li < len(levels) and len(lk) + (new_pos - band_pos) < band_high:
                new_pos = lvl_end[li]
                li += 1
            sel = order[band_pos:new_pos]
            lk = np.concatenate([lk, (w_all[sel] << 32) | eid_all[sel]])
            la = np.concatenate([la, a_all[sel]]); lb = np.concatenate([lb, b_all[sel]])
            band_pos = new_pos
        if not len(lk):
            break
        rounds_fh += 1
        root = find_all(parent)
        parent = root
        ra, rb = root[la], root[lb]
        live = ra != rb
        lk, la, lb, ra, rb = lk[live], la[live], lb[live], ra[live], rb[live]
        if not len(lk):
            continue
        best = np.full(N, INF)
        np.minimum.at(best, ra, lk)
        np.minimum.at(best, rb, lk)
        thr = lastw.astype(np.float64) + (c / size.astype(np.float32)).astype(np.float64)
        ka, kb = best[ra], best[rb]
        wa, wb, w = ka >> 32, kb >> 32, lk >> 32
        dead = (wa > thr[ra]) | (wb > thr[rb])                 # (1)
        pa, pb = ka == lk, kb == lk
        acc = ~dead & ((pa & pb) | (pa & (wb == w)) | (pb & (wa == w)))   # (2), (3)
        mask[(lk[acc] & 0xFFFFFFFF)] = 1
        mutual = acc & pa & pb
        sa, sb = size[ra], size[rb]
        a_hooks = np.where(mutual, (sa < sb) | ((sa == sb) & (ra > rb)), pa)   # only one side of a mutual pick hooks
        frm = np.where(a_hooks, ra, rb)[acc]
        to = np.where(a_hooks, rb, ra)[acc]
        old_size = size.copy()
        parent[frm] = to
        root2 = find_all(parent)
        np.add.at(size, root2[frm], old_size[frm])
        lastw[root2[frm]] = w[acc]
        parent = root2
        keep = ~dead & ~acc
        lk, la, lb = lk[keep], la[keep], lb[keep]
    fh_root = find_all(parent)
    # ---- min-size merge ----
    m = max(2, int(min_size))
    key_all = (w_all << 32) | eid_all
    root = fh_root.copy()
    parent = root.copy()
    pend = np.ones(len(eid_all), bool)
    rounds_merge = 0
    INF = np.iinfo(np.int64).max
    while True:
        root = find_all(parent)
        parent = root
        ra, rb = root[a_all], root[b_all]
        sa, sb = size[ra], size[rb]
        pend &= (ra != rb) & ((sa < m) | (sb < m))
        idx = np.nonzero(pend)[0]
        if not len(idx):
            break
        rounds_merge += 1
        resv = np.full(N, INF)
        np.minimum.at(resv, ra[idx], key_all[idx])
        np.minimum.at(resv, rb[idx], key_all[idx])
        oka = (sa[idx] >= m) | (resv[ra[idx]] == key_all[idx])
        okb = (sb[idx] >= m) | (resv[rb[idx]] == key_all[idx])
        com = idx[oka & okb]
        mask[eid_all[com]] = 2
        cra, crb, csa, csb = ra[com], rb[com], size[ra[com]], size[rb[com]]
        # hook: small under big; both small -> larger id under smaller id
        a_small, b_small = csa < m, csb < m
        a_hooks = (a_small & ~b_small) | (a_small & b_small & (cra > crb))
        frm = np.where(a_hooks, cra, crb)
        to = np.where(a_hooks, crb, cra)
        old = size.copy()
        parent[frm] = to
        np.add.at(size, to, old[frm])
        pend[com] = False
    root = find_all(parent)
    return mask, fh_root, root, dict(rounds_fh=rounds_fh, rounds_merge=rounds_merge)
