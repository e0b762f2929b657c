"""CPU model (numpy) of the GPU forest-construction algorithm, used to validate the parallel
formulation against the sequential reference semantics before/alongside the CUDA kernels.

FH phase  : weight levels in ascending order; within a level, Boruvka rounds over the level's
            edges between *open* components (w <= lastw + c/size), each component choosing its
            minimum edge id; all chosen edges are forest edges.  Thresholds never need storing:
            thr(r) = lastw[r] + f32(c)/f32(size[r]).
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


def build_forest_model(ew, W, H, c, min_size):
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
    for w in np.unique(w_all):
        sel = w_all == w
        pe, pa, pb = eid_all[sel], a_all[sel], b_all[sel]
        while len(pe):
            root = find_all(parent)
            parent = root
            ra, rb = root[pa], root[pb]
            thr = lastw.astype(np.float64) + (c / size.astype(np.float32)).astype(np.float64)
            live = (ra != rb) & (w <= thr[ra]) & (w <= thr[rb])
            pe, pa, pb, ra, rb = pe[live], pa[live], pb[live], ra[live], rb[live]
            if not len(pe):
                break
            rounds_fh += 1
            best = np.full(N, np.iinfo(np.int64).max)
            np.minimum.at(best, ra, pe)
            np.minimum.at(best, rb, pe)
            pick_a = best[ra] == pe
            pick_b = best[rb] == pe
            commit = pick_a | pick_b
            mask[pe[commit]] = 1
            # hooks
            hook_from = []
            hook_to = []
            mutual = pick_a & pick_b
            # a side hooks unless mutual and ra < rb
            ha = pick_a & ~(mutual & (ra < rb))
            hb = pick_b & ~(mutual & (rb < ra))
            old_size = size.copy()
            parent[ra[ha]] = rb[ha]
            parent[rb[hb]] = ra[hb]
            hooked = np.concatenate([ra[ha], rb[hb]])
            root2 = find_all(parent)
            np.add.at(size, root2[hooked], old_size[hooked])
            lastw[root2[hooked]] = w
            parent = root2
            keep = ~commit
            pe, pa, pb = pe[keep], pa[keep], pb[keep]
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
