"""Oracle vs golden vectors produced by the reference's own code (tests/golden/make_golden.py).
Runs anywhere (no /root/reference, no GPU)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_small.npz")


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_forest_matches_reference(oracle, gold, tag):
    c, ms = gold[f"{tag}_params"]
    F = oracle.forest(gold["left"], c=float(c), min_size=int(ms))
    assert np.array_equal(F.tree_start, gold[f"{tag}_tree_start"])
    assert np.array_equal(F.node_pixel, gold[f"{tag}_node_pixel"])
    assert np.array_equal(F.parent, gold[f"{tag}_parent"])
    assert np.array_equal(F.child_count, gold[f"{tag}_child_count"])
    nr = F.parent != np.arange(F.N)
    assert np.array_equal(bits(F.wlut[F.pw][nr]), bits(gold[f"{tag}_weight"][nr]))
    assert np.array_equal(bits(F.w2lut[F.pw][nr]), bits(gold[f"{tag}_weight2"][nr]))
    assert np.array_equal(F.adj_ptr, gold[f"{tag}_adj_ptr"])
    assert np.array_equal(F.adj, gold[f"{tag}_adj"])
    # forest invariants
    assert int((F.mask > 0).sum()) == F.N - F.T
    sizes = np.diff(F.tree_start)
    assert sizes.min() >= max(2, int(ms)) or F.T == 1


def test_plane_init_matches_reference(oracle, gold):
    W, H, D = int(gold["W"]), int(gold["H"]), int(gold["D"])
    assert np.array_equal(bits(oracle.plane_init(W, H, D)), bits(gold["a_abc"]))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_proposals_and_pms_match_reference(oracle, gold, tag):
    W, H, D = int(gold["W"]), int(gold["H"]), int(gold["D"])
    N = W * H
    c, ms = gold[f"{tag}_params"]
    F = oracle.forest(gold["left"], c=float(c), min_size=int(ms))
    lv = oracle.ingest(gold["lv_raw"])
    abc = oracle.plane_init(W, H, D)
    mn = np.full(N, np.finfo(np.float64).max)
    for k, (t, lab) in enumerate(zip(gold[f"{tag}_prop_trees"], gold[f"{tag}_prop_labels"])):
        _, agg = oracle.eval_proposal(F, lv, D, int(t), lab, mn, abc)
        px = F.node_pixel[F.tree_start[t]:F.tree_start[t + 1]]
        assert np.array_equal(bits(agg[px]), bits(gold[f"{tag}_prop_agg"][k][px])), k
    assert np.array_equal(bits(mn), bits(gold[f"{tag}_prop_min"]))
    assert np.array_equal(bits(abc), bits(gold[f"{tag}_prop_abc"]))
    # integer-disparity quirk Q10: proposal 3 is the plane d == 5 exactly => per-node cost 0
    g = oracle.rand_new(1)
    for _ in range(2):
        oracle.mst_pms(F, lv, D, mn, abc, g)
    assert np.array_equal(bits(mn), bits(gold[f"{tag}_pms_min"]))
    assert np.array_equal(bits(abc), bits(gold[f"{tag}_pms_abc"]))
    assert np.array_equal(bits(oracle.label_to_disp(abc, W, H, D)), bits(gold[f"{tag}_pms_disp"]))


def test_lr_check_matches_reference(oracle, gold):
    W, H, D = int(gold["W"]), int(gold["H"]), int(gold["D"])
    lo, mask = oracle.lr_check(gold["lr_left"], gold["lr_right"], W, H, D, 0)
    assert np.array_equal(bits(lo), bits(gold["lr_nofill"]))
    assert np.array_equal(lo[mask == 1], np.zeros(int(mask.sum()), np.float32))
    lf, _ = oracle.lr_check(gold["lr_left"], gold["lr_right"], W, H, D, 1)
    assert np.array_equal(bits(lf), bits(gold["lr_fill"]))


def test_full_pipeline_matches_reference(oracle, gold):
    D = int(gold["D"])
    out = oracle.stereo3dmst(gold["left"], gold["right"], gold["lv_raw"], gold["rv_raw"], D, num_iter=100)
    assert np.array_equal(bits(out["left_disp"]), bits(gold["full_left_disp"]))
    assert np.array_equal(bits(out["right_disp"]), bits(gold["full_right_disp"]))


def test_full_pipeline_matches_reference_many_trees(oracle):
    """tests/golden/ref_medium.npz: the reference's stereo3dmst() on a 200x150 pair that its own literals segment into
    23 / 33 trees, so the 100 x 2 rounds propagate labels between neighbouring trees (ref_small is a single tree)."""
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_medium.npz"))
    D = int(gold["D"])
    L, R = gold["left"], gold["right"]
    assert (oracle.forest(L).T, oracle.forest(R).T) == tuple(gold["trees"])
    lv, rv = oracle.cost_adgrad(L, R, D)
    out = oracle.stereo3dmst(L, R, lv * np.float32(1 / 6.0), rv * np.float32(1 / 6.0), D, num_iter=100)
    assert np.array_equal(bits(out["left_disp"]), bits(gold["full_left_disp"]))
    assert np.array_equal(bits(out["right_disp"]), bits(gold["full_right_disp"]))


def test_tree_filter_bruteforce_identity(oracle, gold):
    """agg(v) = sum_u prod(weights on path u->v) * cost(u)  (SURVEY §4), small tree, 1e-12 rel."""
    D = int(gold["D"])
    F = oracle.forest(gold["left"], c=300.0, min_size=20)
    sizes = np.diff(F.tree_start)
    t = int(np.argmin(np.where(sizes >= 20, sizes, 1 << 30)))
    a, b = F.tree_start[t], F.tree_start[t + 1]
    n = b - a
    rng = np.random.default_rng(0)
    vol = rng.uniform(0, 0.5, (D, F.N)).astype(np.float32)
    disp, best, agg = oracle.aggregate_dense(F, vol, want_agg=True)
    # path products by walking to the root
    par = F.parent[a:b] - a
    w = F.wlut[F.pw[a:b]]
    lev = F.level[a:b]
    P = np.zeros((n, n))
    for u in range(n):
        for v in range(n):
            x, y, prod = u, v, 1.0
            while x != y:
                if lev[x] >= lev[y]:
                    prod *= w[x]; x = par[x]
                else:
                    prod *= w[y]; y = par[y]
            P[u, v] = prod
    pix = F.node_pixel[a:b]
    for d in (0, D // 2):
        want = P.T @ vol[d, pix].astype(np.float64)
        np.testing.assert_allclose(agg[d, pix], want, rtol=1e-12)
    assert np.array_equal(disp, np.argmin(agg, axis=0))  # first minimum == strict '<' ascending d


def test_label_cost_quirks(oracle, gold):
    W, H, D = int(gold["W"]), int(gold["H"]), int(gold["D"])
    N = W * H
    F = oracle.forest(gold["left"])
    lv = oracle.ingest(gold["lv_raw"])
    for lab, expect in ((np.float32([0, 0, 3.0]), 0.0), (np.float32([0, 0, -4.0]), None), (np.float32([0, 0, D + 3.0]), None)):
        mn = np.full(N, np.finfo(np.float64).max)
        abc = np.zeros((N, 3), np.float32)
        _, agg = oracle.eval_proposal(F, lv, D, 0, lab, mn, abc)
        px = F.node_pixel[F.tree_start[0]:F.tree_start[1]]
        if expect is not None:
            assert np.all(agg[px] == 0.0)       # Q10: exactly-integer disparity => cost 0 everywhere
        else:
            vol_half = np.full((D, N), 0.5, np.float32)
            _, agg2 = oracle.eval_proposal(F, vol_half, D, 0, np.float32([0, 0, 2.5]), mn.copy(), abc.copy())
            assert np.array_equal(bits(agg[px]), bits(agg2[px]))  # out of range == constant 0.5
