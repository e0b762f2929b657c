"""Oracle vs the reference's own translation unit (oracle/_ref/libref3dmst.so), randomised.
Skipped where the reference build is absent; the golden-vector test covers that case."""
import tempfile

import numpy as np
import pytest

from stereomatch_b200 import synth


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def test_median_shim_and_oracle_match_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    for shape in ((1, 1), (1, 7), (5, 1), (33, 47)):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        assert np.array_equal(oracle.median3(img), cv2.medianBlur(img, 3))


def test_glibc_rand_restatement(oracle):
    import ctypes
    libc = ctypes.CDLL(None)
    libc.srand(1)
    g = oracle.rand_new(1)
    assert [libc.rand() for _ in range(5000)] == [oracle.lib.orc_rand_next(g) for _ in range(5000)]


def test_segment_graph_header(oracle, ref):
    """include/segment-graph.h on our edge list: same accepted-edge set as the oracle's FH stage."""
    L, _, _ = synth.make_pair(64, 40, 8, seed=3)
    F = oracle.forest(L, c=700.0, min_size=2)
    W, H = 64, 40
    p = np.arange(W * H)
    x, y = p % W, p // W
    a = np.concatenate([p[x < W - 1], p[y < H - 1]])
    b = np.concatenate([p[x < W - 1] + 1, p[y < H - 1] + W])
    eid = np.concatenate([2 * p[x < W - 1], 2 * p[y < H - 1] + 1])
    w = F.ew[eid].astype(np.float64)
    n, ws, as_, bs, mask, comp, _ = ref.segment_graph(W * H, w, a, b, 700.0)
    eid_sorted = np.where(bs == as_ + 1, 2 * as_, 2 * as_ + 1)
    got = np.zeros(2 * W * H, np.uint8)
    got[eid_sorted[mask == 1]] = 1
    assert np.array_equal(got, (F.mask == 1).astype(np.uint8))
    # components agree as partitions
    _, inv_r = np.unique(comp, return_inverse=True)
    _, inv_o = np.unique(F.fh_comp, return_inverse=True)
    assert len(set(zip(inv_r.tolist(), inv_o.tolist()))) == n


@pytest.mark.parametrize("seed,W,H,D,c,ms", [(7, 96, 64, 16, 5000.0, 200), (8, 80, 50, 10, 400.0, 20), (9, 40, 33, 6, 60.0, 5),
                                              (10, 50, 30, 8, 1e12, 2)])
def test_forest_and_init(oracle, ref, seed, W, H, D, c, ms):
    L, _, _ = synth.make_pair(W, H, D, seed=seed)
    F = oracle.forest(L, c=c, min_size=ms)
    V = ref.view(L, D, c=c, min_size=ms)
    assert F.T == V.T
    for k in ("tree_start", "node_pixel", "parent", "child_count", "adj_ptr", "adj"):
        assert np.array_equal(getattr(F, k), getattr(V, k)), k
    m = V.child_count > 0
    assert np.array_equal(F.child_begin[m], V.child_begin[m])
    nr = F.parent != np.arange(F.N)
    assert np.array_equal(bits(F.wlut[F.pw][nr]), bits(V.weight[nr]))
    assert np.array_equal(bits(F.w2lut[F.pw][nr]), bits(V.weight2[nr]))
    assert np.array_equal(bits(oracle.plane_init(W, H, D)), bits(V.abc))
    if c > 1e9:  # c -> inf: a single tree == the (w,a,b)-ordered Kruskal MST
        assert F.T == 1


def test_pms_and_full_pipeline(oracle, ref):
    W, H, D = 88, 56, 14
    L, Rt, _ = synth.make_pair(W, H, D, seed=21)
    N = W * H
    lv_raw, rv_raw = oracle.cost_adgrad(L, Rt, D)
    lv_raw = (lv_raw * np.float32(1 / 6.0)).astype(np.float32)
    rv_raw = (rv_raw * np.float32(1 / 6.0)).astype(np.float32)
    lv = oracle.ingest(lv_raw)
    F = oracle.forest(L, c=800.0, min_size=30)
    ref.srand(1)
    V = ref.view(L, D, c=800.0, min_size=30)
    abc = oracle.plane_init(W, H, D)
    mo = np.full(N, np.finfo(np.float64).max)
    mr = mo.copy()
    agg = np.zeros(N)
    ref.srand(1)
    g = oracle.rand_new(1)
    for it in range(3):
        oracle.mst_pms(F, lv, D, mo, abc, g)
        V.mst_pms(lv, mr, agg)
        assert np.array_equal(bits(mo), bits(mr))
        assert np.array_equal(bits(abc), bits(V.get_abc()))
    with tempfile.TemporaryDirectory() as td:
        ref.srand(1)
        dl, dr = ref.stereo3dmst(td, L, Rt, lv_raw, rv_raw, D)
    out = oracle.stereo3dmst(L, Rt, lv_raw, rv_raw, D, num_iter=100)
    assert np.array_equal(bits(dl), bits(out["left_disp"]))
    assert np.array_equal(bits(dr), bits(out["right_disp"]))
