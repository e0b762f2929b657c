"""GPU parity tests: every CUDA stage, called through the C ABI, against the oracle on the same seeded inputs.
Bit-exact for integer/index work and (exact mode) for the FP64 aggregated costs; run with -m gpu on a B200."""
import os

import numpy as np
import pytest

from stereomatch_b200 import synth

pytestmark = pytest.mark.gpu


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({4: np.uint32, 8: np.uint64}[a.dtype.itemsize])


@pytest.fixture(scope="module")
def api():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from stereomatch_b200 import api as _api
    _api.load_library()
    return _api


CASES = [  # W, H, D, seed, c, min_size, natural
    (96, 64, 16, 7, 5000.0, 200, 0),
    (96, 64, 16, 7, 300.0, 20, 0),
    (64, 64, 8, 1, 50.0, 5, 0),
    (160, 120, 24, 3, 5000.0, 200, 1),
    (320, 200, 40, 5, 5000.0, 200, 0),
    (317, 203, 33, 6, 1000.0, 50, 1),   # ragged sizes, odd D
    (50, 30, 8, 10, 1e12, 2, 0),        # c -> inf: one tree == the (w, edge-id)-ordered Kruskal MST
    (33, 1, 6, 11, 100.0, 2, 0),        # single row
    (1, 37, 6, 12, 100.0, 2, 0),        # single column
]


def make(W, H, D, seed, nat):
    return (synth.make_natural_pair if nat else synth.make_pair)(W, H, max(D, 12), seed=seed)


def check_forest(F, G):
    assert G["T"] == F.T
    assert np.array_equal(G["ew"], F.ew)
    assert np.array_equal(G["mask"], F.mask)
    assert np.array_equal(G["tree_id"], F.tree_id)
    assert np.array_equal(G["tree_start"], F.tree_start)
    assert np.array_equal(G["node_pixel"], F.node_pixel)
    assert np.array_equal(G["parent"], F.parent)
    assert np.array_equal(G["child_count"], F.child_count)
    m = F.child_count > 0
    assert np.array_equal(G["child_begin"][m], F.child_begin[m])
    assert np.array_equal(G["pw"], F.pw)
    assert np.array_equal(G["level"], F.level)
    assert np.array_equal(G["adj_ptr"], F.adj_ptr)
    assert np.array_equal(G["adj"], F.adj)
    assert G["max_depth"] == F.max_depth


@pytest.mark.parametrize("W,H,D,seed,c,ms,nat", CASES)
def test_forest_matches_oracle(api, oracle, W, H, D, seed, c, ms, nat):
    L, R, _ = make(W, H, D, seed, nat)
    eng = api.Stereo3DMST(fh_c=c, min_cc_size=ms)
    eng.set_images(L, R)
    for view, img in ((0, L), (1, R)):
        eng.build_forest(view)
        F = oracle.forest(img, c=c, min_size=ms)
        check_forest(F, eng.get_forest(view))
        # MST property (north star): with c = inf the forest is the unique (w, id)-ordered MST; total weight equal
        assert int(F.ew[F.mask > 0].astype(np.int64).sum()) == int(eng.get_forest(view)["ew"][eng.get_forest(view)["mask"] > 0].astype(np.int64).sum())
    eng.close()


@pytest.mark.parametrize("cluster", [8, 16])
def test_forest_cluster_barrier_variant(api, oracle, cluster):
    """params.fh_cluster: every view's forest kernel runs in one thread-block cluster of 8 / 16 CTAs with the hardware
    cluster barrier instead of the cooperative grid's software barrier — same forest, edge for edge."""
    for (W, H, seed, nat, c, ms) in ((320, 200, 5, 0, 5000.0, 200), (317, 203, 6, 1, 1000.0, 50)):
        L, R, _ = make(W, H, 16, seed, nat)
        eng = api.Stereo3DMST(fh_c=c, min_cc_size=ms, fh_cluster=cluster)
        eng.set_images(L, R)
        eng.build_forest(0); eng.build_forest(1)
        check_forest(oracle.forest(L, c=c, min_size=ms), eng.get_forest(0))
        check_forest(oracle.forest(R, c=c, min_size=ms), eng.get_forest(1))
        eng.close()


def test_forest_without_median(api, oracle):
    L, R, _ = make(80, 60, 16, 4, 0)
    eng = api.Stereo3DMST(median=0, fh_c=800.0, min_cc_size=30)
    eng.set_images(L, R)
    eng.build_forest(0)
    check_forest(oracle.forest(L, c=800.0, min_size=30, median=False), eng.get_forest(0))
    eng.close()


@pytest.mark.parametrize("W,H,D,seed", [(96, 64, 16, 7), (130, 50, 100, 2), (64, 40, 130, 3)])
def test_cost_volume_matches_oracle(api, oracle, W, H, D, seed):
    L, R, _ = make(W, H, 16, seed, 0)
    eng = api.Stereo3DMST()
    eng.set_images(L, R)
    eng.build_forest(0)
    eng.build_forest(1)
    lv, rv = oracle.cost_adgrad(L, R, D)
    eng.build_cost_volume(D, ingest=False)
    assert np.array_equal(bits(eng.get_cost_volume(0)), bits(lv))
    assert np.array_equal(bits(eng.get_cost_volume(1)), bits(rv))
    eng.close()
    eng = api.Stereo3DMST(cost_scale=1 / 6.0)
    eng.set_images(L, R)
    eng.build_forest(0)
    eng.build_forest(1)
    eng.build_cost_volume(D, ingest=True)
    assert np.array_equal(bits(eng.get_cost_volume(0)), bits(oracle.ingest(lv, 0.5, 0.0, 1 / 6.0)))
    # external volume path (mc-cnn layout) with NaN scrub
    ext = lv.copy() * np.float32(0.25)
    ext[1, 5] = np.nan
    eng.set_cost_volume(0, ext, ingest=True)
    assert np.array_equal(bits(eng.get_cost_volume(0)), bits(oracle.ingest(ext, 0.5, 0.0, 1 / 6.0)))
    eng.close()


@pytest.mark.parametrize("W,H,D,seed,c,ms,nat", CASES[:7])
@pytest.mark.parametrize("own_forest,cluster", [(False, -1), (True, -1), (True, 48)])
def test_dense_aggregation_matches_oracle(api, oracle, W, H, D, seed, c, ms, nat, own_forest, cluster):
    """cluster = 48: every tree of >= 48 nodes is walked by a thread-block cluster of 8 CTAs (the giant-tree kernel)."""
    L, R, _ = make(W, H, D, seed, nat)
    F = oracle.forest(L, c=c, min_size=ms)
    lv, _ = oracle.cost_adgrad(L, R, D)
    disp_o, best_o, agg_o = oracle.aggregate_dense(F, lv, want_agg=True)
    eng = api.Stereo3DMST(fh_c=c, min_cc_size=ms, keep_aggregated=1, agg_cluster_nodes=cluster)
    eng.set_images(L, R)
    if own_forest:
        eng.build_forest(0)
    else:
        eng.set_forest(0, W, H, F.tree_start, F.node_pixel, F.parent, F.pw)
    eng.set_cost_volume(0, lv, ingest=False)
    disp, best = eng.aggregate_dense(0)
    agg = eng.get_aggregated(0)
    assert np.array_equal(bits(agg), bits(agg_o))          # aggregated costs: bit-exact (>= the 1e-4 bar)
    assert np.array_equal(disp, disp_o)                     # integer WTA disparities: bit-exact
    assert np.array_equal(bits(best), bits(best_o))
    np.testing.assert_allclose(agg, agg_o, rtol=1e-4)       # the tolerance north_star states, for the record
    eng.close()


@pytest.mark.parametrize("threads,cap", [(32, 1), (64, 2), (256, 4), (512, 64)])
def test_dense_aggregation_config_independent(api, oracle, threads, cap):
    """agg_kernel = 1: the simple level-synchronous kernel (aggregate.cu: any even d0, the fall-back of the dataflow kernel
    for label ranges that do not start on a 16-byte boundary), for several CTA shapes."""
    W, H, D = 200, 120, 70   # two label chunks per lane, a partly filled second chunk
    L, R, _ = make(W, H, 24, 9, 1)
    F = oracle.forest(L)
    lv, _ = oracle.cost_adgrad(L, R, D)
    disp_o, best_o, _ = oracle.aggregate_dense(F, lv)
    eng = api.Stereo3DMST(agg_threads=threads, agg_cache_nodes=cap, agg_kernel=1)
    eng.set_images(L, R)
    eng.build_forest(0)
    eng.set_cost_volume(0, lv, ingest=False)
    disp, best = eng.aggregate_dense(0)
    assert np.array_equal(disp, disp_o) and np.array_equal(bits(best), bits(best_o))
    d2, b2 = eng.aggregate_dense(0, 6, 50)
    do2, bo2, _ = oracle.aggregate_dense(F, lv, 6, 50)
    assert np.array_equal(d2, do2) and np.array_equal(bits(b2), bits(bo2))
    eng.close()


@pytest.mark.parametrize("cluster", [-1, 100])
def test_dense_label_slices_and_ranges(api, oracle, cluster):
    W, H, D = 120, 80, 200   # > 128 labels: several (tree, slice) units + the combine kernel
    L, R, _ = make(W, H, 24, 13, 0)
    F = oracle.forest(L, c=900.0, min_size=40)
    lv, _ = oracle.cost_adgrad(L, R, D)
    eng = api.Stereo3DMST(fh_c=900.0, min_cc_size=40, agg_cluster_nodes=cluster)
    eng.set_images(L, R)
    eng.build_forest(0)
    eng.set_cost_volume(0, lv, ingest=False)
    disp_o, best_o, _ = oracle.aggregate_dense(F, lv)
    disp, best = eng.aggregate_dense(0)
    assert np.array_equal(disp, disp_o) and np.array_equal(bits(best), bits(best_o))
    # label-range sharding: min-loc over the shards == the full run (what the NCCL reduction computes)
    parts = [eng.aggregate_dense(0, d0, d1) for d0, d1 in ((0, 50), (50, 100), (100, 151), (152, 200))]
    po = [oracle.aggregate_dense(F, lv, d0, d1)[:2] for d0, d1 in ((0, 50), (50, 100), (100, 151), (152, 200))]
    for (d, b), (do, bo) in zip(parts, po):
        assert np.array_equal(d, do) and np.array_equal(bits(b), bits(bo))
    eng.close()


def test_pms_apply_matches_oracle(api, oracle):
    W, H, D = 150, 90, 20
    L, R, _ = make(W, H, D, 21, 0)
    N = W * H
    c, ms = 700.0, 30
    F = oracle.forest(L, c=c, min_size=ms)
    lv_raw, _ = oracle.cost_adgrad(L, R, D)
    lv = oracle.ingest(lv_raw, 0.5, 0.0, 1 / 6.0)
    rng = np.random.default_rng(4)
    n = 40 * F.T + 77
    trees = rng.integers(0, F.T, n).astype(np.int32)
    labels = np.stack([rng.uniform(-0.05, 0.05, n), rng.uniform(-0.05, 0.05, n), rng.uniform(-3, D + 3, n)], 1).astype(np.float32)
    labels[5] = (0, 0, 3.0)          # exactly-integer disparity: cost 0 (Q10)
    labels[6] = (np.nan, 0, 3.0)     # NaN plane -> out-of-range cost
    labels[7] = (0, 0, 1e12)
    abc_o = oracle.plane_init(W, H, D)
    mn_o = np.full(N, np.finfo(np.float64).max)
    eng = api.Stereo3DMST(fh_c=c, min_cc_size=ms, cost_scale=1 / 6.0)
    eng.set_images(L, R)
    eng.build_forest(0)
    eng.build_forest(1)
    eng.build_cost_volume(D, ingest=True)
    assert np.array_equal(bits(eng.get_cost_volume(0)), bits(lv))
    eng.set_labels(0, abc_o)
    eng.reset_min_cost(0)
    # two calls: state must carry over exactly like consecutive MST_PMS calls
    h = n // 2
    eng.pms_apply(0, trees[:h], labels[:h])
    eng.pms_apply(0, trees[h:], labels[h:])
    oracle.pms_apply(F, lv, D, trees, labels, mn_o, abc_o)
    assert np.array_equal(bits(eng.get_min_cost(0)), bits(mn_o))
    assert np.array_equal(bits(eng.get_labels(0)), bits(abc_o))
    eng.label_to_disp(0)
    want = oracle.label_to_disp(abc_o, W, H, D) * np.float32(D - 1.0)
    assert np.array_equal(bits(eng.get_disparity(0)), bits(want))
    eng.close()


@pytest.mark.parametrize("cluster,kernel", [(-1, 0), (64, 0), (-1, 1)])
def test_pms_replays_recorded_reference_sequence(api, oracle, cluster, kernel):
    """Injected-proposal parity (north star): record the proposal sequence of oracle MST_PMS iterations and
    replay it on the GPU; labels and min costs must be identical.  (cluster: trees walked by a CTA cluster;
    kernel = 1: the simple level-synchronous proposal kernel.)"""
    W, H, D = 120, 72, 14
    L, R, _ = make(W, H, D, 31, 0)
    N = W * H
    c, ms = 600.0, 40
    F = oracle.forest(L, c=c, min_size=ms)
    lv = oracle.ingest(oracle.cost_adgrad(L, R, D)[0], 0.5, 0.0, 1 / 6.0)
    abc_o = oracle.plane_init(W, H, D)
    mn_o = np.full(N, np.finfo(np.float64).max)
    eng = api.Stereo3DMST(fh_c=c, min_cc_size=ms, cost_scale=1 / 6.0, agg_cluster_nodes=cluster, agg_kernel=kernel)
    eng.set_images(L, R)
    eng.build_forest(0)
    eng.build_forest(1)
    eng.build_cost_volume(D, ingest=True)
    eng.set_labels(0, abc_o)
    eng.reset_min_cost(0)
    g = oracle.rand_new(1)
    for it in range(3):
        n, rt, rl = oracle.mst_pms(F, lv, D, mn_o, abc_o, g, record_cap=64 * F.T)
        eng.pms_apply(0, rt, rl)
        assert np.array_equal(bits(eng.get_min_cost(0)), bits(mn_o)), it
        assert np.array_equal(bits(eng.get_labels(0)), bits(abc_o)), it
    eng.close()


def test_init_labels_and_generator(api, oracle):
    """a6: the library's plane initialisation is the reference's (bit-identical to the oracle, which is pinned to the
    reference TU); a12 with the library's own generator: min_cost never increases, every accepted label is one
    the round proposed for that tree, and a run is reproducible."""
    W, H, D = 128, 80, 16
    L, R, _ = make(W, H, D, 41, 0)
    eng = api.Stereo3DMST(fh_c=600.0, min_cc_size=40, cost_scale=1 / 6.0)
    eng.set_images(L, R)
    eng.build_forest(0); eng.build_forest(1)
    eng.build_cost_volume(D, ingest=True)
    eng.init_labels(0, D)
    assert np.array_equal(bits(eng.get_labels(0)), bits(oracle.plane_init(W, H, D)))
    assert np.all(eng.get_min_cost(0) == np.finfo(np.float64).max)
    eng.pms_iterate(0, 1, seed=3)
    m1 = eng.get_min_cost(0).copy(); l1 = eng.get_labels(0).copy()
    assert np.all(m1 < np.finfo(np.float64).max)
    eng.pms_iterate(0, 2, seed=4)
    m3 = eng.get_min_cost(0)
    assert np.all(m3 <= m1) and np.any(m3 < m1)
    eng.init_labels(0, D)
    eng.pms_iterate(0, 1, seed=3)
    assert np.array_equal(bits(eng.get_min_cost(0)), bits(m1)) and np.array_equal(bits(eng.get_labels(0)), bits(l1))
    eng.label_to_disp(0)
    d = eng.get_disparity(0)
    assert d.min() >= 0.0 and d.max() <= D - 1.0
    eng.close()


@pytest.mark.parametrize("fixture", ["ref_small.npz", "ref_medium.npz"])
def test_reference_pipeline_end_to_end_golden(api, oracle, fixture):
    """a1, north star "the 3DMST pipeline reproduces the reference disparity maps": tests/golden/ref_small.npz holds
    full_left_disp / full_right_disp written by the reference's own stereo3dmst() (its translation unit compiled
    unmodified, tests/golden/make_golden.py).  The oracle re-runs that pipeline (forests, plane init, 100 rounds of MST_PMS
    per view on the reference's RNG streams) and records every round's proposals; the GPU replays each round through the
    C ABI and must hold the oracle's labels and minimum costs after EVERY round, and the reference's maps at the end.
    ref_small is a single tree; ref_medium has 23 / 33 trees (labels propagate between neighbouring trees)."""
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", fixture))
    W, H, D = int(gold["W"]), int(gold["H"]), int(gold["D"])
    N = W * H
    L, R = gold["left"], gold["right"]
    if "lv_raw" in gold.files:
        raws = (gold["lv_raw"], gold["rv_raw"])
    else:  # the fixture's volume is the a2' volume x 1/6 (tests/golden/make_golden.py: make_medium)
        raws = tuple(v * np.float32(1 / 6.0) for v in oracle.cost_adgrad(L, R, D))
    eng = api.Stereo3DMST()                       # the reference's literals: c = 5000, min size 200, gamma = 1/12
    eng.set_images(L, R)
    eng.build_forest(0); eng.build_forest(1)
    g = oracle.rand_new(1)                        # std::rand(), shared by both views (Stereo3DMST.cpp:584)
    vols = (oracle.ingest(raws[0]), oracle.ingest(raws[1]))
    for view, raw in enumerate(raws):
        eng.set_cost_volume(view, raw, ingest=True)
        assert np.array_equal(bits(eng.get_cost_volume(view)), bits(vols[view]))
    forests, states = [], []
    for view, img in enumerate((L, R)):          # :841-847: per view forest, debug colouring (3N random() draws, Q6), plane init
        F = oracle.forest(img)
        oracle.lib.orc_rand_burn(g, 3 * N)
        eng.init_labels(view, D)
        abc = oracle.plane_init(W, H, D)
        assert np.array_equal(bits(eng.get_labels(view)), bits(abc))
        forests.append(F); states.append((np.full(N, np.finfo(np.float64).max), abc))
    n_props = 0
    for view in (0, 1):                           # :858-889: 100 rounds on the left view, then 100 on the right
        F, (mn, abc) = forests[view], states[view]
        for it in range(100):
            n, rt, rl = oracle.mst_pms(F, vols[view], D, mn, abc, g, record_cap=80 * F.T + 64)
            eng.pms_apply(view, rt, rl)
            n_props += n
            if it % 10 == 9 or it < 3:
                assert np.array_equal(bits(eng.get_min_cost(view)), bits(mn)), (view, it)
                assert np.array_equal(bits(eng.get_labels(view)), bits(abc)), (view, it)
        eng.label_to_disp(view)
    eng.lr_check(fill=False)                      # :904
    assert n_props > 900
    assert np.array_equal(bits(eng.get_disparity(0)), bits(gold["full_left_disp"]))
    assert np.array_equal(bits(eng.get_disparity(1)), bits(gold["full_right_disp"]))
    eng.close()


@pytest.fixture(scope="module")
def hdgen(tmp_path_factory):
    """Host build of stereomatch_b200/csrc/hd_math.h (the generator's arithmetic, shared with the kernels)."""
    import ctypes as C
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    out = str(tmp_path_factory.mktemp("hdgen") / "libhdgen.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", out, os.path.join(here, "models", "hd_math_host.cpp")])
    lib = C.CDLL(out)
    p = C.c_void_p
    lib.hd_gen_propagation.argtypes = [C.c_int, p, p, p, p, p, C.c_uint, C.c_uint, p, p]
    lib.hd_gen_refinement.argtypes = [C.c_int, C.c_int, p, p, p, C.c_int, C.c_float, C.c_uint, C.c_uint, p, p]
    lib.hd_gen_refinement.restype = C.c_int
    return lib


@pytest.mark.parametrize("cluster", [-1, 64])
def test_device_generator_equals_its_injected_stream(api, oracle, hdgen, cluster):
    """a12 with the library's own generator (s3dmst_pms_iterate: propagation gather kernel + refinement ladder generated
    inside the proposal kernel, no host round trip).  Its proposal stream is a pure function of (seed, round, tree, slot)
    and of the labels; a host build of the same header regenerates it and injects it through s3dmst_pms_apply into a
    second context: both contexts must hold identical labels and costs after every round — which also pins the
    reference's sequencing: a tree's ladder starts from the label its own propagation proposals left (:584-595)."""
    import ctypes as C
    W, H, D = 160, 96, 20
    L, R, _ = make(W, H, D, 51, 0)
    N = W * H
    def mk():
        e = api.Stereo3DMST(fh_c=600.0, min_cc_size=40, cost_scale=1 / 6.0, agg_cluster_nodes=cluster)
        e.set_images(L, R)
        e.build_forest(0); e.build_forest(1)
        e.build_cost_volume(D, ingest=True)
        e.init_labels(0, D)
        return e
    dev, inj = mk(), mk()
    G = dev.get_forest(0)
    F = oracle.forest(L, c=600.0, min_size=40)
    assert np.array_equal(G["adj_ptr"], F.adj_ptr) and np.array_equal(G["adj"], F.adj)   # device CSR == tree_g (:377-384)
    T, nadj = G["T"], len(G["adj"])
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    seed = 7
    for rnd in range(4):
        dev.pms_iterate(0, 1, seed=seed)
        abc = inj.get_labels(0)                                 # labels at the start of the round
        pt = np.empty(nadj, np.int32); pl = np.empty((nadj, 3), np.float32)
        hdgen.hd_gen_propagation(T, ptr(G["adj_ptr"]), ptr(G["adj"]), ptr(G["tree_start"]), ptr(G["node_pixel"]), ptr(abc), seed, rnd, ptr(pt), ptr(pl))
        inj.pms_apply(0, pt, pl)
        abc = inj.get_labels(0)                                 # ... and after the propagation proposals
        rt = np.empty(64 * T, np.int32); rl = np.empty((64 * T, 3), np.float32)
        n = hdgen.hd_gen_refinement(T, W, ptr(G["tree_start"]), ptr(G["node_pixel"]), ptr(abc), D, 0.1, seed, rnd, ptr(rt), ptr(rl))
        assert n > T                                            # several ladder steps per tree stay inside [0, Dmax]
        inj.pms_apply(0, rt[:n], rl[:n])
        assert np.array_equal(bits(dev.get_min_cost(0)), bits(inj.get_min_cost(0))), rnd
        assert np.array_equal(bits(dev.get_labels(0)), bits(inj.get_labels(0))), rnd
    # the oracle agrees with the injected context on the last round's stream (one more round, checked on the CPU)
    mn_o = inj.get_min_cost(0).copy(); abc_o = inj.get_labels(0).copy()
    vol = inj.get_cost_volume(0)
    dev.pms_iterate(0, 1, seed=seed)
    pt = np.empty(nadj, np.int32); pl = np.empty((nadj, 3), np.float32)
    hdgen.hd_gen_propagation(T, ptr(G["adj_ptr"]), ptr(G["adj"]), ptr(G["tree_start"]), ptr(G["node_pixel"]), ptr(abc_o), seed, 4, ptr(pt), ptr(pl))
    oracle.pms_apply(F, vol, D, pt, pl, mn_o, abc_o)
    rt = np.empty(64 * T, np.int32); rl = np.empty((64 * T, 3), np.float32)
    n = hdgen.hd_gen_refinement(T, W, ptr(G["tree_start"]), ptr(G["node_pixel"]), ptr(abc_o), D, 0.1, seed, 4, ptr(rt), ptr(rl))
    oracle.pms_apply(F, vol, D, rt[:n], rl[:n], mn_o, abc_o)
    assert np.array_equal(bits(dev.get_min_cost(0)), bits(mn_o)) and np.array_equal(bits(dev.get_labels(0)), bits(abc_o))
    dev.close(); inj.close()


def test_run_reference_mode_pipeline(api, oracle):
    """s3dmst_run = the reference's orchestration (:805-904) with the library's generator: equals the stage calls it is
    made of, is reproducible, returns sub-pixel disparities in [0, Dmax-1] with invalid left pixels zeroed, and lands near
    the ground truth of the synthetic pair."""
    W, H, D = 256, 160, 32
    L, R, gt = make(W, H, D, 77, 0)
    eng = api.Stereo3DMST(cost_scale=1 / 6.0, num_iter=12)
    eng.set_images(L, R)
    dl, dr = eng.run(D, seed=5)
    e2 = api.Stereo3DMST(cost_scale=1 / 6.0, num_iter=12)
    e2.set_images(L, R)
    e2.build_forest(0); e2.build_forest(1)
    e2.build_cost_volume(D, ingest=True)
    for v in (0, 1):
        e2.init_labels(v, D)
        e2.pms_iterate(v, 12, seed=5)
        e2.label_to_disp(v)
    e2.lr_check(fill=False)
    assert np.array_equal(bits(dl), bits(e2.get_disparity(0))) and np.array_equal(bits(dr), bits(e2.get_disparity(1)))
    dl2, dr2 = eng.run(D, seed=5)
    assert np.array_equal(bits(dl), bits(dl2)) and np.array_equal(bits(dr), bits(dr2))
    assert dl.min() >= 0.0 and dl.max() <= D - 1.0 and dr.min() >= 0.0 and dr.max() <= D - 1.0
    assert np.any(dl != np.round(dl))                                     # slanted planes: sub-pixel values
    want, _ = oracle.lr_check(oracle.label_to_disp(eng.get_labels(0), W, H, D) * np.float32(D - 1.0), dr, W, H, D, False)
    assert np.array_equal(bits(dl), bits(want))
    valid = dl.reshape(H, W) > 0
    err = np.abs(dl.reshape(H, W) - gt)
    assert valid.mean() > 0.3 and (err[valid] <= 1.0).mean() > 0.5
    eng.close(); e2.close()


def test_batch_mixed_c4_shapes(api, oracle):
    """BASELINE config C4 is a batch of 1920x1080 (D=256) frames plus the FLIR pairs (2048x1536, D=100): contexts of
    different sizes cannot share one aggregation launch, so a caller groups them by shape.  One batch of two full-size
    synthetic C4 frames and one batch of two FLIR-sized frames (the bundled pair and its mirror image): every frame's
    maps equal the single-frame pipeline's, and the forests / winners equal the oracle's on one view of each shape."""
    import cv2
    from oracle.pyoracle import Oracle
    O = Oracle(fast=True)
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    FLl = cv2.imread(os.path.join(g, "flir_000020_left.jpg")); FLr = cv2.imread(os.path.join(g, "flir_000020_right.jpg"))
    groups = [
        (256, [synth.make_pair(1920, 1080, 256, seed=synth.BASE_SEED + 10 + i)[:2] for i in range(2)]),
        (100, [(FLl, FLr), (np.ascontiguousarray(FLr[:, ::-1]), np.ascontiguousarray(FLl[:, ::-1]))]),
    ]
    for D, frames in groups:
        engs = []
        for Li, Ri in frames:
            e = api.Stereo3DMST(fh_ctas=36)
            e.set_images(Li, Ri)
            engs.append(e)
        outs = api.run_dense_batch(engs, D, fill=True)
        Hh, Ww = frames[0][0].shape[:2]
        for (Li, Ri), (dl, dr), e in zip(frames, outs, engs):
            single = api.Stereo3DMST()
            single.set_images(Li, Ri)
            sl, sr = single.run_dense(D, fill=True)
            single.close()
            assert np.array_equal(bits(dl), bits(sl)) and np.array_equal(bits(dr), bits(sr))
        # oracle on the right view of the first frame of the group (forest + aggregated winners)
        Li, Ri = frames[0]
        F = O.forest(Ri)
        G = engs[0].get_forest(1)
        assert np.array_equal(G["mask"], F.mask) and np.array_equal(G["node_pixel"], F.node_pixel) and np.array_equal(G["parent"], F.parent)
        vol = engs[0].get_cost_volume(1)
        do, bo, _ = O.aggregate_dense(F, vol)
        del vol
        assert np.array_equal(bits(outs[0][1]), bits(do.astype(np.float32)))
        disp, best = engs[0].aggregate_dense(1, 0, D)
        assert np.array_equal(disp, do) and np.array_equal(bits(best), bits(bo))
        for e in engs:
            e.close()


def test_golden_reference_proposals(api, oracle):
    """tests/golden/ref_small.npz holds output of the reference's own code: replay its proposals on the GPU."""
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_small.npz"))
    W, H, D = int(gold["W"]), int(gold["H"]), int(gold["D"])
    N = W * H
    for tag in ("a", "b"):
        c, ms = gold[f"{tag}_params"]
        eng = api.Stereo3DMST(fh_c=float(c), min_cc_size=int(ms))
        eng.set_images(gold["left"], gold["right"])
        eng.build_forest(0)
        G = eng.get_forest(0)
        assert np.array_equal(G["node_pixel"], gold[f"{tag}_node_pixel"])
        assert np.array_equal(G["parent"], gold[f"{tag}_parent"])
        assert np.array_equal(G["adj"], gold[f"{tag}_adj"])
        eng.set_cost_volume(0, gold["lv_raw"], ingest=True)
        eng.set_labels(0, gold[f"{tag}_abc"])
        eng.reset_min_cost(0)
        eng.pms_apply(0, gold[f"{tag}_prop_trees"], gold[f"{tag}_prop_labels"])
        assert np.array_equal(bits(eng.get_min_cost(0)), bits(gold[f"{tag}_prop_min"]))
        assert np.array_equal(bits(eng.get_labels(0)), bits(gold[f"{tag}_prop_abc"]))
        eng.close()
    eng = api.Stereo3DMST()
    eng.set_images(gold["left"], gold["right"])
    eng.build_forest(0)
    eng.build_forest(1)
    eng.build_cost_volume(D)
    for fill, key in ((0, "lr_nofill"), (1, "lr_fill")):
        eng.set_disparity(0, gold["lr_left"])
        eng.set_disparity(1, gold["lr_right"])
        eng.lr_check(fill=bool(fill))
        assert np.array_equal(bits(eng.get_disparity(0)), bits(gold[key]))
        assert np.array_equal(bits(eng.get_disparity(1)), bits(gold["lr_right"]))
    eng.close()
    del N


@pytest.mark.parametrize("fill", [False, True])
def test_lr_check_matches_oracle(api, oracle, fill):
    rng = np.random.default_rng(8)
    for (W, H, D) in ((97, 13, 16), (1300, 3, 60), (5, 4, 4), (7001, 2, 24)):   # the last: 2*W ints of shared memory > 48 KB
        L, R, _ = make(W, H, 12, 1, 0)
        eng = api.Stereo3DMST(min_cc_size=2, fh_c=10.0)
        eng.set_images(L, R)
        eng.build_forest(0)
        eng.build_forest(1)
        eng.build_cost_volume(D)
        N = W * H
        for trial in range(4):
            left = rng.uniform(-2, D + 2, N).astype(np.float32)
            right = (left + rng.normal(0, 1.0, N)).astype(np.float32)
            if trial == 1:
                left = np.round(left)
            if trial == 2:
                left[:] = 1e30      # everything invalid
            if trial == 3:
                right = left.copy()  # mostly valid
                left[::7] = np.nan
            eng.set_disparity(0, left)
            eng.set_disparity(1, right)
            eng.lr_check(fill=fill)
            want, _ = oracle.lr_check(left, right, W, H, D, fill)
            assert np.array_equal(bits(eng.get_disparity(0)), bits(want)), (W, H, trial)
        eng.close()


@pytest.mark.parametrize("W,H,D,seed,nat,cluster", [(96, 64, 16, 3, 0, -1), (130, 70, 23, 5, 1, -1), (200, 120, 70, 9, 1, 64),
                                                     (120, 80, 200, 13, 0, -1), (333, 97, 130, 17, 0, 100), (64, 48, 100, 19, 1, -1)])
def test_fused_matching_cost_equals_the_volume_path(api, oracle, W, H, D, seed, nat, cluster):
    """params.fuse_cost (the default): a dense run computes the AD+gradient cost inside the aggregation kernel and builds no
    volume.  The aggregated costs (all of them, keep_aggregated), winners and maps are bit-identical to the run that builds
    the volume first and to the oracle — including labels without a counterpart in the other image (D > W: cost 3.0), the
    image borders, partly filled label slices, several slices per tree and the cluster kernel.  The volume a fused run
    skipped is built on demand and is the oracle's."""
    L, R, _ = make(W, H, min(D, 24), seed, nat)
    lv, rv = oracle.cost_adgrad(L, R, D)
    res = {}
    for fuse in (0, -1):
        eng = api.Stereo3DMST(fuse_cost=fuse, keep_aggregated=1, agg_cluster_nodes=cluster)
        eng.set_images(L, R)
        dl, dr = eng.run_dense(D, fill=True)
        res[fuse] = (dl, dr, eng.get_aggregated(0), eng.get_aggregated(1), eng.get_dense_result(0), eng.get_dense_result(1))
        if fuse == 0:
            assert np.array_equal(bits(eng.get_cost_volume(0)), bits(lv)) and np.array_equal(bits(eng.get_cost_volume(1)), bits(rv))
        eng.close()
    for a, b in zip(res[0], res[-1]):
        if isinstance(a, tuple):
            assert all(np.array_equal(bits(x), bits(y)) for x, y in zip(a, b))
        else:
            assert np.array_equal(bits(a), bits(b))
    for view, (img, vol) in enumerate(((L, lv), (R, rv))):
        F = oracle.forest(img)
        disp_o, best_o, agg_o = oracle.aggregate_dense(F, vol, want_agg=True)
        assert np.array_equal(bits(res[0][2 + view]), bits(agg_o))
        assert np.array_equal(res[0][4 + view][0], disp_o) and np.array_equal(bits(res[0][4 + view][1]), bits(best_o))


@pytest.mark.parametrize("fuse", [0, -1])
def test_run_dense_end_to_end(api, oracle, fuse):
    W, H, D = 256, 160, 48
    L, R, gt = make(W, H, D, 77, 0)
    eng = api.Stereo3DMST(fuse_cost=fuse)
    eng.set_images(L, R)
    dl, dr = eng.run_dense(D, fill=True)
    n0 = eng.launch_count()
    assert n0 > 0
    lv, rv = oracle.cost_adgrad(L, R, D)
    FL, FR = oracle.forest(L), oracle.forest(R)
    dlo = oracle.aggregate_dense(FL, lv)[0].astype(np.float32)
    dro = oracle.aggregate_dense(FR, rv)[0].astype(np.float32)
    want, _ = oracle.lr_check(dlo, dro, W, H, D, True)
    assert np.array_equal(bits(dl), bits(want))
    assert np.array_equal(bits(dr), bits(dro))
    # and it is a usable disparity map: most pixels within 1 px of the ground truth
    err = np.abs(dl.reshape(H, W) - gt)
    assert (err[:, D:] <= 1.0).mean() > 0.6
    # second run on the same context: idempotent
    dl2, dr2 = eng.run_dense(D, fill=True)
    assert np.array_equal(bits(dl), bits(dl2)) and np.array_equal(bits(dr), bits(dr2))
    eng.close()


def test_run_dense_batch_matches_single_frames(api, oracle):
    """Batched pipeline (s3dmst_run_dense_batch): frames on their own contexts/streams, one aggregation launch over all
    frames' trees, reduced forest-kernel grid — results identical to the oracle frame by frame."""
    W, H, D = 200, 120, 40
    frames = [make(W, H, D, 60 + i, i % 2) for i in range(3)]
    engs = []
    for L, R, _ in frames:
        e = api.Stereo3DMST(fh_ctas=24)
        e.set_images(L, R)
        engs.append(e)
    outs = api.run_dense_batch(engs, D, fill=True)
    for (L, R, _), (dl, dr) in zip(frames, outs):
        lv, rv = oracle.cost_adgrad(L, R, D)
        dlo = oracle.aggregate_dense(oracle.forest(L), lv)[0].astype(np.float32)
        dro = oracle.aggregate_dense(oracle.forest(R), rv)[0].astype(np.float32)
        want, _ = oracle.lr_check(dlo, dro, W, H, D, True)
        assert np.array_equal(bits(dl), bits(want)) and np.array_equal(bits(dr), bits(dro))
    # the asynchronous variant: two batches queued back to back into two sets of host buffers, read after sync()
    import torch
    bufs = [[(torch.empty(W * H, dtype=torch.float32).pin_memory().numpy(), torch.empty(W * H, dtype=torch.float32).pin_memory().numpy())
             for _ in engs] for _ in range(2)]
    for k in range(2):
        for e, (L, R, _) in zip(engs, frames if k == 0 else frames[::-1]):
            e.set_images(L, R)
        api.run_dense_batch(engs, D, fill=True, out=bufs[k], wait=False)
    for e in engs:
        e.sync()
    for k in range(2):
        for (dl, dr), (bl, br) in zip(outs if k == 0 else outs[::-1], bufs[k]):
            assert np.array_equal(bits(dl), bits(bl)) and np.array_equal(bits(dr), bits(br))
    # ... and the library's stage timers kept every sample of the calls queued without synchronisation in between
    for e in engs:
        e.stage_total_ms(api.T_AGG, reset=True); e.stage_total_ms(api.T_FOREST, reset=True)
    for k in range(3):
        api.run_dense_batch(engs, D, fill=True, fetch=False)
    tot, n = engs[0].stage_total_ms(api.T_AGG)
    assert n == 3 and tot > 0.0
    assert all(e.stage_total_ms(api.T_FOREST)[1] == 3 for e in engs)
    assert abs(engs[0].stage_ms(api.T_AGG) - tot / 3) < tot   # the latest sample is still readable
    for e, (L, R, _) in zip(engs, frames):
        e.set_images(L, R)
    # a second batch on the same contexts (state reuse) with D not a multiple of the slice width
    outs = api.run_dense_batch(engs[:2], 36, fill=False)
    for (L, R, _), (dl, dr) in zip(frames[:2], outs):
        lv, rv = oracle.cost_adgrad(L, R, 36)
        dlo = oracle.aggregate_dense(oracle.forest(L), lv)[0].astype(np.float32)
        dro = oracle.aggregate_dense(oracle.forest(R), rv)[0].astype(np.float32)
        want, _ = oracle.lr_check(dlo, dro, W, H, 36, False)
        assert np.array_equal(bits(dl), bits(want)) and np.array_equal(bits(dr), bits(dro))
    for e in engs:
        e.close()


def test_reference_signature_wrapper(api, oracle):
    W, H, D = 128, 80, 24
    L, R, _ = make(W, H, D, 5, 0)
    outl = np.zeros((H, W), np.float32)
    outr = np.zeros((H, W), np.float32)
    dl, dr = api.stereo3dmst("img1r.png", "img2r.png", L, R, outl, outr, "ADGRAD", D)
    assert np.array_equal(dl, outl) and np.array_equal(dr, outr)
    with pytest.raises(ValueError):
        api.stereo3dmst("a", "b", L, R, None, None, "nope", D)
    lv, rv = oracle.cost_adgrad(L, R, D)
    dl2, _ = api.stereo3dmst("a", "b", L, R, None, None, "MCCNN_acrt", D, left_volume=lv, right_volume=rv)
    # external-volume path applies the reference ingest (cap 0.5) first
    FL, FR = oracle.forest(L), oracle.forest(R)
    dlo = oracle.aggregate_dense(FL, oracle.ingest(lv))[0].astype(np.float32)
    dro = oracle.aggregate_dense(FR, oracle.ingest(rv))[0].astype(np.float32)
    want, _ = oracle.lr_check(dlo, dro, W, H, D, False)
    assert np.array_equal(bits(dl2.ravel()), bits(want))


def test_full_size_c2_properties(api, oracle):
    """BASELINE config 2 at full size (1280x720, D=128): forest identical to the oracle, dense result
    identical on both views, plus size-independent properties."""
    W, H, D = 1280, 720, 128
    L, R, gt = synth.make_pair(W, H, D)
    eng = api.Stereo3DMST()
    eng.set_images(L, R)
    dl, dr = eng.run_dense(D, fill=False)
    G = eng.get_forest(0)
    from oracle.pyoracle import Oracle
    fast = Oracle(fast=True)
    F = fast.forest(L)
    assert np.array_equal(G["mask"], F.mask) and np.array_equal(G["node_pixel"], F.node_pixel) and np.array_equal(G["parent"], F.parent)
    assert int((G["mask"] > 0).sum()) == W * H - G["T"]                  # forest edge count = N - T
    assert np.diff(G["tree_start"]).min() >= 200                         # every tree >= min size
    lv = eng.get_cost_volume(0)
    disp_o, best_o, _ = fast.aggregate_dense(F, lv)
    disp, best = eng.aggregate_dense(0)
    assert np.array_equal(disp, disp_o) and np.array_equal(bits(best), bits(best_o))
    # linearity of the tree filter: scaling the volume by 2 scales the aggregated minimum by exactly 2
    eng.set_cost_volume(0, lv * np.float32(2.0), ingest=False)
    disp2, best2 = eng.aggregate_dense(0)
    assert np.array_equal(disp2, disp) and np.array_equal(bits(best2), bits(best * 2.0))
    # constant volume: every label ties, lowest d wins everywhere
    eng.set_cost_volume(0, np.full((8, W * H), 0.25, np.float32), ingest=False)
    dconst, _ = eng.aggregate_dense(0)
    assert not dconst.any()
    err = np.abs(dl.reshape(H, W) - gt)
    valid = dl.reshape(H, W) > 0
    # sanity only: pixelwise AD+gradient on 2x2 random dots is ambiguous (the oracle scores 0.34 / 0.57 here)
    assert valid.mean() > 0.2 and (err[valid] <= 1.0).mean() > 0.4
    eng.close()


def test_label_sharded_two_gpus_nccl(api, oracle):
    """BASELINE config C5 in small: one pair, label range split over 2 GPUs (torchrun), MIN-LOC over the library's own NCCL
    communicator; every rank's result must equal the oracle's full-range result bit for bit (tools/label_sharded.py).
    Needs 2 GPUs: on the single-GPU test box tests/test_comm_c.py covers the same entry points with one rank, and
    bench.py --gpus N runs this check on the driver's multi-GPU boxes."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(root, "tools", "label_sharded.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"check": true' in r.stdout


def test_label_sharded_reduce_partial_ranges(api, oracle):
    """s3dmst_reduce_minloc on a one-rank communicator is the identity on (best, disparity); s3dmst_comm_label_range
    partitions like stereomatch_b200.parallel.label_range (the gloo tests' reference)."""
    from stereomatch_b200 import parallel
    W, H, D = 160, 96, 40
    L, R, _ = make(W, H, D, 19, 0)
    eng = api.Stereo3DMST()
    eng.comm_init(eng.comm_unique_id(), 0, 1)
    assert eng.comm_label_range(D) == parallel.label_range(D, 1, 0)
    eng.set_images(L, R)
    eng.build_forest(0); eng.build_forest(1)
    eng.build_cost_volume(D)
    disp, best = eng.aggregate_dense(0, 8, 24)
    eng.reduce_minloc(0)
    d2, b2 = eng.get_dense_result(0)
    assert np.array_equal(disp, d2) and np.array_equal(bits(best), bits(b2))
    lv, _ = oracle.cost_adgrad(L, R, D)
    do, bo, _ = oracle.aggregate_dense(oracle.forest(L), lv, 8, 24)
    assert np.array_equal(d2, do) and np.array_equal(bits(b2), bits(bo))
    eng.close()
    # s3dmst_aggregate_dense_sharded without any cost volume (the aggregation kernel computes the matching cost) == with one
    res = []
    for prebuilt, p2p in ((False, 0), (True, 0), (False, -1)):
        # comm_p2p: 0 = the MIN-LOC as one kernel over peer memory when the ranks can map each other (one rank: itself),
        # -1 = the two NCCL all-reduces
        e = api.Stereo3DMST(comm_p2p=p2p)
        e.comm_init(e.comm_unique_id(), 0, 1)
        assert e.comm_transport() == 0   # decided at the first sharded call
        e.set_images(L, R)
        e.build_forest(0); e.build_forest(1)
        if prebuilt:
            e.build_cost_volume(D)
        e.aggregate_dense_sharded(D)
        e.sync()
        assert e.comm_transport() in ((0,) if p2p < 0 else (0, 1))
        res.append([e.get_dense_result(v) for v in (0, 1)])
        if p2p == 0 and not prebuilt:   # another image size on a live mapping: the buffers are re-mapped
            L2, R2, _ = make(96, 64, 16, 3, 0)
            e.set_images(L2, R2)
            e.build_forest(0); e.build_forest(1)
            e.aggregate_dense_sharded(16)
            e.sync()
            lv2, _ = oracle.cost_adgrad(L2, R2, 16)
            do2, bo2, _ = oracle.aggregate_dense(oracle.forest(L2), lv2)
            d2, b2 = e.get_dense_result(0)
            assert np.array_equal(d2, do2) and np.array_equal(bits(b2), bits(bo2))
        e.close()
    _, rv = oracle.cost_adgrad(L, R, D)
    for v, (img, vol) in enumerate(((L, lv), (R, rv))):
        do, bo, _ = oracle.aggregate_dense(oracle.forest(img), vol)
        for r in res:
            assert np.array_equal(r[v][0], do) and np.array_equal(bits(r[v][1]), bits(bo))


def test_full_size_c3_pms_properties(api, oracle):
    """BASELINE config C3 (1920x1080, 256 disparities, injected proposal sequence) at full size: size-independent
    properties of the proposal path plus oracle checks restricted to whole trees (a tree's update touches only its own
    pixels, so a handful of trees can be checked against the CPU in milliseconds)."""
    import time
    W, H, D = 1920, 1080, 256
    N = W * H
    L, R, _ = synth.make_pair(W, H, D, seed=synth.BASE_SEED + 1)
    eng = api.Stereo3DMST(cost_scale=1 / 6.0)
    eng.set_images(L, R)
    eng.build_forest(0); eng.build_forest(1)
    eng.build_cost_volume(D, ingest=True)
    F = oracle.forest(L)
    G = eng.get_forest(0)
    assert G["T"] == F.T and np.array_equal(G["node_pixel"], F.node_pixel) and np.array_equal(G["parent"], F.parent)
    T = F.T
    rng = np.random.default_rng(synth.BASE_SEED + 2)
    K = 12                                   # labels per tree per iteration, 2 iterations
    trees = np.tile(np.repeat(np.arange(T, dtype=np.int32), K), 2)
    n = trees.size
    labels = np.stack([rng.uniform(-0.02, 0.02, n), rng.uniform(-0.02, 0.02, n), rng.uniform(0, D - 1, n)], 1).astype(np.float32)
    eng.init_labels(0, D)
    abc0 = eng.get_labels(0).copy()
    t0 = time.perf_counter()
    eng.pms_apply(0, trees, labels)
    eng.sync()
    dt = time.perf_counter() - t0
    visits = 2 * K * N
    print(f"C3 pms_apply: {n} proposals, {visits / 1e6:.0f} M node-visits in {dt * 1e3:.1f} ms (incl. upload) = {visits / dt / 1e9:.2f} G node-visits/s; "
          f"device {eng.stage_ms(api.T_PMS):.1f} ms")
    m1 = eng.get_min_cost(0).copy(); l1 = eng.get_labels(0).copy()
    assert np.all(m1 < np.finfo(np.float64).max)
    # (1) idempotence: the same list again cannot improve anything (strict '<')
    eng.pms_apply(0, trees, labels)
    assert np.array_equal(bits(eng.get_min_cost(0)), bits(m1)) and np.array_equal(bits(eng.get_labels(0)), bits(l1))
    # (2) trees are independent: any interleaving that keeps each tree's own order gives the same state
    perm = np.argsort(rng.integers(0, 1 << 30, n), kind="stable")
    perm = perm[np.argsort(trees[perm], kind="stable")]          # grouped by tree ...
    key = rng.permutation(T)[trees[perm]]
    perm = perm[np.argsort(key, kind="stable")]                  # ... trees in a random order, own order kept
    own = np.concatenate([np.sort(perm[trees[perm] == t]) for t in range(0, T, max(1, T // 7))])
    assert np.array_equal(own, np.concatenate([np.nonzero(trees == t)[0] for t in range(0, T, max(1, T // 7))]))
    eng.set_labels(0, abc0); eng.reset_min_cost(0)
    order = np.concatenate([np.nonzero(trees == t)[0] for t in rng.permutation(T)])
    eng.pms_apply(0, trees[order], labels[order])
    assert np.array_equal(bits(eng.get_min_cost(0)), bits(m1)) and np.array_equal(bits(eng.get_labels(0)), bits(l1))
    # (3) oracle on whole trees: the 3 smallest and one mid-sized tree
    vol = eng.get_cost_volume(0)
    sizes = np.diff(np.asarray(F.tree_start))
    pick = list(np.argsort(sizes)[:3]) + [int(np.argsort(sizes)[T // 2])]
    mn_o = np.full(N, np.finfo(np.float64).max); abc_o = abc0.copy()
    for t in pick:
        sel = np.nonzero(trees == t)[0]
        oracle.pms_apply(F, vol, D, trees[sel], labels[sel], mn_o, abc_o)
    tid = np.asarray(F.tree_id)
    pix = np.isin(tid, pick)
    assert np.array_equal(bits(m1[pix]), bits(mn_o[pix])) and np.array_equal(bits(l1.reshape(N, 3)[pix]), bits(abc_o.reshape(N, 3)[pix]))
    # (4) Q10: an exactly-integer in-range disparity costs 0 everywhere, so it wins every pixel with aggregated cost 0
    zl = np.tile(np.array([[0.0, 0.0, 7.0]], np.float32), (T, 1))
    eng.pms_apply(0, np.arange(T, dtype=np.int32), zl)
    assert np.all(eng.get_min_cost(0) == 0.0)
    assert np.array_equal(eng.get_labels(0).reshape(N, 3), np.tile(zl[:1], (N, 1)))
    eng.label_to_disp(0)
    assert np.all(eng.get_disparity(0) == 7.0)
    eng.close()


def test_full_size_c1_flir_pair(api):
    """BASELINE config C1: the bundled FLIR pair 000020 (rectified as src/stereo_Yin.cpp:135-147 does; fixture made by
    tests/golden/make_flir_fixture.py), native 2048x1536, Dmax = 100 (stereo_Yin.cpp:207): forests, cost volumes,
    aggregated WTA and the final disparity maps are bit-identical to the oracle.  A natural image at 3.1 MP: ~500
    trees, the largest > 200 000 nodes and thousands of levels deep — the regime the kernels' far paths exist for."""
    import cv2
    import time
    from oracle.pyoracle import Oracle
    O = Oracle(fast=True)
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    L = cv2.imread(os.path.join(g, "flir_000020_left.jpg")); R = cv2.imread(os.path.join(g, "flir_000020_right.jpg"))
    assert L is not None and L.shape == (1536, 2048, 3) and R.shape == L.shape
    H, W, D = L.shape[0], L.shape[1], 100
    eng = api.Stereo3DMST()
    eng.set_images(L, R)
    t0 = time.perf_counter()
    dl, dr = eng.run_dense(D, fill=True)
    t_gpu = time.perf_counter() - t0
    FL, FR = O.forest(L), O.forest(R)
    for view, F in ((0, FL), (1, FR)):
        check_forest(F, eng.get_forest(view))
    sizes = np.diff(np.asarray(FL.tree_start))
    print(f"FLIR C1: T = {FL.T}/{FR.T} trees, largest {sizes.max()} nodes, depth {FL.max_depth}; GPU run_dense {t_gpu * 1e3:.1f} ms "
          f"(stages {[round(eng.stage_ms(s), 2) for s in range(4)]})")
    lv, rv = O.cost_adgrad(L, R, D)
    assert np.array_equal(bits(eng.get_cost_volume(0)), bits(lv))
    assert np.array_equal(bits(eng.get_cost_volume(1)), bits(rv))
    dlo, blo, _ = O.aggregate_dense(FL, lv)
    dro, bro, _ = O.aggregate_dense(FR, rv)
    del lv, rv
    want, _ = O.lr_check(dlo.astype(np.float32), dro.astype(np.float32), W, H, D, True)
    assert np.array_equal(bits(dr), bits(dro.astype(np.float32)))
    assert np.array_equal(bits(dl), bits(want))
    # the aggregated minima themselves (dense results stay in the context after run_dense)
    for view, (do, bo) in enumerate(((dlo, blo), (dro, bro))):
        disp, best = eng.aggregate_dense(view, 0, D)
        assert np.array_equal(disp, do) and np.array_equal(bits(best), bits(bo))
    eng.close()


def test_reproject_to_3d_matches_opencv(api):
    """SURVEY §8f rank 3 (src/stereo_Yin.cpp:218-243): disparity floor, cv::reprojectImageTo3D(..., Q, true) and the
    packed point-cloud colour — bit-identical to OpenCV (cv2 is the oracle for this row)."""
    import cv2
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    Q = np.load(os.path.join(g, "flir_000020_Q.npy"))
    W, H = 333, 207
    L, R, _ = make(W, H, 16, 77, 1)
    rng = np.random.default_rng(8)
    disp = rng.uniform(0, 99, (H, W)).astype(np.float32)
    disp[rng.random((H, W)) < 0.2] = 0.0                     # what the LR check leaves for invalid pixels
    eng = api.Stereo3DMST()
    eng.set_images(L, R)
    for floor, missing in ((10.0, True), (0.0, True), (10.0, False)):
        eng.set_disparity(0, disp)
        xyz, rgb = eng.reproject_to_3d(Q, disp_floor=floor, handle_missing=missing)
        d = disp.copy()
        d[d < floor] = floor                                # stereo_Yin.cpp:218-222
        want = cv2.reprojectImageTo3D(d, Q, handleMissingValues=missing)
        assert np.array_equal(bits(xyz), bits(want.reshape(-1, 3))), (floor, missing)
        assert np.array_equal(bits(eng.get_disparity(0)), bits(d.ravel()))
        Li = L.reshape(-1, 3).astype(np.uint32)
        assert np.array_equal(rgb, Li[:, 2] * 0x10000 + Li[:, 1] * 0x100 + Li[:, 0])
    eng.close()


def test_fast_mode_fp32_state_within_tolerance(api, oracle):
    """params.exact = 0: the dense aggregation keeps fp32 running sums (12 instead of 20 bytes of HBM traffic per
    pixel-label).  North-star tolerance for aggregated float costs: 1e-4 relative (stated here); the disparity must
    equal the oracle's wherever the oracle's winning margin exceeds that tolerance."""
    W, H, D = 320, 200, 40
    L, R, _ = make(W, H, D, 5, 0)
    F = oracle.forest(L)
    lv, _ = oracle.cost_adgrad(L, R, D)
    do, bo, ao = oracle.aggregate_dense(F, lv, want_agg=True)
    eng = api.Stereo3DMST(exact=0, keep_aggregated=1)
    eng.set_images(L, R)
    eng.build_forest(0); eng.build_forest(1)
    eng.build_cost_volume(D)
    disp, best = eng.aggregate_dense(0)
    agg = eng.get_aggregated(0)
    tol = 1e-4
    assert np.all(np.abs(agg - ao) <= tol * np.abs(ao) + 1e-12)
    assert np.all(np.abs(best - bo) <= tol * np.abs(bo) + 1e-12)
    srt = np.sort(ao, axis=0)
    clear = (srt[1] - srt[0]) > 2 * tol * srt[1]          # the winner is separated from the runner-up by more than the tolerance
    assert clear.mean() > 0.5
    assert np.array_equal(disp[clear], do[clear])
    # wherever the label differs it is a near-tie: the exact cost of the chosen label is within the tolerance of the minimum
    chosen = ao[disp, np.arange(W * H)]
    assert np.all(chosen - bo <= 2 * tol * np.abs(bo) + 1e-12)
    eng.close()


def test_remap_matches_opencv(api, oracle):
    """raw pair -> rectified images on the device == cv::remap(INTER_LINEAR) with the caller's CV_16SC2 maps
    (stereo_Yin.cpp:139-144), bit for bit; then the path runs on them."""
    from oracle import remap_oracle as ro
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "remap_small.npz"))
    e = api.Stereo3DMST()
    # golden vectors made by OpenCV from the reference's calibration (tests/golden/make_remap_fixture.py)
    for lk, rk in (("00", "10"), ("01", "11")):
        e.set_rectify_maps(0, g["xy_" + lk], g["fxy_" + lk])
        e.set_rectify_maps(1, g["xy_" + rk], g["fxy_" + rk])
        e.set_raw_images(g["src_" + lk], g["src_" + rk])
        assert np.array_equal(e.get_image(0), g["exp_" + lk])
        assert np.array_equal(e.get_image(1), g["exp_" + rk])
    # random maps with footprints leaving the source on every side, against the oracle (and OpenCV when it is importable)
    rng = np.random.default_rng(9)
    Hs, Ws, H, W = 70, 90, 64, 96
    srcs = [rng.integers(0, 256, (Hs, Ws, 3), dtype=np.uint8) for _ in range(2)]
    maps = [(np.stack([rng.integers(-4, Ws + 4, (H, W)), rng.integers(-4, Hs + 4, (H, W))], -1).astype(np.int16),
             rng.integers(0, 1024, (H, W)).astype(np.uint16)) for _ in range(2)]
    for v in range(2):
        e.set_rectify_maps(v, *maps[v])
    e.set_raw_images(*srcs)
    for v in range(2):
        want = ro.remap_fixed(srcs[v], *maps[v])
        assert np.array_equal(e.get_image(v), want)
        try:
            import cv2
            assert np.array_equal(want, cv2.remap(srcs[v], maps[v][0], maps[v][1], cv2.INTER_LINEAR))
        except ImportError:
            pass
    # the rectified pair feeds the path like an uploaded one
    left, right = e.get_image(0), e.get_image(1)
    dl, dr = e.run_dense(16, fill=True)
    e2 = api.Stereo3DMST()
    e2.set_images(left, right)
    dl2, dr2 = e2.run_dense(16, fill=True)
    assert np.array_equal(dl, dl2) and np.array_equal(dr, dr2)
    e.close(); e2.close()


def test_weighted_median_matches_oracle(api, oracle):
    """weightedMedianFilter (PatchMatchStereoGPU.cu:2436-2599) on LR-invalid pixels, bit for bit against
    the CPU restatement (oracle/postfilter_oracle.py), incl. windows hanging over every image border."""
    from oracle import postfilter_oracle as po
    W, H, D = 97, 61, 32
    L, R, _ = make(W, H, D, 33, 1)
    rng = np.random.default_rng(12)
    eng = api.Stereo3DMST()
    eng.set_images(L, R)
    for radius, gamma in ((10, 0.1), (3, 0.25)):
        tab = api.wmf_table(gamma)
        assert tab[0] == 1.0 and np.all(np.diff(tab) <= 0)
        disp = np.round(rng.uniform(0, D - 1, (H, W)) * 4).astype(np.float32) / np.float32(4)   # many equal disparities: the stable order matters
        mask = (rng.random((H, W)) < 0.15).astype(np.uint8)
        mask[0, 0] = mask[H - 1, W - 1] = mask[0, W - 1] = 1
        for view, img in ((0, L), (1, R)):
            eng.set_disparity(view, disp)
            eng.weighted_median(view, radius, gamma, mask)
            want = po.weighted_median(disp, mask, img, tab, radius)
            got = eng.get_disparity(view).reshape(H, W)
            assert np.array_equal(bits(got), bits(want)), (radius, view)
            assert np.array_equal(got[mask == 0], disp[mask == 0]) and np.any(got[mask == 1] != disp[mask == 1])
    # default mask = the left-right check's: pipeline -> LR check (no fill) -> weighted median of the invalid pixels
    dl, dr = eng.run_dense(D, fill=False)
    m = eng.get_lr_mask().reshape(H, W)
    assert np.array_equal(m == 1, dl.reshape(H, W) == 0) or np.all((dl.reshape(H, W) == 0)[m == 1])
    eng.weighted_median(0, 10, 0.1)
    want = po.weighted_median(dl.reshape(H, W), m, L, api.wmf_table(0.1), 10)
    assert np.array_equal(bits(eng.get_disparity(0).reshape(H, W)), bits(want))
    eng.close()


def test_norm_factor_matches_oracle(api, oracle):
    """cost_norm_factor (PatchMatchStereoGPU.cu:5333-5429, :5898-5919) = 1 / tree filter of the all-ones volume."""
    W, H, D = 160, 100, 8
    L, R, _ = make(W, H, 16, 35, 1)
    eng = api.Stereo3DMST(fh_c=800.0, min_cc_size=30)
    eng.set_images(L, R)
    eng.build_forest(0); eng.build_forest(1)
    eng.build_cost_volume(D)
    nf = eng.norm_factor(0)
    F = oracle.forest(L, c=800.0, min_size=30)
    _, best, _ = oracle.aggregate_dense(F, np.ones((4, W * H), np.float32))
    assert np.array_equal(bits(nf), bits(1.0 / best))
    assert nf.max() <= 1.0 and nf.min() > 0.0
    # the context's own volume is untouched: the dense result afterwards is the oracle's
    lv, _ = oracle.cost_adgrad(L, R, D)
    do, bo, _ = oracle.aggregate_dense(F, lv)
    disp, bst = eng.aggregate_dense(0)
    assert np.array_equal(disp, do) and np.array_equal(bits(bst), bits(bo))
    eng.close()


@pytest.mark.parametrize("cluster", [-1, 64])
def test_plane_cost_mode_matches_oracle(api, oracle, cluster):
    """North-star item 1, slanted variant (params.pms_cost_mode = 1): proposals are scored by the truncated colour + gradient
    difference to the other image at the sub-pixel match position (pm::PatchMatch, src/pm.cpp:97-154) instead of by a lerp
    in a cost volume — no volume exists.  Gradients bit-identical to the oracle's (= cv2's), an injected proposal list
    gives the oracle's labels and costs on both views, and the library's own generator runs on top of it."""
    W, H, D = 150, 90, 20
    L, R, gt = make(W, H, D, 21, 0)
    N = W * H
    c, ms = 700.0, 30
    rng = np.random.default_rng(4)
    eng = api.Stereo3DMST(fh_c=c, min_cc_size=ms, pms_cost_mode=1, cost_scale=0.25, agg_cluster_nodes=cluster)
    eng.set_images(L, R)
    eng.build_forest(0); eng.build_forest(1)
    eng.prepare_plane_cost(D)
    for view, img in ((0, L), (1, R)):
        assert np.array_equal(bits(eng.get_plane_gradients(view)), bits(oracle.pm_gradients(img)))
    for view, img in ((0, L), (1, R)):
        F = oracle.forest(img, c=c, min_size=ms)
        n = 70 * F.T + 13                      # more than 64 proposals on some trees: several batches
        trees = rng.integers(0, F.T, n).astype(np.int32)
        labels = np.stack([rng.uniform(-0.05, 0.05, n), rng.uniform(-0.05, 0.05, n), rng.uniform(-3, D + 3, n)], 1).astype(np.float32)
        labels[5] = (0, 0, 3.0); labels[6] = (np.nan, 0, 3.0); labels[7] = (0, 0, 1e12); labels[8] = (0, 0, float(D))
        abc_o = oracle.plane_init(W, H, D)
        mn_o = np.full(N, np.finfo(np.float64).max)
        eng.set_labels(view, abc_o)
        eng.reset_min_cost(view)
        eng.pms_apply(view, trees, labels)
        oracle.pms_apply_plane(F, view, L, R, D, trees, labels, mn_o, abc_o, scale=0.25)
        assert np.array_equal(bits(eng.get_min_cost(view)), bits(mn_o)), view
        assert np.array_equal(bits(eng.get_labels(view)), bits(abc_o)), view
    # the whole reference-mode pipeline on this data term (no cost volume anywhere)
    # (out-of-range planes must cost more than a mismatch, as the reference's PLANE_PENALTY = 120 does: 120 x scale)
    e2 = api.Stereo3DMST(pms_cost_mode=1, cost_scale=0.25, oob_cost=30.0, num_iter=16)
    W2, H2, D2 = 256, 160, 32
    L2, R2, gt2 = make(W2, H2, D2, 77, 0)
    e2.set_images(L2, R2)
    dl, dr = e2.run(D2, seed=3)
    valid = dl.reshape(H2, W2) > 0
    err = np.abs(dl.reshape(H2, W2) - gt2)
    assert valid.mean() > 0.3 and (err[valid] <= 1.0).mean() > 0.5
    with pytest.raises(api.S3Error):
        e2.aggregate_dense(0, 0, D2)           # there is no volume to aggregate
    eng.close(); e2.close()
