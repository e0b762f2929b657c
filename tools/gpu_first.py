"""Stage-by-stage GPU diagnostic (development aid): prints mismatch statistics instead of stopping at the first."""
import sys, time, os, faulthandler
faulthandler.enable()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stereomatch_b200 import api, synth
from oracle.pyoracle import Oracle

O = Oracle(fast=True)
def bits(a):
    a = np.ascontiguousarray(a); return a.view({4: np.uint32, 8: np.uint64}[a.dtype.itemsize])
def cmp(name, a, b):
    a = np.asarray(a); b = np.asarray(b)
    if a.shape != b.shape:
        print(f"   {name}: SHAPE {a.shape} vs {b.shape}"); return False
    if a.dtype.kind == 'f':
        ne = bits(a) != bits(b)
    else:
        ne = a != b
    n = int(ne.sum())
    print(f"   {name}: {'OK' if n == 0 else 'MISMATCH %d/%d first@%s' % (n, a.size, np.argwhere(ne)[:3].tolist())}", flush=True)
    return n == 0

def run(W, H, D, seed, c=5000.0, ms=200, nat=0, dense=True):
    print(f"== {W}x{H} D={D} seed={seed} c={c} ms={ms} nat={nat}", flush=True)
    L, R, gt = (synth.make_natural_pair if nat else synth.make_pair)(W, H, max(D, 12), seed=seed)
    eng = api.Stereo3DMST(fh_c=c, min_cc_size=ms, keep_aggregated=1 if W * H * D < 4e7 else 0)
    eng.set_images(L, R)
    t = time.time(); eng.build_forest(0); eng.sync(); print("   forest(left) wall %.1f ms, device %.2f ms" % ((time.time() - t) * 1e3, eng.stage_ms(api.T_FOREST)), flush=True)
    eng.build_forest(1)
    G = eng.get_forest(0)
    F = O.forest(L, c=c, min_size=ms)
    print("   T gpu/oracle", G["T"], F.T, "depth", G["max_depth"], F.max_depth)
    ok = True
    for k in ("ew", "mask", "tree_id", "tree_start", "node_pixel", "parent", "child_count", "pw", "level", "adj_ptr", "adj"):
        ok &= cmp(k, G[k], getattr(F, k))
    if not dense:
        eng.close(); return ok
    lv, rv = O.cost_adgrad(L, R, D)
    eng.build_cost_volume(D)
    ok &= cmp("cost L", eng.get_cost_volume(0), lv)
    ok &= cmp("cost R", eng.get_cost_volume(1), rv)
    if not cmp("forest usable", G["node_pixel"], F.node_pixel):
        eng.set_forest(0, W, H, F.tree_start, F.node_pixel, F.parent, F.pw)
        eng.set_cost_volume(0, lv, ingest=False)
    t = time.time(); disp, best = eng.aggregate_dense(0); print("   agg wall %.1f ms device %.3f ms" % ((time.time() - t) * 1e3, eng.stage_ms(api.T_AGG)))
    t = time.time(); do, bo, ao = O.aggregate_dense(F, lv, want_agg=bool(eng.params.keep_aggregated)); print("   oracle agg %.2f s" % (time.time() - t))
    ok &= cmp("disp", disp, do); ok &= cmp("best", best, bo)
    if eng.params.keep_aggregated:
        ok &= cmp("agg volume", eng.get_aggregated(0), ao)
    eng.close()
    return ok

if __name__ == "__main__":
    ok = True
    ok &= run(96, 64, 16, 7)
    ok &= run(96, 64, 16, 7, 300.0, 20)
    ok &= run(64, 64, 8, 1, 50.0, 5)
    ok &= run(320, 200, 40, 5)
    ok &= run(317, 203, 33, 6, 1000.0, 50, 1)
    ok &= run(1280, 720, 128, 20261018)
    print("ALL OK" if ok else "SOME MISMATCH")
