"""oracle/pyoracle.py — TEST INFRASTRUCTURE ONLY.

ctypes/numpy front-end to oracle/_ref/liboracle.so (the CPU restatement of the reference's
Stereo3DMST path, see oracle/s3dmst_oracle.cpp) and, when it was built, to
oracle/_ref/libref3dmst.so (the reference's own translation unit, see oracle/ref_driver.cpp).

May be imported only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_OUT = os.path.join(_HERE, "_ref")

c_p = C.c_void_p


def build(force: bool = False) -> None:
    """Compile the oracle (and the reference TU if /root/reference is mounted). Outputs: oracle/_ref/."""
    need = force or not os.path.exists(os.path.join(_OUT, "liboracle.so")) or (
        os.path.getmtime(os.path.join(_OUT, "liboracle.so")) < os.path.getmtime(os.path.join(_HERE, "s3dmst_oracle.cpp")))
    if need:
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(c_p)


def _lib(name):
    path = os.path.join(_OUT, name)
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
    return C.CDLL(path)


class Forest:
    """Owning handle on an OrcForest plus numpy copies of every array."""

    def __init__(self, lib, h, W, H):
        self._lib, self._h, self.W, self.H = lib, h, W, H
        N = W * H
        self.N = N
        self.T = lib.orc_forest_num_trees(h)
        nadj = lib.orc_forest_adj_size(h)
        self.ew = np.empty(2 * N, np.uint16)
        self.mask = np.empty(2 * N, np.uint8)
        self.fh_comp = np.empty(N, np.int32)
        self.tree_id = np.empty(N, np.int32)
        self.tree_start = np.empty(self.T + 1, np.int32)
        self.node_pixel = np.empty(N, np.int32)
        self.parent = np.empty(N, np.int32)
        self.child_begin = np.empty(N, np.int32)
        self.child_count = np.empty(N, np.int32)
        self.pw = np.empty(N, np.uint16)
        self.level = np.empty(N, np.int32)
        self.adj_ptr = np.empty(self.T + 1, np.int32)
        self.adj = np.empty(max(nadj, 1), np.int32)
        lib.orc_forest_get(h, _ptr(self.ew), _ptr(self.mask), _ptr(self.fh_comp), _ptr(self.tree_id), _ptr(self.tree_start),
                           _ptr(self.node_pixel), _ptr(self.parent), _ptr(self.child_begin), _ptr(self.child_count),
                           _ptr(self.pw), _ptr(self.level), _ptr(self.adj_ptr), _ptr(self.adj))
        self.adj = self.adj[:nadj]
        self.wlut = np.empty(766, np.float64)
        self.w2lut = np.empty(766, np.float64)
        lib.orc_forest_get_luts(h, _ptr(self.wlut), _ptr(self.w2lut))
        self.max_depth = lib.orc_forest_max_depth(h)

    def __del__(self):
        try:
            self._lib.orc_forest_free(self._h)
        except Exception:
            pass


class Oracle:
    def __init__(self, fast: bool = False):
        build()
        L = self.lib = _lib("liboracle_fast.so" if fast else "liboracle.so")
        L.orc_forest_build.restype = c_p
        L.orc_forest_build.argtypes = [c_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int]
        L.orc_forest_free.argtypes = [c_p]
        for f in (L.orc_forest_num_trees, L.orc_forest_adj_size, L.orc_forest_max_depth):
            f.argtypes = [c_p]
            f.restype = C.c_int
        L.orc_forest_get.argtypes = [c_p] * 14
        L.orc_forest_get_luts.argtypes = [c_p] * 3
        L.orc_median3.argtypes = [c_p, C.c_int, C.c_int, c_p]
        L.orc_edge_weights.argtypes = [c_p, C.c_int, C.c_int, c_p]
        L.orc_plane_init.argtypes = [C.c_int, C.c_int, C.c_int, c_p]
        L.orc_ingest.argtypes = [c_p, C.c_size_t, C.c_float, C.c_float, C.c_float]
        L.orc_cost_adgrad.argtypes = [c_p, c_p, C.c_int, C.c_int, C.c_int, c_p, c_p]
        L.orc_cost_adgrad_range.argtypes = [c_p, c_p, C.c_int, C.c_int, C.c_int, C.c_int, c_p, c_p]
        L.orc_eval_proposal.argtypes = [c_p, c_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, c_p, c_p, c_p]
        L.orc_eval_proposal.restype = C.c_int
        L.orc_pms_apply.argtypes = [c_p, c_p, C.c_int, c_p, c_p, C.c_int, c_p, c_p]
        L.orc_aggregate_label.argtypes = [c_p, c_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, c_p]
        L.orc_aggregate_dense.argtypes = [c_p, c_p, C.c_int, C.c_int, C.c_int, c_p, c_p, c_p]
        L.orc_pm_gradients.argtypes = [c_p, C.c_int, C.c_int, c_p]
        L.orc_pm_plane_cost_map.argtypes = [C.c_int, c_p, c_p, c_p, c_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float,
                                            C.c_float, C.c_float, C.c_float, c_p]
        L.orc_pms_apply_plane.argtypes = [c_p, C.c_int, c_p, c_p, c_p, c_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, c_p, c_p,
                                          C.c_int, c_p, c_p]
        L.orc_rand_new.restype = c_p
        L.orc_rand_new.argtypes = [C.c_uint]
        L.orc_rand_free.argtypes = [c_p]
        L.orc_rand_burn.argtypes = [c_p, C.c_size_t]
        L.orc_rand_next.argtypes = [c_p]
        L.orc_rand_next.restype = C.c_uint
        L.orc_mst_pms.argtypes = [c_p, c_p, C.c_int, c_p, c_p, c_p, c_p, c_p, C.c_int]
        L.orc_mst_pms.restype = C.c_int
        L.orc_label_to_disp.argtypes = [c_p, C.c_int, C.c_int, C.c_int, c_p]
        L.orc_lr_check.argtypes = [c_p, c_p, C.c_int, C.c_int, C.c_int, C.c_int, c_p]
        L.orc_stereo3dmst.argtypes = [c_p, c_p, C.c_int, C.c_int, c_p, c_p, C.c_int, C.c_int, C.c_int] + [c_p] * 6

    # -- stages ---------------------------------------------------------------------------------
    def median3(self, plane):
        plane = np.ascontiguousarray(plane, np.uint8)
        out = np.empty_like(plane)
        self.lib.orc_median3(_ptr(plane), plane.shape[1], plane.shape[0], _ptr(out))
        return out

    def forest(self, bgr, c=5000.0, min_size=200, gamma=1.0 / 12.0, median=True) -> Forest:
        bgr = np.ascontiguousarray(bgr, np.uint8)
        H, W, _ = bgr.shape
        h = self.lib.orc_forest_build(_ptr(bgr), W, H, C.c_float(c), min_size, C.c_float(np.float32(gamma)), int(median))
        f = Forest(self.lib, h, W, H)
        f._bgr = bgr
        return f

    def plane_init(self, W, H, Dmax):
        abc = np.empty((H * W, 3), np.float32)
        self.lib.orc_plane_init(W, H, Dmax, _ptr(abc))
        return abc

    def ingest(self, vol, cap=0.5, offset=0.0, scale=1.0):
        vol = np.ascontiguousarray(vol, np.float32).copy()
        self.lib.orc_ingest(_ptr(vol), vol.size, C.c_float(cap), C.c_float(offset), C.c_float(scale))
        return vol

    def cost_adgrad(self, left_bgr, right_bgr, D):
        left_bgr = np.ascontiguousarray(left_bgr, np.uint8)
        right_bgr = np.ascontiguousarray(right_bgr, np.uint8)
        H, W, _ = left_bgr.shape
        lv = np.empty((D, H * W), np.float32)
        rv = np.empty((D, H * W), np.float32)
        self.lib.orc_cost_adgrad(_ptr(left_bgr), _ptr(right_bgr), W, H, D, _ptr(lv), _ptr(rv))
        return lv, rv

    def cost_adgrad_range(self, left_bgr, right_bgr, d0, d1, views=(0, 1)):
        """Rows d0..d1-1 of the two volumes cost_adgrad returns (None for a view not in `views`): bench.py's sharded CPU arm."""
        left_bgr = np.ascontiguousarray(left_bgr, np.uint8)
        right_bgr = np.ascontiguousarray(right_bgr, np.uint8)
        H, W, _ = left_bgr.shape
        lv = np.empty((d1 - d0, H * W), np.float32) if 0 in views else None
        rv = np.empty((d1 - d0, H * W), np.float32) if 1 in views else None
        self.lib.orc_cost_adgrad_range(_ptr(left_bgr), _ptr(right_bgr), W, H, d0, d1, _ptr(lv), _ptr(rv))
        return lv, rv

    def aggregate_dense(self, F: Forest, vol, d0=0, d1=None, want_agg=False):
        vol = np.ascontiguousarray(vol, np.float32)
        D = vol.shape[0]
        d1 = D if d1 is None else d1
        agg = np.zeros((D, F.N), np.float64) if want_agg else None
        disp = np.empty(F.N, np.int32)
        best = np.empty(F.N, np.float64)
        self.lib.orc_aggregate_dense(F._h, _ptr(vol), D, d0, d1, _ptr(agg), _ptr(disp), _ptr(best))
        return disp, best, agg

    def eval_proposal(self, F: Forest, vol, Dmax, tree, label, min_cost, abc, agg=None):
        agg = np.zeros(F.N, np.float64) if agg is None else agg
        n = self.lib.orc_eval_proposal(F._h, _ptr(vol), Dmax, tree, C.c_float(label[0]), C.c_float(label[1]),
                                       C.c_float(label[2]), _ptr(min_cost), _ptr(abc), _ptr(agg))
        return n, agg

    def pms_apply(self, F: Forest, vol, Dmax, tree_ids, labels, min_cost, abc):
        tree_ids = np.ascontiguousarray(tree_ids, np.int32)
        labels = np.ascontiguousarray(labels, np.float32)
        self.lib.orc_pms_apply(F._h, _ptr(vol), Dmax, _ptr(tree_ids), _ptr(labels), len(tree_ids), _ptr(min_cost), _ptr(abc))

    def pm_gradients(self, bgr):
        """pm.cpp:70-88: Sobel/8 gradients of the BGR2GRAY image, float32 [H, W, 2]."""
        bgr = np.ascontiguousarray(bgr, np.uint8)
        H, W, _ = bgr.shape
        g = np.empty((H, W, 2), np.float32)
        self.lib.orc_pm_gradients(_ptr(bgr), W, H, _ptr(g))
        return g

    def plane_cost_map(self, view, left_bgr, right_bgr, label, Dmax, alpha=0.9, tau_c=10.0, tau_g=2.0, scale=1.0, oob=0.5):
        left_bgr = np.ascontiguousarray(left_bgr, np.uint8); right_bgr = np.ascontiguousarray(right_bgr, np.uint8)
        H, W, _ = left_bgr.shape
        gl, gr = self.pm_gradients(left_bgr), self.pm_gradients(right_bgr)
        out = np.empty(H * W, np.float32)
        self.lib.orc_pm_plane_cost_map(view, _ptr(left_bgr), _ptr(right_bgr), _ptr(gl), _ptr(gr), W, H, C.c_float(label[0]), C.c_float(label[1]),
                                       C.c_float(label[2]), Dmax, C.c_float(alpha), C.c_float(tau_c), C.c_float(tau_g), C.c_float(scale), C.c_float(oob), _ptr(out))
        return out

    def pms_apply_plane(self, F: Forest, view, left_bgr, right_bgr, Dmax, tree_ids, labels, min_cost, abc, alpha=0.9, tau_c=10.0, tau_g=2.0,
                        scale=1.0, oob=0.5):
        """pms_apply with the slanted-plane matching cost computed straight from the images (pm.cpp:97-154)."""
        left_bgr = np.ascontiguousarray(left_bgr, np.uint8); right_bgr = np.ascontiguousarray(right_bgr, np.uint8)
        gl, gr = self.pm_gradients(left_bgr), self.pm_gradients(right_bgr)
        tree_ids = np.ascontiguousarray(tree_ids, np.int32); labels = np.ascontiguousarray(labels, np.float32)
        self.lib.orc_pms_apply_plane(F._h, view, _ptr(left_bgr), _ptr(right_bgr), _ptr(gl), _ptr(gr), Dmax, C.c_float(alpha), C.c_float(tau_c),
                                     C.c_float(tau_g), C.c_float(scale), C.c_float(oob), _ptr(tree_ids), _ptr(labels), len(tree_ids), _ptr(min_cost), _ptr(abc))

    def rand_new(self, seed=1):
        return self.lib.orc_rand_new(seed)

    def mst_pms(self, F: Forest, vol, Dmax, min_cost, abc, grand, record_cap=0):
        rt = np.empty(record_cap, np.int32) if record_cap else None
        rl = np.empty((record_cap, 3), np.float32) if record_cap else None
        n = self.lib.orc_mst_pms(F._h, _ptr(vol), Dmax, _ptr(min_cost), _ptr(abc), grand, _ptr(rt), _ptr(rl), record_cap)
        if record_cap:
            assert n <= record_cap, (n, record_cap)
            return n, rt[:n], rl[:n]
        return n, None, None

    def label_to_disp(self, abc, W, H, Dmax):
        disp = np.empty(W * H, np.float32)
        self.lib.orc_label_to_disp(_ptr(abc), W, H, Dmax, _ptr(disp))
        return disp

    def lr_check(self, left, right, W, H, Dmax, fill):
        left = np.ascontiguousarray(left, np.float32).copy()
        right = np.ascontiguousarray(right, np.float32)
        mask = np.empty(W * H, np.uint8)
        self.lib.orc_lr_check(_ptr(left), _ptr(right), W, H, Dmax, int(fill), _ptr(mask))
        return left, mask

    def stereo3dmst(self, left_bgr, right_bgr, left_vol, right_vol, Dmax, num_iter=100, emulate_gui_rand=True):
        left_bgr = np.ascontiguousarray(left_bgr, np.uint8)
        right_bgr = np.ascontiguousarray(right_bgr, np.uint8)
        H, W, _ = left_bgr.shape
        N = W * H
        lv = np.ascontiguousarray(left_vol, np.float32)
        rv = np.ascontiguousarray(right_vol, np.float32)
        out = dict(left_disp=np.empty(N, np.float32), right_disp=np.empty(N, np.float32),
                   left_abc=np.empty((N, 3), np.float32), right_abc=np.empty((N, 3), np.float32),
                   left_min=np.empty(N, np.float64), right_min=np.empty(N, np.float64))
        self.lib.orc_stereo3dmst(_ptr(left_bgr), _ptr(right_bgr), W, H, _ptr(lv), _ptr(rv), Dmax, num_iter,
                                 int(emulate_gui_rand), _ptr(out["left_disp"]), _ptr(out["right_disp"]),
                                 _ptr(out["left_abc"]), _ptr(out["right_abc"]), _ptr(out["left_min"]), _ptr(out["right_min"]))
        return out


class Ref:
    """The reference's own src/Stereo3DMST.cpp (oracle/_ref/libref3dmst.so). Build container only."""

    @staticmethod
    def available() -> bool:
        return os.path.exists(os.path.join(_OUT, "libref3dmst.so"))

    def __init__(self):
        L = self.lib = _lib("libref3dmst.so")
        self.libc = C.CDLL(None)
        L.ref_segment_graph.argtypes = [C.c_int, C.c_int, c_p, c_p, c_p, c_p, C.c_float, c_p, c_p]
        L.ref_segment_graph.restype = C.c_int
        L.ref_view_build.restype = c_p
        L.ref_view_build.argtypes = [c_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_float]
        L.ref_view_free.argtypes = [c_p]
        L.ref_view_num_trees.argtypes = [c_p]
        L.ref_view_adj_size.argtypes = [c_p]
        L.ref_view_get.argtypes = [c_p] * 12
        L.ref_view_set_abc.argtypes = [c_p, c_p]
        L.ref_eval_proposal.argtypes = [c_p, c_p, C.c_int, C.c_float, C.c_float, C.c_float, c_p, c_p]
        L.ref_mst_pms.argtypes = [c_p, c_p, c_p, c_p]
        L.ref_label_to_disp.argtypes = [c_p, c_p]
        L.ref_lr_check.argtypes = [c_p, c_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_stereo3dmst.argtypes = [C.c_char_p, c_p, c_p, C.c_int, C.c_int, C.c_int, c_p, c_p]
        L.ref_stereo3dmst.restype = C.c_int

    def srand(self, seed=1):
        self.libc.srand(seed)

    def segment_graph(self, nv, w, a, b, c):
        w = np.ascontiguousarray(w, np.float64).copy()
        a = np.ascontiguousarray(a, np.int32).copy()
        b = np.ascontiguousarray(b, np.int32).copy()
        mask = np.zeros(len(w), np.int32)
        comp = np.empty(nv, np.int32)
        csz = np.empty(nv, np.int32)
        n = self.lib.ref_segment_graph(nv, len(w), _ptr(w), _ptr(a), _ptr(b), _ptr(mask), C.c_float(c), _ptr(comp), _ptr(csz))
        return n, w, a, b, mask, comp, csz

    def view(self, bgr, Dmax, c=5000.0, min_size=200, gamma=1.0 / 12.0):
        bgr = np.ascontiguousarray(bgr, np.uint8)
        H, W, _ = bgr.shape
        h = self.lib.ref_view_build(_ptr(bgr), W, H, C.c_float(c), min_size, Dmax, C.c_float(np.float32(gamma)))
        return RefView(self.lib, h, W, H, Dmax)

    def lr_check(self, left, right, W, H, Dmax, fill):
        left = np.ascontiguousarray(left, np.float32).copy()
        right = np.ascontiguousarray(right, np.float32).copy()
        self.lib.ref_lr_check(_ptr(left), _ptr(right), W, H, Dmax, int(fill))
        return left

    def stereo3dmst(self, workdir, left_bgr, right_bgr, left_vol, right_vol, Dmax):
        """Runs the reference entry point. Volumes ([Dmax][H*W] fp32) are written where mc-cnn would put them."""
        left_bgr = np.ascontiguousarray(left_bgr, np.uint8)
        right_bgr = np.ascontiguousarray(right_bgr, np.uint8)
        H, W, _ = left_bgr.shape
        mc = os.path.join(workdir, "mc-cnn-master")
        os.makedirs(mc, exist_ok=True)
        with open(os.path.join(mc, "main.lua"), "w") as f:
            f.write("#!/bin/sh\nexit 0\n")
        os.chmod(os.path.join(mc, "main.lua"), 0o755)
        np.ascontiguousarray(left_vol, np.float32).tofile(os.path.join(mc, "left.bin"))
        np.ascontiguousarray(right_vol, np.float32).tofile(os.path.join(mc, "right.bin"))
        dl = np.empty(W * H, np.float32)
        dr = np.empty(W * H, np.float32)
        rc = self.lib.ref_stereo3dmst(workdir.encode(), _ptr(left_bgr), _ptr(right_bgr), W, H, Dmax, _ptr(dl), _ptr(dr))
        assert rc == 0, rc
        return dl, dr


class RefView:
    def __init__(self, lib, h, W, H, Dmax):
        self.lib, self.h, self.W, self.H, self.Dmax = lib, h, W, H, Dmax
        N = self.N = W * H
        T = self.T = lib.ref_view_num_trees(h)
        nadj = lib.ref_view_adj_size(h)
        self.tree_start = np.empty(T + 1, np.int32)
        self.node_pixel = np.empty(N, np.int32)
        self.parent = np.empty(N, np.int32)
        self.child_begin = np.empty(N, np.int32)
        self.child_count = np.empty(N, np.int32)
        self.weight = np.empty(N, np.float64)
        self.weight2 = np.empty(N, np.float64)
        self.children4 = np.empty((N, 4), np.int32)
        self.adj_ptr = np.empty(T + 1, np.int32)
        self.adj = np.empty(max(nadj, 1), np.int32)
        self.abc = np.empty((N, 3), np.float32)
        lib.ref_view_get(h, _ptr(self.tree_start), _ptr(self.node_pixel), _ptr(self.parent), _ptr(self.child_begin),
                         _ptr(self.child_count), _ptr(self.weight), _ptr(self.weight2), _ptr(self.children4),
                         _ptr(self.adj_ptr), _ptr(self.adj), _ptr(self.abc))
        self.adj = self.adj[:nadj]

    def get_abc(self):
        self.lib.ref_view_get(self.h, _ptr(self.tree_start), _ptr(self.node_pixel), _ptr(self.parent), _ptr(self.child_begin),
                              _ptr(self.child_count), _ptr(self.weight), _ptr(self.weight2), _ptr(self.children4),
                              _ptr(self.adj_ptr), _ptr(np.empty(max(len(self.adj), 1), np.int32)), _ptr(self.abc))
        return self.abc.copy()

    def set_abc(self, abc):
        abc = np.ascontiguousarray(abc, np.float32)
        self.lib.ref_view_set_abc(self.h, _ptr(abc))

    def eval_proposal(self, vol, tree, label, min_cost, agg):
        self.lib.ref_eval_proposal(self.h, _ptr(vol), tree, C.c_float(label[0]), C.c_float(label[1]), C.c_float(label[2]),
                                   _ptr(min_cost), _ptr(agg))

    def mst_pms(self, vol, min_cost, agg):
        self.lib.ref_mst_pms(self.h, _ptr(vol), _ptr(min_cost), _ptr(agg))

    def label_to_disp(self):
        d = np.empty(self.N, np.float32)
        self.lib.ref_label_to_disp(self.h, _ptr(d))
        return d

    def __del__(self):
        try:
            self.lib.ref_view_free(self.h)
        except Exception:
            pass
