"""Host-side mirror of the reference interface for the Stereo3DMST path, over the C ABI (include/s3dmst.h).

The reference exposes one free function (include/Stereo3DMST.h:7)

    stereo3dmst(left_name, right_name, leftImg, rightImg, leftDisp, rightDisp, data_cost="MCCNN_acrt", Dmax=100)

taking cv::Mat images and filling two CV_32F disparity maps.  `stereo3dmst()` below keeps the name, the
argument order and the meaning (numpy arrays stand in for cv::Mat; the two file names are accepted and
ignored exactly as the reference only forwards them to mc-cnn's command line).  `data_cost` selects the
cost source: "MCCNN_acrt"/"MCCNN_fst" expect the caller to pass the mc-cnn volumes
(`left_volume=`, `right_volume=`, float32 [Dmax,H,W] as in left.bin/right.bin), "ADGRAD" builds the truncated
colour+gradient volume on the GPU.  Unknown selectors raise ValueError (the reference prints
"wrong data cost" and returns with the outputs unfilled, Stereo3DMST.cpp:756-759).

`Stereo3DMST` is the stage-level handle used by the parity tests and the bench.  There is no CPU fallback:
if the CUDA library is missing or no GPU is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libs3dmst.so")

c_p = C.c_void_p


class S3Params(C.Structure):
    _fields_ = [("fh_c", C.c_float), ("min_cc_size", C.c_int), ("gamma", C.c_float), ("median", C.c_int),
                ("cost_cap", C.c_float), ("cost_offset", C.c_float), ("cost_scale", C.c_float), ("oob_cost", C.c_float),
                ("num_iter", C.c_int), ("refine_floor", C.c_float), ("exact", C.c_int), ("keep_aggregated", C.c_int),
                ("agg_threads", C.c_int), ("agg_cache_nodes", C.c_int), ("agg_ring_nodes", C.c_int), ("agg_kernel", C.c_int),
                ("fh_ctas", C.c_int), ("fh_threads", C.c_int), ("fuse_cost", C.c_int), ("comm_p2p", C.c_int), ("fh_cluster", C.c_int), ("pms_cost_mode", C.c_int), ("pm_alpha", C.c_float), ("pm_tau_c", C.c_float),
                ("pm_tau_g", C.c_float), ("agg_cluster_nodes", C.c_int)]


# every symbol include/s3dmst.h declares (tests/test_abi.py checks the library exports all of them)
ABI_SYMBOLS = [
    "s3dmst_default_params", "s3dmst_create", "s3dmst_destroy", "s3dmst_last_error", "s3dmst_sync", "s3dmst_set_images", "s3dmst_set_images_async",
    "s3dmst_set_rectify_maps", "s3dmst_set_raw_images", "s3dmst_get_image", "s3dmst_remap_table",
    "s3dmst_build_forest", "s3dmst_forest_info", "s3dmst_get_forest", "s3dmst_set_forest", "s3dmst_build_cost_volume",
    "s3dmst_set_cost_volume", "s3dmst_get_cost_volume", "s3dmst_aggregate_dense", "s3dmst_get_aggregated",
    "s3dmst_dense_result_dev", "s3dmst_minloc_mask", "s3dmst_dense_to_disparity", "s3dmst_set_labels", "s3dmst_get_labels",
    "s3dmst_reset_min_cost", "s3dmst_get_min_cost", "s3dmst_pms_apply", "s3dmst_label_to_disp", "s3dmst_set_disparity",
    "s3dmst_get_disparity", "s3dmst_lr_check", "s3dmst_init_labels", "s3dmst_pms_iterate", "s3dmst_run", "s3dmst_run_dense", "s3dmst_run_dense_batch", "s3dmst_run_dense_batch_async", "s3dmst_batch_front", "s3dmst_batch_back", "s3dmst_reproject_to_3d", "s3dmst_comm_transport", "s3dmst_stage_ms", "s3dmst_stage_total_ms", "s3dmst_launch_count",
    "s3dmst_comm_unique_id", "s3dmst_comm_init", "s3dmst_comm_destroy", "s3dmst_comm_label_range", "s3dmst_reduce_minloc", "s3dmst_aggregate_dense_sharded",
    "s3dmst_comm_minloc_ms", "s3dmst_get_lr_mask", "s3dmst_weighted_median", "s3dmst_wmf_table", "s3dmst_norm_factor",
    "s3dmst_prepare_plane_cost", "s3dmst_get_plane_gradients",
]

_lib = None


def wmf_table(gamma=0.1):
    """float32 [766]: the weights exp(-sqrt(i) * gamma) of the weighted median filter (host-side call, no GPU needed)."""
    tab = np.empty(766, np.float32)
    load_library().s3dmst_wmf_table(C.c_float(gamma), tab.ctypes.data_as(c_p))
    return tab


def remap_table():
    """int16 [1024][4]: the fixed-point bilinear weights the device remap uses (host-side call, no GPU needed)."""
    tab = np.empty((1024, 4), np.int16)
    load_library().s3dmst_remap_table(tab.ctypes.data_as(c_p))
    return tab


def load_library():
    """dlopen the in-tree CUDA library; fails loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -m stereomatch_b200.build` (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.s3dmst_default_params.argtypes = [C.POINTER(S3Params)]
    L.s3dmst_create.argtypes = [C.POINTER(c_p), C.c_int, C.POINTER(S3Params), c_p]
    L.s3dmst_destroy.argtypes = [c_p]
    L.s3dmst_destroy.restype = None
    L.s3dmst_last_error.argtypes = [c_p]
    L.s3dmst_last_error.restype = C.c_char_p
    L.s3dmst_sync.argtypes = [c_p]
    L.s3dmst_set_images.argtypes = [c_p, c_p, c_p, C.c_int, C.c_int, C.c_int]
    L.s3dmst_set_images_async.argtypes = [c_p, c_p, c_p, C.c_int, C.c_int, C.c_int]
    L.s3dmst_set_rectify_maps.argtypes = [c_p, C.c_int, c_p, c_p, C.c_int, C.c_int]
    L.s3dmst_set_raw_images.argtypes = [c_p, c_p, c_p, C.c_int, C.c_int, C.c_int]
    L.s3dmst_get_image.argtypes = [c_p, C.c_int, c_p]
    L.s3dmst_remap_table.argtypes = [c_p]
    L.s3dmst_remap_table.restype = None
    L.s3dmst_build_forest.argtypes = [c_p, C.c_int]
    L.s3dmst_forest_info.argtypes = [c_p, C.c_int, c_p, c_p, c_p]
    L.s3dmst_get_forest.argtypes = [c_p, C.c_int] + [c_p] * 12
    L.s3dmst_set_forest.argtypes = [c_p, C.c_int, C.c_int, C.c_int, C.c_int, c_p, c_p, c_p, c_p]
    L.s3dmst_build_cost_volume.argtypes = [c_p, C.c_int, C.c_int]
    L.s3dmst_set_cost_volume.argtypes = [c_p, C.c_int, c_p, C.c_int, C.c_int]
    L.s3dmst_get_cost_volume.argtypes = [c_p, C.c_int, c_p]
    L.s3dmst_aggregate_dense.argtypes = [c_p, C.c_int, C.c_int, C.c_int, c_p, c_p]
    L.s3dmst_get_aggregated.argtypes = [c_p, C.c_int, c_p]
    L.s3dmst_dense_result_dev.argtypes = [c_p, C.c_int, C.POINTER(c_p), C.POINTER(c_p)]
    L.s3dmst_minloc_mask.argtypes = [c_p, C.c_int, c_p]
    L.s3dmst_dense_to_disparity.argtypes = [c_p, C.c_int]
    L.s3dmst_set_labels.argtypes = [c_p, C.c_int, c_p]
    L.s3dmst_get_labels.argtypes = [c_p, C.c_int, c_p]
    L.s3dmst_reset_min_cost.argtypes = [c_p, C.c_int]
    L.s3dmst_get_min_cost.argtypes = [c_p, C.c_int, c_p]
    L.s3dmst_pms_apply.argtypes = [c_p, C.c_int, c_p, c_p, C.c_size_t]
    L.s3dmst_label_to_disp.argtypes = [c_p, C.c_int]
    L.s3dmst_init_labels.argtypes = [c_p, C.c_int, C.c_int]
    L.s3dmst_pms_iterate.argtypes = [c_p, C.c_int, C.c_int, C.c_uint]
    L.s3dmst_set_disparity.argtypes = [c_p, C.c_int, c_p]
    L.s3dmst_get_disparity.argtypes = [c_p, C.c_int, c_p]
    L.s3dmst_lr_check.argtypes = [c_p, C.c_int]
    L.s3dmst_run_dense.argtypes = [c_p, C.c_int, C.c_int, c_p, c_p]
    L.s3dmst_run.argtypes = [c_p, C.c_int, C.c_uint, C.c_int, c_p, c_p]
    L.s3dmst_reproject_to_3d.argtypes = [c_p, c_p, C.c_float, C.c_int, c_p, c_p]
    L.s3dmst_run_dense_batch.argtypes = [C.POINTER(c_p), C.c_int, C.c_int, C.c_int, C.POINTER(c_p), C.POINTER(c_p)]
    L.s3dmst_run_dense_batch_async.argtypes = [C.POINTER(c_p), C.c_int, C.c_int, C.c_int, C.POINTER(c_p), C.POINTER(c_p)]
    L.s3dmst_batch_front.argtypes = [C.POINTER(c_p), C.c_int, C.c_int]
    L.s3dmst_batch_back.argtypes = [C.POINTER(c_p), C.c_int, C.c_int, C.c_int, C.POINTER(c_p), C.POINTER(c_p)]
    L.s3dmst_stage_ms.argtypes = [c_p, C.c_int]
    L.s3dmst_stage_ms.restype = C.c_double
    L.s3dmst_comm_transport.argtypes = [c_p]
    L.s3dmst_stage_total_ms.argtypes = [c_p, C.c_int, C.POINTER(C.c_int), C.c_int]
    L.s3dmst_stage_total_ms.restype = C.c_double
    L.s3dmst_launch_count.argtypes = [c_p]
    L.s3dmst_launch_count.restype = C.c_longlong
    L.s3dmst_prepare_plane_cost.argtypes = [c_p, C.c_int]
    L.s3dmst_get_plane_gradients.argtypes = [c_p, C.c_int, c_p]
    L.s3dmst_get_lr_mask.argtypes = [c_p, c_p]
    L.s3dmst_weighted_median.argtypes = [c_p, C.c_int, C.c_int, C.c_float, c_p]
    L.s3dmst_wmf_table.argtypes = [C.c_float, c_p]
    L.s3dmst_wmf_table.restype = None
    L.s3dmst_norm_factor.argtypes = [c_p, C.c_int, c_p]
    L.s3dmst_comm_unique_id.argtypes = [c_p]
    L.s3dmst_comm_init.argtypes = [c_p, c_p, C.c_int, C.c_int]
    L.s3dmst_comm_destroy.argtypes = [c_p]
    L.s3dmst_comm_label_range.argtypes = [c_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.s3dmst_reduce_minloc.argtypes = [c_p, C.c_int]
    L.s3dmst_aggregate_dense_sharded.argtypes = [c_p, C.c_int]
    L.s3dmst_comm_minloc_ms.argtypes = [c_p]
    L.s3dmst_comm_minloc_ms.restype = C.c_double
    _lib = L
    return L


def _prefer_torch_nccl():
    """The library binds NCCL at run time by SONAME (libnccl.so.2).  In a Python process that also uses torch, torch's
    bundled copy has to be the one the process holds: had the system copy been loaded first, a later `import torch` would
    be resolved against it and fail on newer symbols.  Importing torch first makes the order right; without torch the
    system copy is used."""
    try:
        import torch  # noqa: F401
    except Exception:
        pass


_cudart_lib = None


def _cudart():
    """The CUDA runtime the library itself links (for plain D2H copies of device pointers the ABI hands out)."""
    global _cudart_lib
    if _cudart_lib is None:
        load_library()
        for name in ("libcudart.so.12", "libcudart.so"):
            try:
                _cudart_lib = C.CDLL(name)
                break
            except OSError:
                continue
        if _cudart_lib is None:
            raise S3Error("libcudart not found")
        _cudart_lib.cudaMemcpy.argtypes = [c_p, c_p, C.c_size_t, C.c_int]
    return _cudart_lib


def default_params() -> S3Params:
    p = S3Params()
    load_library().s3dmst_default_params(C.byref(p))
    return p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(c_p)


T_FOREST, T_COST, T_AGG, T_POST, T_PMS = range(5)


class S3Error(RuntimeError):
    pass


class Stereo3DMST:
    """One context on one GPU (not thread-safe). Stages mirror src/Stereo3DMST.cpp's functions."""

    def __init__(self, device: int = 0, stream=None, **params):
        self.L = load_library()
        p = default_params()
        for k, v in params.items():
            if not hasattr(p, k):
                raise TypeError(f"unknown parameter {k}")
            setattr(p, k, v)
        self.params = p
        h = c_p()
        rc = self.L.s3dmst_create(C.byref(h), device, C.byref(p), c_p(stream) if stream else None)
        if rc != 0:
            raise S3Error(f"s3dmst_create failed ({rc}): {self.L.s3dmst_last_error(None).decode()}")
        self.h = h
        self.W = self.H = self.N = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.s3dmst_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise S3Error(f"s3dmst error {rc}: {self.L.s3dmst_last_error(self.h).decode()}")

    # -- inputs -------------------------------------------------------------------------------------
    def set_images(self, left_bgr, right_bgr, sync=True):
        """sync=False: s3dmst_set_images_async — the arrays (pinned for a truly asynchronous copy) must stay alive and
        untouched until the next synchronising call on this handle."""
        left_bgr = np.ascontiguousarray(left_bgr, np.uint8)
        right_bgr = np.ascontiguousarray(right_bgr, np.uint8)
        if left_bgr.ndim != 3 or left_bgr.shape[2] != 3 or left_bgr.shape != right_bgr.shape:
            raise ValueError("images must be two HxWx3 uint8 BGR arrays of the same size")
        self.H, self.W = left_bgr.shape[:2]
        self.N = self.W * self.H
        fn = self.L.s3dmst_set_images if sync else self.L.s3dmst_set_images_async
        self._ck(fn(self.h, _ptr(left_bgr), _ptr(right_bgr), self.W, self.H, 3 * self.W))
        if not sync:
            self._keep = (left_bgr, right_bgr)

    def set_rectify_maps(self, view, map_xy, map_fxy):
        """The CV_16SC2 / CV_16UC1 pair of cv2.initUndistortRectifyMap(..., cv2.CV_16SC2) for one view (kept on the device)."""
        map_xy = np.ascontiguousarray(map_xy, np.int16)
        map_fxy = np.ascontiguousarray(map_fxy, np.uint16)
        if map_xy.ndim != 3 or map_xy.shape[2] != 2 or map_fxy.shape != map_xy.shape[:2]:
            raise ValueError("maps must be int16 HxWx2 and uint16 HxW")
        self._ck(self.L.s3dmst_set_rectify_maps(self.h, view, _ptr(map_xy), _ptr(map_fxy), map_xy.shape[1], map_xy.shape[0]))
        self._map_size = (map_xy.shape[1], map_xy.shape[0])

    def set_raw_images(self, left_raw, right_raw):
        """Unrectified BGR pair -> rectified images on the device (cv2.remap INTER_LINEAR, bit-identical)."""
        left_raw = np.ascontiguousarray(left_raw, np.uint8)
        right_raw = np.ascontiguousarray(right_raw, np.uint8)
        if left_raw.ndim != 3 or left_raw.shape[2] != 3 or left_raw.shape != right_raw.shape:
            raise ValueError("images must be two HxWx3 uint8 BGR arrays of the same size")
        sh, sw = left_raw.shape[:2]
        self._ck(self.L.s3dmst_set_raw_images(self.h, _ptr(left_raw), _ptr(right_raw), sw, sh, 3 * sw))
        self.W, self.H = self._map_size
        self.N = self.W * self.H

    def get_image(self, view):
        out = np.empty((self.H, self.W, 3), np.uint8)
        self._ck(self.L.s3dmst_get_image(self.h, view, _ptr(out)))
        return out

    # -- forest -------------------------------------------------------------------------------------
    def build_forest(self, view):
        self._ck(self.L.s3dmst_build_forest(self.h, view))

    def forest_info(self, view, want_adj=False):
        t, d, a = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.L.s3dmst_forest_info(self.h, view, C.byref(t), C.byref(d), C.byref(a) if want_adj else None))
        return t.value, d.value, a.value

    def get_forest(self, view):
        T, depth, nadj = self.forest_info(view, want_adj=True)
        N = self.N
        out = dict(ew=np.empty(2 * N, np.uint16), mask=np.empty(2 * N, np.uint8), tree_id=np.empty(N, np.int32),
                   tree_start=np.empty(T + 1, np.int32), node_pixel=np.empty(N, np.int32), parent=np.empty(N, np.int32),
                   child_begin=np.empty(N, np.int32), child_count=np.empty(N, np.int32), pw=np.empty(N, np.uint16),
                   level=np.empty(N, np.int32), adj_ptr=np.empty(T + 1, np.int32), adj=np.empty(max(nadj, 1), np.int32))
        self._ck(self.L.s3dmst_get_forest(self.h, view, *[_ptr(out[k]) for k in (
            "ew", "mask", "tree_id", "tree_start", "node_pixel", "parent", "child_begin", "child_count", "pw", "level",
            "adj_ptr", "adj")]))
        out["adj"] = out["adj"][:nadj]
        out["T"], out["max_depth"] = T, depth
        return out

    def set_forest(self, view, W, H, tree_start, node_pixel, parent, pw):
        tree_start = np.ascontiguousarray(tree_start, np.int32)
        node_pixel = np.ascontiguousarray(node_pixel, np.int32)
        parent = np.ascontiguousarray(parent, np.int32)
        pw = np.ascontiguousarray(pw, np.uint16)
        self._ck(self.L.s3dmst_set_forest(self.h, view, W, H, len(tree_start) - 1, _ptr(tree_start), _ptr(node_pixel),
                                          _ptr(parent), _ptr(pw)))
        self.W, self.H, self.N = W, H, W * H

    # -- cost volume --------------------------------------------------------------------------------
    def build_cost_volume(self, D, ingest=False):
        self._ck(self.L.s3dmst_build_cost_volume(self.h, D, int(ingest)))
        self.D = D

    def set_cost_volume(self, view, vol, ingest=True):
        vol = np.ascontiguousarray(vol, np.float32)
        D = vol.shape[0]
        if vol.size != D * self.N:
            raise ValueError("volume must be float32 [D, H, W]")
        self._ck(self.L.s3dmst_set_cost_volume(self.h, view, _ptr(vol), D, int(ingest)))
        self.D = D

    def get_cost_volume(self, view):
        vol = np.empty((self.D, self.N), np.float32)
        self._ck(self.L.s3dmst_get_cost_volume(self.h, view, _ptr(vol)))
        return vol

    # -- dense mode ---------------------------------------------------------------------------------
    def aggregate_dense(self, view, d0=0, d1=None, fetch=True):
        d1 = self.D if d1 is None else d1
        disp = np.empty(self.N, np.int32) if fetch else None
        best = np.empty(self.N, np.float64) if fetch else None
        self._ck(self.L.s3dmst_aggregate_dense(self.h, view, d0, d1, _ptr(disp), _ptr(best)))
        return disp, best

    def get_aggregated(self, view):
        agg = np.empty((self.D, self.N), np.float64)
        self._ck(self.L.s3dmst_get_aggregated(self.h, view, _ptr(agg)))
        return agg

    def dense_result_dev(self, view):
        a, b = c_p(), c_p()
        self._ck(self.L.s3dmst_dense_result_dev(self.h, view, C.byref(a), C.byref(b)))
        return a.value, b.value

    def minloc_mask(self, view, global_min_dev_ptr):
        self._ck(self.L.s3dmst_minloc_mask(self.h, view, c_p(global_min_dev_ptr)))

    # -- label-range sharding over NCCL (one context per GPU / process) ------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        """128 bytes (ncclUniqueId) made by rank 0; hand them to the other ranks (e.g. torch.distributed broadcast)."""
        _prefer_torch_nccl()
        buf = C.create_string_buffer(128)
        rc = load_library().s3dmst_comm_unique_id(buf)
        if rc != 0:
            raise S3Error(f"s3dmst_comm_unique_id failed ({rc}): NCCL could not be loaded")
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        _prefer_torch_nccl()
        self._ck(self.L.s3dmst_comm_init(self.h, C.c_char_p(unique_id), int(rank), int(nranks)))

    def comm_destroy(self):
        self._ck(self.L.s3dmst_comm_destroy(self.h))

    def comm_label_range(self, D):
        a, b = C.c_int(), C.c_int()
        self._ck(self.L.s3dmst_comm_label_range(self.h, int(D), C.byref(a), C.byref(b)))
        return a.value, b.value

    def reduce_minloc(self, view):
        self._ck(self.L.s3dmst_reduce_minloc(self.h, view))

    def aggregate_dense_sharded(self, D):
        """This rank's label range of both views + the MIN-LOC reduction (s3dmst_aggregate_dense_sharded): asynchronous."""
        self._ck(self.L.s3dmst_aggregate_dense_sharded(self.h, int(D)))

    def comm_minloc_ms(self):
        return self.L.s3dmst_comm_minloc_ms(self.h)

    def get_dense_result(self, view):
        """(disparity int32 [N], best cost float64 [N]) of the last dense / sharded call, copied to the host."""
        import ctypes
        pb, pd = self.dense_result_dev(view)
        disp = np.empty(self.N, np.int32); best = np.empty(self.N, np.float64)
        self.sync()
        cudart = _cudart()
        for dst, src in ((disp, pd), (best, pb)):
            rc = cudart.cudaMemcpy(dst.ctypes.data_as(c_p), c_p(src), ctypes.c_size_t(dst.nbytes), 2)
            if rc != 0:
                raise S3Error(f"cudaMemcpy D2H failed ({rc})")
        return disp, best

    def dense_to_disparity(self, view):
        self._ck(self.L.s3dmst_dense_to_disparity(self.h, view))

    def prepare_plane_cost(self, Dmax):
        """Data term of pms_cost_mode = 1 (slanted-plane colour + gradient cost from the images, pm.cpp:97-154)."""
        self._ck(self.L.s3dmst_prepare_plane_cost(self.h, int(Dmax)))
        self.D = Dmax

    def get_plane_gradients(self, view):
        g = np.empty((self.H, self.W, 2), np.float32)
        self._ck(self.L.s3dmst_get_plane_gradients(self.h, view, _ptr(g)))
        return g

    # -- PatchMatch state ---------------------------------------------------------------------------
    def set_labels(self, view, abc):
        abc = np.ascontiguousarray(abc, np.float32)
        self._ck(self.L.s3dmst_set_labels(self.h, view, _ptr(abc)))

    def get_labels(self, view):
        abc = np.empty((self.N, 3), np.float32)
        self._ck(self.L.s3dmst_get_labels(self.h, view, _ptr(abc)))
        return abc

    def reset_min_cost(self, view):
        self._ck(self.L.s3dmst_reset_min_cost(self.h, view))

    def get_min_cost(self, view):
        mc = np.empty(self.N, np.float64)
        self._ck(self.L.s3dmst_get_min_cost(self.h, view, _ptr(mc)))
        return mc

    def pms_apply(self, view, tree_ids, labels):
        tree_ids = np.ascontiguousarray(tree_ids, np.int32)
        labels = np.ascontiguousarray(labels, np.float32)
        self._ck(self.L.s3dmst_pms_apply(self.h, view, _ptr(tree_ids), _ptr(labels), len(tree_ids)))

    def init_labels(self, view, Dmax):
        """The reference's random plane initialisation (Stereo3DMST.cpp:390-430); resets min_cost."""
        self._ck(self.L.s3dmst_init_labels(self.h, view, int(Dmax)))

    def pms_iterate(self, view, n_iter, seed=1):
        self._ck(self.L.s3dmst_pms_iterate(self.h, view, int(n_iter), int(seed)))

    def label_to_disp(self, view):
        self._ck(self.L.s3dmst_label_to_disp(self.h, view))

    # -- disparity maps -----------------------------------------------------------------------------
    def set_disparity(self, view, disp):
        disp = np.ascontiguousarray(disp, np.float32)
        self._ck(self.L.s3dmst_set_disparity(self.h, view, _ptr(disp)))

    def get_disparity(self, view):
        d = np.empty(self.N, np.float32)
        self._ck(self.L.s3dmst_get_disparity(self.h, view, _ptr(d)))
        return d

    def lr_check(self, fill=False):
        self._ck(self.L.s3dmst_lr_check(self.h, int(fill)))

    def get_lr_mask(self):
        m = np.empty(self.N, np.uint8)
        self._ck(self.L.s3dmst_get_lr_mask(self.h, _ptr(m)))
        return m

    def weighted_median(self, view=0, radius=10, gamma=0.1, mask=None):
        """weightedMedianFilter (PatchMatchStereoGPU.cu:2436-2599) on the pixels of `mask` (None: the LR check's, left view)."""
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
        self._ck(self.L.s3dmst_weighted_median(self.h, view, int(radius), float(gamma), _ptr(mask)))

    def norm_factor(self, view):
        nf = np.empty(self.N, np.float64)
        self._ck(self.L.s3dmst_norm_factor(self.h, view, _ptr(nf)))
        return nf

    def reproject_to_3d(self, Q, disp_floor=10.0, handle_missing=True, want_rgb=True):
        """stereo_Yin.cpp:218-243: floor the left disparity map, cv::reprojectImageTo3D with Q, point-cloud colours."""
        Q = np.ascontiguousarray(Q, np.float64).reshape(16)
        xyz = np.empty((self.N, 3), np.float32)
        rgb = np.empty(self.N, np.uint32) if want_rgb else None
        self._ck(self.L.s3dmst_reproject_to_3d(self.h, _ptr(Q), float(disp_floor), int(handle_missing), _ptr(xyz), _ptr(rgb)))
        return xyz, rgb

    def run_dense(self, D, fill=False, fetch=True):
        dl = np.empty(self.N, np.float32) if fetch else None
        dr = np.empty(self.N, np.float32) if fetch else None
        self._ck(self.L.s3dmst_run_dense(self.h, D, int(fill), _ptr(dl), _ptr(dr)))
        self.D = D
        return dl, dr

    def run(self, Dmax, seed=1, fill=False, fetch=True):
        """The reference's pipeline (plane init, num_iter rounds of MST_PMS per view, LabelToDisp, LR check): s3dmst_run."""
        dl = np.empty(self.N, np.float32) if fetch else None
        dr = np.empty(self.N, np.float32) if fetch else None
        self._ck(self.L.s3dmst_run(self.h, int(Dmax), int(seed), int(fill), _ptr(dl), _ptr(dr)))
        self.D = Dmax
        return dl, dr

    def sync(self):
        self._ck(self.L.s3dmst_sync(self.h))

    def stage_ms(self, stage):
        return self.L.s3dmst_stage_ms(self.h, stage)

    def comm_transport(self):
        """1 = MIN-LOC over peer memory, 0 = NCCL all-reduces, -1 = no communicator (s3dmst_comm_transport)."""
        return self.L.s3dmst_comm_transport(self.h)

    def stage_total_ms(self, stage, reset=False):
        """(accumulated ms, samples) of a stage over every call since the last reset (s3dmst_stage_total_ms)."""
        n = C.c_int()
        ms = self.L.s3dmst_stage_total_ms(self.h, stage, C.byref(n), int(reset))
        return ms, n.value

    def launch_count(self):
        return self.L.s3dmst_launch_count(self.h)


def run_dense_batch(engines, D, fill=False, fetch=True, out=None, wait=True):
    """s3dmst_run_dense_batch over a list of Stereo3DMST handles (one frame each, images already set): forest and cost
    stages of the frames overlap, one aggregation launch covers every frame's trees.  Returns [(left, right), ...]
    (float32 [N]) when fetch, else None.  `out` = optional [(left, right), ...] of preallocated (e.g. pinned) buffers.
    wait=False (s3dmst_run_dense_batch_async): the copies into `out` are only queued; they are complete after
    `sync()` on the frame's handle."""
    n = len(engines)
    L = engines[0].L
    hs = (c_p * n)(*[e.h for e in engines])
    if out is None and fetch:
        out = [(np.empty(e.N, np.float32), np.empty(e.N, np.float32)) for e in engines]
    if out is not None:
        pl = (c_p * n)(*[c_p(_addr(o[0])) for o in out])
        pr = (c_p * n)(*[c_p(_addr(o[1])) for o in out])
    else:
        pl = pr = None
    rc = (L.s3dmst_run_dense_batch if wait else L.s3dmst_run_dense_batch_async)(hs, n, int(D), int(fill), pl, pr)
    if rc != 0:
        raise S3Error(f"s3dmst error {rc}: {L.s3dmst_last_error(engines[0].h).decode()}")
    for e in engines:
        e.D = D
    return out


def batch_front(engines, D):
    """Forests + cost volumes of every frame (first half of run_dense_batch)."""
    n = len(engines)
    hs = (c_p * n)(*[e.h for e in engines])
    rc = engines[0].L.s3dmst_batch_front(hs, n, int(D))
    if rc != 0:
        raise S3Error(f"s3dmst error {rc}: {engines[0].L.s3dmst_last_error(engines[0].h).decode()}")


def batch_back(engines, D, fill=False, out=None):
    """Joint aggregation + LR check (+ copies into `out`) of frames whose batch_front has run."""
    n = len(engines)
    hs = (c_p * n)(*[e.h for e in engines])
    if out is not None:
        pl = (c_p * n)(*[c_p(_addr(o[0])) for o in out])
        pr = (c_p * n)(*[c_p(_addr(o[1])) for o in out])
    else:
        pl = pr = None
    rc = engines[0].L.s3dmst_batch_back(hs, n, int(D), int(fill), pl, pr)
    if rc != 0:
        raise S3Error(f"s3dmst error {rc}: {engines[0].L.s3dmst_last_error(engines[0].h).decode()}")
    for e in engines:
        e.D = D
    return out


def _addr(a):
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return a.ctypes.data


def stereo3dmst(left_name, right_name, leftImg, rightImg, leftDisp=None, rightDisp=None, data_cost="MCCNN_acrt", Dmax=100,
                *, left_volume=None, right_volume=None, proposals=None, mode="dense", fill=False, device=0):
    """Drop-in for the reference's stereo3dmst() (include/Stereo3DMST.h:7).

    Returns (leftDisp, rightDisp) as float32 [H, W] (also written into the arrays passed in, like the
    reference fills its cv::Mat outputs).  mode="dense" runs the dense-label pipeline;
    mode="pms" replays `proposals` = ((left_tree_ids, left_labels), (right_tree_ids, right_labels)), an injected
    proposal sequence, from the reference's random plane initialisation given as `init_labels` in proposals[2:].
    """
    del left_name, right_name  # only forwarded to mc-cnn's command line by the reference (:733-748)
    if data_cost not in ("MCCNN_acrt", "MCCNN_fst", "ADGRAD"):
        raise ValueError("wrong data cost")
    leftImg = np.ascontiguousarray(leftImg, np.uint8)
    rightImg = np.ascontiguousarray(rightImg, np.uint8)
    H, W = leftImg.shape[:2]
    eng = Stereo3DMST(device=device, cost_offset=1.0 if data_cost == "MCCNN_fst" else 0.0,
                      cost_scale=0.5 if data_cost == "MCCNN_fst" else 1.0)
    try:
        eng.set_images(leftImg, rightImg)
        eng.build_forest(0)
        eng.build_forest(1)
        if data_cost == "ADGRAD":
            eng.build_cost_volume(Dmax, ingest=(mode == "pms"))
        else:
            if left_volume is None or right_volume is None:
                raise ValueError("MCCNN_* cost sources need left_volume/right_volume (mc-cnn's left.bin/right.bin)")
            eng.set_cost_volume(0, np.asarray(left_volume, np.float32).reshape(Dmax, -1), ingest=True)
            eng.set_cost_volume(1, np.asarray(right_volume, np.float32).reshape(Dmax, -1), ingest=True)
        if mode == "dense":
            for v in (0, 1):
                eng.aggregate_dense(v, 0, Dmax, fetch=False)
                eng.dense_to_disparity(v)
        elif mode == "pms":
            (lt, ll), (rt, rl), linit, rinit = proposals
            for v, (tids, labs, init) in enumerate(((lt, ll, linit), (rt, rl, rinit))):
                eng.set_labels(v, init)
                eng.reset_min_cost(v)
                eng.pms_apply(v, tids, labs)
                eng.label_to_disp(v)
        else:
            raise ValueError("mode must be 'dense' or 'pms'")
        eng.lr_check(fill=fill)
        dl = eng.get_disparity(0).reshape(H, W)
        dr = eng.get_disparity(1).reshape(H, W)
    finally:
        eng.close()
    if leftDisp is not None:
        leftDisp[...] = dl
    if rightDisp is not None:
        rightDisp[...] = dr
    return dl, dr
