"""ctypes binding of ``libhfnet_b200.so`` (include/hfnet_b200.h).

This is the same C-ABI the reference-side C++ shim (include/HFNetB200Model.h) binds; the Python
host layer in this package (extractor.py, matcher.py, keyframe_database.py, optimizer.py) goes through it and never
falls back to a CPU path: if the library is missing or the device is not a B200 the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional

import numpy as np

HFB_DESC_DIM = 256
HFB_GLOBAL_DIM = 4096
HFB_MAX_LEVELS = 8
# HFB_LIB: A/B experiments against another build of the same library (development only)
LIB_PATH = Path(os.environ.get("HFB_LIB") or Path(__file__).resolve().parent / "libhfnet_b200.so")

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)


class HfbError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"hfnet_b200 status {status}: {message}")
        self.status = status


class hfb_config(C.Structure):
    _fields_ = [("device", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("n_levels", C.c_int32),
                ("scale_factor", C.c_float), ("max_keypoints", C.c_int32), ("max_batch", C.c_int32),
                ("with_global", C.c_int32)]


class hfb_features(C.Structure):
    _fields_ = [("x", _f32p), ("y", _f32p), ("response", _f32p), ("octave", _i32p), ("descriptors", _f32p),
                ("global_descriptor", _f32p), ("n_per_level", C.c_int32 * HFB_MAX_LEVELS), ("n_total", C.c_int32)]


class hfb_lba_problem(C.Structure):
    _fields_ = [("n_cams", C.c_int32), ("n_points", C.c_int32), ("n_edges", C.c_int32), ("poses", _f64p),
                ("fixed", _u8p), ("points", _f64p), ("edge_cam", _i32p), ("edge_point", _i32p), ("obs", _f64p),
                ("inv_sigma2", _f64p), ("K", C.c_float * 4), ("huber_delta", C.c_double)]


class hfb_lba_stats(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("trials", C.c_int32), ("initial_chi2", C.c_double),
                ("final_chi2", C.c_double), ("lambda_", C.c_double), ("n_opt_cams", C.c_int32),
                ("gpu_launches", C.c_int32)]


# name -> (restype, argtypes); every symbol include/hfnet_b200.h declares
SIGNATURES = {
    "hfb_create": (C.c_int, [C.POINTER(hfb_config), C.POINTER(C.c_void_p)]),
    "hfb_destroy": (None, [C.c_void_p]),
    "hfb_last_error": (C.c_char_p, [C.c_void_p]),
    "hfb_version": (C.c_int, []),
    "hfb_sync": (C.c_int, [C.c_void_p]),
    "hfb_stream": (C.c_void_p, [C.c_void_p]),
    "hfb_launch_count": (C.c_uint64, [C.c_void_p]),
    "hfb_host_alloc": (C.c_void_p, [C.c_size_t]),
    "hfb_host_free": (None, [C.c_void_p]),
    "hfb_load_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "hfb_extract": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, _i32p, C.c_float,
                              C.POINTER(hfb_features)]),
    "hfb_extract_level": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                    C.POINTER(hfb_features)]),
    "hfb_extract_batch": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_int32, _i32p, C.c_float,
                                    C.POINTER(hfb_features)]),
    "hfb_extract_batch_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, _i32p, C.c_float]),
    "hfb_extract_match_batch": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_int32, _i32p, C.c_float,
                                          C.POINTER(hfb_features), C.c_int32, C.c_float, _i32p, _f32p]),
    "hfb_extract_match_batch_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, _i32p, C.c_float, C.c_int32, C.c_float]),
    "hfb_fetch_features": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(hfb_features)]),
    "hfb_nms": (C.c_int, [C.c_void_p, _f32p, C.c_int32, C.c_int32, _f32p]),
    "hfb_select_sample": (C.c_int, [C.c_void_p, _f32p, C.c_int32, C.c_int32, _f32p, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_float, _f32p, _f32p, _f32p, _f32p, _i32p]),
    "hfb_resize_linear_u8": (C.c_int, [C.c_void_p, _u8p, C.c_int32, C.c_int32, _u8p, C.c_int32, C.c_int32]),
    "hfb_set_camera": (C.c_int, [C.c_void_p, _f32p, _f32p, C.c_int32]),
    "hfb_undistort_points": (C.c_int, [C.c_void_p, _f32p, _f32p, C.c_int32, _f32p, _f32p]),
    "hfb_fetch_undistorted": (C.c_int, [C.c_void_p, C.c_int32, _f32p, _f32p, C.c_int32]),
    "hfb_image_bounds": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _f32p]),
    "hfb_debug_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32, C.c_int32, _f32p, C.c_size_t,
                                   C.POINTER(C.c_size_t), _i32p]),
    "hfb_debug_gemm": (C.c_int, [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, _f32p,
                                 C.c_int, C.c_int, C.c_int, C.c_int, _f32p]),
    "hfb_match_mutual_l2": (C.c_int, [C.c_void_p, _f32p, C.c_int32, _f32p, C.c_int32, C.c_float, _i32p, _f32p, _i32p]),
    "hfb_match_mutual_cos": (C.c_int, [C.c_void_p, _f32p, C.c_int32, _f32p, C.c_int32, C.c_float, _i32p, _f32p, _i32p]),
    "hfb_match_batch": (C.c_int, [C.c_void_p, C.c_int32, _f32p, C.c_int32, _f32p, C.c_int32, C.c_int32, _i32p, _i32p,
                                  _i32p, _i32p, C.c_float, _i32p, _f32p]),
    "hfb_match_batch_dev": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                      C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p]),
    "hfb_match_projection": (C.c_int, [C.c_void_p, _f32p, C.c_int32, _f32p, _f32p, _i32p, _i32p, _f32p, C.c_int32, _f32p,
                                       _i32p, _u8p, _i32p, _f32p, _i32p]),
    "hfb_match_projection_gated": (C.c_int, [C.c_void_p, _f32p, C.c_int32, _f32p, _f32p, _i32p, _i32p, _f32p, C.c_int32,
                                             _f32p, _i32p, _u8p, _f32p, C.c_float, _i32p, _f32p, _i32p]),
    "hfb_match_projection_frame": (C.c_int, [C.c_void_p, C.c_int32, _i32p, C.c_int32, _f32p, _f32p, _i32p, _i32p, C.c_int32,
                                             _u8p, _f32p, C.c_float, _i32p, _f32p, _i32p]),
    "hfb_match_consecutive_dev": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_float]),
    "hfb_fetch_matches": (C.c_int, [C.c_void_p, C.c_int32, _i32p, _f32p, C.c_int32]),
    "hfb_set_stream_mode": (C.c_int, [C.c_void_p, C.c_int32]),
    "hfb_reset_stream": (C.c_int, [C.c_void_p]),
    "hfb_kfstore_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "hfb_kfstore_destroy": (None, [C.c_void_p]),
    "hfb_kfstore_put_frame": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32]),
    "hfb_kfstore_put": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, _f32p, C.c_int32]),
    "hfb_kfstore_erase": (C.c_int, [C.c_void_p, C.c_int64]),
    "hfb_kfstore_size": (C.c_int32, [C.c_void_p]),
    "hfb_match_kf_neighbours": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, _i64p, C.c_int32, C.c_int32, C.c_float, _i32p, _f32p,
                                          _i32p]),
    "hfb_distinctive_descriptors": (C.c_int, [C.c_void_p, _f32p, _i32p, C.c_int32, _i32p, _f32p]),
    "hfb_match_consecutive": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_float, _i32p, _f32p]),
    "hfb_profile_extract": (C.c_int, [C.c_void_p, C.c_int32, _i32p, C.c_float, C.c_char_p, C.c_size_t]),
    "hfb_kfdb_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "hfb_kfdb_destroy": (None, [C.c_void_p]),
    "hfb_kfdb_add": (C.c_int, [C.c_void_p, _i64p, _f32p, C.c_int32]),
    "hfb_kfdb_add_dev": (C.c_int, [C.c_void_p, _i64p, C.c_void_p, C.c_int32]),
    "hfb_kfdb_erase": (C.c_int, [C.c_void_p, C.c_int64]),
    "hfb_kfdb_clear": (C.c_int, [C.c_void_p]),
    "hfb_kfdb_size": (C.c_int32, [C.c_void_p]),
    "hfb_kfdb_add_tagged": (C.c_int, [C.c_void_p, _i64p, _i64p, _f32p, C.c_int32]),
    "hfb_kfdb_clear_map": (C.c_int, [C.c_void_p, C.c_int64]),
    "hfb_kfdb_query_batch": (C.c_int, [C.c_void_p, _f32p, C.c_int32, C.c_float, C.c_float, C.c_int32, _i64p, _f32p, _i32p,
                                       _f32p]),
    "hfb_kfdb_query_batch_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_float]),
    "hfb_kfdb_query": (C.c_int, [C.c_void_p, _f32p, C.c_float, C.c_float, _i64p, _f32p, C.c_int32, _i32p, _f32p]),
    "hfb_kfdb_scores_of": (C.c_int, [C.c_void_p, _i64p, C.c_int32, _f32p]),
    "hfb_kfdb_scan_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "hfb_kfdb_query_shard": (C.c_int, [C.c_void_p, _f32p, C.c_float, C.c_float, C.c_int32, C.c_void_p]),
    "hfb_kfdb_shard_setup": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "hfb_kfdb_shard_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hfb_kfdb_shard_connect_local": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "hfb_kfdb_query_sharded_begin": (C.c_int, [C.c_void_p, _f32p, C.c_float, C.c_float]),
    "hfb_kfdb_query_sharded_end": (C.c_int, [C.c_void_p, _i64p, _f32p, C.c_int32, _i32p, _f32p, _i32p]),
    "hfb_kfdb_query_sharded": (C.c_int, [C.c_void_p, _f32p, C.c_float, C.c_float, _i64p, _f32p, C.c_int32, _i32p, _f32p,
                                         _i32p]),
    "hfb_lba_optimize": (C.c_int, [C.c_void_p, C.POINTER(hfb_lba_problem), C.c_int32, C.c_double, _u8p, _f64p, _f64p,
                                   _f64p, _u8p, C.POINTER(hfb_lba_stats)]),
    "hfb_pose_optimize": (C.c_int, [C.c_void_p, _f32p, _f64p, C.c_int32, _f64p, _f64p, _f64p, _f64p, _u8p, _i32p, _i32p]),
    "hfb_lba_build_schur": (C.c_int, [C.c_void_p, C.POINTER(hfb_lba_problem), C.c_double, _f64p, _f64p, _f64p, _i32p]),
}

_lib: Optional[C.CDLL] = None


def load(path: Optional[Path] = None) -> C.CDLL:
    """dlopen the library and bind every exported entry point.  Raises if the library or a symbol is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise FileNotFoundError(f"{p} not found: build it with `python -m hfnet_slam_b200.build` "
                                "(or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(str(p))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


class PinnedBuffer:
    """Page-locked host memory (hfb_host_alloc) exposed as numpy arrays: DMA source / target without staging copies."""

    def __init__(self, nbytes: int):
        self.lib = load()
        self.nbytes = int(nbytes)
        self.ptr = self.lib.hfb_host_alloc(self.nbytes)
        if not self.ptr:
            raise HfbError(2, "hfb_host_alloc failed")
        self._buf = (C.c_uint8 * self.nbytes).from_address(self.ptr)

    def array(self, offset: int, shape, dtype) -> np.ndarray:
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        return np.frombuffer(self._buf, dtype=dtype, count=int(np.prod(shape)), offset=offset).reshape(shape)

    def close(self):
        if getattr(self, "ptr", None):
            self._buf = None
            self.lib.hfb_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array over page-locked memory; the buffer lives as long as the array (kept on ``arr.base`` chain)."""
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    buf = PinnedBuffer(max(nbytes, 16))
    arr = buf.array(0, shape, dtype)
    _PINNED_KEEPALIVE[arr.ctypes.data] = buf
    return arr


_PINNED_KEEPALIVE = {}


def as_f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def ptr(a: np.ndarray, typ):
    return a.ctypes.data_as(typ)


class Context:
    """Owns one ``hfb_ctx`` (one CUDA stream + workspaces)."""

    def __init__(self, height: int = 480, width: int = 752, n_levels: int = 1, scale_factor: float = 1.2,
                 max_keypoints: int = 1024, max_batch: int = 1, with_global: bool = True, device: int = 0):
        self.lib = load()
        self.cfg = hfb_config(device, height, width, n_levels, scale_factor, max_keypoints, max_batch,
                              1 if with_global else 0)
        self.handle = C.c_void_p()
        st = self.lib.hfb_create(C.byref(self.cfg), C.byref(self.handle))
        if st != 0:
            msg = self.lib.hfb_last_error(self.handle).decode() if self.handle else "hfb_create failed"
            if self.handle:
                self.lib.hfb_destroy(self.handle)
                self.handle = C.c_void_p()
            raise HfbError(st, msg)
        self.height, self.width, self.n_levels = height, width, n_levels
        self.max_keypoints, self.max_batch, self.with_global = max_keypoints, max_batch, with_global
        self.kp_cap = n_levels * max_keypoints

    def check(self, st: int):
        if st != 0:
            raise HfbError(st, self.lib.hfb_last_error(self.handle).decode())

    def close(self):
        if getattr(self, "handle", None):
            self.lib.hfb_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def sync(self):
        self.check(self.lib.hfb_sync(self.handle))

    @property
    def launch_count(self) -> int:
        return int(self.lib.hfb_launch_count(self.handle))

    @property
    def stream(self) -> int:
        return int(self.lib.hfb_stream(self.handle) or 0)

    def load_weights(self, blob: bytes):
        buf = C.create_string_buffer(blob, len(blob))
        self.check(self.lib.hfb_load_weights(self.handle, C.cast(buf, C.c_void_p), len(blob)))

    # ------------------------------------------------------------------------------------------ extraction
    def _alloc_features(self, n: int = 1, pinned: bool = False, descriptors: bool = True):
        """One contiguous host block per field for n frames (frame b's rows start at b * kp_cap).  pinned=True reuses a
        page-locked block owned by the context (results are valid until the next pinned call).  descriptors=False: the
        struct carries a NULL descriptor pointer -- the local descriptors stay resident in HBM."""
        cap = self.kp_cap
        if pinned:
            cache = self.__dict__.setdefault("_pinned_out", {})
            key = (n, descriptors)
            if key not in cache:
                cache[key] = dict(x=pinned_empty((n, cap), np.float32), y=pinned_empty((n, cap), np.float32),
                                  response=pinned_empty((n, cap), np.float32), octave=pinned_empty((n, cap), np.int32),
                                  descriptors=pinned_empty((n, cap if descriptors else 0, HFB_DESC_DIM), np.float32),
                                  global_descriptor=pinned_empty((n, HFB_GLOBAL_DIM), np.float32))
            arrs = cache[key]
        else:
            arrs = dict(x=np.empty((n, cap), np.float32), y=np.empty((n, cap), np.float32),
                        response=np.empty((n, cap), np.float32), octave=np.empty((n, cap), np.int32),
                        descriptors=np.empty((n, cap if descriptors else 0, HFB_DESC_DIM), np.float32),
                        global_descriptor=np.empty((n, HFB_GLOBAL_DIM), np.float32))
        if pinned and "_feats" in arrs:
            return arrs["_feats"], arrs          # same page-locked arrays, same struct array: nothing to rebuild
        feats = (hfb_features * n)()
        for b in range(n):
            f = feats[b]
            f.x, f.y = ptr(arrs["x"][b], _f32p), ptr(arrs["y"][b], _f32p)
            f.response, f.octave = ptr(arrs["response"][b], _f32p), ptr(arrs["octave"][b], _i32p)
            f.descriptors = ptr(arrs["descriptors"][b], _f32p) if descriptors else None
            f.global_descriptor = ptr(arrs["global_descriptor"][b], _f32p) if self.with_global else None
        if pinned:
            arrs["_feats"] = feats
        return feats, arrs

    @staticmethod
    def _view(f: hfb_features, arrs: dict, b: int, with_global: bool) -> dict:
        n = int(f.n_total)
        out = {k: arrs[k][b, :n] for k in ("x", "y", "response", "octave", "descriptors")}   # descriptors: empty when resident
        out["n_per_level"] = [int(v) for v in f.n_per_level]
        out["global_descriptor"] = arrs["global_descriptor"][b] if with_global else None
        return out

    def _budgets(self, n_per_level):
        return (C.c_int32 * HFB_MAX_LEVELS)(*([int(v) for v in n_per_level] + [0] * (HFB_MAX_LEVELS - len(n_per_level))))

    def extract_batch(self, images, n_per_level, threshold: float, return_block: bool = False, pinned: bool = False):
        """HFextractor::operator() for n frames.  Returns one dict per frame (views into one contiguous host block;
        with return_block=True also the block itself: frame b's descriptors are block['descriptors'][b, :n_b])."""
        imgs = [np.ascontiguousarray(im, dtype=np.uint8) for im in images]
        for im in imgs:
            if im.ndim != 2 or im.shape != (self.height, self.width):
                raise HfbError(1, f"image shape {im.shape} differs from the context's {(self.height, self.width)}")
        n = len(imgs)
        ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
        feats, arrs = self._alloc_features(n, pinned)
        self.check(self.lib.hfb_extract_batch(self.handle, ptrs, n, self.width, self._budgets(n_per_level), threshold,
                                              feats))
        out = [self._view(feats[i], arrs, i, self.with_global) for i in range(n)]
        return (out, arrs) if return_block else out

    def extract_match_batch(self, images, n_per_level, threshold: float, mode: int, thr: float, pinned: bool = False,
                            out=None, descriptors: bool = True):
        """extract_batch + match_consecutive as one call (hfb_extract_match_batch): returns (features, idx, val) with
        idx / val [n][kp_cap] (row b: frame b -> frame b-1).  pinned=True: page-locked outputs written by the graph.
        descriptors=False: keypoints, global descriptors and match rows come back, the 256-d local descriptors stay
        resident in HBM (the matcher, the resident windowed search and the keyframe store read them there)."""
        imgs = [np.ascontiguousarray(im, dtype=np.uint8) for im in images]
        for im in imgs:
            if im.ndim != 2 or im.shape != (self.height, self.width):
                raise HfbError(1, f"image shape {im.shape} differs from the context's {(self.height, self.width)}")
        n = len(imgs)
        ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
        feats, arrs = self._alloc_features(n, pinned, descriptors)
        if out is None:
            mk = pinned_empty if pinned else np.empty
            out = (mk((n, self.kp_cap), np.int32), mk((n, self.kp_cap), np.float32))
        idx, val = out
        self.check(self.lib.hfb_extract_match_batch(self.handle, ptrs, n, self.width, self._budgets(n_per_level),
                                                    threshold, feats, mode, thr, ptr(idx, _i32p), ptr(val, _f32p)))
        return [self._view(feats[i], arrs, i, self.with_global) for i in range(n)], idx, val

    def extract_match_batch_dev(self, d_images_ptr: int, n_images: int, n_per_level, threshold: float, mode: int,
                                thr: float):
        self.check(self.lib.hfb_extract_match_batch_dev(self.handle, C.c_void_p(d_images_ptr), n_images,
                                                        self._budgets(n_per_level), threshold, mode, thr))

    def extract_level(self, level: int, image, n_keypoints: int, threshold: float) -> dict:
        """BaseModel::Detect of one pyramid level on the shared context (level coordinates, octave 0)."""
        im = np.ascontiguousarray(image, dtype=np.uint8)
        feats, arrs = self._alloc_features(1)
        self.check(self.lib.hfb_extract_level(self.handle, level, im.ctypes.data, im.shape[0], im.shape[1], im.shape[1],
                                              n_keypoints, threshold, C.byref(feats[0])))
        return self._view(feats[0], arrs, 0, self.with_global and level == 0)

    def extract(self, image, n_per_level, threshold: float) -> dict:
        return self.extract_batch([image], n_per_level, threshold)[0]

    def extract_batch_dev(self, d_images_ptr: int, n_images: int, n_per_level, threshold: float):
        self.check(self.lib.hfb_extract_batch_dev(self.handle, C.c_void_p(d_images_ptr), n_images,
                                                  self._budgets(n_per_level), threshold))

    def fetch_features(self, image_index: int) -> dict:
        feats, arrs = self._alloc_features(1)
        self.check(self.lib.hfb_fetch_features(self.handle, image_index, C.byref(feats[0])))
        return self._view(feats[0], arrs, 0, self.with_global)

    def match_projection(self, Q, q_uv, q_radius, q_min_level, q_max_level, F, f_xy, f_level, f_skip=None,
                         f_inv_sigma2=None, chi2_max: float = 0.0):
        """Top-4 in-window candidates per query: (idx [nq,4], dist [nq,4], level [nq,4]).  With f_inv_sigma2 the
        candidates also pass Matcher::Fuse's reprojection gate ((du^2 + dv^2) * f_inv_sigma2 <= chi2_max)."""
        q, f = as_f32(Q).reshape(-1, HFB_DESC_DIM), as_f32(F).reshape(-1, HFB_DESC_DIM)
        nq, nf = q.shape[0], f.shape[0]
        uv, rad = as_f32(q_uv).reshape(-1, 2), as_f32(q_radius).reshape(-1)
        mn = np.ascontiguousarray(q_min_level, dtype=np.int32)
        mx = np.ascontiguousarray(q_max_level, dtype=np.int32)
        fxy = as_f32(f_xy).reshape(-1, 2)
        fl = np.ascontiguousarray(f_level, dtype=np.int32)
        fs = np.ascontiguousarray(f_skip, dtype=np.uint8) if f_skip is not None else None
        idx = np.full((nq, 4), -1, np.int32)
        dist = np.full((nq, 4), np.finfo(np.float32).max, np.float32)
        lvl = np.full((nq, 4), -1, np.int32)
        fi = as_f32(f_inv_sigma2).reshape(-1) if f_inv_sigma2 is not None else None
        self.check(self.lib.hfb_match_projection_gated(self.handle, ptr(q, _f32p), nq, ptr(uv, _f32p), ptr(rad, _f32p),
                                                       ptr(mn, _i32p), ptr(mx, _i32p), ptr(f, _f32p), nf, ptr(fxy, _f32p),
                                                       ptr(fl, _i32p), ptr(fs, _u8p) if fs is not None else None,
                                                       ptr(fi, _f32p) if fi is not None else None, float(chi2_max),
                                                       ptr(idx, _i32p), ptr(dist, _f32p), ptr(lvl, _i32p)))
        return idx, dist, lvl

    def set_stream_mode(self, mode: int):
        """0: a batch is B consecutive frames of one stream; 1: one frame of each of B streams (see hfnet_b200.h)."""
        self.check(self.lib.hfb_set_stream_mode(self.handle, mode))

    def reset_stream(self):
        """Forget the previous frame of the streaming association (Tracking::Reset)."""
        self.check(self.lib.hfb_reset_stream(self.handle))

    def match_projection_frame(self, frame_index: int, q_prev_index, q_uv, q_radius, q_min_level, q_max_level, nf: int,
                               f_skip=None, f_inv_sigma2=None, chi2_max: float = 0.0):
        """match_projection on resident descriptors: features = frame ``frame_index`` of the last extraction, queries =
        rows ``q_prev_index`` of its stream's previous frame.  Returns (idx [nq,4], dist [nq,4], level [nq,4])."""
        qi = np.ascontiguousarray(q_prev_index, dtype=np.int32)
        nq = len(qi)
        uv, rad = as_f32(q_uv).reshape(-1, 2), as_f32(q_radius).reshape(-1)
        mn = np.ascontiguousarray(q_min_level, dtype=np.int32)
        mx = np.ascontiguousarray(q_max_level, dtype=np.int32)
        fs = np.ascontiguousarray(f_skip, dtype=np.uint8) if f_skip is not None else None
        fi = as_f32(f_inv_sigma2).reshape(-1) if f_inv_sigma2 is not None else None
        idx = np.full((nq, 4), -1, np.int32)
        dist = np.full((nq, 4), np.finfo(np.float32).max, np.float32)
        lvl = np.full((nq, 4), -1, np.int32)
        self.check(self.lib.hfb_match_projection_frame(self.handle, frame_index, ptr(qi, _i32p), nq, ptr(uv, _f32p),
                                                       ptr(rad, _f32p), ptr(mn, _i32p), ptr(mx, _i32p), nf,
                                                       ptr(fs, _u8p) if fs is not None else None,
                                                       ptr(fi, _f32p) if fi is not None else None, float(chi2_max),
                                                       ptr(idx, _i32p), ptr(dist, _f32p), ptr(lvl, _i32p)))
        return idx, dist, lvl

    def match_consecutive_dev(self, n_images: int, mode: int, thr: float):
        self.check(self.lib.hfb_match_consecutive_dev(self.handle, n_images, mode, thr))

    def distinctive_descriptors(self, descriptors, offsets):
        """MapPoint::ComputeDistinctiveDescriptors for a ragged batch: returns (best_index int32[n], best_median f32[n])."""
        d = as_f32(descriptors).reshape(-1, HFB_DESC_DIM)
        off = np.ascontiguousarray(offsets, dtype=np.int32)
        n = len(off) - 1
        idx = np.full(n, -1, np.int32)
        med = np.zeros(n, np.float32)
        self.check(self.lib.hfb_distinctive_descriptors(self.handle, ptr(d, _f32p), ptr(off, _i32p), n, ptr(idx, _i32p),
                                                        ptr(med, _f32p)))
        return idx, med

    def match_consecutive(self, n_images: int, mode: int, thr: float, out=None):
        """Frame b of the last extraction against frame b-1 (descriptors resident in HBM); returns [n_images][kp_cap]
        match indices / values on the host.  ``out`` = (idx, val) arrays to reuse (e.g. page-locked)."""
        if out is None:
            out = (np.empty((n_images, self.kp_cap), np.int32), np.empty((n_images, self.kp_cap), np.float32))
        idx, val = out
        self.check(self.lib.hfb_match_consecutive(self.handle, n_images, mode, thr, ptr(idx, _i32p), ptr(val, _f32p)))
        return idx, val

    def fetch_matches(self, image_index: int, n: int):
        idx = np.full(n, -1, np.int32)
        val = np.zeros(n, np.float32)
        self.check(self.lib.hfb_fetch_matches(self.handle, image_index, ptr(idx, _i32p), ptr(val, _f32p), n))
        return idx, val

    def profile_extract(self, n_images: int, n_per_level, threshold: float):
        import json
        budgets = (C.c_int32 * HFB_MAX_LEVELS)(*([int(v) for v in n_per_level] + [0] * (HFB_MAX_LEVELS - len(n_per_level))))
        buf = C.create_string_buffer(1 << 16)
        self.check(self.lib.hfb_profile_extract(self.handle, n_images, budgets, threshold, buf, len(buf)))
        return json.loads(buf.value.decode())

    def debug_tensor(self, name: str, image_index: int = 0, level: int = 0) -> np.ndarray:
        n = C.c_size_t()
        dims = (C.c_int32 * 4)()
        st = self.lib.hfb_debug_tensor(self.handle, name.encode(), image_index, level, None, 0, C.byref(n), dims)
        if n.value == 0:
            self.check(st)
        out = np.zeros(n.value, np.float32)
        self.check(self.lib.hfb_debug_tensor(self.handle, name.encode(), image_index, level, ptr(out, _f32p),
                                             out.size, C.byref(n), dims))
        return out.reshape([int(d) for d in dims])

    # ------------------------------------------------------------------------------------------ network tail hooks
    def nms(self, scores: np.ndarray) -> np.ndarray:
        s = as_f32(scores)
        out = np.empty_like(s)
        self.check(self.lib.hfb_nms(self.handle, ptr(s, _f32p), s.shape[0], s.shape[1], ptr(out, _f32p)))
        return out

    def select_sample(self, scores_nms: np.ndarray, desc_map: np.ndarray, n_keypoints: int, threshold: float) -> dict:
        s, d = as_f32(scores_nms), as_f32(desc_map)
        cap = max(n_keypoints, 1)
        x, y, r = (np.zeros(cap, np.float32) for _ in range(3))
        desc = np.zeros((cap, HFB_DESC_DIM), np.float32)
        n = C.c_int32()
        self.check(self.lib.hfb_select_sample(self.handle, ptr(s, _f32p), s.shape[0], s.shape[1], ptr(d, _f32p),
                                              d.shape[0], d.shape[1], n_keypoints, threshold, ptr(x, _f32p),
                                              ptr(y, _f32p), ptr(r, _f32p), ptr(desc, _f32p), C.byref(n)))
        k = n.value
        return {"x": x[:k], "y": y[:k], "response": r[:k], "descriptors": desc[:k]}

    def resize_linear_u8(self, src: np.ndarray, dh: int, dw: int) -> np.ndarray:
        s = np.ascontiguousarray(src, dtype=np.uint8)
        out = np.empty((dh, dw), np.uint8)
        self.check(self.lib.hfb_resize_linear_u8(self.handle, ptr(s, _u8p), s.shape[0], s.shape[1], ptr(out, _u8p),
                                                 dh, dw))
        return out

    # ------------------------------------------------------------------------------------------ calibration (Frame)
    def set_camera(self, K, dist=()) -> None:
        """Frame's mK = (fx, fy, cx, cy) and mDistCoef; with dist[0] != 0 every extraction also produces mvKeysUn."""
        k = as_f32(K).reshape(4)
        d = as_f32(dist).reshape(-1)
        self.check(self.lib.hfb_set_camera(self.handle, ptr(k, _f32p), ptr(d, _f32p) if d.size else None, d.size))

    def undistort_points(self, x, y):
        """cv::undistortPoints(pts, pts, K, dist, noArray(), K) (Frame::UndistortKeyPoints, src/Frame.cc:760-793)."""
        xs, ys = as_f32(x).reshape(-1), as_f32(y).reshape(-1)
        xu, yu = np.empty_like(xs), np.empty_like(ys)
        self.check(self.lib.hfb_undistort_points(self.handle, ptr(xs, _f32p), ptr(ys, _f32p), xs.size, ptr(xu, _f32p),
                                                 ptr(yu, _f32p)))
        return xu, yu

    def fetch_undistorted(self, image_index: int, n: int):
        """mvKeysUn coordinates of the first n keypoints of frame ``image_index`` of the last extraction."""
        xu, yu = np.empty(n, np.float32), np.empty(n, np.float32)
        self.check(self.lib.hfb_fetch_undistorted(self.handle, image_index, ptr(xu, _f32p), ptr(yu, _f32p), n))
        return xu, yu

    def image_bounds(self, width: int, height: int) -> np.ndarray:
        """Frame::ComputeImageBounds: (mnMinX, mnMaxX, mnMinY, mnMaxY)."""
        b = np.empty(4, np.float32)
        self.check(self.lib.hfb_image_bounds(self.handle, width, height, ptr(b, _f32p)))
        return b

    def debug_gemm(self, A: np.ndarray, Wt: np.ndarray, bias=None, relu6=False, conv3x3=False, use_tc=True,
                   BN: int = 0) -> np.ndarray:
        """A: [B,H,W,K] (or [M,K]); Wt: [N, K] (or [N, 9K] for conv3x3).  fp16 operands, fp32 result [.., N]."""
        a = as_f32(A)
        if a.ndim == 2:
            a = a[None, None]
        B, H, W, K = a.shape
        w = as_f32(Wt)
        N = w.shape[0]
        out = np.empty((B * H * W, N), np.float32)
        b = as_f32(bias) if bias is not None else None
        self.check(self.lib.hfb_debug_gemm(self.handle, ptr(a, _f32p), B, H, W, K, ptr(w, _f32p), N,
                                           ptr(b, _f32p) if b is not None else None, int(relu6), int(conv3x3),
                                           int(use_tc), BN, ptr(out, _f32p)))
        return out.reshape(B, H, W, N) if np.asarray(A).ndim == 4 else out

    # ------------------------------------------------------------------------------------------ matching
    def _match(self, fn, A, B, thr):
        a, b = as_f32(A).reshape(-1, HFB_DESC_DIM), as_f32(B).reshape(-1, HFB_DESC_DIM)
        idx = np.full(a.shape[0], -1, np.int32)
        val = np.zeros(a.shape[0], np.float32)
        n = C.c_int32()
        self.check(fn(self.handle, ptr(a, _f32p), a.shape[0], ptr(b, _f32p), b.shape[0], thr, ptr(idx, _i32p),
                      ptr(val, _f32p), C.byref(n)))
        return idx, val, n.value

    def match_mutual_l2(self, A, B, max_dist: float):
        return self._match(self.lib.hfb_match_mutual_l2, A, B, max_dist)

    def match_mutual_cos(self, A, B, min_cos: float):
        return self._match(self.lib.hfb_match_mutual_cos, A, B, min_cos)

    def match_batch(self, mode: int, A_all, B_all, a_off, a_cnt, b_off, b_cnt, thr: float):
        a, b = as_f32(A_all).reshape(-1, HFB_DESC_DIM), as_f32(B_all).reshape(-1, HFB_DESC_DIM)
        tabs = [np.ascontiguousarray(t, dtype=np.int32) for t in (a_off, a_cnt, b_off, b_cnt)]
        idx = np.full(a.shape[0], -1, np.int32)
        val = np.zeros(a.shape[0], np.float32)
        self.check(self.lib.hfb_match_batch(self.handle, mode, ptr(a, _f32p), a.shape[0], ptr(b, _f32p), b.shape[0],
                                            len(tabs[0]), *[ptr(t, _i32p) for t in tabs], thr, ptr(idx, _i32p),
                                            ptr(val, _f32p)))
        return idx, val


class KeyFrameStore:
    """Keyframe local descriptors resident in HBM (hfb_kfstore_*): LocalMapping's neighbour matching without descriptor
    traffic.  ``rows_per_slot`` = the extractor's per-frame keypoint capacity."""

    def __init__(self, ctx: Context, n_slots: int, rows_per_slot: int):
        self.ctx, self.n_slots, self.rows = ctx, n_slots, rows_per_slot
        self.handle = C.c_void_p()
        ctx.check(ctx.lib.hfb_kfstore_create(ctx.handle, n_slots, rows_per_slot, C.byref(self.handle)))

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.hfb_kfstore_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return int(self.ctx.lib.hfb_kfstore_size(self.handle))

    def put_frame(self, ctx: Context, kf_id: int, frame_index: int, n: int):
        """Keyframe = frame ``frame_index`` of ``ctx``'s last extraction (first n keypoints), copied device to device."""
        ctx.check(ctx.lib.hfb_kfstore_put_frame(ctx.handle, self.handle, int(kf_id), frame_index, n))

    def put(self, ctx: Context, kf_id: int, descriptors):
        d = as_f32(descriptors).reshape(-1, HFB_DESC_DIM)
        ctx.check(ctx.lib.hfb_kfstore_put(ctx.handle, self.handle, int(kf_id), ptr(d, _f32p), d.shape[0]))

    def erase(self, kf_id: int):
        self.ctx.lib.hfb_kfstore_erase(self.handle, int(kf_id))

    def match_neighbours(self, ctx: Context, kf_a: int, kf_b, mode: int, thr: float, rows_a: int):
        """(idx [n_b, rows_a], val [n_b, rows_a]): keyframe kf_a against every stored keyframe of kf_b in one launch."""
        ids = np.ascontiguousarray(kf_b, dtype=np.int64)
        idx = np.full((len(ids), rows_a), -1, np.int32)
        val = np.zeros((len(ids), rows_a), np.float32)
        na = C.c_int32()
        ctx.check(ctx.lib.hfb_match_kf_neighbours(ctx.handle, self.handle, int(kf_a), ptr(ids, _i64p), len(ids), mode, thr,
                                                  ptr(idx, _i32p), ptr(val, _f32p), C.byref(na)))
        assert na.value == rows_a or len(ids) == 0, (na.value, rows_a)
        return idx, val
