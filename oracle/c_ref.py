"""ctypes door to the oracle's native pieces (oracle/c/*.c and, where it was built, the reference's own Resampler).
TEST INFRASTRUCTURE ONLY: imported by tests/ and by bench.py's CPU arms, never by hfnet_slam_b200/."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import build_c

_f32p, _f64p = C.POINTER(C.c_float), C.POINTER(C.c_double)
_i32p, _u8p = C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
_lib: Optional[C.CDLL] = None
_ref: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build_c.build_c()))
        _lib.ref_max_threads.restype = C.c_int
    return _lib


def set_threads(n: int) -> int:
    lib().ref_set_threads(int(n))
    return int(lib().ref_max_threads())


def _p(a, t):
    return a.ctypes.data_as(t)


def match_cos_mutual(D1: np.ndarray, D2: np.ndarray, floor: float):
    """src/Matcher.cc:845-889 -> (match12 int32[n1] or -1, cosine f32[n1])."""
    a, b = np.ascontiguousarray(D1, np.float32), np.ascontiguousarray(D2, np.float32)
    idx, val = np.empty(a.shape[0], np.int32), np.empty(a.shape[0], np.float32)
    lib().ref_match_cos_mutual(_p(a, _f32p), a.shape[0], _p(b, _f32p), b.shape[0], a.shape[1], C.c_float(floor),
                               _p(idx, _i32p), _p(val, _f32p))
    return idx, val


def match_bf_l2(A: np.ndarray, B: np.ndarray, max_dist: float):
    """cv::BFMatcher(NORM_L2, crossCheck).match + dist < max_dist (src/Matcher.cc:229-253) -> (match int32[na], dist f32[na])."""
    a, b = np.ascontiguousarray(A, np.float32), np.ascontiguousarray(B, np.float32)
    idx, val = np.empty(a.shape[0], np.int32), np.empty(a.shape[0], np.float32)
    lib().ref_match_bf_l2(_p(a, _f32p), a.shape[0], _p(b, _f32p), b.shape[0], a.shape[1], C.c_float(max_dist),
                          _p(idx, _i32p), _p(val, _f32p))
    return idx, val


def kfdb_scores(q: np.ndarray, db: np.ndarray) -> np.ndarray:
    """src/KeyFrameDatabase.cc:86-96, the literal per-keyframe loop in fp32."""
    qq, d = np.ascontiguousarray(q, np.float32), np.ascontiguousarray(db, np.float32)
    out = np.empty(d.shape[0], np.float32)
    lib().ref_kfdb_scores(_p(qq, _f32p), _p(d, _f32p), d.shape[0], d.shape[1], _p(out, _f32p))
    return out


def lba_optimize(problem: dict, iterations: int = 10, user_lambda_init: float = 0.0, huber_delta: float = float(np.sqrt(5.991))):
    """The C restatement of oracle/lba_ref.optimize on the flat problem dict of hfnet_slam_b200.synthetic.lba_problem."""
    poses = np.ascontiguousarray(problem["poses"], np.float64)
    fixed = np.ascontiguousarray(problem["fixed"], np.uint8)
    points = np.ascontiguousarray(problem["points"], np.float64)
    cam = np.ascontiguousarray(problem["cam_idx"], np.int32)
    pt = np.ascontiguousarray(problem["pt_idx"], np.int32)
    obs = np.ascontiguousarray(problem["obs"], np.float64)
    is2 = np.ascontiguousarray(problem["inv_sigma2"], np.float64)
    K = np.ascontiguousarray(problem["K"], np.float32)
    po, pp = np.empty_like(poses), np.empty_like(points)
    chi2, depth, stats = np.empty(len(cam), np.float64), np.empty(len(cam), np.uint8), np.zeros(5, np.float64)
    lib().ref_lba_optimize(poses.shape[0], points.shape[0], len(cam), _p(poses, _f64p), _p(fixed, _u8p), _p(points, _f64p),
                           _p(cam, _i32p), _p(pt, _i32p), _p(obs, _f64p), _p(is2, _f64p), _p(K, _f32p),
                           C.c_double(huber_delta), int(iterations), C.c_double(user_lambda_init), _p(po, _f64p),
                           _p(pp, _f64p), _p(chi2, _f64p), _p(depth, _u8p), _p(stats, _f64p))
    return dict(poses=po, points=pp, chi2=chi2, depth_positive=depth.astype(bool), iterations=int(stats[0]),
                trials=int(stats[1]), initial_chi2=float(stats[2]), final_chi2=float(stats[3]), lambda_=float(stats[4]))


def reference_resampler():
    """The reference's own Resampler compiled from /root/reference (oracle/_ref), or None where it was never built."""
    global _ref
    if _ref is None:
        p = build_c.build_ref()
        if p is None or not p.exists():
            return None
        _ref = C.CDLL(str(p))
    return _ref


def resample_reference(data: np.ndarray, warp: np.ndarray) -> Optional[np.ndarray]:
    """data [H,W,C] f32 (NHWC, batch 1), warp [N,2] (x, y) -> [N,C] through the reference's Resampler; None if unavailable."""
    r = reference_resampler()
    if r is None:
        return None
    d, w = np.ascontiguousarray(data, np.float32), np.ascontiguousarray(warp, np.float32)
    out = np.empty((w.shape[0], d.shape[2]), np.float32)
    r.ref_resampler(_p(d, _f32p), _p(w, _f32p), _p(out, _f32p), 1, d.shape[0], d.shape[1], d.shape[2], w.shape[0])
    return out


def pose_optimize(K, pose0, Xw, obs, inv_sigma2, huber_delta: float = float(np.sqrt(5.991))):
    """The C restatement of oracle/lba_ref.pose_optimization -> dict(pose, outlier, n_inliers, trials)."""
    k = np.ascontiguousarray(K, np.float32)
    p0 = np.ascontiguousarray(pose0, np.float64).reshape(7)
    X = np.ascontiguousarray(Xw, np.float64).reshape(-1, 3)
    o = np.ascontiguousarray(obs, np.float64).reshape(-1, 2)
    s2 = np.ascontiguousarray(inv_sigma2, np.float64).reshape(-1)
    out, flags = np.zeros(7, np.float64), np.zeros(max(len(X), 1), np.uint8)
    ninl, ntr = C.c_int(), C.c_int()
    lib().ref_pose_optimize(_p(k, _f32p), _p(p0, _f64p), len(X), _p(X, _f64p), _p(o, _f64p), _p(s2, _f64p),
                            C.c_double(huber_delta), _p(out, _f64p), _p(flags, _u8p), C.byref(ninl), C.byref(ntr))
    return dict(pose=out, outlier=flags[:len(X)].astype(bool), n_inliers=int(ninl.value), trials=int(ntr.value))
