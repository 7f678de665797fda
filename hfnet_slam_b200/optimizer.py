"""Host-side mirror of ``Optimizer::LocalBundleAdjustment``'s numeric core (src/Optimizer.cc:1116-1498) over the GPU
kernels (``hfb_lba_*``).  Graph selection (which keyframes / points are local, src/Optimizer.cc:1120-1262) and the
write-back under the map mutex (:1464-1497) walk the caller's map and are not part of this path; the caller passes
the flat problem (poses, fixed mask, points, edges sorted by point)."""
from __future__ import annotations

import ctypes as C
from typing import Dict

import numpy as np

from .lib import Context, _f64p, _i32p, _u8p, hfb_lba_problem, hfb_lba_stats, ptr

HUBER_MONO = float(np.sqrt(5.991))   # thHuberMono, src/Optimizer.cc:1206
CHI2_MONO = 5.991                    # src/Optimizer.cc:1425


def _pack(d: Dict[str, np.ndarray], huber_delta: float):
    keep = dict(
        poses=np.ascontiguousarray(d["poses"], np.float64), fixed=np.ascontiguousarray(d["fixed"], np.uint8),
        points=np.ascontiguousarray(d["points"], np.float64), cam=np.ascontiguousarray(d["cam_idx"], np.int32),
        pt=np.ascontiguousarray(d["pt_idx"], np.int32), obs=np.ascontiguousarray(d["obs"], np.float64),
        is2=np.ascontiguousarray(d["inv_sigma2"], np.float64))
    p = hfb_lba_problem()
    p.n_cams, p.n_points, p.n_edges = keep["poses"].shape[0], keep["points"].shape[0], keep["cam"].shape[0]
    p.poses, p.fixed, p.points = ptr(keep["poses"], _f64p), ptr(keep["fixed"], _u8p), ptr(keep["points"], _f64p)
    p.edge_cam, p.edge_point = ptr(keep["cam"], _i32p), ptr(keep["pt"], _i32p)
    p.obs, p.inv_sigma2 = ptr(keep["obs"], _f64p), ptr(keep["is2"], _f64p)
    for i in range(4):
        p.K[i] = float(d["K"][i])
    p.huber_delta = huber_delta
    return p, keep


def local_bundle_adjustment(ctx: Context, problem: Dict[str, np.ndarray], iterations: int = 10,
                            lambda_init: float = 0.0, stop: bool = False, huber_delta: float = HUBER_MONO) -> dict:
    """optimizer.optimize(10) + the outlier test of src/Optimizer.cc:1411-1431.  ``problem`` keys: poses [n,7]
    (qx qy qz qw tx ty tz), fixed [n], points [m,3], cam_idx, pt_idx (non-decreasing), obs [e,2], inv_sigma2 [e], K[4]."""
    p, keep = _pack(problem, huber_delta)
    poses = np.zeros_like(keep["poses"])
    points = np.zeros_like(keep["points"])
    chi2 = np.zeros(p.n_edges, np.float64)
    depth = np.zeros(p.n_edges, np.uint8)
    stats = hfb_lba_stats()
    flag = np.array([1 if stop else 0], np.uint8)
    ctx.check(ctx.lib.hfb_lba_optimize(ctx.handle, C.byref(p), iterations, lambda_init, ptr(flag, _u8p),
                                       ptr(poses, _f64p), ptr(points, _f64p), ptr(chi2, _f64p), ptr(depth, _u8p),
                                       C.byref(stats)))
    dp = depth.astype(bool)
    return dict(poses=poses, points=points, chi2=chi2, depth_positive=dp, outlier=(chi2 > CHI2_MONO) | ~dp,
                iterations=int(stats.iterations), trials=int(stats.trials), initial_chi2=float(stats.initial_chi2),
                final_chi2=float(stats.final_chi2), lambda_=float(stats.lambda_), n_opt=int(stats.n_opt_cams),
                gpu_launches=int(stats.gpu_launches))


def build_schur(ctx: Context, problem: Dict[str, np.ndarray], lam: float, huber_delta: float = HUBER_MONO):
    """One linearisation + Schur reduction (parity hook): returns (Hschur [6n,6n], bschur [6n], robust chi2, n_opt)."""
    p, keep = _pack(problem, huber_delta)
    n_opt = int((keep["fixed"] == 0).sum())
    Hs = np.zeros((6 * n_opt, 6 * n_opt), np.float64)
    bs = np.zeros(6 * n_opt, np.float64)
    chi, n = C.c_double(), C.c_int32()
    ctx.check(ctx.lib.hfb_lba_build_schur(ctx.handle, C.byref(p), lam, ptr(Hs, _f64p), ptr(bs, _f64p), C.byref(chi),
                                          C.byref(n)))
    return Hs, bs, float(chi.value), int(n.value)


def pose_optimization(ctx: Context, K, pose, Xw, obs, inv_sigma2) -> dict:
    """Optimizer::PoseOptimization (src/Optimizer.cc:814-1114), monocular: returns dict(pose, outlier, n_inliers, trials)."""
    from .lib import _f32p
    k = np.ascontiguousarray(K, np.float32)
    p0 = np.ascontiguousarray(pose, np.float64).reshape(7)
    X = np.ascontiguousarray(Xw, np.float64).reshape(-1, 3)
    o = np.ascontiguousarray(obs, np.float64).reshape(-1, 2)
    s2 = np.ascontiguousarray(inv_sigma2, np.float64).reshape(-1)
    n = X.shape[0]
    out = np.zeros(7, np.float64)
    flags = np.zeros(max(n, 1), np.uint8)
    ninl, ntr = C.c_int32(), C.c_int32()
    ctx.check(ctx.lib.hfb_pose_optimize(ctx.handle, ptr(k, _f32p), ptr(p0, _f64p), n, ptr(X, _f64p), ptr(o, _f64p),
                                        ptr(s2, _f64p), ptr(out, _f64p), ptr(flags, _u8p), C.byref(ninl), C.byref(ntr)))
    return dict(pose=out, outlier=flags[:n].astype(bool), n_inliers=int(ninl.value), trials=int(ntr.value))
