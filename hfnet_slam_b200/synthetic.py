"""Seeded synthetic workloads of BASELINE.json's configs.  numpy only, no oracle imports.

There is no network for datasets or checkpoints, so every benchmark and parity input is generated here:
C1 descriptor pairs, C4 keyframe database, C3 local-BA problems (EuRoC pinhole intrinsics, Examples/Monocular/EuRoC.yaml:23-26).
"""
from __future__ import annotations

import numpy as np

EUROC_K = np.array([458.654, 457.296, 367.215, 248.375], np.float32)


def _unit_rows(x: np.ndarray) -> np.ndarray:
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)


def descriptor_pair(na: int = 1000, nb: int = 1000, dim: int = 256, n_true: int = 300, noise: float = 0.035,
                    seed: int = 0):
    """C1: unit rows; B[0:n_true] = normalise(A[100:100+n_true] + noise*N(0,1)); rest random."""
    rng = np.random.default_rng(seed)
    A = _unit_rows(rng.normal(size=(na, dim)))
    B = _unit_rows(rng.normal(size=(nb, dim)))
    n_true = min(n_true, nb, max(na - 100, 0))
    if n_true > 0:
        B[:n_true] = _unit_rows(A[100:100 + n_true] + noise * rng.normal(size=(n_true, dim)))
    return A, B


def keyframe_db(n: int = 50000, dim: int = 4096, n_planted: int = 200, noise: float = 0.3, seed: int = 2,
                n_queries: int = 1):
    """C4: unit rows, ``n_planted`` near-duplicates of earlier rows; queries = planted rows + small noise."""
    rng = np.random.default_rng(seed)
    db = rng.standard_normal(size=(n, dim), dtype=np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    n_planted = min(n_planted, n // 2)
    src = rng.choice(n // 2, n_planted, replace=False)
    dst = n // 2 + rng.choice(n - n // 2, n_planted, replace=False)
    for s, d in zip(src, dst):
        v = db[s] + noise / np.sqrt(dim) * rng.standard_normal(dim).astype(np.float32)
        db[d] = v / np.linalg.norm(v)
    qi = src[:n_queries] if n_planted else rng.choice(n, n_queries)
    q = db[qi] + 0.1 / np.sqrt(dim) * rng.standard_normal(size=(len(qi), dim)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return db.astype(np.float32), q.astype(np.float32), qi


def _rot(rv: np.ndarray) -> np.ndarray:
    th = np.linalg.norm(rv)
    if th < 1e-12:
        return np.eye(3)
    k = rv / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def _rot_to_quat(R: np.ndarray) -> np.ndarray:
    w = np.sqrt(max(1e-16, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    return np.array([(R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w), w])


def lba_problem(n_opt: int = 20, n_fixed: int = 40, n_points: int = 3000, seed: int = 3, pixel_noise: float = 1.0,
                outlier_frac: float = 0.02, pose_noise: float = 0.01, point_noise: float = 0.03,
                width: int = 752, height: int = 480, K: np.ndarray = EUROC_K, max_obs_per_point: int = 12):
    """C3-shaped local BA: cameras on a smooth forward trajectory looking at a landmark cloud; observations are the
    true projections + Gaussian pixel noise (+ a few gross outliers), octave drawn from {0..3} so that
    invSigma2 = 1/1.2^(2*octave) (src/Optimizer.cc:1316-1317, HFextractor.cc:92-103).  Initial poses/points are the
    truth perturbed.  Returns a dict of flat arrays (the C-ABI layout)."""
    rng = np.random.default_rng(seed)
    n_cam = n_opt + n_fixed
    fx, fy, cx, cy = [float(v) for v in K]
    # camera centres along a gentle arc; fixed cams are the older part of the trajectory
    s = np.linspace(0.0, 1.0, n_cam)
    centres = np.stack([4.0 * s, 0.3 * np.sin(3 * s), 0.2 * np.cos(2 * s)], axis=1)
    poses = np.zeros((n_cam, 7))
    Rs, ts = [], []
    for i in range(n_cam):
        Rwc = _rot(np.array([0.02 * np.sin(5 * s[i]), 0.3 * s[i] - 0.15, 0.01 * np.cos(3 * s[i])]))
        Rcw = Rwc.T
        tcw = -Rcw @ centres[i]
        Rs.append(Rcw); ts.append(tcw)
    pts = np.stack([rng.uniform(-3, 8, n_points), rng.uniform(-2.5, 2.5, n_points), rng.uniform(4, 14, n_points)], 1)
    cam_idx, pt_idx, obs, inv_s2 = [], [], [], []
    for p in range(n_points):
        cams = rng.permutation(n_cam)
        cnt = 0
        for c in cams:
            Xc = Rs[c] @ pts[p] + ts[c]
            if Xc[2] < 0.5:
                continue
            u, v = fx * Xc[0] / Xc[2] + cx, fy * Xc[1] / Xc[2] + cy
            if not (0 <= u < width and 0 <= v < height):
                continue
            octave = int(rng.integers(0, 4))
            sig = 1.2 ** octave
            du, dv = rng.normal(0, pixel_noise * sig, 2)
            if rng.random() < outlier_frac:
                du, dv = rng.uniform(-40, 40, 2)
            cam_idx.append(c); pt_idx.append(p); obs.append((u + du, v + dv)); inv_s2.append(1.0 / (sig * sig))
            cnt += 1
            if cnt >= max_obs_per_point:
                break
    cam_idx = np.array(cam_idx, np.int32); pt_idx = np.array(pt_idx, np.int32)
    # drop points with < 2 observations, re-index
    counts = np.bincount(pt_idx, minlength=n_points)
    keep_pt = counts >= 2
    remap = -np.ones(n_points, np.int64); remap[keep_pt] = np.arange(keep_pt.sum())
    e_keep = keep_pt[pt_idx]
    cam_idx, pt_idx = cam_idx[e_keep], remap[pt_idx[e_keep]].astype(np.int32)
    obs = np.array(obs)[e_keep]; inv_s2 = np.array(inv_s2)[e_keep]
    pts = pts[keep_pt]
    # the reference's fixed keyframes come first in time; optimisable = the most recent n_opt
    fixed = np.zeros(n_cam, bool); fixed[:n_fixed] = True
    for i in range(n_cam):
        R, t = Rs[i], ts[i]
        if not fixed[i]:
            R = _rot(rng.normal(0, pose_noise, 3)) @ R
            t = t + rng.normal(0, pose_noise, 3)
        poses[i, :4] = _rot_to_quat(R)
        poses[i, 4:] = t
    pts_init = pts + rng.normal(0, point_noise, pts.shape)
    # sort edges by point (the device layout: one point's edges are contiguous)
    order = np.lexsort((cam_idx, pt_idx))
    return dict(poses=poses, fixed=fixed, points=pts_init, cam_idx=cam_idx[order], pt_idx=pt_idx[order],
                obs=obs[order], inv_sigma2=inv_s2[order], K=np.asarray(K, np.float32),
                true_points=pts)


def pose_problem(n: int = 400, seed: int = 11, pixel_noise: float = 1.0, outlier_frac: float = 0.12,
                 pose_noise: float = 0.02, width: int = 752, height: int = 480, K: np.ndarray = EUROC_K):
    """Tracking-shaped motion-only BA (Optimizer::PoseOptimization, src/Optimizer.cc:814): one frame observing n fixed
    map points; observations = true projection + octave-scaled pixel noise, a fraction replaced by gross mismatches;
    the initial pose is the truth perturbed (the motion-model prediction)."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = [float(v) for v in K]
    Rcw = _rot(np.array([0.05, -0.1, 0.02]))
    tcw = np.array([0.3, -0.1, 0.2])
    u = rng.uniform(5, width - 5, n); v = rng.uniform(5, height - 5, n); z = rng.uniform(1.5, 12.0, n)
    Xc = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1)
    Xw = (Xc - tcw) @ Rcw                    # Rcw^T (Xc - t)
    octave = rng.integers(0, 4, n)
    sig = 1.2 ** octave
    obs = np.stack([u, v], 1) + rng.normal(0, pixel_noise, (n, 2)) * sig[:, None]
    bad = rng.random(n) < outlier_frac
    obs[bad] += rng.uniform(-60, 60, (int(bad.sum()), 2))
    R0 = _rot(rng.normal(0, pose_noise, 3)) @ Rcw
    t0 = tcw + rng.normal(0, pose_noise, 3)
    pose0 = np.concatenate([_rot_to_quat(R0), t0])
    return dict(K=np.asarray(K, np.float32), pose0=pose0, Xw=Xw, obs=obs, inv_sigma2=1.0 / (sig * sig),
                true_pose=np.concatenate([_rot_to_quat(Rcw), tcw]), planted_outlier=bad)
