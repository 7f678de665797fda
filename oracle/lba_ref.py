"""Oracle: float64 CPU restatement of ``Optimizer::LocalBundleAdjustment``'s numeric core.  TEST INFRASTRUCTURE ONLY.

Follows (all under /root/reference):
* edge error / depth test           include/OptimizableTypes.h:99-110, src/CameraModels/Pinhole.cpp:35-49
* Jacobians                         src/OptimizableTypes.cpp:139-159, src/CameraModels/Pinhole.cpp:71-81
* Huber + weighted quadratic form   Thirdparty/g2o/g2o/core/robust_kernel_impl.cpp:78-91,
                                    Thirdparty/g2o/g2o/core/base_edge.h:96-102,
                                    Thirdparty/g2o/g2o/core/base_binary_edge.hpp:55-121
* Schur complement + back-subst.    Thirdparty/g2o/g2o/core/block_solver.hpp:354-486, setLambda :564-589
* Levenberg-Marquardt control       Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:61-185
                                    (tau 1e-5, good-step scale in [1/3, 2/3], <=10 trials, 3-strike 0.1% stop)
* outer loop / 10 iterations        Thirdparty/g2o/g2o/core/sparse_optimizer.cpp:353-416, src/Optimizer.cc:1411
* pose / point updates              Thirdparty/g2o/g2o/types/types_six_dof_expmap.h:73-76, se3quat.h:223-257,
                                    Thirdparty/g2o/g2o/types/types_sba.h:52-56
* outlier flags                     src/Optimizer.cc:1417-1431  (chi2 > 5.991 or depth <= 0)

The reduced camera system is solved with a dense Cholesky (scipy) instead of g2o's sparse LDLT
(Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h:94-124): same linear system, different elimination order.
g2o itself cannot be compiled here (needs Eigen), so this restatement is PARITY UNPINNED against the binary; it is
pinned against finite-difference Jacobians and a dense normal-equation solve in tests/test_oracle_pins.py.

Quirk kept on purpose: ``e->chi2()`` after ``optimize`` reads the error cached by the LAST ``computeActiveErrors``,
i.e. of the last LM *trial* even if that trial was rejected and popped (Optimizer.cc:1425 never recomputes).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import scipy.linalg

HUBER_MONO = np.sqrt(5.991)       # src/Optimizer.cc:1206 thHuberMono
CHI2_MONO = 5.991                 # src/Optimizer.cc:1425


@dataclass
class Problem:
    poses: np.ndarray          # [n_cam,7] float64: qx qy qz qw tx ty tz  (Tcw, g2o::SE3Quat)
    fixed: np.ndarray          # [n_cam] bool
    points: np.ndarray         # [n_pt,3] float64
    cam_idx: np.ndarray        # [n_edge] int32
    pt_idx: np.ndarray         # [n_edge] int32
    obs: np.ndarray            # [n_edge,2] float64
    inv_sigma2: np.ndarray     # [n_edge] float64
    K: np.ndarray              # fx fy cx cy (float32 in the reference, promoted)
    huber_delta: float = float(HUBER_MONO)


# ----------------------------------------------------------------------------------------------------- SE3 helpers
def quat_to_rot(q: np.ndarray) -> np.ndarray:
    """Eigen quaternion (x,y,z,w) -> rotation matrix, [...,4] -> [...,3,3]."""
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - (tyy + tzz); R[..., 0, 1] = txy - twz; R[..., 0, 2] = txz + twy
    R[..., 1, 0] = txy + twz; R[..., 1, 1] = 1 - (txx + tzz); R[..., 1, 2] = tyz - twx
    R[..., 2, 0] = txz - twy; R[..., 2, 1] = tyz + twx; R[..., 2, 2] = 1 - (txx + tyy)
    return R


def rot_to_quat(R: np.ndarray) -> np.ndarray:
    """Eigen's Quaternion(Matrix3) (Shepperd), returns (x,y,z,w)."""
    t = R[0, 0] + R[1, 1] + R[2, 2]
    q = np.empty(4)
    if t > 0:
        t = np.sqrt(t + 1.0)
        q[3] = 0.5 * t
        t = 0.5 / t
        q[0] = (R[2, 1] - R[1, 2]) * t
        q[1] = (R[0, 2] - R[2, 0]) * t
        q[2] = (R[1, 0] - R[0, 1]) * t
    else:
        i = 0
        if R[1, 1] > R[0, 0]:
            i = 1
        if R[2, 2] > R[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        t = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        q[i] = 0.5 * t
        t = 0.5 / t
        q[3] = (R[k, j] - R[j, k]) * t
        q[j] = (R[j, i] + R[i, j]) * t
        q[k] = (R[k, i] + R[i, k]) * t
    return q


def quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz])


def normalize_rotation(q: np.ndarray) -> np.ndarray:
    """se3quat.h normalizeRotation(): flip to w >= 0, normalise."""
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q)


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def se3_exp(update: np.ndarray):
    """se3quat.h:223-257.  update = [omega(3), upsilon(3)] -> (quat xyzw, t)."""
    omega, upsilon = update[:3], update[3:]
    theta = np.linalg.norm(omega)
    Om = skew(omega)
    if theta < 0.00001:
        R = np.eye(3) + Om + Om @ Om
        V = R
    else:
        Om2 = Om @ Om
        R = np.eye(3) + np.sin(theta) / theta * Om + (1 - np.cos(theta)) / (theta * theta) * Om2
        V = np.eye(3) + (1 - np.cos(theta)) / (theta * theta) * Om + (theta - np.sin(theta)) / (theta ** 3) * Om2
    return normalize_rotation(rot_to_quat(R)), V @ upsilon


def pose_oplus(pose: np.ndarray, update: np.ndarray) -> np.ndarray:
    """types_six_dof_expmap.h:73-76: estimate = exp(update) * estimate  (left-multiplicative)."""
    qe, te = se3_exp(update)
    q = quat_mul(qe, pose[:4])
    t = te + quat_to_rot(qe) @ pose[4:]
    return np.concatenate([normalize_rotation(q), t])


# --------------------------------------------------------------------------------------------------- edge arithmetic
def edge_errors(pr: Problem, poses: np.ndarray, points: np.ndarray):
    """Returns err [n,2], chi2 [n] (= e^T Omega e), rho0 [n] (robustified), depth [n], Xc [n,3]."""
    R = quat_to_rot(poses[:, :4])
    Xc = np.einsum("nij,nj->ni", R[pr.cam_idx], points[pr.pt_idx]) + poses[pr.cam_idx, 4:]
    fx, fy, cx, cy = [float(v) for v in pr.K]
    proj = np.stack([fx * Xc[:, 0] / Xc[:, 2] + cx, fy * Xc[:, 1] / Xc[:, 2] + cy], axis=1)
    err = pr.obs - proj
    chi2 = pr.inv_sigma2 * (err * err).sum(axis=1)
    dsqr = pr.huber_delta * pr.huber_delta
    rho0 = np.where(chi2 <= dsqr, chi2, 2 * np.sqrt(np.maximum(chi2, 1e-300)) * pr.huber_delta - dsqr)
    return err, chi2, rho0, Xc[:, 2].copy(), Xc


def jacobians(pr: Problem, poses: np.ndarray, Xc: np.ndarray):
    """J_point [n,2,3] (= -Jproj R) and J_pose [n,2,6] (= -Jproj [-[X]x | I]); OptimizableTypes.cpp:139-159."""
    fx, fy = float(pr.K[0]), float(pr.K[1])
    x, y, z = Xc[:, 0], Xc[:, 1], Xc[:, 2]
    n = Xc.shape[0]
    Jp = np.zeros((n, 2, 3))
    Jp[:, 0, 0] = fx / z
    Jp[:, 0, 2] = -fx * x / (z * z)
    Jp[:, 1, 1] = fy / z
    Jp[:, 1, 2] = -fy * y / (z * z)
    Jp = -Jp
    R = quat_to_rot(poses[:, :4])[pr.cam_idx]
    J_point = np.einsum("nij,njk->nik", Jp, R)
    D = np.zeros((n, 3, 6))
    D[:, 0, 1] = z; D[:, 0, 2] = -y; D[:, 0, 3] = 1
    D[:, 1, 0] = -z; D[:, 1, 2] = x; D[:, 1, 4] = 1
    D[:, 2, 0] = y; D[:, 2, 1] = -x; D[:, 2, 5] = 1
    J_pose = np.einsum("nij,njk->nik", Jp, D)
    return J_point, J_pose


@dataclass
class System:
    opt_cams: np.ndarray       # camera indices that are optimised (array order)
    cam_slot: np.ndarray       # [n_cam] slot in the reduced system or -1
    Hpp: np.ndarray            # [n_opt,6,6]
    bp: np.ndarray             # [n_opt,6]
    Hll: np.ndarray            # [n_pt,3,3]
    bl: np.ndarray             # [n_pt,3]
    Hpl: np.ndarray            # [n_edge,6,3]  (zero for edges to fixed cameras)
    chi2: np.ndarray
    rho0: np.ndarray


def build_system(pr: Problem, poses: np.ndarray, points: np.ndarray) -> System:
    """computeActiveErrors + BlockSolver::buildSystem (block_solver.hpp:502-560) with the robust branch of
    constructQuadraticForm (base_binary_edge.hpp:93-113): weightedOmega = rho'(chi2) * Omega, b -= J^T (rho' Omega e)."""
    err, chi2, rho0, _, Xc = edge_errors(pr, poses, points)
    J_point, J_pose = jacobians(pr, poses, Xc)
    dsqr = pr.huber_delta * pr.huber_delta
    w = np.where(chi2 <= dsqr, 1.0, pr.huber_delta / np.sqrt(np.maximum(chi2, 1e-300)))
    wo = w * pr.inv_sigma2                                        # weighted information (scalar * I2)
    omega_r = -(wo[:, None] * err)                                # -(Omega e) * rho'
    n_cam, n_pt = poses.shape[0], points.shape[0]
    opt_cams = np.flatnonzero(~pr.fixed)
    cam_slot = np.full(n_cam, -1, np.int64)
    cam_slot[opt_cams] = np.arange(opt_cams.size)
    Hll = np.zeros((n_pt, 3, 3)); bl = np.zeros((n_pt, 3))
    np.add.at(Hll, pr.pt_idx, wo[:, None, None] * np.einsum("nki,nkj->nij", J_point, J_point))
    np.add.at(bl, pr.pt_idx, np.einsum("nki,nk->ni", J_point, omega_r))
    Hpp = np.zeros((opt_cams.size, 6, 6)); bp = np.zeros((opt_cams.size, 6))
    slot = cam_slot[pr.cam_idx]
    live = slot >= 0
    np.add.at(Hpp, slot[live], (wo[:, None, None] * np.einsum("nki,nkj->nij", J_pose, J_pose))[live])
    np.add.at(bp, slot[live], np.einsum("nki,nk->ni", J_pose, omega_r)[live])
    Hpl = wo[:, None, None] * np.einsum("nki,nkj->nij", J_pose, J_point)      # B^T Omega' A  (6x3)
    Hpl[~live] = 0
    return System(opt_cams, cam_slot, Hpp, bp, Hll, bl, Hpl, chi2, rho0)


def schur(pr: Problem, sy: System, lam: float):
    """BlockSolver::solve marginalisation (block_solver.hpp:380-439) with lambda added to every diagonal
    (setLambda :564-589).  Returns dense Hschur [6n,6n], bschur [6n], Dinv [n_pt,3,3]."""
    n_opt = sy.opt_cams.size
    Hs = np.zeros((6 * n_opt, 6 * n_opt))
    for i in range(n_opt):
        Hs[6 * i:6 * i + 6, 6 * i:6 * i + 6] = sy.Hpp[i] + lam * np.eye(6)
    D = sy.Hll + lam * np.eye(3)[None]
    Dinv = np.linalg.inv(D)
    bs = sy.bp.reshape(-1).copy()
    slot = sy.cam_slot[pr.cam_idx]
    order = np.argsort(pr.pt_idx, kind="stable")
    pts_sorted = pr.pt_idx[order]
    starts = np.flatnonzero(np.r_[True, pts_sorted[1:] != pts_sorted[:-1]])
    ends = np.r_[starts[1:], pts_sorted.size]
    for s, e in zip(starts, ends):
        eids = order[s:e]
        eids = eids[slot[eids] >= 0]
        if eids.size == 0:
            continue
        p = pr.pt_idx[eids[0]]
        B = sy.Hpl[eids]                                  # [k,6,3]
        BD = B @ Dinv[p]                                  # [k,6,3]
        sl = slot[eids]
        db = Dinv[p] @ sy.bl[p]
        for a in range(eids.size):
            ia = sl[a]
            bs[6 * ia:6 * ia + 6] -= B[a] @ db
            for b in range(eids.size):
                ib = sl[b]
                Hs[6 * ia:6 * ia + 6, 6 * ib:6 * ib + 6] -= BD[a] @ B[b].T
    return Hs, bs, Dinv


def solve_reduced(Hs: np.ndarray, bs: np.ndarray):
    """Dense Cholesky stand-in for LinearSolverEigen.  Returns (ok, x)."""
    if Hs.shape[0] == 0:
        return True, np.zeros(0)
    try:
        c = scipy.linalg.cho_factor(Hs, lower=True, check_finite=True)
        return True, scipy.linalg.cho_solve(c, bs)
    except (np.linalg.LinAlgError, ValueError):
        return False, np.zeros_like(bs)


def back_substitute(pr: Problem, sy: System, Dinv: np.ndarray, xp: np.ndarray) -> np.ndarray:
    """block_solver.hpp:461-481: xl = Dinv (bl - Hpl^T xp)."""
    cl = sy.bl.copy()
    slot = sy.cam_slot[pr.cam_idx]
    live = slot >= 0
    xp6 = xp.reshape(-1, 6)
    contrib = np.einsum("nij,ni->nj", sy.Hpl[live], xp6[slot[live]])
    np.subtract.at(cl, pr.pt_idx[live], contrib)
    return np.einsum("nij,nj->ni", Dinv, cl)


@dataclass
class Result:
    poses: np.ndarray
    points: np.ndarray
    chi2: np.ndarray               # per-edge chi2 as cached by the last computeActiveErrors
    depth_positive: np.ndarray     # at the final estimate
    outlier: np.ndarray
    iterations: int
    lambdas: list = field(default_factory=list)
    chis: list = field(default_factory=list)
    trials: int = 0


def optimize(pr: Problem, iterations: int = 10, user_lambda_init: float = 0.0,
             stop_flag: Optional[list] = None) -> Result:
    """g2o SparseOptimizer::optimize(iterations) with OptimizationAlgorithmLevenberg + BlockSolver_6_3 (Schur)."""
    poses = pr.poses.copy()
    points = pr.points.copy()
    tau, good_lo, good_hi, max_trials = 1e-5, 1.0 / 3.0, 2.0 / 3.0, 10
    lam, ni, n_bad = 0.0, 2.0, 0
    last_chi2 = edge_errors(pr, poses, points)[1]
    res = Result(poses, points, last_chi2, None, None, 0)
    terminate = lambda: bool(stop_flag and stop_flag[0])
    it_done = 0
    for it in range(iterations):
        if terminate():
            break
        sy = build_system(pr, poses, points)
        last_chi2 = sy.chi2
        current_chi = float(sy.rho0.sum())
        ini_chi = current_chi
        if it == 0:
            if user_lambda_init > 0:
                lam = user_lambda_init
            else:
                max_diag = 0.0
                if sy.Hpp.size:
                    max_diag = max(max_diag, float(np.abs(np.einsum("nii->ni", sy.Hpp)).max()))
                if sy.Hll.size:
                    max_diag = max(max_diag, float(np.abs(np.einsum("nii->ni", sy.Hll)).max()))
                lam = tau * max_diag
            ni, n_bad = 2.0, 0
        rho, qmax = 0.0, 0
        while True:
            Hs, bs, Dinv = schur(pr, sy, lam)
            ok2, xp = solve_reduced(Hs, bs)
            xl = back_substitute(pr, sy, Dinv, xp)
            new_poses = poses.copy()
            for s, c in enumerate(sy.opt_cams):
                new_poses[c] = pose_oplus(poses[c], xp[6 * s:6 * s + 6])
            new_points = points + xl
            _, chi2_t, rho0_t, _, _ = edge_errors(pr, new_poses, new_points)
            last_chi2 = chi2_t
            temp_chi = float(rho0_t.sum()) if ok2 else np.finfo(np.float64).max
            x = np.concatenate([xp, xl.reshape(-1)])
            b = np.concatenate([sy.bp.reshape(-1), sy.bl.reshape(-1)])
            scale = float(np.sum(x * (lam * x + b))) + 1e-3
            rho = (current_chi - temp_chi) / scale
            res.trials += 1
            if rho > 0 and np.isfinite(temp_chi):
                alpha = 1.0 - (2 * rho - 1) ** 3
                alpha = min(alpha, good_hi)
                lam *= max(good_lo, alpha)
                ni = 2.0
                current_chi = temp_chi
                poses, points = new_poses, new_points
            else:
                lam *= ni
                ni *= 2
            qmax += 1
            if not (rho < 0 and qmax < max_trials and not terminate()):
                break
        res.lambdas.append(lam)
        res.chis.append(current_chi)
        it_done += 1
        if qmax == max_trials or rho == 0:
            break
        if (ini_chi - current_chi) * 1e3 < ini_chi:
            n_bad += 1
        else:
            n_bad = 0
        if n_bad >= 3:
            break
    _, _, _, depth, _ = edge_errors(pr, poses, points)
    res.poses, res.points, res.chi2 = poses, points, last_chi2
    res.depth_positive = depth > 0.0
    res.outlier = (last_chi2 > CHI2_MONO) | ~res.depth_positive
    res.iterations = it_done
    return res


# ====================================================================================================================
# Optimizer::PoseOptimization (src/Optimizer.cc:814-1114), monocular branch: one VertexSE3Expmap, unary
# EdgeSE3ProjectXYZOnlyPose edges (include/OptimizableTypes.h:30-56, src/OptimizableTypes.cpp:49-64) with fixed Xw.
# Four rounds of optimize(10); every round restarts from the frame's pose (:1007-1008), uses only the current inliers
# (level 0, :1010), re-classifies all edges with chi2 > 5.991 (:1024-1036, `float` compare) and after the third round
# drops the Huber kernel (:1038-1039).  Returns the last round's pose.
def _pose_edges(K, pose, Xw, obs, inv_sigma2):
    R = quat_to_rot(pose[:4])
    Xc = Xw @ R.T + pose[4:]
    fx, fy, cx, cy = [float(v) for v in K]
    proj = np.stack([fx * Xc[:, 0] / Xc[:, 2] + cx, fy * Xc[:, 1] / Xc[:, 2] + cy], axis=1)
    err = obs - proj
    chi2 = inv_sigma2 * (err * err).sum(axis=1)
    return err, chi2, Xc


def _pose_jac(K, Xc):
    fx, fy = float(K[0]), float(K[1])
    x, y, z = Xc[:, 0], Xc[:, 1], Xc[:, 2]
    n = Xc.shape[0]
    Jp = np.zeros((n, 2, 3))
    Jp[:, 0, 0] = -fx / z
    Jp[:, 0, 2] = fx * x / (z * z)
    Jp[:, 1, 1] = -fy / z
    Jp[:, 1, 2] = fy * y / (z * z)
    D = np.zeros((n, 3, 6))
    D[:, 0, 1] = z; D[:, 0, 2] = -y; D[:, 0, 3] = 1
    D[:, 1, 0] = -z; D[:, 1, 2] = x; D[:, 1, 4] = 1
    D[:, 2, 0] = y; D[:, 2, 1] = -x; D[:, 2, 5] = 1
    return np.einsum("nij,njk->nik", Jp, D)


def pose_optimization(K, pose0, Xw, obs, inv_sigma2, huber_delta: float = float(HUBER_MONO)):
    """Returns (pose [7], outlier [n] bool, n_inliers, stats dict)."""
    pose0 = np.asarray(pose0, np.float64)
    n = Xw.shape[0]
    outlier = np.zeros(n, bool)
    chi2_mono = np.float32(5.991)
    tau, good_lo, good_hi, max_trials = 1e-5, 1.0 / 3.0, 2.0 / 3.0, 10
    dsqr = huber_delta * huber_delta
    pose = pose0.copy()
    trials_total, iters_total = 0, 0
    n_bad = 0

    def rho_sum(chi2, robust):
        if not robust:
            return float(chi2.sum())
        return float(np.where(chi2 <= dsqr, chi2, 2 * np.sqrt(np.maximum(chi2, 1e-300)) * huber_delta - dsqr).sum())

    for rnd in range(4):
        robust = rnd < 3
        act = ~outlier
        pose = pose0.copy()
        cached = np.zeros(n)                       # chi2 as cached by the last computeActiveErrors (active edges)
        lam, ni, nbad_lm = 0.0, 2.0, 0
        for it in range(10):
            if act.sum() == 0:
                break
            err, chi2, Xc = _pose_edges(K, pose, Xw[act], obs[act], inv_sigma2[act])
            cached[act] = chi2
            current_chi = rho_sum(chi2, robust)
            ini_chi = current_chi
            J = _pose_jac(K, Xc)
            w = np.where(chi2 <= dsqr, 1.0, huber_delta / np.sqrt(np.maximum(chi2, 1e-300))) if robust else np.ones_like(chi2)
            wo = w * inv_sigma2[act]
            H = np.einsum("n,nki,nkj->ij", wo, J, J)
            b = np.einsum("nki,nk->i", J, -(wo[:, None] * err))
            if it == 0:
                lam = tau * float(np.abs(np.diag(H)).max())
                ni, nbad_lm = 2.0, 0
            rho, qmax = 0.0, 0
            while True:
                ok2, x = solve_reduced(H + lam * np.eye(6), b)
                new_pose = pose_oplus(pose, x)
                _, chi2_t, _ = _pose_edges(K, new_pose, Xw[act], obs[act], inv_sigma2[act])
                cached[act] = chi2_t
                temp_chi = rho_sum(chi2_t, robust) if ok2 else np.finfo(np.float64).max
                scale = float(np.sum(x * (lam * x + b))) + 1e-3
                rho = (current_chi - temp_chi) / scale
                trials_total += 1
                if rho > 0 and np.isfinite(temp_chi):
                    alpha = min(1.0 - (2 * rho - 1) ** 3, good_hi)
                    lam *= max(good_lo, alpha)
                    ni = 2.0
                    current_chi = temp_chi
                    pose = new_pose
                else:
                    lam *= ni
                    ni *= 2
                qmax += 1
                if not (rho < 0 and qmax < max_trials):
                    break
            iters_total += 1
            if qmax == max_trials or rho == 0:
                break
            if (ini_chi - current_chi) * 1e3 < ini_chi:
                nbad_lm += 1
            else:
                nbad_lm = 0
            if nbad_lm >= 3:
                break
        # classification (:1016-1040): outlier edges are re-evaluated at the round's final pose, inliers keep the cache
        _, chi2_all, _ = _pose_edges(K, pose, Xw, obs, inv_sigma2)
        chi2_used = np.where(outlier, chi2_all, cached)
        outlier = chi2_used.astype(np.float32) > chi2_mono
        n_bad = int(outlier.sum())
        if n < 10:
            break
    return pose, outlier, n - n_bad, dict(trials=trials_total, iterations=iters_total)
