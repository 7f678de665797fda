"""Local-BA kernels through the C-ABI against oracle/lba_ref.py (float64).  Tolerances: reduced system relative 1e-9,
final poses 1e-6 (rad / m), points 1e-6 m, identical LM iteration / trial counts and outlier flags."""
import numpy as np
import pytest

from hfnet_slam_b200 import synthetic
from hfnet_slam_b200.optimizer import local_bundle_adjustment, build_schur
from oracle import lba_ref

pytestmark = pytest.mark.gpu


def _problem(**kw):
    d = synthetic.lba_problem(**kw)
    return d, lba_ref.Problem(d["poses"], d["fixed"], d["points"], d["cam_idx"], d["pt_idx"], d["obs"], d["inv_sigma2"], d["K"])


@pytest.mark.parametrize("kw", [dict(n_opt=6, n_fixed=4, n_points=300, seed=1), dict(n_opt=20, n_fixed=40, n_points=3000, seed=3)])
def test_build_schur_matches_oracle(small_ctx, kw):
    d, pr = _problem(**kw)
    lam = 3.7
    Hs, bs, chi, n_opt = build_schur(small_ctx, d, lam)
    sy = lba_ref.build_system(pr, pr.poses, pr.points)
    Hr, br, _ = lba_ref.schur(pr, sy, lam)
    assert n_opt == kw["n_opt"]
    assert abs(chi - sy.rho0.sum()) <= 1e-9 * abs(sy.rho0.sum())
    assert np.abs(Hs - Hr).max() <= 1e-9 * np.abs(Hr).max()
    assert np.abs(bs - br).max() <= 1e-9 * np.abs(br).max()
    assert np.array_equal(Hs, Hs.T)


@pytest.mark.parametrize("kw", [dict(n_opt=6, n_fixed=4, n_points=300, seed=1), dict(n_opt=20, n_fixed=40, n_points=3000, seed=3),
                                dict(n_opt=3, n_fixed=1, n_points=60, seed=9, outlier_frac=0.1)])
def test_optimize_matches_oracle(small_ctx, kw):
    d, pr = _problem(**kw)
    ref = lba_ref.optimize(pr, 10)
    out = local_bundle_adjustment(small_ctx, d, iterations=10)
    assert out["iterations"] == ref.iterations and out["trials"] == ref.trials
    assert np.abs(out["poses"] - ref.poses).max() <= 1e-6
    assert np.abs(out["points"] - ref.points).max() <= 1e-6
    assert np.abs(out["chi2"] - ref.chi2).max() <= 1e-6 * max(1.0, np.abs(ref.chi2).max())
    assert np.array_equal(out["depth_positive"], ref.depth_positive)
    assert np.array_equal(out["outlier"], ref.outlier)
    assert out["final_chi2"] < out["initial_chi2"]
    # fixed cameras untouched
    assert np.array_equal(out["poses"][d["fixed"]], d["poses"][d["fixed"]])


def test_bit_reproducible(small_ctx):
    d, _ = _problem(n_opt=8, n_fixed=6, n_points=500, seed=4)
    a = local_bundle_adjustment(small_ctx, d, iterations=5)
    b = local_bundle_adjustment(small_ctx, d, iterations=5)
    assert np.array_equal(a["poses"], b["poses"]) and np.array_equal(a["points"], b["points"])


def test_stop_flag_and_zero_iterations(small_ctx):
    d, pr = _problem(n_opt=4, n_fixed=2, n_points=100, seed=5)
    out = local_bundle_adjustment(small_ctx, d, iterations=0)
    assert out["iterations"] == 0 and np.array_equal(out["poses"], d["poses"])
    out = local_bundle_adjustment(small_ctx, d, iterations=10, stop=True)
    assert out["iterations"] == 0


def test_unsorted_edges_rejected(small_ctx):
    d, _ = _problem(n_opt=4, n_fixed=2, n_points=100, seed=5)
    d["pt_idx"] = d["pt_idx"][::-1].copy()
    with pytest.raises(Exception):
        local_bundle_adjustment(small_ctx, d, iterations=1)


def test_device_solver_equals_host_solve_mode(small_ctx):
    """The reduced camera system solved on the device (one-CTA packed Cholesky + pose update) against the host-solve mode
    (HFB_LBA_HOST_SOLVE=1: dense Cholesky on the host, the round-1 path): same LM path, estimates to 1e-9."""
    import os
    d, _ = _problem(n_opt=20, n_fixed=40, n_points=3000, seed=3)
    a = local_bundle_adjustment(small_ctx, d, iterations=10)
    os.environ["HFB_LBA_HOST_SOLVE"] = "1"
    try:
        b = local_bundle_adjustment(small_ctx, d, iterations=10)
    finally:
        del os.environ["HFB_LBA_HOST_SOLVE"]
    assert a["iterations"] == b["iterations"] and a["trials"] == b["trials"]
    assert np.abs(a["poses"] - b["poses"]).max() <= 1e-9 and np.abs(a["points"] - b["points"]).max() <= 1e-9
    assert a["gpu_launches"] > b["gpu_launches"] - 20      # the device path adds one solve launch per trial
    # a system larger than one CTA's shared memory (33 optimisable cameras -> N = 198 > 192) falls back to the host solve
    big, prb = _problem(n_opt=33, n_fixed=10, n_points=800, seed=6)
    ref = lba_ref.optimize(prb, 4)
    out = local_bundle_adjustment(small_ctx, big, iterations=4)
    assert out["iterations"] == ref.iterations and np.abs(out["poses"] - ref.poses).max() <= 1e-6
