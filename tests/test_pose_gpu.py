"""Motion-only BA (Optimizer::PoseOptimization, src/Optimizer.cc:814-1114) through the C-ABI against
oracle/lba_ref.pose_optimization (float64).  Bar: identical outlier flags, inlier count and LM trial count; pose within
1e-8 (unit quaternion components / metres) -- the only difference is the summation order of the 6x6 normal equations."""
import numpy as np
import pytest

from hfnet_slam_b200 import synthetic
from hfnet_slam_b200.optimizer import pose_optimization
from oracle import lba_ref

pytestmark = pytest.mark.gpu


def _ref(d):
    K = d["K"].astype(np.float64)
    return lba_ref.pose_optimization(K, d["pose0"], d["Xw"], d["obs"], d["inv_sigma2"])


@pytest.mark.parametrize("kw", [dict(n=400, seed=11), dict(n=1500, seed=12, outlier_frac=0.3), dict(n=37, seed=13),
                                dict(n=257, seed=14, pose_noise=0.08), dict(n=12, seed=15, outlier_frac=0.0)])
def test_pose_optimization_matches_oracle(small_ctx, kw):
    d = synthetic.pose_problem(**kw)
    pose_r, out_r, ninl_r, st = _ref(d)
    out = pose_optimization(small_ctx, d["K"], d["pose0"], d["Xw"], d["obs"], d["inv_sigma2"])
    assert out["trials"] == st["trials"]
    assert np.array_equal(out["outlier"], out_r)
    assert out["n_inliers"] == ninl_r
    assert np.abs(out["pose"] - pose_r).max() <= 1e-8
    # and it actually recovers the pose
    assert np.abs(out["pose"] - d["true_pose"]).max() < np.abs(d["pose0"] - d["true_pose"]).max()


def test_fewer_than_ten_edges_runs_one_round(small_ctx):
    """optimizer.edges().size() < 10 -> break after the first round (src/Optimizer.cc:1108-1109)."""
    d = synthetic.pose_problem(n=7, seed=16, outlier_frac=0.0)
    pose_r, out_r, ninl_r, st = _ref(d)
    out = pose_optimization(small_ctx, d["K"], d["pose0"], d["Xw"], d["obs"], d["inv_sigma2"])
    assert out["trials"] == st["trials"] and st["iterations"] <= 10
    assert np.array_equal(out["outlier"], out_r) and out["n_inliers"] == ninl_r
    assert np.abs(out["pose"] - pose_r).max() <= 1e-8


def test_empty_and_reproducible(small_ctx):
    d = synthetic.pose_problem(n=300, seed=17)
    e = pose_optimization(small_ctx, d["K"], d["pose0"], d["Xw"][:0], d["obs"][:0], d["inv_sigma2"][:0])
    assert e["n_inliers"] == 0 and np.array_equal(e["pose"], d["pose0"])
    a = pose_optimization(small_ctx, d["K"], d["pose0"], d["Xw"], d["obs"], d["inv_sigma2"])
    b = pose_optimization(small_ctx, d["K"], d["pose0"], d["Xw"], d["obs"], d["inv_sigma2"])
    assert np.array_equal(a["pose"], b["pose"]) and np.array_equal(a["outlier"], b["outlier"])
