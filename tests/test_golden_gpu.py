"""The CUDA path (through the C-ABI) against the committed golden fixtures of tests/golden/."""
from pathlib import Path

import numpy as np
import pytest

from hfnet_slam_b200 import synthetic, weights
from hfnet_slam_b200.keyframe_database import KeyFrameDatabase
from hfnet_slam_b200.lib import Context
from hfnet_slam_b200.optimizer import local_bundle_adjustment
from tests.test_golden_cpu import tail_inputs

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"


def test_match_golden(small_ctx):
    g = np.load(G / "match.npz")
    na, nb, nt, seed = g["params"]
    A, B = synthetic.descriptor_pair(int(na), int(nb), n_true=int(nt), seed=int(seed))
    idx, val, n = small_ctx.match_mutual_l2(A, B, 0.6)
    got = np.array([(i, idx[i]) for i in np.flatnonzero(idx >= 0)], np.int32).reshape(-1, 2)
    assert np.array_equal(got, g["bow_pairs"])
    idx, val, n = small_ctx.match_mutual_cos(A, B, 0.71875)
    got = np.array([(i, idx[i]) for i in np.flatnonzero(idx >= 0)], np.int32).reshape(-1, 2)
    assert np.array_equal(got, g["tri_pairs"]) and np.allclose(val[got[:, 0]], g["tri_cos"], atol=2e-6)


def test_tail_golden(small_ctx):
    g = np.load(G / "tail.npz")
    s, dm = tail_inputs()
    nms = small_ctx.nms(s)
    assert np.array_equal(np.argwhere(nms > 0).astype(np.int16), g["nms_nonzero"])
    f = small_ctx.select_sample(nms, dm, 60, 0.05)
    for k in ("x", "y", "response", "descriptors"):
        assert np.array_equal(f[k], g[k]), k


def test_pyramid_golden(small_ctx):
    g = np.load(G / "pyramid.npz")
    cur = weights.synthetic_image(120, 188, seed=3, n_corners=30)
    for k in ("l1", "l2", "l3"):
        cur = small_ctx.resize_linear_u8(cur, *g[k].shape)
        assert np.array_equal(cur, g[k]), k


def test_kfdb_golden(small_ctx):
    g = np.load(G / "kfdb.npz")
    n, dim, npl, seed = g["params"]
    db, q, _ = synthetic.keyframe_db(int(n), int(dim), n_planted=int(npl), seed=int(seed))
    kf = KeyFrameDatabase(small_ctx, capacity=int(n))
    kf.add_many(np.arange(int(n), dtype=np.int64), db)
    cand, sc, best = kf.query(q[0])
    assert np.array_equal(cand.astype(np.int32), g["cand"]) and abs(best - float(g["best"])) <= 2e-6
    assert np.abs(kf.scores_of(np.arange(int(n), dtype=np.int64)) - g["scores"]).max() <= 2e-6


def test_lba_golden(small_ctx):
    g = np.load(G / "lba.npz")
    no, nf, npts, seed = g["params"]
    d = synthetic.lba_problem(n_opt=int(no), n_fixed=int(nf), n_points=int(npts), seed=int(seed))
    out = local_bundle_adjustment(small_ctx, d, iterations=10)
    assert [out["iterations"], out["trials"]] == g["iterations"].tolist()
    assert np.abs(out["poses"] - g["poses"]).max() <= 1e-6 and np.abs(out["points"] - g["points"]).max() <= 1e-6
    assert np.array_equal(out["outlier"], g["outlier"])


def test_hfnet_golden(native_lib, weights_blob):
    g = np.load(G / "hfnet.npz")
    im = weights.synthetic_image(64, 96, seed=2, n_corners=12)
    with Context(height=64, width=96, n_levels=1, max_keypoints=256, max_batch=1) as ctx:
        ctx.load_weights(weights_blob)
        out = ctx.extract(im, [200], 0.01)
        sc = ctx.debug_tensor("scores_dense")[0, :, :, 0]
        assert np.abs(sc - g["scores_dense"].astype(np.float32)).max() <= 4e-3
        gd = out["global_descriptor"]
        assert float(gd @ g["global_descriptor"]) / np.linalg.norm(gd) >= 0.999


def test_undistort_golden(small_ctx):
    """The device kernel against cv2.undistortPoints' own outputs (tests/golden/undistort.npz), bit for bit."""
    from tests.test_golden_cpu import undistort_inputs
    g = np.load(G / "undistort.npz")
    x, y = undistort_inputs()
    for c, xy, b in zip(g["cams"], g["xy_un"], g["bounds"]):
        small_ctx.set_camera(c[:4], c[4:] if c[8] != 0 else c[4:8])
        ux, uy = small_ctx.undistort_points(x, y)
        assert np.array_equal(ux, xy[:, 0]) and np.array_equal(uy, xy[:, 1])
        assert np.array_equal(small_ctx.image_bounds(752, 480), b)
