"""The oracle against the committed golden fixtures (tests/golden/, produced by make_golden.py)."""
from pathlib import Path

import numpy as np
import torch

from hfnet_slam_b200 import synthetic, weights
from oracle import hfnet_ref, kfdb_ref, lba_ref, match_ref, select_ref

G = Path(__file__).resolve().parent / "golden"


def tail_inputs():
    rng = np.random.default_rng(7)
    s = (rng.random((96, 128), dtype=np.float32) ** 6)
    s[10:13, 20:22] = 0.9
    dm = rng.normal(size=(12, 16, 256)).astype(np.float32)
    dm /= np.linalg.norm(dm, axis=-1, keepdims=True)
    return s, dm


def test_match_golden():
    g = np.load(G / "match.npz")
    na, nb, nt, seed = g["params"]
    A, B = synthetic.descriptor_pair(int(na), int(nb), n_true=int(nt), seed=int(seed))
    ia, ib, _ = match_ref.search_by_bow(A, B, 0.6)
    assert np.array_equal(np.stack([ia, ib], 1), g["bow_pairs"])
    i1, i2, c = match_ref.search_for_triangulation_core(A, B)
    assert np.array_equal(np.stack([i1, i2], 1), g["tri_pairs"]) and np.allclose(c, g["tri_cos"], atol=1e-6)


def test_tail_golden():
    g = np.load(G / "tail.npz")
    s, dm = tail_inputs()
    nms = hfnet_ref.simple_nms(torch.from_numpy(s)[None], 4, 2)[0].numpy()
    assert np.array_equal(np.argwhere(nms > 0).astype(np.int16), g["nms_nonzero"])
    f = select_ref.local_features(nms, dm, 60, 0.05)
    for k in ("x", "y", "response", "descriptors"):
        assert np.array_equal(f[k], g[k]), k


def test_pyramid_golden():
    g = np.load(G / "pyramid.npz")
    img = weights.synthetic_image(120, 188, seed=3, n_corners=30)
    cur = img
    for k, (h, w) in zip(("l1", "l2", "l3"), select_ref.level_sizes(120, 188, 4, 1.2)[1:]):
        cur = select_ref.resize_linear_u8(cur, h, w)
        assert np.array_equal(cur, g[k]), k


def test_kfdb_golden():
    g = np.load(G / "kfdb.npz")
    n, dim, npl, seed = g["params"]
    db, q, _ = synthetic.keyframe_db(int(n), int(dim), n_planted=int(npl), seed=int(seed))
    sc = kfdb_ref.scores(q[0], db)
    assert np.allclose(sc, g["scores"], atol=1e-7)
    sel, best = kfdb_ref.candidate_set(sc, 0.8)
    assert np.array_equal(sel.astype(np.int32), g["cand"]) and abs(best - float(g["best"])) < 1e-7


def test_lba_golden():
    g = np.load(G / "lba.npz")
    no, nf, npts, seed = g["params"]
    d = synthetic.lba_problem(n_opt=int(no), n_fixed=int(nf), n_points=int(npts), seed=int(seed))
    pr = lba_ref.Problem(d["poses"], d["fixed"], d["points"], d["cam_idx"], d["pt_idx"], d["obs"], d["inv_sigma2"], d["K"])
    r = lba_ref.optimize(pr, 10)
    assert [r.iterations, r.trials] == g["iterations"].tolist()
    assert np.allclose(r.poses, g["poses"], atol=1e-9) and np.allclose(r.points, g["points"], atol=1e-9)
    assert np.array_equal(r.outlier, g["outlier"])


def test_hfnet_golden():
    g = np.load(G / "hfnet.npz")
    wd = weights.synthetic(seed=0)
    im = weights.synthetic_image(64, 96, seed=2, n_corners=12)
    o = hfnet_ref.forward(im, wd, want_global=True)
    assert np.abs(o["scores_dense"][0] - g["scores_dense"].astype(np.float32)).max() < 2e-3
    gd = o["global_descriptor"][0]
    assert float(gd @ g["global_descriptor"]) > 0.9999


def undistort_inputs():
    rng = np.random.default_rng(17)
    return rng.uniform(-20, 772, 256).astype(np.float32), rng.uniform(-20, 500, 256).astype(np.float32)


def test_undistort_golden():
    """tests/golden/undistort.npz holds cv2.undistortPoints' own outputs: the oracle must reproduce them bit for bit."""
    g = np.load(G / "undistort.npz")
    x, y = undistort_inputs()
    for c, xy, b in zip(g["cams"], g["xy_un"], g["bounds"]):
        dist = c[4:] if c[8] != 0 else c[4:8]
        ux, uy = select_ref.undistort_points(x, y, c[:4], dist)
        assert np.array_equal(ux, xy[:, 0]) and np.array_equal(uy, xy[:, 1])
        assert np.array_equal(select_ref.image_bounds(752, 480, c[:4], dist), b)
