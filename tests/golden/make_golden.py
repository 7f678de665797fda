"""Generates tests/golden/*.npz from the CPU oracle (and cv2 where the reference calls OpenCV).  Inputs are seeded
(hfnet_slam_b200.synthetic / weights), so only the expected OUTPUTS are stored.  Run from the repository root:

    python tests/golden/make_golden.py

The reference itself cannot run in this image (no TensorRT / OpenCV C++ / Eigen), so these fixtures freeze the oracle;
the oracle in turn is pinned in tests/test_oracle_pins.py."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
OUT = Path(__file__).resolve().parent


def undistort_fixture(cv2):
    """Frame::UndistortKeyPoints / ComputeImageBounds: outputs of cv2.undistortPoints ITSELF (the function the reference
    calls, src/Frame.cc:778,809) for EuRoC cam0 and a 5-coefficient fisheye-ish lens on seeded points."""
    rng = np.random.default_rng(17)
    x = rng.uniform(-20, 772, 256).astype(np.float32)
    y = rng.uniform(-20, 500, 256).astype(np.float32)
    cams = np.array([[458.654, 457.296, 367.215, 248.375, -0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0],
                     [380.0, 381.0, 376.0, 240.0, -0.31, 0.12, 0.0007, -0.0004, -0.02]], np.float32)
    out, bounds = [], []
    for c in cams:
        Km = np.array([[c[0], 0, c[2]], [0, c[1], c[3]], [0, 0, 1]], np.float32)
        d = c[4:] if c[8] != 0 else c[4:8]
        out.append(cv2.undistortPoints(np.stack([x, y], 1).reshape(-1, 1, 2), Km, d, None, Km).reshape(-1, 2))
        cr = cv2.undistortPoints(np.array([[[0, 0]], [[752, 0]], [[0, 480]], [[752, 480]]], np.float32), Km, d, None, Km).reshape(4, 2)
        bounds.append([min(cr[0, 0], cr[2, 0]), max(cr[1, 0], cr[3, 0]), min(cr[0, 1], cr[1, 1]), max(cr[2, 1], cr[3, 1])])
    np.savez_compressed(OUT / "undistort.npz", seed=np.array([17]), cams=cams, xy_un=np.stack(out).astype(np.float32),
                        bounds=np.array(bounds, np.float32))


def main():
    import cv2
    import torch
    from hfnet_slam_b200 import synthetic, weights
    from oracle import hfnet_ref, kfdb_ref, lba_ref, match_ref, select_ref

    if len(sys.argv) > 1 and sys.argv[1] == "undistort":     # only this fixture (the others stay byte-identical)
        undistort_fixture(cv2)
        return
    undistort_fixture(cv2)

    # --- matcher (C1-shaped, smaller): cv2.BFMatcher is the function the reference calls
    A, B = synthetic.descriptor_pair(400, 380, n_true=150, seed=1)
    ms = cv2.BFMatcher(cv2.NORM_L2, crossCheck=True).match(A, B)
    bow = np.array(sorted((m.queryIdx, m.trainIdx) for m in ms if m.distance < 0.6), np.int32)
    i1, i2, c = match_ref.search_for_triangulation_core(A, B)
    np.savez_compressed(OUT / "match.npz", params=np.array([400, 380, 150, 1]), bow_pairs=bow,
                        tri_pairs=np.stack([i1, i2], 1).astype(np.int32), tri_cos=c)

    # --- network tail: NMS + select + sample on a seeded map
    rng = np.random.default_rng(7)
    s = (rng.random((96, 128), dtype=np.float32) ** 6)
    s[10:13, 20:22] = 0.9
    dm = rng.normal(size=(12, 16, 256)).astype(np.float32)
    dm /= np.linalg.norm(dm, axis=-1, keepdims=True)
    nms = hfnet_ref.simple_nms(torch.from_numpy(s)[None], 4, 2)[0].numpy()
    f = select_ref.local_features(nms, dm, 60, 0.05)
    np.savez_compressed(OUT / "tail.npz", seed=np.array([7]), nms_nonzero=np.argwhere(nms > 0).astype(np.int16),
                        x=f["x"], y=f["y"], response=f["response"], descriptors=f["descriptors"])

    # --- pyramid
    img = weights.synthetic_image(120, 188, seed=3, n_corners=30)
    pyr = select_ref.compute_pyramid(img, 4, 1.2)
    np.savez_compressed(OUT / "pyramid.npz", l1=pyr[1], l2=pyr[2], l3=pyr[3])

    # --- keyframe DB
    db, q, qi = synthetic.keyframe_db(600, 4096, n_planted=60, seed=4)
    sc = kfdb_ref.scores(q[0], db)
    sel, best = kfdb_ref.candidate_set(sc, 0.8)
    np.savez_compressed(OUT / "kfdb.npz", params=np.array([600, 4096, 60, 4]), scores=sc, cand=sel.astype(np.int32),
                        best=np.float32(best))

    # --- local BA
    d = synthetic.lba_problem(n_opt=5, n_fixed=4, n_points=150, seed=6)
    pr = lba_ref.Problem(d["poses"], d["fixed"], d["points"], d["cam_idx"], d["pt_idx"], d["obs"], d["inv_sigma2"], d["K"])
    r = lba_ref.optimize(pr, 10)
    np.savez_compressed(OUT / "lba.npz", params=np.array([5, 4, 150, 6]), poses=r.poses, points=r.points, chi2=r.chi2,
                        outlier=r.outlier, iterations=np.array([r.iterations, r.trials]), chis=np.array(r.chis))

    # --- encoder on a 64x96 frame
    wd = weights.synthetic(seed=0)
    im = weights.synthetic_image(64, 96, seed=2, n_corners=12)
    o = hfnet_ref.forward(im, wd, want_global=True)
    np.savez_compressed(OUT / "hfnet.npz", scores_dense=o["scores_dense"][0].astype(np.float16),
                        global_descriptor=o["global_descriptor"][0],
                        desc_map_rows=o["local_descriptor_map"][0].reshape(-1, 256)[::7].astype(np.float16))
    for p in sorted(OUT.glob("*.npz")):
        print(p.name, p.stat().st_size)


if __name__ == "__main__":
    main()
