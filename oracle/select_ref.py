"""Oracle: CPU restatement of the reference's post-network stages.  TEST INFRASTRUCTURE ONLY.

Follows src/Extractors/HFNetRTModel.cc:139-196 (``GetLocalFeaturesFromTensor``), the bilinear ``Resampler`` in
src/Extractors/BaseModel.cc:491-562, the per-level budget and pyramid of src/Extractors/HFextractor.cc:108-119,159-173
and the level concatenation of src/Extractors/HFextractor.cc:255-284.

Pins (checked in tests/test_oracle_pins.py): ``l2_normalize_rows`` against ``cv2.normalize`` (the reference calls
``cv::normalize``), ``compute_pyramid`` *is* ``cv2.resize(INTER_LINEAR)`` (the reference calls ``cv::resize``);
the NMS restatement lives in hfnet_ref.simple_nms and is pinned against a brute-force window scan.

Tie-break: ``std::nth_element`` leaves both the order of the kept keypoints and the choice among equal responses at
the cut unspecified (HFNetRTModel.cc:172-179).  We define: keep the k largest by (response desc, scan index asc)
where the scan index is the reference's visit order (column-major: ``col * H + row``, HFNetRTModel.cc:155-168), and
emit them in that order.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def features_per_level(nfeatures: int, nlevels: int, scale_factor: float) -> List[int]:
    """src/Extractors/HFextractor.cc:108-119 (float arithmetic, cvRound = round-half-to-even)."""
    factor = np.float32(1.0) / np.float32(scale_factor)
    n = np.float32(nfeatures) * (np.float32(1) - factor) / (np.float32(1) - np.float32(float(factor) ** nlevels))
    out, total = [], 0
    for _ in range(nlevels - 1):
        v = int(np.rint(np.float32(n)))
        out.append(v)
        total += v
        n = np.float32(n * factor)
    out.append(max(nfeatures - total, 0))
    return out


def level_sizes(height: int, width: int, nlevels: int, scale_factor: float) -> List[Tuple[int, int]]:
    """(rows, cols) per level: cvRound(cols * invScale), src/Extractors/HFextractor.cc:92-103,159-173."""
    sizes = [(height, width)]
    sf = np.float32(1.0)
    for _ in range(1, nlevels):
        sf = np.float32(sf * np.float32(scale_factor))
        inv = np.float32(1.0) / sf
        sizes.append((int(np.rint(np.float32(height) * inv)), int(np.rint(np.float32(width) * inv))))
    return sizes


def compute_pyramid(image: np.ndarray, nlevels: int, scale_factor: float) -> List[np.ndarray]:
    """src/Extractors/HFextractor.cc:159-173: each level is cv::resize(INTER_LINEAR) of the PREVIOUS level."""
    import cv2
    pyr = [image]
    for (h, w) in level_sizes(image.shape[0], image.shape[1], nlevels, scale_factor)[1:]:
        pyr.append(cv2.resize(pyr[-1], (w, h), interpolation=cv2.INTER_LINEAR))
    return pyr


def resize_linear_u8(src: np.ndarray, dh: int, dw: int) -> np.ndarray:
    """Restatement of OpenCV's 8-bit INTER_LINEAR resize (fixed-point, 11-bit coefficients):
    fx = (dx+0.5)*scale-0.5; coefficient = saturate_cast<short>(frac*2048) (round-half-even); horizontal pass keeps
    int32 = s0*a0+s1*a1, vertical pass out = (((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2.
    Pinned against cv2.resize in tests/test_oracle_pins.py."""
    sh, sw = src.shape

    def coeffs(dn, sn):
        scale = sn / dn
        idx = np.empty(dn, np.int64)
        a = np.empty((dn, 2), np.int32)
        for d in range(dn):
            f = np.float32((d + 0.5) * scale - 0.5)
            s = int(np.floor(f))
            f = np.float32(f - s)
            if s < 0:
                s, f = 0, np.float32(0)
            if s >= sn - 1:
                s, f = sn - 1, np.float32(0)
            idx[d] = s
            a1 = int(np.rint(np.float32(f) * np.float32(2048)))
            a0 = int(np.rint((np.float32(1) - np.float32(f)) * np.float32(2048)))
            a[d] = (a0, a1)
        return idx, a

    xi, xa = coeffs(dw, sw)
    yi, ya = coeffs(dh, sh)
    s32 = src.astype(np.int32)
    x1 = np.minimum(xi + 1, sw - 1)
    rows = s32[:, xi] * xa[:, 0][None, :] + s32[:, x1] * xa[:, 1][None, :]       # [sh, dw]
    y1 = np.minimum(yi + 1, sh - 1)
    r0 = rows[yi] >> 4
    r1 = rows[y1] >> 4
    out = (((ya[:, 0][:, None] * r0) >> 16) + ((ya[:, 1][:, None] * r1) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def select_topk(scores_nms: np.ndarray, n_keypoints: int, threshold: float):
    """HFNetRTModel.cc:150-179.  Returns (cols int32[n], rows int32[n], response f32[n]) ordered by
    (response desc, col*H+row asc)."""
    H, W = scores_nms.shape
    s_cm = np.ascontiguousarray(scores_nms.T)                 # column-major visit order
    flat = s_cm.reshape(-1)
    cand = np.flatnonzero(flat >= np.float32(threshold))      # ascending scan index
    resp = flat[cand]
    order = np.lexsort((cand, -resp.astype(np.float64)))      # primary: response desc, secondary: scan index asc
    order = order[:n_keypoints]
    idx = cand[order]
    return (idx // H).astype(np.int32), (idx % H).astype(np.int32), resp[order].astype(np.float32)


def resample_bilinear(desc_map: np.ndarray, warp_xy: np.ndarray) -> np.ndarray:
    """src/Extractors/BaseModel.cc:491-562, expression order of :540-550 kept, every product/sum rounded to fp32
    (no FMA contraction): ((dx*dy)*f00 + ((1-dx)*(1-dy))*f11) + (dx*(1-dy))*f01) + ((1-dx)*dy)*f10."""
    Hd, Wd, C = desc_map.shape
    f32 = np.float32
    n = warp_xy.shape[0]
    out = np.zeros((n, C), f32)
    x = warp_xy[:, 0].astype(f32)
    y = warp_xy[:, 1].astype(f32)
    inside = (x > f32(-1)) & (y > f32(-1)) & (x < f32(Wd)) & (y < f32(Hd))
    fx = np.floor(x).astype(np.int64)
    fy = np.floor(y).astype(np.int64)
    cx, cy = fx + 1, fy + 1
    dx = (cx.astype(f32) - x).astype(f32)
    dy = (cy.astype(f32) - y).astype(f32)

    def get(xx, yy):
        ok = (xx >= 0) & (yy >= 0) & (xx <= Wd - 1) & (yy <= Hd - 1)
        v = desc_map[np.clip(yy, 0, Hd - 1), np.clip(xx, 0, Wd - 1)]
        return np.where(ok[:, None], v, f32(0)).astype(f32)

    one = f32(1)
    w00 = (dx * dy).astype(f32)[:, None]
    w11 = ((one - dx) * (one - dy)).astype(f32)[:, None]
    w01 = (dx * (one - dy)).astype(f32)[:, None]
    w10 = ((one - dx) * dy).astype(f32)[:, None]
    a = (w00 * get(fx, fy)).astype(f32)
    b = (w11 * get(cx, cy)).astype(f32)
    c = (w01 * get(fx, cy)).astype(f32)
    d = (w10 * get(cx, fy)).astype(f32)
    val = (((a + b).astype(f32) + c).astype(f32) + d).astype(f32)
    out[inside] = val[inside]
    return out


def l2_normalize_rows(m: np.ndarray) -> np.ndarray:
    """HFNetRTModel.cc:192-195 ``cv::normalize(row, row)`` (NORM_L2, alpha 1): the norm is accumulated in double,
    the scale 1/norm is rounded to fp32 and the row is multiplied in fp32 (``convertTo`` with a float scale).
    Bit-identical to cv2.normalize on 1000 random rows (tests/test_oracle_pins.py)."""
    m = np.asarray(m, np.float32)
    nrm = np.sqrt((m.astype(np.float64) ** 2).sum(axis=1))
    scale = np.where(nrm > np.finfo(np.float64).eps, 1.0 / np.maximum(nrm, 1e-300), 0.0).astype(np.float32)
    return (m * scale[:, None]).astype(np.float32)


def local_features(scores_nms: np.ndarray, desc_map: np.ndarray, n_keypoints: int, threshold: float):
    """HFNetRTModel.cc:139-196 for one level.  Returns dict(x, y, response, descriptors)."""
    H, W = scores_nms.shape
    Hd, Wd, _ = desc_map.shape
    f32 = np.float32
    scale_w = f32(f32(Wd) - f32(1)) / f32(f32(W) - f32(1))
    scale_h = f32(f32(Hd) - f32(1)) / f32(f32(H) - f32(1))
    cols, rows, resp = select_topk(scores_nms, n_keypoints, threshold)
    warp = np.stack([(scale_w * cols.astype(f32)).astype(f32), (scale_h * rows.astype(f32)).astype(f32)], axis=1)
    desc = l2_normalize_rows(resample_bilinear(desc_map, warp))
    return {"x": cols.astype(f32), "y": rows.astype(f32), "response": resp, "descriptors": desc}


def concat_levels(per_level: list, scale_factor: float):
    """src/Extractors/HFextractor.cc:272-281: octave = level, pt *= mvScaleFactor[level], vconcat descriptors."""
    xs, ys, rs, octs, ds = [], [], [], [], []
    sf = np.float32(1.0)
    for level, f in enumerate(per_level):
        if level > 0:
            sf = np.float32(sf * np.float32(scale_factor))
        xs.append((f["x"] * sf).astype(np.float32))
        ys.append((f["y"] * sf).astype(np.float32))
        rs.append(f["response"])
        octs.append(np.full(len(f["x"]), level, np.int32))
        ds.append(f["descriptors"])
    return {"x": np.concatenate(xs), "y": np.concatenate(ys), "response": np.concatenate(rs),
            "octave": np.concatenate(octs), "descriptors": np.concatenate(ds, axis=0)}


def image_bounds(width, height, K, dist):
    """Frame::ComputeImageBounds (src/Frame.cc:796-825): (mnMinX, mnMaxX, mnMinY, mnMaxY) from the undistorted image corners
    (0,0), (cols,0), (0,rows), (cols,rows); the plain rectangle without distortion."""
    dist = np.asarray(dist, np.float32).reshape(-1)
    if dist.size == 0 or dist[0] == 0.0:
        return np.array([0.0, width, 0.0, height], np.float32)
    cx = np.array([0, width, 0, width], np.float32)
    cy = np.array([0, 0, height, height], np.float32)
    ux, uy = undistort_points(cx, cy, K, dist)
    return np.array([min(ux[0], ux[2]), max(ux[1], ux[3]), min(uy[0], uy[1]), max(uy[2], uy[3])], np.float32)


def undistort_points(x, y, K, dist):
    """Frame::UndistortKeyPoints (src/Frame.cc:760-793): ``cv::undistortPoints(mat, mat, K, mDistCoef, cv::Mat(), mK)`` on the
    N x 2 float keypoint coordinates; the identity when ``dist[0] == 0`` (:762-766).  OpenCV is a third-party dependency of
    the reference (CMakeLists.txt: OpenCV 4.x; not under /root/reference), so this restates its published algorithm
    (``cvUndistortPointsInternal``, calib3d/src/undistort.dispatch.cpp: normalise, 5 fixed-point iterations of the inverse
    Brown-Conrady model in double -- the 6-argument overload's criteria are TermCriteria(COUNT, 5, 0.01) -- re-project with
    P = K, round to float) and is pinned bit for bit to ``cv2.undistortPoints`` in tests/test_oracle_pins.py.
    K = (fx, fy, cx, cy) float32, dist = (k1, k2, p1, p2[, k3]) float32."""
    x = np.asarray(x, np.float32)
    y = np.asarray(y, np.float32)
    dist = np.asarray(dist, np.float32).reshape(-1)
    if dist.size == 0 or dist[0] == 0.0:
        return x.copy(), y.copy()
    fx, fy, cx, cy = [float(np.float32(v)) for v in K]
    k = np.zeros(12, np.float64)
    k[:dist.size] = dist.astype(np.float64)
    ifx, ify = 1.0 / fx, 1.0 / fy
    u, v = x.astype(np.float64), y.astype(np.float64)
    xn = (u - cx) * ifx
    yn = (v - cy) * ify
    x0, y0 = xn.copy(), yn.copy()
    alive = np.ones(xn.shape, bool)
    for _ in range(5):
        r2 = xn * xn + yn * yn
        icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2)
        gave_up = alive & (icdist < 0)       # OpenCV: x = (u - cx) * ifx; y = (v - cy) * ify; break
        dx = 2 * k[2] * xn * yn + k[3] * (r2 + 2 * xn * xn) + k[8] * r2 + k[9] * r2 * r2
        dy = k[2] * (r2 + 2 * yn * yn) + 2 * k[3] * xn * yn + k[10] * r2 + k[11] * r2 * r2
        step = alive & ~gave_up
        xn = np.where(step, (x0 - dx) * icdist, np.where(gave_up, x0, xn))
        yn = np.where(step, (y0 - dy) * icdist, np.where(gave_up, y0, yn))
        alive = step
    xx = fx * xn + 0.0 * yn + cx
    yy = 0.0 * xn + fy * yn + cy
    ww = 1.0 / (0.0 * xn + 0.0 * yn + 1.0)
    return (xx * ww).astype(np.float32), (yy * ww).astype(np.float32)
