"""Oracle: CPU restatement of the 256-d local-descriptor matching flavours.  TEST INFRASTRUCTURE ONLY.

Follows src/Matcher.cc:
* ``descriptor_distance``            :1893-1900  (``(a-b).norm()``, fp32)
* ``search_by_bow``                  :220-263, :561-621  (cv::BFMatcher(NORM_L2, crossCheck=true) then ``dist < TH_LOW``)
* ``search_for_triangulation_core``  :845-889   (sgemm ``D1*D2^T``, row argmax above ``1-0.5*TH_HIGH^2``, column
                                                 cross-check; strict ``>`` so the lowest index wins ties)
* ``best2_masked`` / ``accept_projection``  :40-125 (``SearchByProjection(F, MPs)`` best / second best over the
                                                 candidate window, level-aware ratio test)
and the GEMM + ratio + mutual variant of Examples/Utility/test_match_local_feats.cc:48-121 (``search_ratio_mutual``).

Pin: ``search_by_bow`` is checked pair-for-pair against ``cv2.BFMatcher(cv2.NORM_L2, crossCheck=True)`` (the
function the reference calls) in tests/test_oracle_pins.py.  The floating-point summation order of Eigen's sgemm
is unpinned; index outputs are exact away from fp32 ties.
"""
from __future__ import annotations

import numpy as np

TH_HIGH = np.float32(0.75)   # Matcher.cc:33
TH_LOW = np.float32(0.6)     # Matcher.cc:34
COS_FLOOR = np.float32(-0.5 * 0.75 * 0.75 + 1)   # Matcher.cc:851  = 0.71875


def descriptor_distance(a: np.ndarray, b: np.ndarray) -> np.float32:
    d = (np.asarray(a, np.float32) - np.asarray(b, np.float32)).astype(np.float32)
    return np.float32(np.sqrt(np.sum(d * d, dtype=np.float32)))


def l2_distance_matrix(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """Explicit ||a_i - b_j||_2 in fp32-of-fp64 (what OpenCV's batchDistance NORM_L2 evaluates, up to rounding)."""
    A64, B64 = A.astype(np.float64), B.astype(np.float64)
    d2 = (A64 * A64).sum(1)[:, None] + (B64 * B64).sum(1)[None, :] - 2.0 * (A64 @ B64.T)
    return np.sqrt(np.maximum(d2, 0.0)).astype(np.float32)


def search_by_bow(A: np.ndarray, B: np.ndarray, max_dist: float = float(TH_LOW)):
    """Mutual nearest neighbours by L2 distance, kept iff dist < max_dist (strict, Matcher.cc:253).
    Returns (idxA int32[m], idxB int32[m], dist f32[m]) sorted by idxA."""
    if A.shape[0] == 0 or B.shape[0] == 0:
        z = np.zeros(0, np.int32)
        return z, z.copy(), np.zeros(0, np.float32)
    D = l2_distance_matrix(A, B)
    nn_ab = D.argmin(axis=1)          # first minimum on ties
    nn_ba = D.argmin(axis=0)
    ia = np.arange(A.shape[0])
    mutual = nn_ba[nn_ab] == ia
    dist = D[ia, nn_ab]
    keep = mutual & (dist < np.float32(max_dist))
    return ia[keep].astype(np.int32), nn_ab[keep].astype(np.int32), dist[keep]


def search_for_triangulation_core(D1: np.ndarray, D2: np.ndarray, floor: float = float(COS_FLOOR)):
    """Matcher.cc:845-889 without the geometric filters.  Returns (idx1, idx2, cos) sorted by idx1."""
    if D1.shape[0] == 0 or D2.shape[0] == 0:
        z = np.zeros(0, np.int32)
        return z, z.copy(), np.zeros(0, np.float32)
    S = (D1.astype(np.float64) @ D2.astype(np.float64).T).astype(np.float32)
    fl = np.float32(floor)
    Sm = np.where(S > fl, S, -np.inf)
    row_best = Sm.argmax(axis=1)                    # first max on ties == strict '>' scan
    row_ok = np.isfinite(Sm[np.arange(S.shape[0]), row_best])
    col_best = Sm.argmax(axis=0)
    i = np.arange(S.shape[0])
    keep = row_ok & (col_best[row_best] == i)
    return i[keep].astype(np.int32), row_best[keep].astype(np.int32), S[i[keep], row_best[keep]]


def search_ratio_mutual(D1: np.ndarray, D2: np.ndarray, ratio: float, threshold: float, mutual: bool = True):
    """Examples/Utility/test_match_local_feats.cc:48-121: dist = 2*(1 - d1.d2); best < threshold, best < ratio*second,
    optional cross-check (strict '<', first index wins)."""
    S = (D1.astype(np.float64) @ D2.astype(np.float64).T).astype(np.float32)
    dist = (np.float32(2) * (np.float32(1) - S)).astype(np.float32)
    out = []
    col_best = dist.argmin(axis=0)
    for i in range(dist.shape[0]):
        row = dist[i]
        j = int(row.argmin())
        b1 = row[j]
        if row.shape[0] > 1:
            rest = np.delete(row, j)
            b2 = rest.min()
        else:
            b2 = np.float32(np.finfo(np.float32).max)
        if b1 < np.float32(threshold) and b1 < np.float32(ratio) * b2:
            if (not mutual) or col_best[j] == i:
                out.append((i, j, b1))
    if not out:
        z = np.zeros(0, np.int32)
        return z, z.copy(), np.zeros(0, np.float32)
    a = np.array(out)
    return a[:, 0].astype(np.int32), a[:, 1].astype(np.int32), a[:, 2].astype(np.float32)


def best2_masked(Q: np.ndarray, F: np.ndarray, cand_ptr: np.ndarray, cand_idx: np.ndarray, f_level: np.ndarray):
    """Best / second-best L2 distance of each query over its ragged candidate list (Matcher.cc:78-117 inner loop).
    cand_ptr int[M+1], cand_idx int[...] (visit order kept).  Returns best_idx, best_dist, best_level, second_dist,
    second_level (idx -1 / dist FLT_MAX / level -1 where absent)."""
    M = Q.shape[0]
    fmax = np.finfo(np.float32).max
    bi = np.full(M, -1, np.int32)
    bd = np.full(M, fmax, np.float32)
    bl = np.full(M, -1, np.int32)
    sd = np.full(M, fmax, np.float32)
    sl = np.full(M, -1, np.int32)
    for m in range(M):
        for idx in cand_idx[cand_ptr[m]:cand_ptr[m + 1]]:
            d = np.float32(np.sqrt(((Q[m].astype(np.float64) - F[idx].astype(np.float64)) ** 2).sum()))
            if d < bd[m]:
                sd[m], sl[m] = bd[m], bl[m]
                bd[m], bl[m], bi[m] = d, f_level[idx], idx
            elif d < sd[m]:
                sd[m], sl[m] = d, f_level[idx]
    return bi, bd, bl, sd, sl


def accept_projection(bd, bl, sd, sl, ratio: float, th_high: float = float(TH_HIGH)):
    """Matcher.cc:119-125: accept iff best <= TH_HIGH and not (same level and best > ratio*second)."""
    ok = bd <= np.float32(th_high)
    reject = (bl == sl) & (bd > np.float32(ratio) * sd)
    return ok & ~reject


def search_by_projection_map_points(Q: np.ndarray, uv: np.ndarray, radius: np.ndarray, min_level: np.ndarray,
                                    max_level: np.ndarray, F: np.ndarray, f_xy: np.ndarray, f_level: np.ndarray,
                                    occupied=None, ratio: float = 0.8, th_high: float = float(TH_HIGH)) -> np.ndarray:
    """Matcher::SearchByProjection(F, vpMapPoints, th, ...) (src/Matcher.cc:40-210), the descriptor + bookkeeping part:
    map points in order; window of Frame::GetFeaturesInArea (|x-u| < r, |y-v| < r, octave in [min, max], max < 0 =
    unbounded, src/Frame.cc:659-725); features that already carry a map point with observations are skipped (:84-86) --
    which includes every feature claimed by an earlier map point of this call (:137); best / second best with their
    levels (:88-117); accepted iff best <= TH_HIGH and not (bestLevel == bestLevel2 and best > ratio * second) (:119-125).
    Returns match[i] = feature index or -1."""
    M, N = Q.shape[0], F.shape[0]
    taken = np.zeros(N, bool) if occupied is None else np.array(occupied, bool, copy=True)
    out = np.full(M, -1, np.int32)
    fmax = np.finfo(np.float32).max
    for i in range(M):
        r = np.float32(radius[i])
        ok = (np.abs(f_xy[:, 0] - np.float32(uv[i, 0])) < r) & (np.abs(f_xy[:, 1] - np.float32(uv[i, 1])) < r)
        ok &= f_level >= min_level[i]
        if max_level[i] >= 0:
            ok &= f_level <= max_level[i]
        bd, bl, bi, sd, sl = fmax, -1, -1, fmax, -1
        for j in np.flatnonzero(ok):
            if taken[j]:
                continue
            d = descriptor_distance(Q[i], F[j])
            if d < bd:
                sd, sl = bd, bl
                bd, bl, bi = d, int(f_level[j]), int(j)
            elif d < sd:
                sd, sl = d, int(f_level[j])
        if bd <= np.float32(th_high):
            if bl == sl and bd > np.float32(ratio) * np.float32(sd):
                continue
            out[i] = bi
            taken[bi] = True
    return out


def search_for_initialization(d1: np.ndarray, xy1: np.ndarray, oct1: np.ndarray, d2: np.ndarray, xy2: np.ndarray,
                              oct2: np.ndarray, prev_matched: np.ndarray, nnratio: float = 0.9, window: float = 100.0):
    """Matcher::SearchForInitialization (src/Matcher.cc:486-559), literal: level-0 keypoints of frame 1 in order, window
    search around ``prev_matched`` (Frame::GetFeaturesInArea with minLevel = maxLevel = 0, src/Frame.cc:659-725: the grid
    only pre-filters ``|dx| < r and |dy| < r``), candidates whose recorded match distance is <= the new distance are
    skipped, best <= TH_LOW and best < nnratio * second, a claimed feature is taken over by the better match.
    Returns (matches12 int32[N1], n_matches, updated prev_matched).  Candidates are visited in ascending index here, by
    grid cell in the reference: only exact fp32 distance ties could tell the difference."""
    n1, n2 = d1.shape[0], d2.shape[0]
    fmax = np.finfo(np.float32).max
    m12 = np.full(n1, -1, np.int32)
    m21 = np.full(n2, -1, np.int32)
    matched_dist = np.full(n2, fmax, np.float32)
    n = 0
    r = np.float32(window)
    for i1 in range(n1):
        if oct1[i1] > 0:
            continue
        u, v = prev_matched[i1]
        ok = (np.abs(xy2[:, 0] - u) < r) & (np.abs(xy2[:, 1] - v) < r) & (oct2 == 0)
        best, best2, bidx = fmax, fmax, -1
        for i2 in np.flatnonzero(ok):
            dist = descriptor_distance(d1[i1], d2[i2])
            if matched_dist[i2] <= dist:
                continue
            if dist < best:
                best2, best, bidx = best, dist, int(i2)
            elif dist < best2:
                best2 = dist
        if best <= TH_LOW and best < np.float32(best2) * np.float32(nnratio):
            if m21[bidx] >= 0:
                m12[m21[bidx]] = -1
                n -= 1
            m12[i1] = bidx
            m21[bidx] = i1
            matched_dist[bidx] = best
            n += 1
    pm = np.array(prev_matched, np.float32, copy=True)
    hit = m12 >= 0
    pm[hit] = xy2[m12[hit]]
    return m12, n, pm


def search_by_projection_last_frame(Tcw: np.ndarray, K, bounds, scale_factors, last_points_w: np.ndarray,
                                    last_valid: np.ndarray, last_octave: np.ndarray, last_desc: np.ndarray,
                                    cur_desc: np.ndarray, cur_xy: np.ndarray, cur_octave: np.ndarray,
                                    cur_occupied: np.ndarray, th: float, th_high: float = float(TH_HIGH)):
    """Matcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono=true) (src/Matcher.cc:1574-1650), literal, monocular:
    every map point of the last frame (in order, skipping outliers / empty slots = ``last_valid`` False) is projected
    with the current pose estimate (Pinhole::project, src/CameraModels/Pinhole.cpp:35-41), searched in the window
    ``th * scale[octave]`` over octaves [oct-1, oct+1] (Frame::GetFeaturesInArea), skipping features that already carry
    a map point with observations (``cur_occupied``, extended as this loop assigns), best distance only, accepted iff
    ``best <= TH_HIGH``.  Tcw: 3x4 [R|t]; K = (fx, fy, cx, cy); bounds = (minX, maxX, minY, maxY).
    Returns (assigned int32[N_cur]: index of the last-frame map point or -1, n_matches)."""
    fx, fy, cx, cy = [np.float32(v) for v in K]
    mnx, mxx, mny, mxy = [np.float32(v) for v in bounds]
    R, t = Tcw[:, :3].astype(np.float32), Tcw[:, 3].astype(np.float32)
    occ = np.array(cur_occupied, bool, copy=True)
    assigned = np.full(cur_desc.shape[0], -1, np.int32)
    n = 0
    for i in range(last_points_w.shape[0]):
        if not last_valid[i]:
            continue
        xc = (R @ last_points_w[i].astype(np.float32) + t).astype(np.float32)
        invz = np.float32(1.0) / xc[2]
        if invz < 0:
            continue
        u = fx * xc[0] / xc[2] + cx
        v = fy * xc[1] / xc[2] + cy
        if u < mnx or u > mxx or v < mny or v > mxy:
            continue
        o = int(last_octave[i])
        r = np.float32(th) * np.float32(scale_factors[o])
        ok = (np.abs(cur_xy[:, 0] - u) < r) & (np.abs(cur_xy[:, 1] - v) < r) & (cur_octave >= o - 1) & (cur_octave <= o + 1)
        best, bidx = np.finfo(np.float32).max, -1
        for i2 in np.flatnonzero(ok):
            if occ[i2]:
                continue
            d = descriptor_distance(last_desc[i], cur_desc[i2])
            if d < best:
                best, bidx = d, int(i2)
        if best <= np.float32(th_high):
            assigned[bidx] = i
            occ[bidx] = True          # the assigned map point has observations: later points skip this feature
            n += 1
    return assigned, n


def fuse(Tcw: np.ndarray, Ow: np.ndarray, K, bounds, scale_factors, log_scale_factor: float, mp_pos: np.ndarray,
         mp_normal: np.ndarray, mp_min_dist: np.ndarray, mp_max_dist: np.ndarray, mp_desc: np.ndarray,
         mp_skip: np.ndarray, kf_desc: np.ndarray, kf_xy: np.ndarray, kf_octave: np.ndarray, th: float = 3.0,
         th_low: float = float(TH_LOW), chi2: float = 5.99):
    """Matcher::Fuse(pKF, vpMapPoints, th) (src/Matcher.cc:1046-1250), monocular, up to the map bookkeeping: for every map
    point (``mp_skip`` = null / bad / already in the keyframe) -- positive depth, inside the image, distance within the
    scale-invariance range [GetMinDistanceInvariance, GetMaxDistanceInvariance] = [mfMinDistance / 1.2f, 1.2f *
    mfMaxDistance] (src/MapPoint.cc:504-516; [0, 10000] for a single-level pyramid), viewing angle below 60 degrees
    (PO . Pn >= 0.5 dist), predicted level (MapPoint::PredictScale, src/MapPoint.cc:518-534:
    ceil(log(mfMaxDistance / dist) / logScaleFactor) on the RAW mfMaxDistance, clamped; 0 for a single level), window
    th * scale[level], keyframe features of level [pred-1, pred] that pass the reprojection gate
    e2 * invLevelSigma2 <= chi2, least descriptor distance, accepted iff <= TH_LOW.
    ``mp_min_dist`` / ``mp_max_dist`` are the map points' raw mfMinDistance / mfMaxDistance.
    Returns (best_idx int32[M] or -1, best_dist f32[M]): the feature each map point would be fused into (the reference
    then Replace()s or AddObservation()s, which is map data-model code outside the path)."""
    fx, fy, cx, cy = [np.float32(v) for v in K]
    mnx, mxx, mny, mxy = [np.float32(v) for v in bounds]
    R, t = Tcw[:, :3].astype(np.float32), Tcw[:, 3].astype(np.float32)
    sf = np.asarray(scale_factors, np.float32)
    inv_sigma2 = (np.float32(1.0) / (sf * sf)).astype(np.float32)
    n_levels = len(sf)
    M = mp_pos.shape[0]
    best_idx = np.full(M, -1, np.int32)
    best_dist = np.full(M, np.finfo(np.float32).max, np.float32)
    for i in range(M):
        if mp_skip[i]:
            continue
        pw = mp_pos[i].astype(np.float32)
        pc = (R @ pw + t).astype(np.float32)
        if pc[2] < 0:
            continue
        u = fx * pc[0] / pc[2] + cx
        v = fy * pc[1] / pc[2] + cy
        if not (mnx <= u < mxx and mny <= v < mxy):           # KeyFrame::IsInImage (src/KeyFrame.cc:812-815)
            continue
        PO = (pw - Ow.astype(np.float32)).astype(np.float32)
        dist3d = np.float32(np.sqrt(np.sum(PO * PO, dtype=np.float32)))
        if n_levels <= 1:
            min_inv, max_inv = np.float32(0), np.float32(10000)
        else:
            min_inv = np.float32(mp_min_dist[i]) / np.float32(1.2)
            max_inv = np.float32(1.2) * np.float32(mp_max_dist[i])
        if dist3d < min_inv or dist3d > max_inv:
            continue
        if np.float32(PO @ mp_normal[i].astype(np.float32)) < np.float32(0.5) * dist3d:
            continue
        if n_levels <= 1:
            lvl = 0
        else:
            ratio = np.float32(mp_max_dist[i]) / dist3d
            lvl = int(np.ceil(np.log(ratio) / np.float32(log_scale_factor)))
            lvl = min(max(lvl, 0), n_levels - 1)
        r = np.float32(th) * sf[lvl]
        ok = (np.abs(kf_xy[:, 0] - u) < r) & (np.abs(kf_xy[:, 1] - v) < r)
        for j in np.flatnonzero(ok):
            kl = int(kf_octave[j])
            if kl < lvl - 1 or kl > lvl:
                continue
            ex, ey = u - kf_xy[j, 0], v - kf_xy[j, 1]
            e2 = np.float32(ex * ex + ey * ey)
            if e2 * inv_sigma2[kl] > np.float32(chi2):
                continue
            d = descriptor_distance(mp_desc[i], kf_desc[j])
            if d < best_dist[i]:
                best_dist[i], best_idx[i] = d, int(j)
        if not best_dist[i] <= np.float32(th_low):
            best_idx[i] = -1
    return best_idx, best_dist


# ---------------------------------------------------------------------------------------------------------------------
# The remaining members of the projection family, transcribed loop for loop.  Shared helpers:
def _predict_scale(max_dist, dist, log_scale_factor, n_levels):
    """MapPoint::PredictScale (src/MapPoint.cc:518-552): ceil(log(mfMaxDistance / dist) / mfLogScaleFactor), clamped;
    0 for a single-level pyramid."""
    if n_levels <= 1:
        return 0
    ratio = np.float32(max_dist) / np.float32(dist)
    lvl = int(np.ceil(np.log(ratio) / np.float32(log_scale_factor)))
    return min(max(lvl, 0), n_levels - 1)


def _invariance(min_dist, max_dist, n_levels):
    """MapPoint::GetMinDistanceInvariance / GetMaxDistanceInvariance (src/MapPoint.cc:504-516)."""
    if n_levels <= 1:
        return np.float32(0), np.float32(10000)
    return np.float32(min_dist) / np.float32(1.2), np.float32(1.2) * np.float32(max_dist)


def search_by_projection_keyframe(Tcw, K, bounds, scale_factors, log_scale_factor, mp_pos, mp_min_dist, mp_max_dist,
                                  mp_desc, mp_skip, cur_desc, cur_xy, cur_octave, cur_occupied, th: float,
                                  threshold: float):
    """Matcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, threshold) (src/Matcher.cc:1723-1805), used by
    relocalisation: the keyframe's map points (``mp_skip`` = null / bad / already found) are projected with the frame's
    pose; frame bounds [mnMinX, mnMaxX] x [mnMinY, mnMaxY] (both ends inside, :1747-1750; NO positive-depth and NO
    viewing-angle test in this variant), invariance range, predicted level, window th * scale[level] over octaves
    [pred-1, pred+1] (Frame::GetFeaturesInArea with level bounds), features that already carry a map point skipped --
    including the ones claimed earlier in this call (:1795) --, nearest descriptor accepted iff <= threshold.
    Returns (assigned: map-point index per frame feature or -1, n_matches)."""
    fx, fy, cx, cy = [np.float32(v) for v in K]
    mnx, mxx, mny, mxy = [np.float32(v) for v in bounds]
    Tcw = np.asarray(Tcw, np.float32)
    R, t = Tcw[:, :3], Tcw[:, 3]
    Ow = (-(R.T @ t)).astype(np.float32)
    sf = np.asarray(scale_factors, np.float32)
    nl = len(sf)
    occ = np.array(cur_occupied, bool, copy=True)
    assigned = np.full(cur_desc.shape[0], -1, np.int32)
    n = 0
    for i in range(mp_pos.shape[0]):
        if mp_skip[i]:
            continue
        pw = mp_pos[i].astype(np.float32)
        pc = (R @ pw + t).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            u = fx * pc[0] / pc[2] + cx
            v = fy * pc[1] / pc[2] + cy
        if u < mnx or u > mxx or v < mny or v > mxy or not (np.isfinite(u) and np.isfinite(v)):
            continue
        PO = (pw - Ow).astype(np.float32)
        dist3d = np.float32(np.sqrt(np.sum(PO * PO, dtype=np.float32)))
        mn, mx = _invariance(mp_min_dist[i], mp_max_dist[i], nl)
        if dist3d < mn or dist3d > mx:
            continue
        lvl = _predict_scale(mp_max_dist[i], dist3d, log_scale_factor, nl)
        r = np.float32(th) * sf[lvl]
        ok = (np.abs(cur_xy[:, 0] - u) < r) & (np.abs(cur_xy[:, 1] - v) < r) & (cur_octave >= lvl - 1) & (cur_octave <= lvl + 1)
        best, bidx = np.finfo(np.float32).max, -1
        for j in np.flatnonzero(ok):
            if occ[j]:
                continue
            d = descriptor_distance(mp_desc[i], cur_desc[j])
            if d < best:
                best, bidx = d, int(j)
        if best <= np.float32(threshold):
            assigned[bidx] = i
            occ[bidx] = True
            n += 1
    return assigned, n


def search_by_projection_sim3(Tcw, Ow, K, bounds, scale_factors, log_scale_factor, mp_pos, mp_normal, mp_min_dist,
                              mp_max_dist, mp_desc, mp_skip, kf_desc, kf_xy, kf_octave, kf_matched, th: float,
                              threshold: float):
    """Matcher::SearchByProjection(pKF, Scw, vpPoints[, vpPointsKFs], vpMatched[, vpMatchedKF], th, threshold)
    (src/Matcher.cc:265-367 and :369-484; the second only also records the source keyframe of each point): Tcw / Ow from
    the Sim3 as at :275-276; per candidate point (``mp_skip`` = bad / already in vpMatched): positive depth,
    KeyFrame::IsInImage (min <= . < max), invariance range, viewing angle (PO . Pn >= 0.5 dist), predicted level, window
    th * scale[level] over ALL octaves, then octave in [pred-1, pred], keyframe features that are already matched skipped
    -- including those claimed earlier in this call --, nearest descriptor accepted iff <= threshold.
    Returns (matched: point index per keyframe feature or -1 (only the NEW matches), n_matches)."""
    fx, fy, cx, cy = [np.float32(v) for v in K]
    mnx, mxx, mny, mxy = [np.float32(v) for v in bounds]
    Tcw = np.asarray(Tcw, np.float32)
    R, t = Tcw[:, :3], Tcw[:, 3]
    Ow = np.asarray(Ow, np.float32)
    sf = np.asarray(scale_factors, np.float32)
    nl = len(sf)
    taken = np.array(kf_matched, bool, copy=True)
    matched = np.full(kf_desc.shape[0], -1, np.int32)
    n = 0
    for i in range(mp_pos.shape[0]):
        if mp_skip[i]:
            continue
        pw = mp_pos[i].astype(np.float32)
        pc = (R @ pw + t).astype(np.float32)
        if pc[2] < 0:
            continue
        u = fx * pc[0] / pc[2] + cx
        v = fy * pc[1] / pc[2] + cy
        if not (mnx <= u < mxx and mny <= v < mxy):
            continue
        PO = (pw - Ow).astype(np.float32)
        dist = np.float32(np.sqrt(np.sum(PO * PO, dtype=np.float32)))
        mn, mx = _invariance(mp_min_dist[i], mp_max_dist[i], nl)
        if dist < mn or dist > mx:
            continue
        if np.float32(PO @ mp_normal[i].astype(np.float32)) < np.float32(0.5) * dist:
            continue
        lvl = _predict_scale(mp_max_dist[i], dist, log_scale_factor, nl)
        r = np.float32(th) * sf[lvl]
        ok = (np.abs(kf_xy[:, 0] - u) < r) & (np.abs(kf_xy[:, 1] - v) < r)
        best, bidx = np.finfo(np.float32).max, -1
        for j in np.flatnonzero(ok):
            if taken[j]:
                continue
            kl = int(kf_octave[j])
            if kl < lvl - 1 or kl > lvl:
                continue
            d = descriptor_distance(mp_desc[i], kf_desc[j])
            if d < best:
                best, bidx = d, int(j)
        if best <= np.float32(threshold):
            matched[bidx] = i
            taken[bidx] = True
            n += 1
    return matched, n


def _sim3_directed(Tsrc_w, S_dst_src, K, bounds, sf, log_scale_factor, mp_pos, mp_min, mp_max, mp_desc, mp_valid, already,
                   dst_desc, dst_xy, dst_oct, th):
    """One direction of SearchBySim3 (src/Matcher.cc:1393-1466): the source keyframe's map points through the source pose
    and the similarity into the destination camera; best descriptor in the window, octave in [pred-1, pred], <= TH_HIGH."""
    fx, fy, cx, cy = [np.float32(v) for v in K]
    mnx, mxx, mny, mxy = [np.float32(v) for v in bounds]
    Rs, ts = np.asarray(Tsrc_w, np.float32)[:, :3], np.asarray(Tsrc_w, np.float32)[:, 3]
    s, Rd, td = np.float32(S_dst_src[0]), np.asarray(S_dst_src[1], np.float32), np.asarray(S_dst_src[2], np.float32)
    nl = len(sf)
    out = np.full(mp_pos.shape[0], -1, np.int32)
    for i in range(mp_pos.shape[0]):
        if not mp_valid[i] or already[i]:
            continue
        pw = mp_pos[i].astype(np.float32)
        p_src = (Rs @ pw + ts).astype(np.float32)
        p = (s * (Rd @ p_src) + td).astype(np.float32)
        if p[2] < 0:
            continue
        invz = np.float32(1.0) / p[2]
        u = fx * (p[0] * invz) + cx
        v = fy * (p[1] * invz) + cy
        if not (mnx <= u < mxx and mny <= v < mxy):
            continue
        dist3d = np.float32(np.sqrt(np.sum(p * p, dtype=np.float32)))
        mn, mx = _invariance(mp_min[i], mp_max[i], nl)
        if dist3d < mn or dist3d > mx:
            continue
        lvl = _predict_scale(mp_max[i], dist3d, log_scale_factor, nl)
        r = np.float32(th) * sf[lvl]
        ok = (np.abs(dst_xy[:, 0] - u) < r) & (np.abs(dst_xy[:, 1] - v) < r)
        best, bidx = np.finfo(np.float32).max, -1
        for j in np.flatnonzero(ok):
            kl = int(dst_oct[j])
            if kl < lvl - 1 or kl > lvl:
                continue
            d = descriptor_distance(mp_desc[i], dst_desc[j])
            if d < best:
                best, bidx = d, int(j)
        if best <= TH_HIGH:
            out[i] = bidx
    return out


def search_by_sim3(K, bounds, scale_factors, log_scale_factor, T1w, T2w, S12, S21, mp1_pos, mp1_min, mp1_max, mp1_desc,
                   mp1_valid, already1, mp2_pos, mp2_min, mp2_max, mp2_desc, mp2_valid, already2, desc1, xy1, oct1, desc2,
                   xy2, oct2, th: float):
    """Matcher::SearchBySim3(pKF1, pKF2, vpMatches12, S12, th) (src/Matcher.cc:1355-1572): feature i of a keyframe carries
    map point i (``mp*_valid`` = non-null and not bad; ``already*`` = vbAlreadyMatched*); KF1's points go through T1w and
    S21 = (scale, R, t) into camera 2 and are matched against KF2's features, KF2's through T2w and S12 against KF1's;
    a pair is kept iff both directions agree (:1553-1567).  Returns (match12: feature of KF2 per feature of KF1 or -1, n)."""
    sf = np.asarray(scale_factors, np.float32)
    m1 = _sim3_directed(T1w, S21, K, bounds, sf, log_scale_factor, mp1_pos, mp1_min, mp1_max, mp1_desc, mp1_valid, already1,
                        desc2, xy2, oct2, th)
    m2 = _sim3_directed(T2w, S12, K, bounds, sf, log_scale_factor, mp2_pos, mp2_min, mp2_max, mp2_desc, mp2_valid, already2,
                        desc1, xy1, oct1, th)
    out = np.full(len(m1), -1, np.int32)
    n = 0
    for i1 in range(len(m1)):
        i2 = m1[i1]
        if i2 >= 0 and m2[i2] == i1:
            out[i1] = i2
            n += 1
    return out, n
