"""Host-side mirror of the brute-force flavours of ``Matcher`` (include/Matcher.h:43-89, src/Matcher.cc) over the GPU
distance + arg-max kernel.  Constants as in src/Matcher.cc:33-34.  The geometric filters around the descriptor search
(epipolar test of SearchForTriangulation :894-909, map-point bookkeeping of SearchByBoW :236-262) use the caller's
map and stay on the host; these methods return the descriptor-level correspondences those loops consume."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from .lib import Context

TH_HIGH = 0.75
TH_LOW = 0.6
COS_FLOOR = float(np.float32(-0.5 * 0.75 * 0.75 + 1))   # src/Matcher.cc:851


class Matcher:
    def __init__(self, ctx: Context, nnratio: float = 0.6, check_orientation: bool = True):
        self.ctx, self.nnratio, self.check_orientation = ctx, nnratio, check_orientation

    @staticmethod
    def descriptor_distance(a: np.ndarray, b: np.ndarray) -> float:
        """Matcher::DescriptorDistance (src/Matcher.cc:1893-1900)."""
        d = np.asarray(a, np.float32) - np.asarray(b, np.float32)
        return float(np.sqrt(np.sum(d * d, dtype=np.float32)))

    def search_by_bow(self, desc1: np.ndarray, desc2: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """cv::BFMatcher(NORM_L2, crossCheck=true).match + dist < TH_LOW (src/Matcher.cc:229-253, 574-610).
        Returns (idx1, idx2, distance) sorted by idx1."""
        idx, val, _ = self.ctx.match_mutual_l2(desc1, desc2, TH_LOW)
        i = np.flatnonzero(idx >= 0)
        return i.astype(np.int32), idx[i], val[i]

    def search_for_triangulation(self, desc1: np.ndarray, desc2: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """The descriptor stage of SearchForTriangulation (src/Matcher.cc:845-889): (idx1, idx2, cosine)."""
        idx, val, _ = self.ctx.match_mutual_cos(desc1, desc2, COS_FLOOR)
        i = np.flatnonzero(idx >= 0)
        return i.astype(np.int32), idx[i], val[i]

    def search_for_triangulation_batch(self, desc1: np.ndarray, neighbours: Sequence[np.ndarray]) -> List[tuple]:
        """One CreateNewMapPoints (src/LocalMapping.cc:513-893): the current keyframe against <= 30 covisible
        keyframes in ONE launch."""
        n = len(neighbours)
        if n == 0:
            return []
        a = np.ascontiguousarray(desc1, np.float32).reshape(-1, 256)
        A_all = np.concatenate([a] * n) if a.shape[0] else a
        cnt_a = np.full(n, a.shape[0], np.int32)
        cnt_b = np.array([nb.shape[0] for nb in neighbours], np.int32)
        B_all = np.concatenate([np.ascontiguousarray(nb, np.float32).reshape(-1, 256) for nb in neighbours])
        a_off = (np.arange(n) * a.shape[0]).astype(np.int32)
        b_off = np.concatenate([[0], np.cumsum(cnt_b)[:-1]]).astype(np.int32)
        idx, val = self.ctx.match_batch(1, A_all, B_all, a_off, cnt_a, b_off, cnt_b, COS_FLOOR)
        out = []
        for p in range(n):
            sl = slice(a_off[p], a_off[p] + cnt_a[p])
            i = np.flatnonzero(idx[sl] >= 0)
            out.append((i.astype(np.int32), idx[sl][i], val[sl][i]))
        return out


    def search_by_projection(self, mp_desc: np.ndarray, proj_uv: np.ndarray, radius: np.ndarray, min_level: np.ndarray,
                             max_level: np.ndarray, frame_desc: np.ndarray, frame_xy: np.ndarray,
                             frame_octave: np.ndarray, occupied=None, ratio: float = 0.8, th_high: float = TH_HIGH):
        """Matcher::SearchByProjection(F, vpMapPoints, ratio, th) (src/Matcher.cc:40-210), descriptor + bookkeeping part:
        map point i (descriptor, projection (u, v), window radius, octave range) is matched to the nearest unclaimed
        frame feature in its window if best <= TH_HIGH and not (same octave as the second best and best > ratio *
        second).  Map points are processed in order and claim their feature, exactly like the reference's loop; the
        distances come from the GPU (top-4 candidates per map point), the claiming runs here.
        Returns match[i] = feature index or -1."""
        idx, dist, lvl = self.ctx.match_projection(mp_desc, proj_uv, radius, min_level, max_level, frame_desc, frame_xy,
                                                   frame_octave, occupied)
        taken = set()
        out = np.full(idx.shape[0], -1, np.int32)
        fmax = np.finfo(np.float32).max
        for i in range(idx.shape[0]):
            cand = [(dist[i, k], lvl[i, k], idx[i, k]) for k in range(idx.shape[1]) if idx[i, k] >= 0]
            free = [c for c in cand if int(c[2]) not in taken]
            if len(free) < 2 and len(cand) == idx.shape[1]:
                # the device list may be truncated (features claimed by earlier map points took its slots): the
                # reference would go on to the 5th, 6th ... nearest feature, so this window is re-scanned exactly
                free = self._rescan_window(np.asarray(mp_desc, np.float32)[i], proj_uv[i], radius[i], min_level[i],
                                           max_level[i], frame_desc, frame_xy, frame_octave, occupied, taken)
            if not free:
                continue
            bd, bl, bi = free[0]
            sd, sl = (free[1][0], free[1][1]) if len(free) > 1 else (np.float32(fmax), -1)
            if bd <= np.float32(th_high):
                if bl == sl and bd > np.float32(ratio) * sd:
                    continue
                out[i] = bi
                taken.add(int(bi))
        return out

    @staticmethod
    def _rescan_window(q, uv, r, min_level, max_level, frame_desc, frame_xy, frame_octave, occupied, taken):
        """Exact host scan of one search window (Frame::GetFeaturesInArea + Matcher::DescriptorDistance): the unclaimed
        in-window features as (distance, octave, index), nearest first (ties: lower index)."""
        fd = np.asarray(frame_desc, np.float32)
        xy = np.asarray(frame_xy, np.float32)
        octv = np.asarray(frame_octave, np.int32)
        ok = (np.abs(xy[:, 0] - np.float32(uv[0])) < np.float32(r)) & (np.abs(xy[:, 1] - np.float32(uv[1])) < np.float32(r))
        ok &= octv >= int(min_level)
        if int(max_level) >= 0:
            ok &= octv <= int(max_level)
        if occupied is not None:
            ok &= ~np.asarray(occupied, bool)
        ii = [int(j) for j in np.flatnonzero(ok) if int(j) not in taken]
        if not ii:
            return []
        dd = np.sqrt(((fd[ii] - np.asarray(q, np.float32)) ** 2).sum(1, dtype=np.float32)).astype(np.float32)
        order = np.lexsort((np.asarray(ii), dd))
        return [(dd[o], int(octv[ii[o]]), ii[o]) for o in order]

    def compute_distinctive_descriptors(self, observations):
        """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:331-400) for a batch of map points: ``observations``
        is a list of [N_i][256] arrays (the descriptors of each point's observations).  Returns (index, median): the row
        of each point with the least median L2 distance to the rest (-1 for a point without observations)."""
        sizes = [int(np.asarray(o).shape[0]) for o in observations]
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        rows = [np.asarray(o, np.float32).reshape(-1, 256) for o in observations if np.asarray(o).shape[0]]
        desc = np.concatenate(rows) if rows else np.zeros((0, 256), np.float32)
        return self.ctx.distinctive_descriptors(desc, off)

    def search_for_initialization(self, d1, xy1, oct1, d2, xy2, oct2, prev_matched, nnratio: float = 0.9,
                                  window: float = 100.0):
        """Matcher::SearchForInitialization (src/Matcher.cc:486-559): the distances come from the windowed kernel (top-4
        level-0 candidates of every level-0 keypoint of frame 1 around its previously matched position), the sequential
        claiming / take-over bookkeeping of the reference's loop is replayed here.  A keypoint whose four nearest
        candidates are all already claimed by better matches is re-scanned exactly over its whole window (rare).
        Returns (matches12, n_matches, updated prev_matched)."""
        d1, d2 = np.asarray(d1, np.float32), np.asarray(d2, np.float32)
        xy1, xy2 = np.asarray(xy1, np.float32), np.asarray(xy2, np.float32)
        oct1, oct2 = np.asarray(oct1, np.int32), np.asarray(oct2, np.int32)
        prev = np.asarray(prev_matched, np.float32)
        n1, n2 = d1.shape[0], d2.shape[0]
        fmax = np.finfo(np.float32).max
        m12 = np.full(n1, -1, np.int32)
        m21 = np.full(n2, -1, np.int32)
        matched = np.full(n2, fmax, np.float32)
        q = np.flatnonzero(oct1 == 0)
        n = 0
        if len(q) and n2:
            zeros = np.zeros(len(q), np.int32)
            idx, dist, _ = self.ctx.match_projection(d1[q], prev[q], np.full(len(q), window, np.float32), zeros, zeros,
                                                     d2, xy2, oct2)
            r = np.float32(window)
            for row, i1 in enumerate(q):
                cand = [(dist[row, k], int(idx[row, k])) for k in range(idx.shape[1]) if idx[row, k] >= 0]
                free = [(dd, i2) for dd, i2 in cand if not matched[i2] <= dd]
                if len(free) < 2 and len(cand) == idx.shape[1]:
                    # the list may be truncated: exact scan of the whole window for this keypoint
                    u, v = prev[i1]
                    ok = (np.abs(xy2[:, 0] - u) < r) & (np.abs(xy2[:, 1] - v) < r) & (oct2 == 0)
                    ii = np.flatnonzero(ok)
                    dd = np.sqrt(((d2[ii] - d1[i1]) ** 2).sum(1, dtype=np.float32)).astype(np.float32)
                    free = sorted((float(x), int(i)) for x, i in zip(dd, ii) if not matched[i] <= x)
                if not free:
                    continue
                best, bidx = free[0]
                best2 = free[1][0] if len(free) > 1 else fmax
                if best <= np.float32(TH_LOW) and best < np.float32(best2) * np.float32(nnratio):
                    if m21[bidx] >= 0:
                        m12[m21[bidx]] = -1
                        n -= 1
                    m12[i1] = bidx
                    m21[bidx] = i1
                    matched[bidx] = best
                    n += 1
        pm = prev.copy()
        hit = m12 >= 0
        pm[hit] = xy2[m12[hit]]
        return m12, n, pm

    def search_by_projection_last_frame(self, Tcw, K, bounds, scale_factors, last_points_w, last_valid, last_octave,
                                        last_desc, cur_desc, cur_xy, cur_octave, cur_occupied, th: float,
                                        th_high: float = TH_HIGH):
        """Matcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono=true) (src/Matcher.cc:1574-1650): projection of
        the last frame's map points with the current pose estimate on the host (a few thousand 3x4 products), windowed
        distances on the device (top-4 per map point over octaves [oct-1, oct+1]), the reference's sequential claiming
        replayed here (exact re-scan when all four candidates are already taken).
        Returns (assigned: last-frame map-point index per current feature or -1, n_matches)."""
        Tcw = np.asarray(Tcw, np.float32)
        fx, fy, cx, cy = [np.float32(v) for v in K]
        mnx, mxx, mny, mxy = [np.float32(v) for v in bounds]
        P = np.asarray(last_points_w, np.float32)
        xc = (P @ Tcw[:, :3].T + Tcw[:, 3]).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            invz = np.float32(1.0) / xc[:, 2]
            u = fx * xc[:, 0] / xc[:, 2] + cx
            v = fy * xc[:, 1] / xc[:, 2] + cy
        lo = np.asarray(last_octave, np.int32)
        ok = np.asarray(last_valid, bool) & ~(invz < 0) & ~((u < mnx) | (u > mxx) | (v < mny) | (v > mxy))
        q = np.flatnonzero(ok)
        cur_desc = np.asarray(cur_desc, np.float32)
        cur_xy = np.asarray(cur_xy, np.float32)
        cur_octave = np.asarray(cur_octave, np.int32)
        occ = np.array(cur_occupied, bool, copy=True)
        assigned = np.full(cur_desc.shape[0], -1, np.int32)
        n = 0
        if len(q) == 0 or cur_desc.shape[0] == 0:
            return assigned, n
        rad = (np.float32(th) * np.asarray(scale_factors, np.float32)[lo[q]]).astype(np.float32)
        uv = np.stack([u[q], v[q]], 1).astype(np.float32)
        ld = np.asarray(last_desc, np.float32)
        idx, dist, _ = self.ctx.match_projection(ld[q], uv, rad, lo[q] - 1, lo[q] + 1, cur_desc, cur_xy, cur_octave,
                                                 occ.astype(np.uint8))
        for row, i in enumerate(q):
            cand = [(dist[row, k], int(idx[row, k])) for k in range(idx.shape[1]) if idx[row, k] >= 0]
            free = [(d, i2) for d, i2 in cand if not occ[i2]]
            if not free and len(cand) == idx.shape[1]:   # list possibly truncated: exact scan of this window
                r = rad[row]
                w = (np.abs(cur_xy[:, 0] - uv[row, 0]) < r) & (np.abs(cur_xy[:, 1] - uv[row, 1]) < r) & \
                    (cur_octave >= lo[i] - 1) & (cur_octave <= lo[i] + 1) & ~occ
                ii = np.flatnonzero(w)
                dd = np.sqrt(((cur_desc[ii] - ld[i]) ** 2).sum(1, dtype=np.float32)).astype(np.float32)
                free = sorted((float(x), int(j)) for x, j in zip(dd, ii))
            if free and free[0][0] <= np.float32(th_high):
                b = free[0][1]
                assigned[b] = i
                occ[b] = True
                n += 1
        return assigned, n



    def fuse(self, Tcw, Ow, K, bounds, scale_factors, log_scale_factor, mp_pos, mp_normal, mp_min_dist, mp_max_dist,
             mp_desc, mp_skip, kf_desc, kf_xy, kf_octave, th: float = 3.0, th_low: float = TH_LOW, chi2: float = 5.99):
        """Matcher::Fuse(pKF, vpMapPoints, th) (src/Matcher.cc:1046-1250), monocular, up to the map bookkeeping: geometry
        (projection, distance range, viewing angle, PredictScale) on the host, the gated window search on the device (every
        map point is independent here, so the best in-window candidate is the answer).  ``mp_min_dist`` / ``mp_max_dist``
        are the raw mfMinDistance / mfMaxDistance: the range gate uses the invariance distances min / 1.2f and 1.2f * max
        (MapPoint::GetMin/MaxDistanceInvariance, src/MapPoint.cc:504-516), PredictScale the raw maximum (:518-534).
        Returns (best_idx, best_dist): the keyframe feature each map point would be fused into, or -1."""
        Tcw = np.asarray(Tcw, np.float32)
        fx, fy, cx, cy = [np.float32(v) for v in K]
        mnx, mxx, mny, mxy = [np.float32(v) for v in bounds]
        P = np.asarray(mp_pos, np.float32)
        sf = np.asarray(scale_factors, np.float32)
        pc = (P @ Tcw[:, :3].T + Tcw[:, 3]).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            u = fx * pc[:, 0] / pc[:, 2] + cx
            v = fy * pc[:, 1] / pc[:, 2] + cy
            PO = (P - np.asarray(Ow, np.float32)).astype(np.float32)
            d3 = np.sqrt(np.sum(PO * PO, axis=1, dtype=np.float32)).astype(np.float32)
            ratio = (np.asarray(mp_max_dist, np.float32) / d3).astype(np.float32)
            lvl = np.ceil(np.log(ratio) / np.float32(log_scale_factor))
        if len(sf) <= 1:
            lvl = np.zeros_like(d3)
            min_inv, max_inv = np.zeros_like(d3), np.full_like(d3, 10000.0)
        else:
            min_inv = (np.asarray(mp_min_dist, np.float32) / np.float32(1.2)).astype(np.float32)
            max_inv = (np.float32(1.2) * np.asarray(mp_max_dist, np.float32)).astype(np.float32)
        ok = ~np.asarray(mp_skip, bool) & ~(pc[:, 2] < 0) & (u >= mnx) & (u < mxx) & (v >= mny) & (v < mxy)
        ok &= ~((d3 < min_inv) | (d3 > max_inv))
        ok &= ~(np.sum(PO * np.asarray(mp_normal, np.float32), axis=1, dtype=np.float32) < np.float32(0.5) * d3)
        M = P.shape[0]
        best_idx = np.full(M, -1, np.int32)
        best_dist = np.full(M, np.finfo(np.float32).max, np.float32)
        q = np.flatnonzero(ok)
        if len(q) == 0 or np.asarray(kf_desc).shape[0] == 0:
            return best_idx, best_dist
        lv = np.clip(np.nan_to_num(lvl[q], nan=0.0, posinf=len(sf) - 1, neginf=0.0), 0, len(sf) - 1).astype(np.int32)
        rad = (np.float32(th) * sf[lv]).astype(np.float32)
        ko = np.asarray(kf_octave, np.int32)
        inv_sigma2 = (np.float32(1.0) / (sf * sf)).astype(np.float32)[ko]
        idx, dist, _ = self.ctx.match_projection(np.asarray(mp_desc, np.float32)[q], np.stack([u[q], v[q]], 1), rad,
                                                 lv - 1, lv, kf_desc, kf_xy, ko, None, inv_sigma2, chi2)
        hit = (idx[:, 0] >= 0) & (dist[:, 0] <= np.float32(th_low))
        best_dist[q] = np.where(idx[:, 0] >= 0, dist[:, 0], best_dist[q])
        best_idx[q[hit]] = idx[hit, 0]
        return best_idx, best_dist

    # ------------------------------------------------------------------------------------------------ m4, remaining variants
    @staticmethod
    def _predict_levels(max_dist, dist, log_scale_factor, n_levels):
        """MapPoint::PredictScale (src/MapPoint.cc:518-552), vectorised in fp32; 0 for a single-level pyramid."""
        if n_levels <= 1:
            return np.zeros(len(dist), np.int32)
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = (np.asarray(max_dist, np.float32) / dist).astype(np.float32)
            lvl = np.ceil(np.log(ratio) / np.float32(log_scale_factor))
        return np.clip(np.nan_to_num(lvl, nan=0.0, posinf=n_levels - 1, neginf=0.0), 0, n_levels - 1).astype(np.int32)

    @staticmethod
    def _invariance(min_dist, max_dist, n_levels):
        """MapPoint::GetMin/MaxDistanceInvariance (src/MapPoint.cc:504-516)."""
        if n_levels <= 1:
            return np.zeros(len(min_dist), np.float32), np.full(len(min_dist), 10000.0, np.float32)
        return ((np.asarray(min_dist, np.float32) / np.float32(1.2)).astype(np.float32),
                (np.float32(1.2) * np.asarray(max_dist, np.float32)).astype(np.float32))

    def _claim_best(self, q, mp_desc, uv, rad, mn, mx, feat_desc, feat_xy, feat_oct, occupied, threshold):
        """Sequential best-1 claiming shared by the variants below: queries ``q`` in order, each takes the nearest free
        in-window feature if its distance <= threshold (device top-4 lists; exact re-scan when a list is exhausted)."""
        occ = np.array(occupied, bool, copy=True)
        owner = np.full(feat_desc.shape[0], -1, np.int32)
        n = 0
        if len(q) == 0 or feat_desc.shape[0] == 0:
            return owner, n
        idx, dist, _ = self.ctx.match_projection(mp_desc[q], uv, rad, mn, mx, feat_desc, feat_xy, feat_oct, occ.astype(np.uint8))
        for row, i in enumerate(q):
            cand = [(dist[row, k], int(idx[row, k])) for k in range(idx.shape[1]) if idx[row, k] >= 0]
            free = [(d, j) for d, j in cand if not occ[j]]
            if not free and len(cand) == idx.shape[1]:
                free = [(d, j) for d, _, j in self._rescan_window(mp_desc[i], uv[row], rad[row], mn[row], mx[row], feat_desc,
                                                                   feat_xy, feat_oct, occ, set())]
            if free and free[0][0] <= np.float32(threshold):
                j = free[0][1]
                owner[j] = i
                occ[j] = True
                n += 1
        return owner, n

    def search_by_projection_keyframe(self, Tcw, K, bounds, scale_factors, log_scale_factor, mp_pos, mp_min_dist, mp_max_dist,
                                      mp_desc, mp_skip, cur_desc, cur_xy, cur_octave, cur_occupied, th: float, threshold: float):
        """Matcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, threshold) (src/Matcher.cc:1723-1805,
        relocalisation): geometry on the host, window search over octaves [pred-1, pred+1] on the device, sequential
        claiming here.  Returns (assigned: map-point index per frame feature or -1, n_matches)."""
        Tcw = np.asarray(Tcw, np.float32)
        fx, fy, cx, cy = [np.float32(v) for v in K]
        mnx, mxx, mny, mxy = [np.float32(v) for v in bounds]
        P = np.asarray(mp_pos, np.float32)
        sf = np.asarray(scale_factors, np.float32)
        R, t = Tcw[:, :3], Tcw[:, 3]
        Ow = (-(R.T @ t)).astype(np.float32)
        pc = (P @ R.T + t).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            u = fx * pc[:, 0] / pc[:, 2] + cx
            v = fy * pc[:, 1] / pc[:, 2] + cy
        PO = (P - Ow).astype(np.float32)
        d3 = np.sqrt(np.sum(PO * PO, axis=1, dtype=np.float32)).astype(np.float32)
        mn_inv, mx_inv = self._invariance(mp_min_dist, mp_max_dist, len(sf))
        ok = ~np.asarray(mp_skip, bool) & np.isfinite(u) & np.isfinite(v) & ~((u < mnx) | (u > mxx) | (v < mny) | (v > mxy))
        ok &= ~((d3 < mn_inv) | (d3 > mx_inv))
        q = np.flatnonzero(ok)
        lv = self._predict_levels(np.asarray(mp_max_dist, np.float32)[q], d3[q], log_scale_factor, len(sf))
        rad = (np.float32(th) * sf[lv]).astype(np.float32)
        uv = np.stack([u[q], v[q]], 1).astype(np.float32)
        return self._claim_best(q, np.asarray(mp_desc, np.float32), uv, rad, lv - 1, lv + 1, np.asarray(cur_desc, np.float32),
                                np.asarray(cur_xy, np.float32), np.asarray(cur_octave, np.int32), cur_occupied, threshold)

    def search_by_projection_sim3(self, Tcw, Ow, K, bounds, scale_factors, log_scale_factor, mp_pos, mp_normal, mp_min_dist,
                                  mp_max_dist, mp_desc, mp_skip, kf_desc, kf_xy, kf_octave, kf_matched, th: float,
                                  threshold: float):
        """Matcher::SearchByProjection(pKF, Scw, vpPoints[, vpPointsKFs], vpMatched[, vpMatchedKF], th, threshold)
        (src/Matcher.cc:265-367, :369-484; loop detection / place recognition): Tcw / Ow derived from the Sim3 by the
        caller as at :275-276.  Returns (matched: point index per keyframe feature or -1 for the NEW matches, n_matches);
        the second overload's vpMatchedKF is the caller's lookup of the point's source keyframe."""
        Tcw = np.asarray(Tcw, np.float32)
        fx, fy, cx, cy = [np.float32(v) for v in K]
        mnx, mxx, mny, mxy = [np.float32(v) for v in bounds]
        P = np.asarray(mp_pos, np.float32)
        sf = np.asarray(scale_factors, np.float32)
        pc = (P @ Tcw[:, :3].T + Tcw[:, 3]).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            u = fx * pc[:, 0] / pc[:, 2] + cx
            v = fy * pc[:, 1] / pc[:, 2] + cy
        PO = (P - np.asarray(Ow, np.float32)).astype(np.float32)
        d3 = np.sqrt(np.sum(PO * PO, axis=1, dtype=np.float32)).astype(np.float32)
        mn_inv, mx_inv = self._invariance(mp_min_dist, mp_max_dist, len(sf))
        ok = ~np.asarray(mp_skip, bool) & ~(pc[:, 2] < 0) & (u >= mnx) & (u < mxx) & (v >= mny) & (v < mxy)
        ok &= ~((d3 < mn_inv) | (d3 > mx_inv))
        ok &= ~(np.sum(PO * np.asarray(mp_normal, np.float32), axis=1, dtype=np.float32) < np.float32(0.5) * d3)
        q = np.flatnonzero(ok)
        lv = self._predict_levels(np.asarray(mp_max_dist, np.float32)[q], d3[q], log_scale_factor, len(sf))
        rad = (np.float32(th) * sf[lv]).astype(np.float32)
        uv = np.stack([u[q], v[q]], 1).astype(np.float32)
        return self._claim_best(q, np.asarray(mp_desc, np.float32), uv, rad, lv - 1, lv, np.asarray(kf_desc, np.float32),
                                np.asarray(kf_xy, np.float32), np.asarray(kf_octave, np.int32), kf_matched, threshold)

    def _sim3_directed(self, Tsrc_w, S_dst_src, K, bounds, sf, log_scale_factor, mp_pos, mp_min, mp_max, mp_desc, mp_valid,
                       already, dst_desc, dst_xy, dst_oct, th):
        """One direction of SearchBySim3 (src/Matcher.cc:1393-1466): independent queries -> the device's best in-window
        candidate is the answer."""
        fx, fy, cx, cy = [np.float32(v) for v in K]
        mnx, mxx, mny, mxy = [np.float32(v) for v in bounds]
        Ts = np.asarray(Tsrc_w, np.float32)
        s, Rd, td = np.float32(S_dst_src[0]), np.asarray(S_dst_src[1], np.float32), np.asarray(S_dst_src[2], np.float32)
        P = np.asarray(mp_pos, np.float32)
        p_src = (P @ Ts[:, :3].T + Ts[:, 3]).astype(np.float32)
        p = (s * (p_src @ Rd.T) + td).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            invz = (np.float32(1.0) / p[:, 2]).astype(np.float32)
            u = fx * (p[:, 0] * invz) + cx
            v = fy * (p[:, 1] * invz) + cy
        d3 = np.sqrt(np.sum(p * p, axis=1, dtype=np.float32)).astype(np.float32)
        mn_inv, mx_inv = self._invariance(mp_min, mp_max, len(sf))
        ok = np.asarray(mp_valid, bool) & ~np.asarray(already, bool) & ~(p[:, 2] < 0) & (u >= mnx) & (u < mxx) & (v >= mny) & (v < mxy)
        ok &= ~((d3 < mn_inv) | (d3 > mx_inv))
        out = np.full(P.shape[0], -1, np.int32)
        q = np.flatnonzero(ok)
        if len(q) == 0 or np.asarray(dst_desc).shape[0] == 0:
            return out
        lv = self._predict_levels(np.asarray(mp_max, np.float32)[q], d3[q], log_scale_factor, len(sf))
        rad = (np.float32(th) * sf[lv]).astype(np.float32)
        idx, dist, _ = self.ctx.match_projection(np.asarray(mp_desc, np.float32)[q], np.stack([u[q], v[q]], 1), rad, lv - 1, lv,
                                                 dst_desc, dst_xy, dst_oct)
        hit = (idx[:, 0] >= 0) & (dist[:, 0] <= np.float32(TH_HIGH))
        out[q[hit]] = idx[hit, 0]
        return out

    def search_by_sim3(self, K, bounds, scale_factors, log_scale_factor, T1w, T2w, S12, S21, mp1_pos, mp1_min, mp1_max, mp1_desc,
                       mp1_valid, already1, mp2_pos, mp2_min, mp2_max, mp2_desc, mp2_valid, already2, desc1, xy1, oct1, desc2,
                       xy2, oct2, th: float):
        """Matcher::SearchBySim3(pKF1, pKF2, vpMatches12, S12, th) (src/Matcher.cc:1355-1572): both directed searches on
        the device, the agreement check (:1553-1567) here.  S12 / S21 = (scale, R, t).  Returns (match12, n)."""
        sf = np.asarray(scale_factors, np.float32)
        d1, d2 = np.asarray(desc1, np.float32), np.asarray(desc2, np.float32)
        x1, x2 = np.asarray(xy1, np.float32), np.asarray(xy2, np.float32)
        o1, o2 = np.asarray(oct1, np.int32), np.asarray(oct2, np.int32)
        m1 = self._sim3_directed(T1w, S21, K, bounds, sf, log_scale_factor, mp1_pos, mp1_min, mp1_max, mp1_desc, mp1_valid,
                                 already1, d2, x2, o2, th)
        m2 = self._sim3_directed(T2w, S12, K, bounds, sf, log_scale_factor, mp2_pos, mp2_min, mp2_max, mp2_desc, mp2_valid,
                                 already2, d1, x1, o1, th)
        out = np.full(len(m1), -1, np.int32)
        i1 = np.flatnonzero(m1 >= 0)
        agree = i1[m2[m1[i1]] == i1]
        out[agree] = m1[agree]
        return out, int(len(agree))
