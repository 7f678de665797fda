"""Windowed matching: hfb_match_projection + Matcher.search_by_projection against a restatement of
Matcher::SearchByProjection(F, MPs) (src/Matcher.cc:40-210) and Frame::GetFeaturesInArea (src/Frame.cc:659-725).
Indices exact, distances within 2e-6."""
import numpy as np
import pytest

from hfnet_slam_b200.matcher import Matcher
from oracle import match_ref

pytestmark = pytest.mark.gpu


def _scene(nq, nf, seed, W=752, H=480):
    rng = np.random.default_rng(seed)
    F = rng.normal(size=(nf, 256)).astype(np.float32)
    F /= np.linalg.norm(F, axis=1, keepdims=True)
    fxy = np.stack([rng.uniform(0, W, nf), rng.uniform(0, H, nf)], 1).astype(np.float32)
    flev = rng.integers(0, 4, nf).astype(np.int32)
    src = rng.integers(0, nf, nq)
    Q = F[src] + 0.04 * rng.normal(size=(nq, 256)).astype(np.float32)
    Q[nq // 2:] = rng.normal(size=(nq - nq // 2, 256))            # half the map points have no true partner
    Q = (Q / np.linalg.norm(Q, axis=1, keepdims=True)).astype(np.float32)
    uv = (fxy[src] + rng.normal(0, 3.0, (nq, 2))).astype(np.float32)
    pred = flev[src].copy()
    rad = (rng.uniform(4.0, 40.0, nq) * (1.2 ** pred)).astype(np.float32)
    return Q, uv, rad, (pred - 1).astype(np.int32), pred.astype(np.int32), F, fxy, flev


def _window(u, v, r, mn, mx, fxy, flev, skip=None):
    """Frame::GetFeaturesInArea predicate (the grid cells only pre-filter it), ascending feature index."""
    ok = (np.abs(fxy[:, 0] - u) < r) & (np.abs(fxy[:, 1] - v) < r) & (flev >= mn)
    if mx >= 0:
        ok &= flev <= mx
    if skip is not None:
        ok &= ~skip
    return np.flatnonzero(ok)


@pytest.mark.parametrize("nq,nf,seed", [(900, 1200, 0), (1, 1, 1), (300, 4000, 2), (130, 129, 3)])
def test_top2_in_window_matches_oracle(small_ctx, nq, nf, seed):
    Q, uv, rad, mn, mx, F, fxy, flev = _scene(nq, nf, seed)
    idx, dist, lvl = small_ctx.match_projection(Q, uv, rad, mn, mx, F, fxy, flev)
    ptr, cand = [0], []
    for i in range(nq):
        c = _window(uv[i, 0], uv[i, 1], rad[i], mn[i], mx[i], fxy, flev)
        cand.extend(c.tolist())
        ptr.append(len(cand))
    bi, bd, bl, sd, sl = match_ref.best2_masked(Q, F, np.array(ptr), np.array(cand, np.int64), flev)
    assert np.array_equal(idx[:, 0], bi), f"best index differs for rows {np.flatnonzero(idx[:, 0] != bi)[:8]}"
    has = bi >= 0
    assert np.abs(dist[has, 0] - bd[has]).max(initial=0) <= 2e-6 and np.array_equal(lvl[has, 0], bl[has])
    has2 = sl >= 0
    assert np.abs(dist[has2, 1] - sd[has2]).max(initial=0) <= 2e-6 and np.array_equal(lvl[has2, 1], sl[has2])
    assert (idx[~has2, 1] == -1).all()
    # lists are sorted and duplicate-free
    d = np.where(idx >= 0, dist.astype(np.float64), 1e30)
    assert (np.diff(d, axis=1) >= 0).all()
    assert has.sum() > 0 or nq == 1


def test_search_by_projection_sequential_claims(small_ctx):
    Q, uv, rad, mn, mx, F, fxy, flev = _scene(700, 900, 5)
    Q[10] = Q[3]; uv[10] = uv[3]; rad[10] = rad[3]; mn[10] = mn[3]; mx[10] = mx[3]      # two map points want one feature
    occupied = np.zeros(900, bool); occupied[::17] = True                               # features that already have a map point
    got = Matcher(small_ctx).search_by_projection(Q, uv, rad, mn, mx, F, fxy, flev, occupied=occupied, ratio=0.8)
    # restatement of src/Matcher.cc:78-125 with the occupancy test of :84-86
    exp = match_ref.search_by_projection_map_points(Q, uv, rad, mn, mx, F, fxy, flev, occupied=occupied, ratio=0.8)
    assert np.array_equal(got, exp), f"rows {np.flatnonzero(got != exp)[:10]}"
    assert got[3] >= 0 and got[10] != got[3]
    assert (got >= 0).sum() > 100


def test_projection_empty_and_no_window(small_ctx):
    Q, uv, rad, mn, mx, F, fxy, flev = _scene(50, 60, 7)
    idx, dist, lvl = small_ctx.match_projection(Q, uv, np.zeros(50, np.float32), mn, mx, F, fxy, flev)
    assert (idx == -1).all()
    idx, dist, lvl = small_ctx.match_projection(Q[:0], uv[:0], rad[:0], mn[:0], mx[:0], F, fxy, flev)
    assert idx.shape == (0, 4)
    idx, dist, lvl = small_ctx.match_projection(Q, uv, rad, mn, mx, F[:0], fxy[:0], flev[:0])
    assert (idx == -1).all()
    # unbounded max level (-1) and min level 0 == no level check
    idx2, _, _ = small_ctx.match_projection(Q, uv, rad * 100, np.zeros(50, np.int32), -np.ones(50, np.int32), F, fxy, flev)
    D = np.sqrt(((Q[:, None, :].astype(np.float64) - F[None].astype(np.float64)) ** 2).sum(-1))
    assert np.array_equal(idx2[:, 0], D.argmin(1))


@pytest.mark.parametrize("seed,shift", [(0, 6.0), (1, 25.0)])
def test_search_for_initialization_matches_oracle(small_ctx, seed, shift):
    """Matcher::SearchForInitialization (src/Matcher.cc:486-559): two frames of the same scene, the second shifted and
    jittered; level-0 keypoints only, window 100, nnratio 0.9; matches12, count and updated vbPrevMatched identical."""
    rng = np.random.default_rng(seed)
    n1 = 700
    d1 = rng.normal(size=(n1, 256)).astype(np.float32)
    d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
    xy1 = np.stack([rng.uniform(0, 752, n1), rng.uniform(0, 480, n1)], 1).astype(np.float32)
    oct1 = (rng.uniform(size=n1) < 0.25).astype(np.int32) * rng.integers(1, 4, n1).astype(np.int32)
    keep = rng.uniform(size=n1) < 0.8                     # 80 % of the keypoints reappear in frame 2 ...
    d2 = d1[keep] + 0.02 * rng.normal(size=(keep.sum(), 256)).astype(np.float32)
    xy2 = xy1[keep] + np.float32(shift) + rng.normal(0, 1.5, (keep.sum(), 2)).astype(np.float32)
    oct2 = oct1[keep]
    extra = 250                                           # ... next to unrelated ones, some of them near-duplicates
    de = rng.normal(size=(extra, 256)).astype(np.float32)
    de[:60] = d2[:60] + 0.025 * rng.normal(size=(60, 256)).astype(np.float32)
    d2 = np.concatenate([d2, de]).astype(np.float32)
    d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    xe = np.stack([rng.uniform(0, 752, extra), rng.uniform(0, 480, extra)], 1).astype(np.float32)
    xe[:60] = xy2[:60] + rng.normal(0, 8.0, (60, 2)).astype(np.float32)
    xy2 = np.concatenate([xy2, xe]).astype(np.float32)
    oct2 = np.concatenate([oct2, np.zeros(extra, np.int32)])
    prev = xy1.copy()
    m = Matcher(small_ctx)
    got, n_got, pm_got = m.search_for_initialization(d1, xy1, oct1, d2, xy2, oct2, prev, 0.9, 100.0)
    ref, n_ref, pm_ref = match_ref.search_for_initialization(d1, xy1, oct1, d2, xy2, oct2, prev, 0.9, 100.0)
    assert n_ref > 200, "the scene should produce a few hundred initial matches"
    assert np.array_equal(got, ref) and n_got == n_ref
    assert np.array_equal(pm_got, pm_ref)
    assert (got[oct1 > 0] == -1).all()


@pytest.mark.parametrize("seed", [0, 1])
def test_search_by_projection_last_frame_matches_oracle(small_ctx, seed):
    """Matcher::SearchByProjection(CurrentFrame, LastFrame, th, mono) (src/Matcher.cc:1574-1650): synthetic scene, small
    inter-frame motion, EuRoC intrinsics; the assignment of last-frame map points to current features is identical."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = 458.654, 457.296, 367.215, 248.375
    W, H = 752, 480
    n_last = 600
    Pw = np.stack([rng.uniform(-4, 4, n_last), rng.uniform(-3, 3, n_last), rng.uniform(3, 12, n_last)], 1).astype(np.float32)
    Pw[:15, 2] = -2.0                                              # behind the camera
    ang = 0.02
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]], np.float32)
    Tcw = np.concatenate([R, np.array([[0.05], [-0.02], [0.1]], np.float32)], 1)
    last_valid = rng.uniform(size=n_last) < 0.9
    last_oct = rng.integers(0, 4, n_last).astype(np.int32)
    ld = rng.normal(size=(n_last, 256)).astype(np.float32)
    ld /= np.linalg.norm(ld, axis=1, keepdims=True)
    xc = Pw @ Tcw[:, :3].T + Tcw[:, 3]
    with np.errstate(divide="ignore", invalid="ignore"):
        uv = np.stack([fx * xc[:, 0] / xc[:, 2] + cx, fy * xc[:, 1] / xc[:, 2] + cy], 1).astype(np.float32)
    vis = rng.uniform(size=n_last) < 0.85
    cd = ld[vis] + 0.02 * rng.normal(size=(vis.sum(), 256)).astype(np.float32)
    cxy = uv[vis] + rng.normal(0, 2.0, (vis.sum(), 2)).astype(np.float32)
    coct = np.clip(last_oct[vis] + rng.integers(-1, 2, vis.sum()), 0, 3).astype(np.int32)
    extra = 300
    cd = np.concatenate([cd, rng.normal(size=(extra, 256))]).astype(np.float32)
    cd /= np.linalg.norm(cd, axis=1, keepdims=True)
    cxy = np.concatenate([cxy, np.stack([rng.uniform(0, W, extra), rng.uniform(0, H, extra)], 1)]).astype(np.float32)
    coct = np.concatenate([coct, rng.integers(0, 4, extra)]).astype(np.int32)
    occupied = rng.uniform(size=cd.shape[0]) < 0.05
    scale = (1.2 ** np.arange(4)).astype(np.float32)
    args = (Tcw, (fx, fy, cx, cy), (0.0, float(W), 0.0, float(H)), scale, Pw, last_valid, last_oct, ld, cd, cxy, coct,
            occupied, 15.0)
    got, n_got = Matcher(small_ctx).search_by_projection_last_frame(*args)
    ref, n_ref = match_ref.search_by_projection_last_frame(*args)
    assert n_ref > 250, "most visible map points should be re-found"
    assert n_got == n_ref and np.array_equal(got, ref)


def _fuse_scene(seed):
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = 458.654, 457.296, 367.215, 248.375
    W, H = 752, 480
    M = 500
    Pw = np.stack([rng.uniform(-5, 5, M), rng.uniform(-4, 4, M), rng.uniform(-1, 12, M)], 1).astype(np.float32)
    ang = -0.03
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]], np.float32)
    t = np.array([0.1, 0.05, 0.2], np.float32)
    Tcw = np.concatenate([R, t[:, None]], 1)
    Ow = (-R.T @ t).astype(np.float32)
    scale = (1.2 ** np.arange(4)).astype(np.float32)
    PO = Pw - Ow
    d3 = np.linalg.norm(PO, axis=1).astype(np.float32)
    normal = (PO / d3[:, None] + 0.5 * rng.normal(size=(M, 3))).astype(np.float32)      # some fail the 60 degree test
    normal /= np.linalg.norm(normal, axis=1, keepdims=True)
    ref_level = rng.integers(0, 4, M)
    max_d = (d3 * rng.uniform(0.9, 1.2 ** 3, M)).astype(np.float32)                      # some are out of range
    min_d = (max_d / 1.2 ** 4 * rng.uniform(0.8, 1.3, M)).astype(np.float32)
    md = rng.normal(size=(M, 256)).astype(np.float32)
    md /= np.linalg.norm(md, axis=1, keepdims=True)
    pc = Pw @ R.T + t
    with np.errstate(divide="ignore", invalid="ignore"):
        uv = np.stack([fx * pc[:, 0] / pc[:, 2] + cx, fy * pc[:, 1] / pc[:, 2] + cy], 1).astype(np.float32)
    uv = np.nan_to_num(uv, nan=-1e4, posinf=1e4, neginf=-1e4)
    # keyframe features: noisy re-observations (some several pixels off: they meet the chi^2 gate) + clutter
    kd = np.concatenate([md + 0.02 * rng.normal(size=md.shape), rng.normal(size=(400, 256))]).astype(np.float32)
    kd /= np.linalg.norm(kd, axis=1, keepdims=True)
    kxy = np.concatenate([uv + rng.normal(0, 1.6, uv.shape), np.stack([rng.uniform(0, W, 400), rng.uniform(0, H, 400)], 1)])
    koct = rng.integers(0, 4, len(kd)).astype(np.int32)
    skip = rng.uniform(size=M) < 0.1
    return dict(Tcw=Tcw, Ow=Ow, K=(fx, fy, cx, cy), bounds=(0.0, float(W), 0.0, float(H)), scale_factors=scale,
                log_scale_factor=float(np.log(1.2)), mp_pos=Pw, mp_normal=normal, mp_min_dist=min_d, mp_max_dist=max_d,
                mp_desc=md, mp_skip=skip, kf_desc=kd, kf_xy=kxy.astype(np.float32), kf_octave=koct)


@pytest.mark.parametrize("seed", [0, 1])
def test_fuse_matches_oracle(small_ctx, seed):
    """Matcher::Fuse(pKF, vpMapPoints, th = 3) (src/Matcher.cc:1046-1250) up to the map bookkeeping: the keyframe feature
    every map point would be fused into is identical; distances within 2e-6."""
    sc = _fuse_scene(seed)
    got_i, got_d = Matcher(small_ctx).fuse(**sc)
    ref_i, ref_d = match_ref.fuse(**sc)
    assert (ref_i >= 0).sum() > 40, "the scene should fuse a fair number of points"
    assert np.array_equal(got_i, ref_i)
    hit = ref_i >= 0
    assert np.abs(got_d[hit] - ref_d[hit]).max() <= 2e-6


# ---------------------------------------------------------------------------------------------------------------------
# the remaining members of the projection family
def _kf_args(seed):
    sc = _fuse_scene(seed)
    rng = np.random.default_rng(100 + seed)
    occ = rng.uniform(size=sc["kf_desc"].shape[0]) < 0.1
    return sc, occ


@pytest.mark.parametrize("seed", [0, 1])
def test_search_by_projection_keyframe_matches_oracle(small_ctx, seed):
    """Matcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, threshold) (src/Matcher.cc:1723-1805)."""
    sc, occ = _kf_args(seed)
    args = (sc["Tcw"], sc["K"], sc["bounds"], sc["scale_factors"], sc["log_scale_factor"], sc["mp_pos"], sc["mp_min_dist"],
            sc["mp_max_dist"], sc["mp_desc"], sc["mp_skip"], sc["kf_desc"], sc["kf_xy"], sc["kf_octave"], occ, 10.0, 0.75)
    got, n_got = Matcher(small_ctx).search_by_projection_keyframe(*args)
    ref, n_ref = match_ref.search_by_projection_keyframe(*args)
    assert n_ref > 100 and n_got == n_ref and np.array_equal(got, ref)


@pytest.mark.parametrize("seed", [0, 1])
def test_search_by_projection_sim3_matches_oracle(small_ctx, seed):
    """Matcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, threshold) and its vpMatchedKF overload
    (src/Matcher.cc:265-484)."""
    sc, occ = _kf_args(seed)
    args = (sc["Tcw"], sc["Ow"], sc["K"], sc["bounds"], sc["scale_factors"], sc["log_scale_factor"], sc["mp_pos"], sc["mp_normal"],
            sc["mp_min_dist"], sc["mp_max_dist"], sc["mp_desc"], sc["mp_skip"], sc["kf_desc"], sc["kf_xy"], sc["kf_octave"], occ,
            8.0, 0.6)
    got, n_got = Matcher(small_ctx).search_by_projection_sim3(*args)
    ref, n_ref = match_ref.search_by_projection_sim3(*args)
    assert n_ref > 40 and n_got == n_ref and np.array_equal(got, ref)
    assert not (got[occ] >= 0).any(), "features that were already matched are never claimed"


def _sim3_scene(seed):
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = 458.654, 457.296, 367.215, 248.375
    W, H, M = 752, 480, 450

    def rot(ax, ang):
        c, s = np.cos(ang), np.sin(ang)
        return {0: np.array([[1, 0, 0], [0, c, -s], [0, s, c]]), 1: np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])}[ax].astype(np.float32)

    T1w = np.concatenate([rot(1, 0.02), np.array([[0.05], [0.0], [0.1]], np.float32)], 1)
    T2w = np.concatenate([rot(1, -0.04) @ rot(0, 0.01), np.array([[-0.2], [0.03], [0.05]], np.float32)], 1)
    R12 = (T1w[:, :3] @ T2w[:, :3].T).astype(np.float32)
    t12 = (T1w[:, 3] - R12 @ T2w[:, 3]).astype(np.float32)
    S12 = (np.float32(1.0), R12, t12)
    S21 = (np.float32(1.0), R12.T.copy(), (-(R12.T @ t12)).astype(np.float32))
    scale = (1.2 ** np.arange(4)).astype(np.float32)

    def kf(T):
        Pw = np.stack([rng.uniform(-5, 5, M), rng.uniform(-4, 4, M), rng.uniform(2, 12, M)], 1).astype(np.float32)
        pc = Pw @ T[:, :3].T + T[:, 3]
        d3 = np.linalg.norm(pc, axis=1).astype(np.float32)
        mx = (d3 * rng.uniform(0.9, 1.2 ** 3, M)).astype(np.float32)
        mn = (mx / 1.2 ** 4 * rng.uniform(0.8, 1.3, M)).astype(np.float32)
        desc = rng.normal(size=(M, 256)).astype(np.float32)
        desc /= np.linalg.norm(desc, axis=1, keepdims=True)
        uv = np.stack([fx * pc[:, 0] / pc[:, 2] + cx, fy * pc[:, 1] / pc[:, 2] + cy], 1).astype(np.float32)
        return Pw, mn, mx, desc, uv

    P1, mn1, mx1, md1, uv1 = kf(T1w)
    # keyframe 2 re-observes the first 300 points of keyframe 1 (as its own map points, at its own feature slots) + clutter
    P2, mn2, mx2, md2, uv2 = kf(T2w)
    perm = rng.permutation(M)[:300]
    P2[:300], mn2[:300], mx2[:300] = P1[perm], mn1[perm], mx1[perm]
    md2[:300] = md1[perm] + 0.02 * rng.normal(size=(300, 256)).astype(np.float32)
    md2 /= np.linalg.norm(md2, axis=1, keepdims=True)
    pc2 = P2 @ T2w[:, :3].T + T2w[:, 3]
    uv2 = np.stack([fx * pc2[:, 0] / pc2[:, 2] + cx, fy * pc2[:, 1] / pc2[:, 2] + cy], 1).astype(np.float32)
    xy1 = (uv1 + rng.normal(0, 1.0, uv1.shape)).astype(np.float32)
    xy2 = (uv2 + rng.normal(0, 1.0, uv2.shape)).astype(np.float32)
    d1 = (md1 + 0.02 * rng.normal(size=md1.shape)).astype(np.float32)
    d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
    d2 = (md2 + 0.02 * rng.normal(size=md2.shape)).astype(np.float32)
    d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    o1 = rng.integers(0, 4, M).astype(np.int32)
    o2 = rng.integers(0, 4, M).astype(np.int32)

    def pred(mx, pc):     # MapPoint::PredictScale of the shared points in the OTHER camera: their features sit at that octave
        lv = np.ceil(np.log(mx / np.linalg.norm(pc, axis=1)) / np.log(1.2))
        return np.clip(lv, 0, 3).astype(np.int32)

    o2[:300] = pred(mx1[perm], P1[perm] @ T2w[:, :3].T + T2w[:, 3]) - (rng.uniform(size=300) < 0.3)
    o1[perm] = pred(mx2[:300], P2[:300] @ T1w[:, :3].T + T1w[:, 3]) - (rng.uniform(size=300) < 0.3)
    o1, o2 = np.clip(o1, 0, 3).astype(np.int32), np.clip(o2, 0, 3).astype(np.int32)
    v1, v2 = rng.uniform(size=M) > 0.1, rng.uniform(size=M) > 0.1
    a1, a2 = rng.uniform(size=M) < 0.05, rng.uniform(size=M) < 0.05
    return ((fx, fy, cx, cy), (0.0, float(W), 0.0, float(H)), scale, float(np.log(1.2)), T1w, T2w, S12, S21, P1, mn1, mx1, md1, v1, a1,
            P2, mn2, mx2, md2, v2, a2, d1, xy1, o1, d2, xy2, o2, 7.5)


@pytest.mark.parametrize("seed", [3, 4])
def test_search_by_sim3_matches_oracle(small_ctx, seed):
    """Matcher::SearchBySim3 (src/Matcher.cc:1355-1572): two directed windowed searches + agreement."""
    args = _sim3_scene(seed)
    got, n_got = Matcher(small_ctx).search_by_sim3(*args)
    ref, n_ref = match_ref.search_by_sim3(*args)
    assert n_ref > 40 and n_got == n_ref and np.array_equal(got, ref)
