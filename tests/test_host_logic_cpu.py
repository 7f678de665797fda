"""Host-side logic that needs no GPU: weight blob format, shard-record merge (incl. a 2-rank gloo all-gather),
synthetic workload generators, golden fixtures against the oracle."""
import os
import struct
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from hfnet_slam_b200 import synthetic, weights
from hfnet_slam_b200.keyframe_database import merge_shard_records, parse_shard_record
from oracle import kfdb_ref

ROOT = Path(__file__).resolve().parents[1]


def test_weight_blob_roundtrip():
    wd = weights.synthetic(seed=3, n_clusters=16)
    blob = weights.pack(wd, n_clusters=16)
    back, nc, dm = weights.unpack(blob)
    assert nc == 16 and abs(dm - 0.75) < 1e-7
    assert set(back) == set(wd) and all(np.array_equal(back[k], wd[k]) for k in wd)
    assert blob[:8] == b"HFB2WTS1" and len(blob) == weights.HEADER_BYTES + 4 * sum(v.size for v in wd.values())
    with pytest.raises(ValueError):
        weights.unpack(b"XXXXXXXX" + blob[8:])


def test_fold_bn():
    rng = np.random.default_rng(0)
    w = rng.normal(size=(9, 4, 6))
    g, b, m, v = rng.uniform(0.5, 1.5, 6), rng.normal(size=6), rng.normal(size=6), rng.uniform(0.5, 1.5, 6)
    wf, bf = weights.fold_bn(w, g, b, m, v)
    x = rng.normal(size=(5, 9, 4))
    y = np.einsum("nkc,kco->no", x, w)
    ref = (y - m) * g / np.sqrt(v + 1e-3) + b
    assert np.allclose(np.einsum("nkc,kco->no", x, wf) + bf, ref, atol=1e-5)


def make_record(ids, scores, rel, floor, k):
    """CPU restatement of hfb_kfdb_query_shard (record layout of include/hfnet_b200.h) from oracle scores."""
    best = np.float32(scores.max()) if len(scores) else np.float32(0)
    thr = max(np.float32(floor), np.float32(best * np.float32(rel)))
    sel = np.flatnonzero(scores > thr)
    order = sorted(sel, key=lambda i: (-scores[i], ids[i]))
    count = min(len(order), k)
    rec = bytearray(16 + 16 * k)
    struct.pack_into("<fiii", rec, 0, float(best), count, int(len(order) > k), 0)
    for j in range(count):
        struct.pack_into("<fiq", rec, 16 + 16 * j, float(scores[order[j]]), 0, int(ids[order[j]]))
    return bytes(rec)


def test_shard_merge_equals_unsharded():
    db, q, _ = synthetic.keyframe_db(3000, 256, n_planted=300, noise=0.2, seed=1)
    ids = np.arange(3000, dtype=np.int64)
    sc = kfdb_ref.scores(q[0], db)
    sel, best = kfdb_ref.candidate_set(sc, 0.8)
    for world in (1, 2, 4, 8):
        recs = [make_record(ids[ids % world == r], sc[ids % world == r], 0.8, 0.0, 64) for r in range(world)]
        m_ids, m_sc, m_best, ov = merge_shard_records(recs)
        assert not ov and m_best == float(best)
        assert m_ids.tolist() == sel.tolist() and np.array_equal(m_sc, sc[sel])
    # overflow is flagged when a shard has more rows above the global bar than slots
    sc2 = np.linspace(0.5, 0.9, 10).astype(np.float32)            # all ten rows are above 0.5 * best
    recs = [make_record(np.arange(10, dtype=np.int64), sc2, 0.5, 0.0, 4)]
    assert merge_shard_records(recs, rel=0.5)[3] is True
    b, i, s, o = parse_shard_record(recs[0])
    assert len(i) == 4 and o and np.all(np.diff(s) <= 0) and i.tolist() == [9, 8, 7, 6]


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from hfnet_slam_b200 import synthetic
from hfnet_slam_b200.keyframe_database import merge_shard_records
from oracle import kfdb_ref
from tests.test_host_logic_cpu import make_record
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
db, q, _ = synthetic.keyframe_db(2000, 256, n_planted=200, noise=0.2, seed=7)
ids = np.arange(2000, dtype=np.int64)
mine = ids % world == rank                                  # row-shard by id % world (SURVEY.md 8e)
rec = make_record(ids[mine], kfdb_ref.scores(q[0], db[mine]), 0.8, 0.0, 64)
t = torch.frombuffer(bytearray(rec), dtype=torch.uint8)
out = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(out, t)                                     # the ONE collective of the sharded query
m_ids, m_sc, m_best, ov = merge_shard_records([bytes(o.numpy()) for o in out])
sc = kfdb_ref.scores(q[0], db)
sel, best = kfdb_ref.candidate_set(sc, 0.8)
assert not ov and m_ids.tolist() == sel.tolist() and m_best == float(best), (rank, len(m_ids), len(sel))
print("rank", rank, "ok", len(sel))
dist.destroy_process_group()
"""


def test_sharded_query_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=str(ROOT)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=240, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_synthetic_workloads_are_deterministic():
    A1, B1 = synthetic.descriptor_pair(seed=0)
    A2, B2 = synthetic.descriptor_pair(seed=0)
    assert np.array_equal(A1, A2) and np.array_equal(B1, B2)
    assert np.allclose(np.linalg.norm(A1, axis=1), 1, atol=1e-6)
    d = synthetic.lba_problem(n_opt=4, n_fixed=3, n_points=80, seed=1)
    assert np.all(np.diff(d["pt_idx"]) >= 0) and d["fixed"].sum() == 3
    assert np.bincount(d["pt_idx"]).min() >= 2


def test_bench_reference_arm_prints_contract_line():
    import json
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


class _FakeProjectionCtx:
    """numpy stand-in for Context.match_projection (top-k in-window candidates, ascending distance): lets the host-side
    claiming loops of hfnet_slam_b200/matcher.py run without a GPU.  ``k`` smaller than the device's 4 forces the
    'list exhausted -> exact re-scan' branch."""

    def __init__(self, k=4):
        self.k = k

    def match_projection(self, Q, q_uv, q_radius, q_min_level, q_max_level, F, f_xy, f_level, f_skip=None):
        Q, F = np.asarray(Q, np.float32), np.asarray(F, np.float32)
        nq = Q.shape[0]
        idx = np.full((nq, self.k), -1, np.int32)
        dist = np.full((nq, self.k), np.finfo(np.float32).max, np.float32)
        lvl = np.full((nq, self.k), -1, np.int32)
        for i in range(nq):
            ok = (np.abs(f_xy[:, 0] - q_uv[i, 0]) < q_radius[i]) & (np.abs(f_xy[:, 1] - q_uv[i, 1]) < q_radius[i]) & \
                 (f_level >= q_min_level[i])
            if q_max_level[i] >= 0:
                ok &= f_level <= q_max_level[i]
            if f_skip is not None:
                ok &= ~np.asarray(f_skip, bool)
            ii = np.flatnonzero(ok)
            d = np.sqrt(((F[ii] - Q[i]) ** 2).sum(1, dtype=np.float32)).astype(np.float32)
            o = np.argsort(d, kind="stable")[:self.k]
            idx[i, :len(o)], dist[i, :len(o)], lvl[i, :len(o)] = ii[o], d[o], f_level[ii[o]]
        return idx, dist, lvl


def _two_frames(seed):
    rng = np.random.default_rng(seed)
    n1 = 300
    d1 = rng.normal(size=(n1, 256)).astype(np.float32)
    d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
    xy1 = np.stack([rng.uniform(0, 752, n1), rng.uniform(0, 480, n1)], 1).astype(np.float32)
    oct1 = (rng.uniform(size=n1) < 0.25).astype(np.int32)
    # every keypoint reappears three times with decreasing similarity: claimed candidates and take-overs are frequent
    d2 = np.concatenate([d1 + s * rng.normal(size=d1.shape).astype(np.float32) for s in (0.015, 0.02, 0.03)])
    d2 = (d2 / np.linalg.norm(d2, axis=1, keepdims=True)).astype(np.float32)
    xy2 = np.concatenate([xy1 + rng.normal(0, 3.0, xy1.shape).astype(np.float32) for _ in range(3)]).astype(np.float32)
    oct2 = np.concatenate([oct1] * 3)
    return d1, xy1, oct1, d2, xy2, oct2


@pytest.mark.parametrize("k", [4, 1])
def test_search_for_initialization_replay_equals_oracle(k):
    """Host replay of Matcher::SearchForInitialization (src/Matcher.cc:486-559) over top-k candidate lists == the literal
    restatement, including the exact re-scan when a list is exhausted (k = 1 exercises it constantly)."""
    from hfnet_slam_b200.matcher import Matcher
    from oracle import match_ref
    d1, xy1, oct1, d2, xy2, oct2 = _two_frames(0)
    got, n_got, pm_got = Matcher(_FakeProjectionCtx(k)).search_for_initialization(d1, xy1, oct1, d2, xy2, oct2, xy1.copy(),
                                                                                  0.9, 100.0)
    ref, n_ref, pm_ref = match_ref.search_for_initialization(d1, xy1, oct1, d2, xy2, oct2, xy1.copy(), 0.9, 100.0)
    assert n_ref > 100
    assert n_got == n_ref and np.array_equal(got, ref) and np.array_equal(pm_got, pm_ref)


@pytest.mark.parametrize("k", [4, 1])
def test_search_by_projection_last_frame_replay_equals_oracle(k):
    """Host replay of Matcher::SearchByProjection(CurrentFrame, LastFrame) (src/Matcher.cc:1574-1650) == the restatement."""
    from hfnet_slam_b200.matcher import Matcher
    from oracle import match_ref
    rng = np.random.default_rng(3)
    fx, fy, cx, cy = 458.654, 457.296, 367.215, 248.375
    n_last = 250
    Pw = np.stack([rng.uniform(-4, 4, n_last), rng.uniform(-3, 3, n_last), rng.uniform(3, 12, n_last)], 1).astype(np.float32)
    Tcw = np.concatenate([np.eye(3, dtype=np.float32), np.array([[0.05], [-0.02], [0.1]], np.float32)], 1)
    last_valid = rng.uniform(size=n_last) < 0.9
    last_oct = rng.integers(0, 4, n_last).astype(np.int32)
    ld = rng.normal(size=(n_last, 256)).astype(np.float32)
    ld /= np.linalg.norm(ld, axis=1, keepdims=True)
    xc = Pw @ Tcw[:, :3].T + Tcw[:, 3]
    uv = np.stack([fx * xc[:, 0] / xc[:, 2] + cx, fy * xc[:, 1] / xc[:, 2] + cy], 1).astype(np.float32)
    cd = np.concatenate([ld + s * rng.normal(size=ld.shape).astype(np.float32) for s in (0.02, 0.03)])
    cd = (cd / np.linalg.norm(cd, axis=1, keepdims=True)).astype(np.float32)
    cxy = np.concatenate([uv + rng.normal(0, 2.0, uv.shape).astype(np.float32) for _ in range(2)]).astype(np.float32)
    coct = np.concatenate([last_oct, last_oct])
    occupied = rng.uniform(size=cd.shape[0]) < 0.1
    scale = (1.2 ** np.arange(4)).astype(np.float32)
    args = (Tcw, (fx, fy, cx, cy), (0.0, 752.0, 0.0, 480.0), scale, Pw, last_valid, last_oct, ld, cd, cxy, coct, occupied, 15.0)
    got, n_got = Matcher(_FakeProjectionCtx(k)).search_by_projection_last_frame(*args)
    ref, n_ref = match_ref.search_by_projection_last_frame(*args)
    assert n_ref > 100 and n_got == n_ref and np.array_equal(got, ref)


@pytest.mark.parametrize("k", [4, 2])
def test_search_by_projection_contended_window_rescans_exactly(k):
    """More map points than the device's top-k contend for ONE window: once the candidate list of a point is
    eaten by earlier claims the host mirror must re-scan the window exactly -- the reference simply walks on to the 5th,
    6th ... nearest feature (src/Matcher.cc:78-125) -- instead of dropping the point or passing the ratio test against
    FLT_MAX."""
    from hfnet_slam_b200.matcher import Matcher
    from oracle import match_ref
    rng = np.random.default_rng(11)
    nq, nf = 10, 14
    proto = rng.normal(size=256).astype(np.float32)
    Q = (proto + 0.05 * rng.normal(size=(nq, 256))).astype(np.float32)
    F = (proto + 0.05 * rng.normal(size=(nf, 256))).astype(np.float32)
    Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    F /= np.linalg.norm(F, axis=1, keepdims=True)
    uv = np.full((nq, 2), 100.0, np.float32)
    fxy = (100.0 + rng.uniform(-5, 5, (nf, 2))).astype(np.float32)
    rad = np.full(nq, 15.0, np.float32)
    flev = rng.integers(0, 2, nf).astype(np.int32)
    mn, mx = np.zeros(nq, np.int32), np.full(nq, -1, np.int32)
    occupied = np.zeros(nf, bool)
    occupied[3] = True
    for ratio, least in ((0.9, 4), (1.0, 9)):
        got = Matcher(_FakeProjectionCtx(k)).search_by_projection(Q, uv, rad, mn, mx, F, fxy, flev, occupied=occupied,
                                                                  ratio=ratio)
        ref = match_ref.search_by_projection_map_points(Q, uv, rad, mn, mx, F, fxy, flev, occupied=occupied, ratio=ratio)
        assert np.array_equal(got, ref), ratio
        assert (ref >= 0).sum() >= least, "the contending points should still find free features"
        assert len(set(ref[ref >= 0].tolist())) == (ref >= 0).sum()


def test_fuse_host_geometry_equals_oracle():
    """Host side of Matcher::Fuse (projection, distance range, viewing angle, PredictScale, gate parameters) over the numpy
    stand-in for the device search == the literal restatement (src/Matcher.cc:1046-1250)."""
    from hfnet_slam_b200.matcher import Matcher
    from oracle import match_ref
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    from test_projection_gpu import _fuse_scene

    class Ctx(_FakeProjectionCtx):
        def match_projection(self, Q, q_uv, q_radius, q_min_level, q_max_level, F, f_xy, f_level, f_skip=None,
                             f_inv_sigma2=None, chi2_max=0.0):
            Q, F = np.asarray(Q, np.float32), np.asarray(F, np.float32)
            nq = Q.shape[0]
            idx = np.full((nq, 4), -1, np.int32)
            dist = np.full((nq, 4), np.finfo(np.float32).max, np.float32)
            lvl = np.full((nq, 4), -1, np.int32)
            for i in range(nq):
                ex, ey = q_uv[i, 0] - f_xy[:, 0], q_uv[i, 1] - f_xy[:, 1]
                ok = (np.abs(ex) < q_radius[i]) & (np.abs(ey) < q_radius[i]) & (f_level >= q_min_level[i]) & \
                     (f_level <= q_max_level[i]) & ~((ex * ex + ey * ey).astype(np.float32) * f_inv_sigma2 > np.float32(chi2_max))
                ii = np.flatnonzero(ok)
                d = np.sqrt(((F[ii] - Q[i]) ** 2).sum(1, dtype=np.float32)).astype(np.float32)
                o = np.argsort(d, kind="stable")[:4]
                idx[i, :len(o)], dist[i, :len(o)], lvl[i, :len(o)] = ii[o], d[o], f_level[ii[o]]
            return idx, dist, lvl

    sc = _fuse_scene(0)
    got_i, _ = Matcher(Ctx()).fuse(**sc)
    ref_i, _ = match_ref.fuse(**sc)
    assert (ref_i >= 0).sum() > 40
    assert np.array_equal(got_i, ref_i)
    # single-level pyramid: PredictScale returns 0 and the invariance range is [0, 10000] (src/MapPoint.cc:504-534)
    one = dict(sc, scale_factors=sc["scale_factors"][:1], kf_octave=np.zeros_like(sc["kf_octave"]))
    g1, _ = Matcher(Ctx()).fuse(**one)
    r1, _ = match_ref.fuse(**one)
    assert np.array_equal(g1, r1) and (r1 >= 0).sum() > 40


def test_fuse_distance_gate_and_predicted_level_pins():
    """Pins of the oracle's Fuse geometry against hand-evaluated reference expressions: the range gate uses
    GetMin/MaxDistanceInvariance = mfMinDistance / 1.2f and 1.2f * mfMaxDistance, PredictScale the RAW mfMaxDistance
    (src/MapPoint.cc:504-534) -- with scaleFactor 1.2 the two choices differ by exactly one level."""
    from oracle import match_ref
    sf = (1.2 ** np.arange(4)).astype(np.float32)
    Tcw = np.concatenate([np.eye(3, dtype=np.float32), np.zeros((3, 1), np.float32)], 1)
    Ow = np.zeros(3, np.float32)
    K, bounds = (400.0, 400.0, 320.0, 240.0), (0.0, 640.0, 0.0, 480.0)
    desc = np.zeros((1, 256), np.float32); desc[0, 0] = 1

    def run(dist3d, min_d, max_d, kf_level):
        pos = np.array([[0.0, 0.0, dist3d]], np.float32)
        return match_ref.fuse(Tcw, Ow, K, bounds, sf, float(np.log(1.2)), pos, np.array([[0, 0, 1.0]], np.float32),
                              np.array([min_d], np.float32), np.array([max_d], np.float32), desc, np.zeros(1, bool), desc,
                              np.array([[320.0, 240.0]], np.float32), np.array([kf_level], np.int32))[0][0]

    # dist 11 > mfMaxDistance 10 but < 1.2 * 10: inside the invariance range; ratio < 1 -> level clamps to 0
    assert run(11.0, 4.0, 10.0, 0) == 0
    assert run(12.5, 4.0, 10.0, 0) == -1                   # beyond 1.2 * mfMaxDistance
    assert run(3.5, 4.0, 10.0, 3) == 0                     # >= mfMinDistance / 1.2 = 3.33; ratio 2.857 -> ceil(5.76) -> clamp 3
    assert run(3.2, 4.0, 10.0, 3) == -1                    # below mfMinDistance / 1.2
    # dist 9: ratio 10/9 -> ceil(log(1.111)/log(1.2)) = ceil(0.578) = 1 -> octaves [0, 1] pass, 2 does not
    assert run(9.0, 4.0, 10.0, 1) == 0 and run(9.0, 4.0, 10.0, 0) == 0 and run(9.0, 4.0, 10.0, 2) == -1
    # with the invariance maximum (12) in PredictScale the level would be ceil(log(12/9)/log 1.2) = 2: octave 2 would pass




def test_onnx_initialiser_import_round_trip(tmp_path):
    """hfnet_slam_b200/onnx_import.py: the seeded weights written as a minimal ONNX graph (Conv nodes with OIHW weights +
    biases in a SHUFFLED-between-branches but topologically valid order, centroids, MatMul weight) come back as exactly the
    blob tensors they started from; a wrong shape fails loudly."""
    from hfnet_slam_b200 import onnx_import as oi
    wd = weights.synthetic(seed=3)
    exp = oi.expected_convs()
    c1, blocks = weights.architecture(0.75)
    convs = []
    for base, shape, dw in exp:
        w = wd[base + ".w"]
        if base == "vlad.memberships":
            w = w.reshape(1 * shape[1], shape[0])               # [c_global][C] == HWIO-flattened 1x1
        convs.append((base, oi.to_oihw(w, shape[2], shape[1], dw), wd[base + ".b"]))
    # heads interleaved with the late backbone layers, as a topological sort of the forked graph may emit them
    head = [c for c in convs if c[0].startswith(("desc.", "det."))]
    rest = [c for c in convs if not c[0].startswith(("desc.", "det."))]
    k = next(i for i, c in enumerate(rest) if c[0] == "l9.dw")
    order = rest[:k] + head[:2] + rest[k:k + 5] + head[2:] + rest[k + 5:]
    extra = {"vlad/clusters": wd["vlad.clusters"].reshape(1, 1, 1, *wd["vlad.clusters"].shape),
             "dimensionality_reduction/weights": wd["fc.w"], "dimensionality_reduction/biases": wd["fc.b"]}
    path = tmp_path / "synthetic_hfnet.onnx"
    oi.write_model(path, order, extra)
    got = oi.convert(path)
    for name, _ in weights.tensor_specs():
        assert np.array_equal(got[name], wd[name]), name
    assert weights.pack(got) == weights.pack(wd)
    # a truncated graph (one Conv missing) is an error, not a silent mis-assignment
    oi.write_model(path, [c for c in order if c[0] != "l10.project"], extra)
    with pytest.raises(ValueError):
        oi.convert(path)


def test_remaining_projection_variants_replay_equals_oracle():
    """Host sides of SearchByProjection(F, KF), the Sim3 projection search and SearchBySim3 (src/Matcher.cc:1723-1805,
    :265-484, :1355-1572) over the numpy stand-in for the device candidates == the literal restatements; k = 2 forces the
    exhausted-list re-scan."""
    from hfnet_slam_b200.matcher import Matcher
    from oracle import match_ref
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    from test_projection_gpu import _kf_args, _sim3_scene
    for k in (4, 2):
        sc, occ = _kf_args(0)
        a = (sc["Tcw"], sc["K"], sc["bounds"], sc["scale_factors"], sc["log_scale_factor"], sc["mp_pos"], sc["mp_min_dist"],
             sc["mp_max_dist"], sc["mp_desc"], sc["mp_skip"], sc["kf_desc"], sc["kf_xy"], sc["kf_octave"], occ, 10.0, 0.75)
        got, n = Matcher(_FakeProjectionCtx(k)).search_by_projection_keyframe(*a)
        ref, nr = match_ref.search_by_projection_keyframe(*a)
        assert nr > 100 and n == nr and np.array_equal(got, ref)
        b = (sc["Tcw"], sc["Ow"], sc["K"], sc["bounds"], sc["scale_factors"], sc["log_scale_factor"], sc["mp_pos"], sc["mp_normal"],
             sc["mp_min_dist"], sc["mp_max_dist"], sc["mp_desc"], sc["mp_skip"], sc["kf_desc"], sc["kf_xy"], sc["kf_octave"], occ,
             8.0, 0.6)
        got, n = Matcher(_FakeProjectionCtx(k)).search_by_projection_sim3(*b)
        ref, nr = match_ref.search_by_projection_sim3(*b)
        assert nr > 40 and n == nr and np.array_equal(got, ref)
    c = _sim3_scene(3)
    got, n = Matcher(_FakeProjectionCtx(4)).search_by_sim3(*c)
    ref, nr = match_ref.search_by_sim3(*c)
    assert nr > 40 and n == nr and np.array_equal(got, ref)
