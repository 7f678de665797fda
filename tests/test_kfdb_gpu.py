"""Keyframe-database scan through the C-ABI against oracle/kfdb_ref.py: scores within 2e-6 (fp32 vs fp64 norm),
identical candidate sets except rows within 5e-6 of the relative threshold (summation-order noise, written here),
identical final loop/merge candidates."""
import numpy as np
import pytest

from hfnet_slam_b200 import synthetic
from hfnet_slam_b200.keyframe_database import KeyFrameDatabase
from oracle import kfdb_ref

pytestmark = pytest.mark.gpu
TOL = 2e-6


def _covis(n):
    return lambda kf, k: [j for j in range(kf - 5, kf + 6) if j != kf and 0 <= j < n][:k]


@pytest.mark.parametrize("n,seed", [(3000, 2), (257, 3), (1, 4)])
def test_query_matches_oracle(small_ctx, n, seed):
    db, q, qi = synthetic.keyframe_db(n, 4096, n_planted=min(50, n // 2), seed=seed, n_queries=1)
    ids = np.arange(n, dtype=np.int64) * 3 + 7
    kf = KeyFrameDatabase(small_ctx, capacity=n + 8)
    kf.add_many(ids, db)
    cand, scores, best = kf.query(q[0], rel=0.8, floor=0.0)
    sc_ref = kfdb_ref.scores(q[0], db)
    sel, best_ref = kfdb_ref.candidate_set(sc_ref, 0.8)
    assert abs(best - best_ref) <= TOL
    thr = 0.8 * best_ref
    sure = {int(ids[i]) for i in sel if sc_ref[i] > thr + 5e-6}
    maybe = {int(ids[i]) for i in np.flatnonzero(np.abs(sc_ref - thr) <= 5e-6)}
    got = set(int(c) for c in cand)
    assert sure <= got <= (sure | maybe), f"{len(got)} vs {len(sure)}"
    assert list(cand) == sorted(cand)
    all_sc = kf.scores_of(ids)
    assert np.abs(all_sc - sc_ref).max() <= TOL
    assert kf.scores_of(np.array([-5], np.int64))[0] == -1


def test_detect_n_best_candidates(small_ctx):
    n = 2000
    db, q, qi = synthetic.keyframe_db(n, 4096, n_planted=100, seed=5, n_queries=3)
    ids = np.arange(n, dtype=np.int64)
    map_of = {int(i): (0 if i < 1500 else 1) for i in ids}
    kf = KeyFrameDatabase(small_ctx, capacity=n)
    kf.add_many(ids, db)
    for k in range(3):
        loop, merge = kf.detect_n_best_candidates(q[k], query_map=0, map_of=map_of, covisibles=_covis(n), n_candidates=3)
        l_ref, m_ref, _ = kfdb_ref.detect_n_best_candidates(q[k], ids, db, map_of, 0, _covis(n), 3)
        assert set(loop) == set(l_ref) and set(merge) == set(m_ref)
        reloc = kf.detect_relocalization_candidates(q[k], query_map=0, map_of=map_of, covisibles=_covis(n))
        r_ref = kfdb_ref.detect_relocalization_candidates(q[k], ids, db, map_of, 0, _covis(n))
        assert set(reloc) == set(r_ref)


def test_add_erase_clear(small_ctx):
    db, q, _ = synthetic.keyframe_db(64, 4096, n_planted=8, seed=6)
    ids = np.arange(64, dtype=np.int64)
    kf = KeyFrameDatabase(small_ctx, capacity=64)
    kf.add_many(ids[:40], db[:40])
    for i in range(40, 64):
        kf.add(int(ids[i]), db[i])
    assert len(kf) == 64
    with pytest.raises(Exception):
        kf.add(0, db[0])                      # duplicate id
    with pytest.raises(Exception):
        kf.add(1000, db[0])                   # capacity
    kf.erase(5); kf.erase(63); kf.erase(12345)
    assert len(kf) == 62
    keep = np.array([i for i in range(64) if i not in (5, 63)])
    cand, scores, best = kf.query(q[0])
    sc_ref = kfdb_ref.scores(q[0], db[keep])
    assert np.abs(kf.scores_of(ids[keep]) - sc_ref).max() <= TOL
    assert kf.scores_of(np.array([5], np.int64))[0] == -1
    kf.clear()
    assert len(kf) == 0
    cand, scores, best = kf.query(q[0])
    assert len(cand) == 0 and best == 0.0


def test_shard_records_merge_to_global_set(small_ctx):
    """Row-shard by id % world, one fixed-size record per shard, merged == unsharded candidate set."""
    from hfnet_slam_b200.keyframe_database import merge_shard_records
    n, world = 4000, 4
    db, q, _ = synthetic.keyframe_db(n, 4096, n_planted=200, seed=8)
    ids = np.arange(n, dtype=np.int64)
    full = KeyFrameDatabase(small_ctx, capacity=n)
    full.add_many(ids, db)
    cand, scores, best = full.query(q[0])
    recs = []
    for r in range(world):
        sh = KeyFrameDatabase(small_ctx, capacity=n // world + 1)
        m = ids % world == r
        sh.add_many(ids[m], db[m])
        recs.append(sh.query_shard(q[0], k=64))
    m_ids, m_scores, m_best, overflow = merge_shard_records(recs, rel=0.8, floor=0.0)
    assert not overflow and m_best == best
    assert list(m_ids) == list(cand)
    assert np.array_equal(m_scores, scores)


@pytest.fixture(scope="module")
def c4_db():
    """BASELINE.json configs[3] / SURVEY.md 8(d)-C4: 50 000 x 4096 fp32 unit rows, 200 planted near-duplicates, queries =
    planted rows + noise."""
    return synthetic.keyframe_db(50000, 4096, n_planted=200, seed=2, n_queries=64)


def _check_candidates(cand, best, sc_ref, ids):
    sel, best_ref = kfdb_ref.candidate_set(sc_ref, 0.8)
    assert abs(best - best_ref) <= TOL
    thr = 0.8 * best_ref
    sure = {int(ids[i]) for i in sel if sc_ref[i] > thr + 5e-6}
    maybe = {int(ids[i]) for i in np.flatnonzero(np.abs(sc_ref - thr) <= 5e-6)}
    got = set(int(c) for c in cand)
    assert sure <= got <= (sure | maybe), f"{len(got)} vs {len(sure)}"


def test_c4_50k_rows_single_query(small_ctx, c4_db):
    db, q, qi = c4_db
    n = db.shape[0]
    ids = np.arange(n, dtype=np.int64)
    kf = KeyFrameDatabase(small_ctx, capacity=n)
    kf.add_many(ids, db)
    for k in (0, 17):
        cand, scores, best = kf.query(q[k])
        sc_ref = kfdb_ref.scores(q[k], db)
        _check_candidates(cand, best, sc_ref, ids)
        assert int(qi[k]) in set(int(c) for c in cand), "the planted source row is a candidate"
        assert np.abs(kf.scores_of(ids[::97]) - sc_ref[::97]).max() <= TOL
    kf.close()


def test_c4_50k_rows_eight_way_shard_merge(small_ctx, c4_db):
    """Row-shard by id % 8, one fixed-size record per shard, merged == the unsharded candidate set at the full C4 size."""
    from hfnet_slam_b200.keyframe_database import merge_shard_records
    db, q, _ = c4_db
    n, world = db.shape[0], 8
    ids = np.arange(n, dtype=np.int64)
    shards = []
    for r in range(world):
        sh = KeyFrameDatabase(small_ctx, capacity=n // world + 1)
        m = ids % world == r
        sh.add_many(ids[m], db[m])
        shards.append(sh)
    for k in (0, 33):
        sc_ref = kfdb_ref.scores(q[k], db)
        recs = [sh.query_shard(q[k], k=64) for sh in shards]
        m_ids, m_scores, m_best, overflow = merge_shard_records(recs, rel=0.8, floor=0.0)
        assert not overflow
        _check_candidates(m_ids, m_best, sc_ref, ids)
        assert np.abs(m_scores - sc_ref[m_ids]).max() <= TOL
    for sh in shards:
        sh.close()


def test_c4_50k_rows_q64_batch_equals_single_queries(small_ctx, c4_db):
    """BASELINE.json configs[3] with Q = 64: one tensor-core pass + exact re-scoring == 64 single queries (ids, scores and
    best scores bit-identical: both paths score a pair with the same arithmetic) == the oracle's candidate sets."""
    db, q, qi = c4_db
    n = db.shape[0]
    ids = np.arange(n, dtype=np.int64)
    kf = KeyFrameDatabase(small_ctx, capacity=n)
    kf.add_many(ids, db)
    out = kf.query_batch(q, cap=256)
    assert len(out) == 64
    for k in (0, 1, 31, 63):
        cand, scores, best = kf.query(q[k])
        assert np.array_equal(out[k][0], cand) and np.array_equal(out[k][1], scores) and out[k][2] == best
    for k in (5, 40):
        _check_candidates(out[k][0], out[k][2], kfdb_ref.scores(q[k], db), ids)
    # a random (unplanted) query: best score 0 -> no candidates, like the single-query path
    rq = np.random.default_rng(1).standard_normal((3, 4096)).astype(np.float32)
    rq /= np.linalg.norm(rq, axis=1, keepdims=True)
    rq[2] = db[123]                                  # an exact duplicate: distance 0, score 1
    for k, (c, s, b) in enumerate(kf.query_batch(rq)):
        c1, s1, b1 = kf.query(rq[k])
        assert np.array_equal(c, c1) and np.array_equal(s, s1) and b == b1
    assert kf.query_batch(rq)[2][2] == 1.0
    kf.close()


@pytest.mark.parametrize("n,nq", [(1, 2), (130, 3), (2000, 70)])
def test_query_batch_small_and_ragged(small_ctx, n, nq):
    db, q, _ = synthetic.keyframe_db(n, 4096, n_planted=min(40, n // 2), seed=9, n_queries=min(nq, max(n // 2, 1)))
    rng = np.random.default_rng(2)
    extra = rng.standard_normal((nq - len(q), 4096)).astype(np.float32) if nq > len(q) else np.zeros((0, 4096), np.float32)
    if len(extra):
        extra /= np.linalg.norm(extra, axis=1, keepdims=True)
    Q = np.concatenate([q, extra])[:nq]
    ids = np.arange(n, dtype=np.int64) * 2 + 1
    kf = KeyFrameDatabase(small_ctx, capacity=n + 3)
    kf.add_many(ids, db)
    for k, (c, s, b) in enumerate(kf.query_batch(Q, cap=64)):
        c1, s1, b1 = kf.query(Q[k])
        assert np.array_equal(c, c1) and np.array_equal(s, s1) and b == b1, k
    kf.close()


def test_clear_map(small_ctx):
    """KeyFrameDatabase::clearMap (src/KeyFrameDatabase.cc:54-68): every keyframe of one map leaves the database."""
    db, q, _ = synthetic.keyframe_db(90, 4096, n_planted=10, seed=6)
    ids = np.arange(90, dtype=np.int64) + 100
    maps = (np.arange(90) % 3).astype(np.int64)
    kf = KeyFrameDatabase(small_ctx, capacity=90)
    kf.add_tagged(ids, maps, db)
    kf.clear_map(1)
    assert len(kf) == 60
    keep = maps != 1
    cand, scores, best = kf.query(q[0])
    sc_ref = kfdb_ref.scores(q[0], db[keep])
    assert np.abs(kf.scores_of(ids[keep]) - sc_ref).max() <= TOL
    assert (kf.scores_of(ids[~keep]) == -1).all()
    kf.clear_map(7)
    assert len(kf) == 60
    kf.close()


def test_device_side_shard_exchange_equals_unsharded(native_lib, c4_db):
    """The host-free sharded query (device record + peer-memory exchange + device merge; here 8 shard objects with one
    context each in ONE process, inboxes connected by direct pointers) at the full C4 size: every shard returns the
    unsharded candidate set, scores and best."""
    from hfnet_slam_b200.lib import Context
    db, q, _ = c4_db
    n, world = db.shape[0], 8
    ids = np.arange(n, dtype=np.int64)
    ctxs = [Context(height=64, width=64, n_levels=1, max_keypoints=64, max_batch=1, with_global=False) for _ in range(world + 1)]
    full = KeyFrameDatabase(ctxs[world], capacity=n)
    full.add_many(ids, db)
    shards = []
    for r in range(world):
        sh = KeyFrameDatabase(ctxs[r], capacity=n // world + 1)
        m = ids % world == r
        sh.add_many(ids[m], db[m])
        shards.append(sh)
    KeyFrameDatabase.connect_shards_local(shards, k=64)
    for k in (0, 33, 7):
        cand, scores, best = full.query(q[k])
        for sh in shards:                      # all ranks enqueue, then all collect (they wait for one another on the device)
            sh.query_sharded_begin(q[k])
        for sh in shards:
            m_ids, m_sc, m_best, ov = sh.query_sharded_end()
            assert not ov and m_best == best
            assert np.array_equal(m_ids, cand) and np.array_equal(m_sc, scores)
    for sh in shards:
        sh.close()
    full.close()
    for c in ctxs:
        c.close()


_TWO_RANK_SCRIPT = r"""
import os, sys
sys.path.insert(0, os.environ["HFB_ROOT"])
import numpy as np, torch, torch.distributed as dist
from hfnet_slam_b200 import synthetic
from hfnet_slam_b200.keyframe_database import KeyFrameDatabase
from hfnet_slam_b200.lib import Context
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
dev = rank % torch.cuda.device_count()
db, q, _ = synthetic.keyframe_db(6000, 4096, n_planted=120, seed=8, n_queries=6)
ids = np.arange(6000, dtype=np.int64)
ctx = Context(height=64, width=64, n_levels=1, max_keypoints=64, max_batch=1, with_global=False, device=dev)
full = KeyFrameDatabase(ctx, capacity=6000); full.add_many(ids, db)
sh = KeyFrameDatabase(ctx, capacity=6000 // world + 1)
m = ids % world == rank
sh.add_many(ids[m], db[m])
sh.connect_shards(dist, rank, world, k=64)         # IPC handles exchanged once, peers' inboxes mapped
for k in range(6):
    cand, scores, best = full.query(q[k])
    m_ids, m_sc, m_best, ov = sh.query_sharded(q[k])
    assert not ov and m_best == best and np.array_equal(m_ids, cand) and np.array_equal(m_sc, scores), (rank, k)
dist.barrier()
sh.close(); full.close(); ctx.close()
print("RANK_OK", rank)
"""


def test_device_side_shard_exchange_two_processes_ipc(native_lib, tmp_path):
    """The multi-process form of the exchange: two ranks (torchrun, gloo rendezvous; both on the visible GPU(s)) map each
    other's inbox with CUDA IPC and answer six collective queries identically to an unsharded database."""
    import os, subprocess, sys
    from pathlib import Path
    script = tmp_path / "two_rank.py"
    script.write_text(_TWO_RANK_SCRIPT)
    env = dict(os.environ, HFB_ROOT=str(Path(__file__).resolve().parents[1]))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("RANK_OK") == 2


def test_repeated_queries_leave_no_state_behind(small_ctx):
    """A query costs no memset: the scan finds best / count zeroed by the PREVIOUS query's last block.  Alternating
    queries with a high best (planted near-duplicate), a low best (random direction), long candidate lists (> 64: the
    second fetch), erasures and insertions in between must each equal the oracle on the current rows."""
    n = 1500
    db, q, qi = synthetic.keyframe_db(n, 4096, n_planted=40, seed=21, n_queries=2)
    rng = np.random.default_rng(3)
    rand_q = rng.standard_normal(4096).astype(np.float32)
    rand_q /= np.linalg.norm(rand_q)
    for k, sc in enumerate(np.linspace(0.1, 0.9, 300)):          # 300 rows at graded distances from query 0
        v = q[0] + np.float32(sc / 64.0) * rng.standard_normal(4096).astype(np.float32)
        db[100 + k] = v / np.linalg.norm(v)
    ids = np.arange(n, dtype=np.int64) + 11
    kf = KeyFrameDatabase(small_ctx, capacity=n + 4)
    kf.add_many(ids, db)
    live = np.ones(n, bool)

    def check(query, rel, floor):
        cand, scores, best = kf.query(query, rel=rel, floor=floor)
        sc_ref = kfdb_ref.scores(query, db[live])
        lids = ids[live]
        best_ref = float(sc_ref.max())
        assert abs(best - best_ref) <= TOL, (best, best_ref)
        thr = max(floor, rel * best_ref)
        sure = {int(i) for i in lids[sc_ref > thr + 5e-6]}
        maybe = {int(i) for i in lids[np.abs(sc_ref - thr) <= 5e-6]}
        got = {int(c) for c in cand}
        assert sure <= got <= (sure | maybe), (len(got), len(sure))
        return len(got), best

    n_hi, b_hi = check(q[0], 0.8, 0.0)
    n_lo, b_lo = check(rand_q, 0.8, 0.0)            # a much lower best right after a high one
    assert b_lo < 0.5 * b_hi
    n_long, _ = check(q[0], 0.05, 0.0)              # hundreds of candidates: longer than the head of the list
    assert n_long > 64
    for victim in (int(ids[qi[0]]), int(ids[100]), int(ids[3]), int(ids[n - 1])):
        kf.erase(victim)
        live[victim - 11] = False
    n_hi2, b_hi2 = check(q[0], 0.8, 0.0)
    assert b_hi2 <= b_hi
    check(q[1], 0.8, 0.5)
    check(rand_q, 0.8, 0.0)
    check(q[0], 0.3, 0.0)
    kf.close()
