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
