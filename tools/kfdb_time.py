"""Device-timed keyframe-database scans on a resident 50 k x 4096 database: Q = 1 (exact HBM-streaming kernel) and the
Q = 64 tensor-core pass (hfb_kfdb_query_batch_dev: scan + marking + exact re-scoring + selection).
  python tools/kfdb_time.py [rows]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from hfnet_slam_b200.keyframe_database import KeyFrameDatabase
from hfnet_slam_b200.lib import Context, _i64p, ptr

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
dev = torch.device("cuda", 0)
ctx = Context(height=64, width=64, n_levels=1, max_keypoints=64, max_batch=1, with_global=False)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
rows = torch.randn(n, 4096, device=dev)
rows /= rows.norm(dim=1, keepdim=True)
kf = KeyFrameDatabase(ctx, capacity=n)
ids = np.arange(n, dtype=np.int64)
ctx.check(ctx.lib.hfb_kfdb_add_dev(kf.handle, ptr(ids, _i64p), rows.data_ptr(), n))


def ev(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


dq = rows[5:6].clone()
dsc = torch.empty(n, device=dev)
db_ = torch.empty(1, device=dev)
t1 = ev(lambda: ctx.check(ctx.lib.hfb_kfdb_scan_dev(kf.handle, dq.data_ptr(), 1, dsc.data_ptr(), db_.data_ptr())), 20)
print(f"Q=1 : {1e3 * t1:8.1f} us  {n * 4096 * 4 / t1 / 1e6:7.1f} GB/s")
for Q in (64, 128):
    q = rows[:Q] + 0.002 * torch.randn(Q, 4096, device=dev)
    q /= q.norm(dim=1, keepdim=True)
    t = ev(lambda: ctx.check(ctx.lib.hfb_kfdb_query_batch_dev(kf.handle, q.data_ptr(), Q, 0.8, 0.0)), 10)
    print(f"Q={Q}: {1e3 * t:8.1f} us  = {t / t1:5.2f} single-query passes, {Q / t * 1e3:9.0f} queries/s, "
          f"{2.0 * Q * n * 4096 / t / 1e9:6.1f} TFLOP/s (tf32), rows streamed at {(Q + 63) // 64 * n * 4096 * 4 / t / 1e6:7.1f} GB/s")
kf.close()
ctx.close()
