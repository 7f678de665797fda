"""Local BA + pose optimisation timings alone (bench.py's extra.lba block)."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench
from hfnet_slam_b200.lib import Context

with Context(height=64, width=64, n_levels=1, max_keypoints=64, max_batch=1, with_global=False) as ctx:
    r = bench.extra_lba(ctx, cpu=False)
    r2 = bench.extra_lba(ctx, cpu=False)
print(json.dumps({k: r2[k] for k in ("ms_per_iter", "ms_total", "iterations", "trials", "gpu_launches", "pose_optimization_ms")}))
