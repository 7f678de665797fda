"""Device-timed extract(+associate) step for a batch (graph replay, frames resident) + the per-launch profile."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from bench import H, W, NKP, THR, synthetic_frames
from hfnet_slam_b200 import weights
from hfnet_slam_b200.lib import Context

B = int(os.environ.get("BATCH", "8"))
reps = int(os.environ.get("REPS", "200"))
ctx = Context(height=H, width=W, n_levels=1, max_keypoints=NKP, max_batch=B, with_global=os.environ.get("GLOBAL", "1") == "1")
ctx.load_weights(weights.synthetic_blob(seed=0))
ring = [torch.from_numpy(np.stack(synthetic_frames(B, 100 * i))).cuda() for i in range(8)]
stream = torch.cuda.ExternalStream(ctx.stream)


def step(i):
    if os.environ.get("MATCH", "1") == "1":
        ctx.extract_match_batch_dev(ring[i % len(ring)].data_ptr(), B, [NKP], THR, 0, 0.6)
    else:
        ctx.extract_batch_dev(ring[i % len(ring)].data_ptr(), B, [NKP], THR)


for i in range(10):
    step(i)
ctx.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(stream):
    e0.record()
    for i in range(reps):
        step(i)
    e1.record()
ctx.sync()
ms = e0.elapsed_time(e1) / reps
print(f"STEP batch={B} ms_per_step={ms:.4f} frames_per_s={B / ms * 1e3:.0f}")
if os.environ.get("PROFILE", "1") == "1":
    ctx.profile_extract(B, [NKP], THR)
    prof = ctx.profile_extract(B, [NKP], THR)
    tot = 0.0
    for r in prof:
        tot += r["ms"]
        print(f"  {r['name']:<26s} {1e3 * r['ms']:7.1f} us")
    print(f"  sum {1e3 * tot:.1f} us")
ctx.close()
