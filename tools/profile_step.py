"""Minimal driver for ncu: a few steps of the bench workload (extract batch + consecutive match) with plain launches."""
import os
import sys
from pathlib import Path

os.environ.setdefault("HFB_NO_GRAPH", "1")
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from bench import H, W, NKP, THR, synthetic_frames
from hfnet_slam_b200 import weights
from hfnet_slam_b200.lib import Context

B = int(os.environ.get("BATCH", "8"))
steps = int(os.environ.get("STEPS", "3"))
ctx = Context(height=H, width=W, n_levels=1, max_keypoints=NKP, max_batch=B, with_global=True)
ctx.load_weights(weights.synthetic_blob(seed=0))
d = torch.from_numpy(np.stack(synthetic_frames(B, 0))).cuda()
for _ in range(steps):
    ctx.extract_batch_dev(d.data_ptr(), B, [NKP], THR)
    ctx.match_consecutive_dev(B, 0, 0.6)
    ctx.sync()
print("launches", ctx.launch_count)
ctx.close()
