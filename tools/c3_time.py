"""C3 (BASELINE.json configs[2]) stage timings alone + an A/B of the concurrent pyramid levels (HFB_FORK_LEVELS)."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

import bench
from hfnet_slam_b200 import weights
from hfnet_slam_b200.lib import Context


def extract_once(env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        with Context(height=bench.H, width=bench.W, n_levels=4, scale_factor=1.2, max_keypoints=675, max_batch=2,
                     with_global=True) as ctx:
            ctx.load_weights(weights.synthetic_blob(seed=0))
            imgs = [weights.synthetic_image(bench.H, bench.W, seed=s, n_corners=200) for s in (3, 4)]
            outs = []
            for rep in range(3):          # warm run, captured graph, replay
                outs.append(ctx.extract_batch(imgs, bench.C3_BUDGETS, 0.01))
            return outs
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


if os.environ.get("AB", "1") == "1":
    a = extract_once({"HFB_FORK_LEVELS": "0"})
    b = extract_once({"HFB_FORK_LEVELS": "1"})
    for rep in range(3):
        for i in range(2):
            for k in ("x", "y", "response", "octave", "descriptors", "global_descriptor"):
                assert np.array_equal(a[0][i][k], b[rep][i][k]), (rep, i, k)
            assert a[0][i]["n_per_level"] == b[rep][i]["n_per_level"]
    print("levels A/B identical:", a[0][0]["n_per_level"], a[0][1]["n_per_level"])
dev = torch.device("cuda:0")
for th in (2, 1):
    print(json.dumps(bench.c3_gpu(torch, dev, int(os.environ.get("FRAMES", "120")), th)))
