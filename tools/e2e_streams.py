"""End-to-end (host buffers) throughput with S camera streams per GPU: S contexts, S host threads, each thread calling
the synchronous hfb_extract_match_batch on its own stream of 8-frame batches (ctypes releases the GIL in the call)."""
import os
import sys
import threading
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from bench import H, W, NKP, THR, synthetic_frames
from hfnet_slam_b200 import weights
from hfnet_slam_b200.lib import Context, pinned_empty

B, NB = 8, int(os.environ.get("NB", "48"))
blob = weights.synthetic_blob(seed=0)
frames = synthetic_frames(B * NB, 0)
ring = pinned_empty((NB, B, H, W), np.uint8)
for i in range(NB):
    ring[i] = np.stack(frames[i * B:(i + 1) * B])
for S in (1, 2, 3, 4):
    ctxs = [Context(height=H, width=W, n_levels=1, max_keypoints=NKP, max_batch=B, with_global=True) for _ in range(S)]
    for c in ctxs:
        c.load_weights(blob)
    outs = [(pinned_empty((B, c.kp_cap), np.int32), pinned_empty((B, c.kp_cap), np.float32)) for c in ctxs]

    def work(s, reps):
        c, o = ctxs[s], outs[s]
        for _ in range(reps):
            for i in range(s, NB, S):
                c.extract_match_batch([ring[i, b] for b in range(B)], [NKP], THR, 0, 0.6, pinned=True, out=o)

    def run(reps):
        th = [threading.Thread(target=work, args=(s, reps)) for s in range(S)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0

    run(2)
    dt = run(6)
    print(f"streams={S}: {6 * NB * B / dt:.0f} frames/s  ({1e3 * dt / (6 * NB):.3f} ms per 8-frame call)")
    for c in ctxs:
        c.close()
