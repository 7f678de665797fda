"""Device-resident throughput with S contexts (S camera streams) per GPU enqueued round-robin from one host thread."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from bench import H, W, NKP, THR, synthetic_frames
from hfnet_slam_b200 import weights
from hfnet_slam_b200.lib import Context

B, NB = int(os.environ.get("BATCH", "8")), 48
blob = weights.synthetic_blob(seed=0)
d_ring = torch.from_numpy(np.stack(synthetic_frames(B * NB, 0)).reshape(NB, B, H, W)).cuda()
for S in [int(s) for s in os.environ.get("S", "1 2 3 4 6").split()]:
    ctxs = [Context(height=H, width=W, n_levels=1, max_keypoints=NKP, max_batch=B, with_global=True) for _ in range(S)]
    for c in ctxs:
        c.load_weights(blob)
    streams = [torch.cuda.ExternalStream(c.stream) for c in ctxs]

    def run(reps):
        for _ in range(reps):
            for i in range(NB):
                ctxs[i % S].extract_match_batch_dev(d_ring[i].data_ptr(), B, [NKP], THR, 0, 0.6)

    run(2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    run(6)
    for s in streams[1:]:
        ev = torch.cuda.Event()
        ev.record(s)
        streams[0].wait_event(ev)
    e1.record(streams[0])
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"contexts={S} batch={B}: {6 * NB * B / ms * 1e3:.0f} frames/s ({ms / (6 * NB):.4f} ms per call)")
    for c in ctxs:
        c.close()
