"""Device-timed matcher kernel chain (prep + tcgen05 contraction/arg-max + finalize) on resident descriptors:
B pairs of n x n 256-d rows through hfb_match_batch_dev, CUDA events on the library's stream.
  python tools/match_time.py [pairs ...]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from hfnet_slam_b200 import synthetic
from hfnet_slam_b200.lib import Context

pairs = [int(a) for a in sys.argv[1:]] or [1, 8, 30, 256]
n = 1000
dev = torch.device("cuda", 0)
ctx = Context(height=64, width=64, n_levels=1, max_keypoints=8192, max_batch=1, with_global=False)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
A, B = synthetic.descriptor_pair(n, n, n_true=300, seed=0)
for P in pairs:
    dA = torch.from_numpy(np.concatenate([A] * P)).to(dev)
    dB = torch.from_numpy(np.concatenate([B] * P)).to(dev)
    off = (np.arange(P) * n).astype(np.int32)
    cnt = np.full(P, n, np.int32)
    tab = torch.from_numpy(np.concatenate([off, cnt, off, cnt])).to(dev)
    idx = torch.empty(P * n, dtype=torch.int32, device=dev)
    val = torch.empty(P * n, dtype=torch.float32, device=dev)

    def run():
        ctx.check(ctx.lib.hfb_match_batch_dev(ctx.handle, 0, dA.data_ptr(), P * n, dB.data_ptr(), P * n, P, tab.data_ptr(),
                                              n, n, 0.6, idx.data_ptr(), val.data_ptr()))
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    e0.record(stream)
    for _ in range(reps):
        run()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nm = int((idx.view(P, n)[0] >= 0).sum())
    fl = 2.0 * P * n * n * 256
    print(f"pairs {P:4d}: {1e3 * ms:8.1f} us per call, {1e3 * ms / P:7.2f} us per pair, {fl / ms / 1e9:7.1f} TFLOP/s algorithmic "
          f"(K = 256), matches in pair 0: {nm}")
ctx.close()
