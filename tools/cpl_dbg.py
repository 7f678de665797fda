"""Checks the raw expand accumulators the channel-per-lane kernel dumps (HFB_CPL_DBG=<layer>, CTA 0 / tile 0 / chunk 0)
against numpy on the device's own input activation."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np

L = int(os.environ["HFB_CPL_DBG"])
os.environ["HFB_NO_GRAPH"] = "1"
os.environ["HFB_CPL_LAYERS"] = str(1 << L)
from hfnet_slam_b200 import weights
from hfnet_slam_b200.lib import Context

H, W = 480, 752
wd = weights.synthetic(seed=0)
_, blocks = weights.architecture(0.75)
blk = blocks[L - 2]
ctx = Context(height=H, width=W, n_levels=1, max_keypoints=1000, max_batch=1, with_global=True)
ctx.load_weights(weights.pack(wd))
img = weights.synthetic_image(H, W, seed=10, n_corners=200)
os.makedirs("gpurun_out", exist_ok=True)
ctx.extract_batch([img], [1000], 0.01)
x = ctx.debug_tensor(f"layer_{L - 1}")[0]            # [Hi][Wi][Cin]
We, be = wd[f"l{L}.expand.w"].astype(np.float16).astype(np.float32), wd[f"l{L}.expand.b"]
S = blk.stride
TH = int(os.environ.get("TH", "4" if S == 2 else "8"))
TW = 16 if S == 1 else 8
IW, IH = (TW - 1) * S + 3, (TH - 1) * S + 3
NRO = TH // 4
NRI = (NRO - 1) * S + 3
Hi, Wi = x.shape[:2]
def pad_before(n, s):
    out = (n + s - 1) // s
    tot = max((out - 1) * s + 3 - n, 0)
    return tot // 2
pt, pl = pad_before(Hi, S), pad_before(Wi, S)
d = np.fromfile("gpurun_out/cpl_dbg.bin", dtype=np.float32).reshape(4, 128, 4, 18)[:, :, :, :] if NRI == 4 else \
    np.fromfile("gpurun_out/cpl_dbg.bin", dtype=np.float32)[:4 * 128 * NRI * 18].reshape(4, 128, NRI, 18)
exp = np.zeros_like(d)
for rg in range(4):
    for i in range(NRI):
        iy = -pt + rg * NRO * S + i
        for k in range(IW):
            ix = -pl + k
            if 0 <= iy < Hi and 0 <= ix < Wi:
                e = x[iy, ix] @ We[:, :128] + be[:128].astype(np.float16).astype(np.float32)
                exp[rg, :len(e), i, k] = e
err = np.abs(d - exp)
print("pad", pt, pl, "dump shape", d.shape, "max |expected|", np.abs(exp).max(), "max err", err.max())
for rg in range(4):
    print("rg", rg, "err per row", [float(err[rg, :, i].max()) for i in range(NRI)])
ch = 5
np.set_printoptions(precision=3, suppress=True, linewidth=200)
print("got  ch5 rg0:\n", d[0, ch])
print("want ch5 rg0:\n", exp[0, ch])
print("got  ch40 rg1:\n", d[1, 40])
print("want ch40 rg1:\n", exp[1, 40])
