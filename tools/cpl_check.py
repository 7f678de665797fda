"""Per-layer comparison of the fused (channel-per-lane) blocks against the three-kernel path on the same frames."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np

from hfnet_slam_b200 import weights
from hfnet_slam_b200.lib import Context

H, W = int(os.environ.get("CH", "480")), int(os.environ.get("CW", "752"))
B = int(os.environ.get("BATCH", "2"))
blob = weights.synthetic_blob(seed=0)
imgs = [weights.synthetic_image(H, W, seed=10 + i, n_corners=200) for i in range(B)]


def run(env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    ctx = Context(height=H, width=W, n_levels=1, max_keypoints=1000, max_batch=B, with_global=True)
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    ctx.load_weights(blob)
    ctx.extract_batch(imgs, [1000], 0.01)
    out = {}
    for L in range(2, 19):
        out[L] = np.stack([ctx.debug_tensor(f"layer_{L}", image_index=b)[0] for b in range(B)])
    out["g"] = np.stack([ctx.debug_tensor("global_descriptor", image_index=b).reshape(-1) for b in range(B)])
    ctx.close()
    return out


ref = run({"HFB_FUSED": "0", "HFB_STEM": "0"})
new = run({"HFB_TRACE": "1"})
bad = 0
for L in range(2, 19):
    a, b = new[L], ref[L]
    scale = np.abs(b).max() + 1e-9
    err = np.abs(a - b).max() / scale
    where = np.unravel_index(np.abs(a - b).argmax(), a.shape)
    flag = "" if err < 2e-2 else "  <-- MISMATCH"
    bad += err >= 2e-2
    print(f"layer_{L}: shape {a.shape} rel max err {err:.3e} at {where} nan={np.isnan(a).sum()}{flag}")
cos = (new["g"] * ref["g"]).sum(1) / (np.linalg.norm(new["g"], axis=1) * np.linalg.norm(ref["g"], axis=1))
print("global cos", cos)
print("CPL_CHECK", "FAIL" if bad else "OK")
