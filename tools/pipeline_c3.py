"""BASELINE.json configs[2] shape, chained through the public host API on one B200: a synthetic 752x480 monocular
sequence with the SLAM extractor settings (4 pyramid levels, 675 keypoints, threshold 0.01, EuRoC.yaml:67-80), per frame
extraction + association with the previous frame + windowed projection search + two motion-only pose optimisations;
every 6th frame a "keyframe": mutual-NN matching against 10 neighbour keyframes (one batched call), keyframe-database
insert + query, one local BA (20 optimisable + 40 fixed keyframes, 3000 landmarks, 10 LM iterations) -- the schedule of
SURVEY.md 8(d)-C3.  The geometry problems are the seeded synthetic ones of hfnet_slam_b200/synthetic.py (there is no map:
Tracking / LocalMapping themselves are out of scope), so this measures the throughput of the device path under the
reference's per-frame call pattern, not SLAM accuracy.   python tools/pipeline_c3.py [n_frames]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np

from hfnet_slam_b200 import synthetic, weights
from hfnet_slam_b200.keyframe_database import KeyFrameDatabase
from hfnet_slam_b200.lib import Context, pinned_empty
from hfnet_slam_b200.optimizer import local_bundle_adjustment, pose_optimization

N = int(sys.argv[1]) if len(sys.argv) > 1 else 120
H, W = 480, 752
budgets = [217, 181, 151, 126]          # 675 features over 4 levels (HFextractor.cc:108-119)
ctx = Context(height=H, width=W, n_levels=4, scale_factor=1.2, max_keypoints=675, max_batch=1, with_global=True)
ctx.load_weights(weights.synthetic_blob(seed=0))
base = weights.synthetic_image(H, W, seed=1, n_corners=300)
frame = pinned_empty((H, W), np.uint8)
kf = KeyFrameDatabase(ctx, capacity=4096)
pose_p = synthetic.pose_problem(n=300, seed=11)
lba_p = synthetic.lba_problem(n_opt=20, n_fixed=40, n_points=3000, seed=3)
t = dict(extract=0.0, match_prev=0.0, projection=0.0, pose=0.0, kf_match=0.0, kfdb=0.0, lba=0.0)
prev = None
kfs = []
n_kf = 0


def tick(key, t0):
    t[key] += time.perf_counter() - t0


for warm in (True, False):
    n = 72 if warm else N   # the warm-up also fills the 10-neighbour window (buffers reach their steady size)
    for k in t:
        t[k] = 0.0
    wall0 = time.perf_counter()
    for i in range(n):
        frame[...] = np.roll(base, (3 * i, 5 * i), axis=(0, 1))
        t0 = time.perf_counter()
        f = ctx.extract_batch([frame], budgets, 0.01, pinned=True)[0]
        desc, xy, octv = f["descriptors"].copy(), np.stack([f["x"], f["y"]], 1), f["octave"].copy()
        tick("extract", t0)
        if prev is not None and len(desc) and len(prev[0]):
            t0 = time.perf_counter()
            ctx.match_mutual_l2(desc, prev[0], 0.6)                                   # SearchByBoW flavour vs last frame
            tick("match_prev", t0)
            t0 = time.perf_counter()
            q = min(400, len(prev[0]))                                                # SearchByProjection(F, LastFrame):
            rad = (15.0 * 1.2 ** prev[2][:q]).astype(np.float32)                      # windowed candidates on the device
            ctx.match_projection(prev[0][:q], prev[1][:q], rad, prev[2][:q] - 1, prev[2][:q] + 1, desc, xy, octv)
            tick("projection", t0)
        t0 = time.perf_counter()
        for _ in range(2):                                                            # TrackWithMotionModel + TrackLocalMap
            pose_optimization(ctx, pose_p["K"], pose_p["pose0"], pose_p["Xw"], pose_p["obs"], pose_p["inv_sigma2"])
        tick("pose", t0)
        if i % 6 == 0:
            n_kf += 1
            if kfs:
                t0 = time.perf_counter()
                nb = kfs[-10:]
                A = np.concatenate([desc] * len(nb))
                Bm = np.concatenate(nb)
                a_off = (np.arange(len(nb)) * len(desc)).astype(np.int32)
                a_cnt = np.full(len(nb), len(desc), np.int32)
                b_cnt = np.array([len(x) for x in nb], np.int32)
                b_off = (np.cumsum(b_cnt) - b_cnt).astype(np.int32)
                ctx.match_batch(1, A, Bm, a_off, a_cnt, b_off, b_cnt, 0.71875)        # SearchForTriangulation flavour
                tick("kf_match", t0)
            t0 = time.perf_counter()
            kf.add(n_kf, f["global_descriptor"])
            if n_kf > 1:
                kf.query(f["global_descriptor"])
            tick("kfdb", t0)
            t0 = time.perf_counter()
            local_bundle_adjustment(ctx, lba_p, iterations=10)
            tick("lba", t0)
            kfs.append(desc)
        prev = (desc, xy, octv)
    wall = time.perf_counter() - wall0
n_key = (N + 5) // 6
print(f"frames {N}, keyframes {n_key}, wall {wall:.3f} s -> {N / wall:.1f} frames/s (tracking + mapping work serialised on one thread)")
for k, v in t.items():
    per = n_key if k in ("kf_match", "kfdb", "lba") else N
    print(f"  {k:11s} {1e3 * v / per:8.3f} ms per {'keyframe' if per == n_key else 'frame'}   ({100 * v / wall:4.1f} % of wall)")
kf.close()
ctx.close()
