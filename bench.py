#!/usr/bin/env python
"""Benchmark of the HFNet-SLAM per-frame front-end on B200 (BASELINE.json metric "frames/sec extract+match 752x480;
local-BA ms/iter; loop-DB queries/sec").

One step = one pass of the hot path over ``--batches`` consecutive batches of ``--batch`` synthetic 752x480 frames of
one stream: HF-Net extraction (1 level, 1000 keypoints + 256-d local + 4096-d global, BASELINE.json configs[1]) and the
mutual-NN L2 match of every frame against the previous frame of the stream (configs[0]'s 1000 x 1000 brute-force match;
frame 0 of a batch continues from the last frame of the previous batch).  The default 50 batches x 8 frames walk a
144 MB ring of distinct frames (larger than the 126 MB L2), so a step is ~40 ms of sustained load.
``value`` = frames/s with the u8 frames already in HBM; ``e2e`` = the same work through the host-buffer C-ABI call the
reference shim binds (pinned H2D of the frames, D2H of keypoints / descriptors / matches inside the timed region).
The other parts of the metric and of BASELINE.json's configs are reported under ``extra``: ``c3`` (configs[2]: 4-level
extract + association + windowed search + 2 pose optimisations per frame, neighbour matching + database query + local BA
per keyframe, with its own CPU arm), ``match_c1`` (device-timed matcher roofline for 1 / 30 / 256 pairs), ``lba``
(ms/iter next to the single-threaded C restatement), ``loopdb`` (50 k x 4096 database: Q = 1 / 64, sharded over the
ranks), ``library_baseline`` (the same network as cuDNN fp16 channels_last on the same GPU).

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                      (CPU arm: the oracle port on the host cores)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

H, W, NKP, THR = 480, 752, 1000, 0.01
METRIC = "frames/sec extract+match 752x480"
L2_FLUSH_BYTES = 256 << 20
C3_BUDGETS = [217, 181, 151, 126]        # 675 features over 4 levels (src/Extractors/HFextractor.cc:108-119)


def workload_config(batch: int, batches: int, world: int = 1, streams: int = 1) -> dict:
    """The ``config`` object both arms print (the reference arm runs a bounded sample of the same workload)."""
    return {"workload": "HF-Net extract single 752x480 grayscale -> 1000 kpts + 256-d local + 4096-d global + mutual-NN L2 "
                        "match vs the previous frame of the stream (BASELINE.json configs[1] + configs[0])",
            "frames_per_step": batch * batches, "frames_per_call": batch, "threshold": THR, "levels": 1,
            "weights": "seeded random init",
            "l2": "inputs larger than L2 (ring of distinct frames, 144 MB at the defaults) + 256 MiB flush between steps",
            "streams_per_gpu": streams,
            "streams_note": "camera streams served concurrently by one GPU: one library context + CUDA stream each, every call "
                            "carries frames_per_call consecutive frames of its stream (the CPU arm is one stream)",
            "parallelism": f"replicas x{world}"}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=float(d["hbm_gbs"]), tf=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    tf_burst=float(d["bf16_tflops"]), src="measured (MEASURED_PEAKS.json; tensor = sustained)")
    return dict(hbm=6650.0, tf=1400.0, tf_burst=1590.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms DURING the timed device loop."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0=None, t1=None):
        sm, mx, reasons, pw = [], 0.0, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if t0 is not None and not (t0 <= ts <= t1):
                continue
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                pw.append(float(r[6]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw) if pw else None}


def synthetic_frames(n: int, seed0: int, h: int = H, w: int = W):
    from hfnet_slam_b200 import weights
    base = [weights.synthetic_image(h, w, seed=seed0 + i, n_corners=300) for i in range(min(n, 4))]
    out = []
    for i in range(n):      # cheap variations of a few rendered frames (shifted + flipped), distinct content per frame
        im = np.roll(base[i % len(base)], (7 * i, 13 * i), axis=(0, 1))
        out.append(np.ascontiguousarray(im[:, ::-1] if (i // len(base)) % 2 else im))
    return out


# ---------------------------------------------------------------------------------------------------- reference arm
def cpu_frame_pass(wd, prev_desc, img, cv2):
    """The reference's per-frame path restated on the CPU (oracle/): fp32 network forward (torch CPU, all threads),
    CPU threshold / top-k / resample / normalise, cv::BFMatcher(NORM_L2, crossCheck) + dist < 0.6 vs the previous frame."""
    from oracle import hfnet_ref, select_ref
    r = hfnet_ref.forward(img, wd, want_global=True)
    f = select_ref.local_features(r["scores_dense_nms"][0], r["local_descriptor_map"][0], NKP, THR)
    n = 0
    if prev_desc is not None and len(prev_desc) and len(f["descriptors"]):
        ms = cv2.BFMatcher(cv2.NORM_L2, crossCheck=True).match(f["descriptors"], prev_desc)
        n = sum(1 for m in ms if m.distance < 0.6)
    return f["descriptors"], n


def run_cpu(frames: int, warm: int):
    import cv2
    import torch
    from hfnet_slam_b200 import weights
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cv2.setNumThreads(cores)
    wd = weights.synthetic(seed=0)
    imgs = synthetic_frames(frames + warm + 1, 100)
    prev, _ = cpu_frame_pass(wd, None, imgs[0], cv2)
    for i in range(warm):
        prev, _ = cpu_frame_pass(wd, prev, imgs[1 + i], cv2)
    t0 = time.perf_counter()
    for i in range(frames):
        prev, _ = cpu_frame_pass(wd, prev, imgs[1 + warm + i], cv2)
    dt = time.perf_counter() - t0
    return frames / dt, cores, dt


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = args.ref_frames
    fps, cores, dt = run_cpu(args.steps * per_step, min(args.warmup, 2))
    sample = (f"{per_step} frames per step ({args.steps * per_step} frames in all) of the same workload: oracle fp32 forward "
              f"(torch CPU, {cores} threads) + CPU select/resample + cv2.BFMatcher crossCheck")
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.batch, args.batches, max(args.gpus, 1), args.streams),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------- extras (B200 arm)
def event_ms(torch, stream, fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def extra_match_c1(torch, dev, pk, cpu: bool):
    """BASELINE.json configs[0]: 1000 x 1000 brute-force 256-d match.  Device-timed kernel chain (prep + tcgen05
    contraction / arg-max + finalize, descriptors resident) for 1 / 30 / 256 pairs, the host-pointer
    calls, and the CPU baselines."""
    from hfnet_slam_b200 import synthetic
    from hfnet_slam_b200.lib import Context
    out = {}
    n = 1000
    A, Bd = synthetic.descriptor_pair(n, n, n_true=300, seed=0)
    with Context(height=64, width=64, n_levels=1, max_keypoints=8192, max_batch=1, with_global=False,
                 device=dev.index or 0) as c:
        stream = torch.cuda.ExternalStream(c.stream, device=dev)
        rows = []
        for P in (1, 30, 256):
            dA = torch.from_numpy(np.concatenate([A] * P)).to(dev)
            dB = torch.from_numpy(np.concatenate([Bd] * P)).to(dev)
            off = (np.arange(P) * n).astype(np.int32)
            cnt = np.full(P, n, np.int32)
            tab = torch.from_numpy(np.concatenate([off, cnt, off, cnt])).to(dev)
            idx = torch.empty(P * n, dtype=torch.int32, device=dev)
            val = torch.empty(P * n, dtype=torch.float32, device=dev)

            def run():
                c.check(c.lib.hfb_match_batch_dev(c.handle, 0, dA.data_ptr(), P * n, dB.data_ptr(), P * n, P, tab.data_ptr(),
                                                  n, n, 0.6, idx.data_ptr(), val.data_ptr()))
            ms = event_ms(torch, stream, run, 30 if P < 256 else 10)
            fl = 2.0 * P * n * n * 256
            by = 2.0 * P * n * 256 * 4 + P * n * 8
            rows.append({"pairs": P, "us_per_call": 1e3 * ms, "us_per_pair": 1e3 * ms / P,
                         "algorithmic_tflops": fl / (ms / 1e3) / 1e12, "tensor_frac": fl / (ms / 1e3) / 1e12 / pk["tf"],
                         "hbm_frac": by / (ms / 1e3) / 1e9 / pk["hbm"], "matches_pair0": int((idx[:n] >= 0).sum())})
            del dA, dB
        out["device_chain"] = rows
        out["device_chain_note"] = ("CUDA events on the library stream around hfb_match_batch_dev (memset + prep + contraction/arg-max + "
                                    "finalize); flops = 2*pairs*1000*1000*256 (algorithmic K = 256; the split product executes 3x that), "
                                    "bytes = fp32 descriptors in + match rows out; fractions of the measured sustained bf16 / HBM peaks")
        for _ in range(3):
            c.match_mutual_l2(A, Bd, 0.6)
        t0 = time.perf_counter()
        for _ in range(20):
            c.match_mutual_l2(A, Bd, 0.6)
        out["single_pair_host_ms"] = 1e3 * (time.perf_counter() - t0) / 20
        npair = 30
        A30, B30 = np.concatenate([A] * npair), np.concatenate([Bd] * npair)
        off = (np.arange(npair) * 1000).astype(np.int32)
        cnt = np.full(npair, 1000, np.int32)
        for _ in range(2):
            c.match_batch(0, A30, B30, off, cnt, off, cnt, 0.6)
        t0 = time.perf_counter()
        for _ in range(5):
            c.match_batch(0, A30, B30, off, cnt, off, cnt, 0.6)
        out["pairs_per_s_batch30_host"] = npair / ((time.perf_counter() - t0) / 5)
        out["host_note"] = "hfb_match_mutual_l2 / hfb_match_batch with pageable host descriptors (2 MB per pair up)"
    if cpu:
        import cv2
        from oracle import c_ref
        cores = os.cpu_count() or 1
        cv2.setNumThreads(cores)
        bf = cv2.BFMatcher(cv2.NORM_L2, crossCheck=True)
        bf.match(A, Bd)
        t0 = time.perf_counter()
        for _ in range(3):
            bf.match(A, Bd)
        base = {"cores": cores, "cv2_bfmatcher_crosscheck_ms": 1e3 * (time.perf_counter() - t0) / 3}
        for thr in sorted({cores, max(cores // 2, 1), 1}, reverse=True):
            c_ref.set_threads(thr)
            c_ref.match_cos_mutual(A, Bd, 0.71875)
            t0 = time.perf_counter()
            for _ in range(3):
                c_ref.match_cos_mutual(A, Bd, 0.71875)
            base[f"c_sgemm_argmax_crosscheck_ms_{thr}thr"] = 1e3 * (time.perf_counter() - t0) / 3
        c_ref.set_threads(cores)
        base["note"] = ("cv2 = the function src/Matcher.cc:229-249 calls; C = oracle/c restatement of src/Matcher.cc:845-889 (the "
                        "reference gives Eigen half the cores, src/System.cc:45-48)")
        out["cpu_baseline"] = base
    return out


def extra_lba(ctx, cpu: bool):
    from hfnet_slam_b200 import synthetic
    from hfnet_slam_b200.optimizer import local_bundle_adjustment, pose_optimization
    prob = synthetic.lba_problem(n_opt=20, n_fixed=40, n_points=3000, seed=3)
    local_bundle_adjustment(ctx, prob, iterations=2)
    t0 = time.perf_counter()
    out = local_bundle_adjustment(ctx, prob, iterations=10)
    dt = time.perf_counter() - t0
    r = {"ms_per_iter": 1e3 * dt / max(out["iterations"], 1), "ms_total": 1e3 * dt, "iterations": out["iterations"],
         "trials": out["trials"], "edges": int(len(prob["cam_idx"])), "chi2": [out["initial_chi2"], out["final_chi2"]],
         "gpu_launches": out["gpu_launches"],
         "problem": "20 optimisable + 40 fixed keyframes, 3000 landmarks (SURVEY.md 8(d) C3)"}
    pp = synthetic.pose_problem(n=300, seed=11)
    pose_optimization(ctx, pp["K"], pp["pose0"], pp["Xw"], pp["obs"], pp["inv_sigma2"])
    t0 = time.perf_counter()
    for _ in range(20):
        po = pose_optimization(ctx, pp["K"], pp["pose0"], pp["Xw"], pp["obs"], pp["inv_sigma2"])
    r["pose_optimization_ms"] = 1e3 * (time.perf_counter() - t0) / 20
    if cpu:
        from oracle import c_ref
        c_ref.lba_optimize(prob, 2)
        t0 = time.perf_counter()
        c = c_ref.lba_optimize(prob, 10)
        dtc = time.perf_counter() - t0
        t0 = time.perf_counter()
        for _ in range(20):
            c_ref.pose_optimize(pp["K"], pp["pose0"], pp["Xw"], pp["obs"], pp["inv_sigma2"])
        r["cpu_baseline"] = {"ms_per_iter": 1e3 * dtc / max(c["iterations"], 1), "ms_total": 1e3 * dtc, "iterations": c["iterations"],
                             "trials": c["trials"], "cores": 1, "kind": "port",
                             "pose_optimization_ms": 1e3 * (time.perf_counter() - t0) / 20,
                             "note": "oracle/c/lba_ref.c: single-threaded double-precision restatement (g2o is built without "
                                     "OpenMP, Thirdparty/g2o/CMakeLists.txt:48); same LM path as the device "
                                     f"(iterations {c['iterations']} / {out['iterations']}, trials {c['trials']} / {out['trials']})"}
    return r


def c3_gpu(torch, dev, n_frames: int, threads: int = 2):
    """BASELINE.json configs[2] call pattern through the public host API on one B200 (see module docstring).
    threads = 2: tracking (extract + associate + windowed search + 2 pose optimisations per frame) on the calling thread
    and local mapping (neighbour matching + place recognition + local BA per keyframe) on a second host thread with its
    own library context, as the reference runs them (src/System.cc starts LocalMapping::Run on its own thread; keyframes
    are handed over through a queue, src/LocalMapping.cc:InsertKeyFrame).  threads = 1: everything serialised."""
    import queue
    import threading
    from hfnet_slam_b200 import synthetic, weights
    from hfnet_slam_b200.keyframe_database import KeyFrameDatabase
    from hfnet_slam_b200.lib import Context, KeyFrameStore, pinned_empty
    from hfnet_slam_b200.optimizer import local_bundle_adjustment, pose_optimization
    ctx = Context(height=H, width=W, n_levels=4, scale_factor=1.2, max_keypoints=675, max_batch=1, with_global=True,
                  device=dev.index or 0)
    # LocalMapping's context: matcher / database / BA workspaces only (no extraction)
    ctx_map = ctx if threads < 2 else Context(height=64, width=64, n_levels=1, max_keypoints=64, max_batch=1,
                                              with_global=False, device=dev.index or 0)
    store = KeyFrameStore(ctx, n_slots=16, rows_per_slot=ctx.kp_cap)     # keyframe descriptors stay in HBM
    ctx.load_weights(weights.synthetic_blob(seed=0))
    base = weights.synthetic_image(H, W, seed=1, n_corners=300)
    frame = pinned_empty((H, W), np.uint8)
    match_out = (pinned_empty((1, ctx.kp_cap), np.int32), pinned_empty((1, ctx.kp_cap), np.float32))
    kf = KeyFrameDatabase(ctx_map, capacity=4096)
    pose_p = synthetic.pose_problem(n=300, seed=11)
    lba_p = synthetic.lba_problem(n_opt=20, n_fixed=40, n_points=3000, seed=3)
    t = {}
    prev, kfs, n_kf = None, [], 0

    def tick(key, t0):
        t[key] = t.get(key, 0.0) + time.perf_counter() - t0

    def map_keyframe(job):
        kid, n_desc, gdesc = job
        t0 = time.perf_counter()
        if kfs:
            store.match_neighbours(ctx_map, kid, kfs[-10:], 1, 0.71875, n_desc)   # SearchForTriangulation flavour
        if len(kfs) >= 12:
            store.erase(kfs[-12])
        tick("kf_match", t0)
        t0 = time.perf_counter()
        kf.add(kid, gdesc)
        if kid > 1:
            kf.query(gdesc)
        tick("kfdb", t0)
        t0 = time.perf_counter()
        local_bundle_adjustment(ctx_map, lba_p, iterations=10)
        tick("lba", t0)
        kfs.append(kid)

    jobs = queue.Queue()

    def mapping_thread():
        while True:
            job = jobs.get()
            if job is None:
                jobs.task_done()
                return
            map_keyframe(job)
            jobs.task_done()

    worker = None
    if threads >= 2:
        worker = threading.Thread(target=mapping_thread, daemon=True)
        worker.start()
    wall = 0.0
    for warm in (True, False):
        n = 72 if warm else n_frames      # the warm-up also fills the 10-neighbour window
        t.clear()
        wall0 = time.perf_counter()
        for i in range(n):
            frame[...] = np.roll(base, (3 * i, 5 * i), axis=(0, 1))
            t0 = time.perf_counter()
            f, midx, _ = ctx.extract_match_batch([frame], C3_BUDGETS, 0.01, 0, 0.6, pinned=True, out=match_out)
            f = f[0]
            desc, xy, octv = f["descriptors"].copy(), np.stack([f["x"], f["y"]], 1), f["octave"].copy()
            tick("extract+associate", t0)
            if prev is not None and len(desc) and len(prev[0]):
                t0 = time.perf_counter()
                q = min(400, len(prev[0]))                                       # SearchByProjection(F, LastFrame) on the
                rad = (15.0 * 1.2 ** prev[2][:q]).astype(np.float32)             # descriptors resident in HBM
                ctx.match_projection_frame(0, np.arange(q, dtype=np.int32), prev[1][:q], rad, prev[2][:q] - 1,
                                           prev[2][:q] + 1, nf=len(desc))
                tick("projection", t0)
            t0 = time.perf_counter()
            for _ in range(2):                                                   # TrackWithMotionModel + TrackLocalMap
                pose_optimization(ctx, pose_p["K"], pose_p["pose0"], pose_p["Xw"], pose_p["obs"], pose_p["inv_sigma2"])
            tick("pose", t0)
            if i % 6 == 0:
                n_kf += 1
                t0 = time.perf_counter()
                store.put_frame(ctx, n_kf, 0, len(desc))                     # device to device, prepared once
                tick("kf_insert", t0)
                job = (n_kf, len(desc), f["global_descriptor"].copy())
                if worker is not None:
                    jobs.put(job)                                            # LocalMapping::InsertKeyFrame
                else:
                    map_keyframe(job)
            prev = (desc, xy, octv)
        jobs.join()                                                          # local mapping has caught up
        wall = time.perf_counter() - wall0
    if worker is not None:
        jobs.put(None)
        worker.join()
    n_key = (n_frames + 5) // 6
    stages = {k: (1e3 * v / (n_key if k in ("kf_match", "kfdb", "lba", "kf_insert") else n_frames)) for k, v in t.items()}
    kf.close()
    store.close()
    if ctx_map is not ctx:
        ctx_map.close()
    ctx.close()
    return {"frames": n_frames, "keyframes": n_key, "frames_per_s": n_frames / wall, "host_threads": threads,
            "stage_ms": stages, "stage_ms_note": "kf_insert / kf_match / kfdb / lba per keyframe, the others per frame; with two "
                                               "host threads the per-keyframe stages run beside the per-frame ones"}


def c5_stream(torch, dev, n_frames: int, barrier, maxred, world: int):
    """BASELINE.json configs[4]: one 512 x 512 TUM-VI-shaped camera stream per GPU (Examples/Monocular/TUM-VI.yaml:66-67 -> 850
    features, 4 levels, 1.2, threshold 0.02), frame by frame through the host API: 4-level extraction + association with
    the previous frame (one call, frame from page-locked memory, keypoints / descriptors / global descriptor / match row
    back) + the masked best-2 search of SearchByProjection(CurrentFrame, LastFrame) on the resident descriptors.  All ranks
    run their stream at the same time; the figure is frames/s over all streams (slowest rank's wall clock)."""
    from hfnet_slam_b200 import weights
    from hfnet_slam_b200.extractor import features_per_level
    from hfnet_slam_b200.lib import Context, pinned_empty
    Hc = Wc = 512
    budgets = features_per_level(850, 4, 1.2)
    ctx = Context(height=Hc, width=Wc, n_levels=4, scale_factor=1.2, max_keypoints=850, max_batch=1, with_global=True,
                  device=dev.index or 0)
    ctx.load_weights(weights.synthetic_blob(seed=0))
    base = weights.synthetic_image(Hc, Wc, seed=12, n_corners=250)
    frame = pinned_empty((Hc, Wc), np.uint8)
    match_out = (pinned_empty((1, ctx.kp_cap), np.int32), pinned_empty((1, ctx.kp_cap), np.float32))
    prev, n_kp, n_cand, dt = None, 0, 0, 0.0
    for warm in (True, False):
        n = 10 if warm else n_frames
        if not warm:
            barrier()
        t0 = time.perf_counter()
        for i in range(n):
            frame[...] = np.roll(base, (2 * i, 3 * i), axis=(0, 1))
            f, midx, _ = ctx.extract_match_batch([frame], budgets, 0.02, 0, 0.6, pinned=True, out=match_out)
            f = f[0]
            xy, octv = np.stack([f["x"], f["y"]], 1), f["octave"].copy()
            if prev is not None and len(xy) and len(prev[0]):
                q = len(prev[0])
                uv = prev[0] + np.float32([3.0, 2.0])
                rad = (np.float32(15.0) * np.float32(1.2) ** prev[1]).astype(np.float32)
                idx, _, _ = ctx.match_projection_frame(0, np.arange(q, dtype=np.int32), uv, rad, prev[1] - 1, prev[1] + 1, nf=len(xy))
                if not warm:
                    n_cand += int((idx[:, 0] >= 0).sum())
            if not warm:
                n_kp += len(xy)
            prev = (xy, octv)
        if not warm:
            dt = maxred(time.perf_counter() - t0)
    ctx.close()
    return {"frames_per_s": world * n_frames / dt, "streams": world, "ms_per_frame_per_stream": 1e3 * dt / n_frames,
            "keypoints_per_frame": n_kp / n_frames, "windowed_matches_per_frame": n_cand / max(n_frames - 0, 1),
            "frames": n_frames, "note": "512x512, 4 levels, 850 keypoints, threshold 0.02, one stream per GPU; host API, "
                                        "one frame per call (H2D of the frame, D2H of the features and match row inside)"}


def c3_cpu(n_frames: int):
    """The same per-frame / per-keyframe schedule restated on the CPU (oracle/): the reference's CPU stages as they are
    (cv::resize pyramid, threshold / top-k / Resampler, cv::BFMatcher, C restatements of the matcher, database scan, pose
    optimisation and local BA) plus the fp32 network on torch-CPU in place of the TensorRT engine."""
    import cv2
    import torch
    from hfnet_slam_b200 import synthetic, weights
    from oracle import c_ref, hfnet_ref, select_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cv2.setNumThreads(cores)
    c_ref.set_threads(cores)
    wd = weights.synthetic(seed=0)
    base = weights.synthetic_image(H, W, seed=1, n_corners=300)
    pose_p = synthetic.pose_problem(n=300, seed=11)
    lba_p = synthetic.lba_problem(n_opt=20, n_fixed=40, n_points=3000, seed=3)
    db = np.zeros((0, 4096), np.float32)
    t = {}
    prev, kfs = None, []

    def tick(key, t0):
        t[key] = t.get(key, 0.0) + time.perf_counter() - t0

    def extract(img):
        pyr = select_ref.compute_pyramid(img, 4, 1.2)
        per, g = [], None
        for l, im in enumerate(pyr):
            r = hfnet_ref.forward(im, wd, want_global=(l == 0))
            if l == 0:
                g = r["global_descriptor"][0]
            per.append(select_ref.local_features(r["scores_dense_nms"][0], r["local_descriptor_map"][0], C3_BUDGETS[l], 0.01))
        f = select_ref.concat_levels(per, 1.2)
        return f, g

    extract(base)                                                              # warm-up (thread pools, allocations)
    wall0 = time.perf_counter()
    for i in range(n_frames):
        img = np.roll(base, (3 * i, 5 * i), axis=(0, 1))
        t0 = time.perf_counter()
        f, g = extract(img)
        desc, xy, octv = f["descriptors"], np.stack([f["x"], f["y"]], 1), f["octave"]
        tick("extract", t0)
        if prev is not None and len(desc) and len(prev[0]):
            t0 = time.perf_counter()
            cv2.BFMatcher(cv2.NORM_L2, crossCheck=True).match(desc, prev[0])
            tick("associate", t0)
            t0 = time.perf_counter()
            q = min(400, len(prev[0]))                                         # masked best-2 over the windows
            rad = (15.0 * 1.2 ** prev[2][:q]).astype(np.float32)
            d2 = 2.0 - 2.0 * prev[0][:q] @ desc.T
            ok = (np.abs(xy[None, :, 0] - prev[1][:q, 0:1]) < rad[:, None]) & (np.abs(xy[None, :, 1] - prev[1][:q, 1:2]) < rad[:, None]) & \
                 (octv[None] >= prev[2][:q, None] - 1) & (octv[None] <= prev[2][:q, None] + 1)
            np.partition(np.where(ok, d2, np.inf), 1, axis=1)[:, :2]
            tick("projection", t0)
        t0 = time.perf_counter()
        for _ in range(2):
            c_ref.pose_optimize(pose_p["K"], pose_p["pose0"], pose_p["Xw"], pose_p["obs"], pose_p["inv_sigma2"])
        tick("pose", t0)
        if i % 6 == 0:
            if kfs:
                t0 = time.perf_counter()
                for nb in kfs[-10:]:
                    c_ref.match_cos_mutual(desc, nb, 0.71875)
                tick("kf_match", t0)
            t0 = time.perf_counter()
            db = np.concatenate([db, g[None]])
            if len(db) > 1:
                c_ref.set_threads(1)                                           # the reference scans under the database mutex
                c_ref.kfdb_scores(g, db)
                c_ref.set_threads(cores)
            tick("kfdb", t0)
            t0 = time.perf_counter()
            c_ref.lba_optimize(lba_p, 10)
            tick("lba", t0)
            kfs.append(desc)
        prev = (desc, xy, octv)
    wall = time.perf_counter() - wall0
    n_key = (n_frames + 5) // 6
    stages = {k: (1e3 * v / (n_key if k in ("kf_match", "kfdb", "lba") else n_frames)) for k, v in t.items()}
    return {"frames": n_frames, "keyframes": n_key, "frames_per_s": n_frames / wall, "cores": cores, "kind": "port",
            "stage_ms": stages,
            "note": "keyframe neighbour window is still filling in this short sample (<= 10 neighbours); network = oracle fp32 "
                    "on torch-CPU (the reference runs it on TensorRT/GPU: README.md:17 claims 50 FPS on an RTX 2070)"}


def extra_loopdb(torch, dist, ctx, dev, pk, world, rank, n_db, barrier, maxred, stream, cpu: bool):
    """BASELINE.json configs[3]: 4096-d search over a 50 k-keyframe database, rows sharded by id % world."""
    from hfnet_slam_b200.keyframe_database import KeyFrameDatabase
    from hfnet_slam_b200.lib import _i64p, ptr
    rows = torch.randn(n_db // world, 4096, device=dev)
    rows /= rows.norm(dim=1, keepdim=True)
    kf = KeyFrameDatabase(ctx, capacity=rows.shape[0])
    ids = (np.arange(rows.shape[0], dtype=np.int64) * world + rank)
    ctx.check(ctx.lib.hfb_kfdb_add_dev(kf.handle, ptr(ids, _i64p), rows.data_ptr(), rows.shape[0]))
    q = rows[5].cpu().numpy() + 0.002 * np.random.default_rng(0).standard_normal(4096).astype(np.float32)
    q /= np.linalg.norm(q)
    out = {"rows_total": n_db, "rows_per_gpu": int(rows.shape[0])}
    nq = 50
    if hasattr(kf, "connect_shards") and world > 1:
        kf.connect_shards(dist, rank, world)
    sharded = getattr(kf, "query_sharded", None) if world > 1 else None
    run_q = sharded if sharded else kf.query      # sharded: scan + device record + peer exchange + device merge, one D2H
    for _ in range(3):
        run_q(q)
    barrier()
    t0 = time.perf_counter()
    for _ in range(nq):
        run_q(q)
    barrier()
    dt = maxred(time.perf_counter() - t0)
    out["queries_per_sec_e2e"] = nq / dt
    out["e2e_note"] = "host query in, global candidate list out (collective over the ranks when sharded)"
    out["collective"] = ("none" if world == 1 else
                         "peer-memory record exchange inside the library: 16+16*64 B per rank written into every peer's inbox "
                         "over NVLink (CUDA IPC mapping) + system-scope flag, merged on the device")
    # device-only scan (HBM roofline of the scan kernel), Q = 1 and Q = 64
    dq = torch.from_numpy(q).to(dev)
    dsc = torch.empty(rows.shape[0], device=dev)
    dbest = torch.empty(1, device=dev)

    def scan1():
        ctx.check(ctx.lib.hfb_kfdb_scan_dev(kf.handle, dq.data_ptr(), 1, dsc.data_ptr(), dbest.data_ptr()))
    scan_ms = event_ms(torch, stream, scan1, 20)
    scan_gbs = rows.shape[0] * 4096 * 4 / (scan_ms / 1e3) / 1e9
    out.update({"scan_ms": scan_ms, "scan_gbs": scan_gbs, "scan_hbm_frac": scan_gbs / pk["hbm"]})
    Q = 64
    dq64 = (rows[:Q] + 0.002 * torch.randn(Q, 4096, device=dev))
    dq64 /= dq64.norm(dim=1, keepdim=True)

    def scan64():
        ctx.check(ctx.lib.hfb_kfdb_query_batch_dev(kf.handle, dq64.data_ptr(), Q, 0.8, 0.0))
    ms64 = event_ms(torch, stream, scan64, 10, warm=3)
    t0 = time.perf_counter()
    res64 = kf.query_batch(dq64.cpu().numpy(), cap=256)
    host64 = time.perf_counter() - t0
    out["q64"] = {"queries": Q, "device_ms": ms64, "queries_per_sec_device": Q / (ms64 / 1e3),
                  "passes_over_rows_equiv": ms64 / scan_ms, "rows_gbs": rows.shape[0] * 4096 * 4 / (ms64 / 1e3) / 1e9,
                  "rows_hbm_frac": rows.shape[0] * 4096 * 4 / (ms64 / 1e3) / 1e9 / pk["hbm"],
                  "tf32_tflops": 2.0 * Q * rows.shape[0] * 4096 / (ms64 / 1e3) / 1e12,
                  "host_call_ms": 1e3 * host64, "candidates_query0": int(len(res64[0][0])),
                  "note": "hfb_kfdb_query_batch: kind::tf32 tcgen05 pass over the rows (selection) + exact fp32 re-scoring of the "
                          "marked pairs + per-query candidate lists, device-timed; results identical to 64 single queries"}
    if cpu:
        from oracle import c_ref
        dbh = rows.cpu().numpy()
        base = {}
        for thr in (1, os.cpu_count() or 1):
            c_ref.set_threads(thr)
            c_ref.kfdb_scores(q, dbh)
            t0 = time.perf_counter()
            for _ in range(2):
                c_ref.kfdb_scores(q, dbh)
            base[f"scan_ms_{thr}thr"] = 1e3 * (time.perf_counter() - t0) / 2
        base["note"] = "oracle/c restatement of the per-keyframe loop (src/KeyFrameDatabase.cc:86-96); the reference scans under the database mutex (1 thread)"
        out["cpu_baseline"] = base
        del dbh
    kf.close()
    del rows
    return out


def host_transfer_ceiling(torch, dev, barrier, maxred, d2h_bytes: int, h2d_bytes: int, reps: int = 200):
    """What the box's host path gives every rank when all ranks move one call's results device to host and one call's frames
    host to device at the same time (page-locked buffers, both directions concurrently, nothing else running): the upper
    bound of the end-to-end arm at this GPU count (tools/pcie_bw.py is the stand-alone version)."""
    h_out = [torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
    h_in = [torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
    d_out = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    d_in = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def loop(n):
        for i in range(n):
            with torch.cuda.stream(s1):
                h_out[i & 1].copy_(d_out, non_blocking=True)
            with torch.cuda.stream(s2):
                d_in.copy_(h_in[i & 1], non_blocking=True)
        torch.cuda.synchronize(dev)

    loop(10)
    barrier()
    t0 = time.perf_counter()
    loop(reps)
    dt = maxred(time.perf_counter() - t0)
    return {"gbs_per_gpu": reps * (d2h_bytes + h2d_bytes) / dt / 1e9, "ms_per_call_transfers": 1e3 * dt / reps,
            "bytes_per_call": d2h_bytes + h2d_bytes,
            "note": "every rank copies one call's outputs D2H and inputs H2D concurrently (pinned memory, slowest rank); the "
                    "end-to-end arm cannot finish a call faster than this on this box"}


# ---------------------------------------------------------------------------------------------------- B200 arm
def main_gpu(args):
    import torch
    from hfnet_slam_b200 import weights
    from hfnet_slam_b200.lib import Context, pinned_empty

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # every rank on its own slice of the host cores (the ranks otherwise share cores 0..n for their launch threads)
        try:
            ncpu = os.cpu_count() or 1
            per = max(ncpu // world, 1)
            os.sched_setaffinity(0, set(range(rank * per, min((rank + 1) * per, ncpu))))
        except (AttributeError, OSError):
            pass
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, NB = args.batch, args.batches
    pk = peaks()
    cpu_arms = rank == 0 and world == 1 and not args.skip_cpu

    # S camera streams per GPU: one context (weights, activations, CUDA stream, captured graph) each.  The kernels of
    # different contexts overlap on the device: the late layers' 64..96-CTA grids and every launch's ramp / tail leave
    # SMs idle that another stream's kernels fill.
    S = max(1, min(args.streams, NB))
    blob = weights.synthetic_blob(seed=0)
    ctxs = []
    for _ in range(S):
        c = Context(height=H, width=W, n_levels=1, max_keypoints=NKP, max_batch=B, with_global=True, device=local)
        c.load_weights(blob)
        ctxs.append(c)
    ctx = ctxs[0]
    frames = synthetic_frames(B * NB, 1000 * rank)
    pinned_ring = pinned_empty((NB, B, H, W), np.uint8)              # e2e arm: frames arrive in page-locked host memory
    for i in range(NB):                                              # (slots of one capture ring: one H2D transfer per call)
        pinned_ring[i] = np.stack(frames[i * B:(i + 1) * B])
    d_ring = torch.from_numpy(pinned_ring).to(dev)                   # resident in HBM before the timed region
    del frames
    d_ptrs = [d_ring[i].data_ptr() for i in range(NB)]
    host_batches = [[pinned_ring[i, b] for b in range(B)] for i in range(NB)]
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    streams = [torch.cuda.ExternalStream(c.stream, device=dev) for c in ctxs]
    stream = streams[0]
    budgets = [NKP]

    def join_streams():
        """stream 0 waits for the work enqueued on the other contexts' streams (device-side, no host synchronisation)"""
        for st in streams[1:]:
            e = torch.cuda.Event()
            e.record(st)
            stream.wait_event(e)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_dev():
        # per batch: extraction + frame-to-previous-frame association as one enqueue; the matching runs on the main stream
        # while the global branch (layer_8 .. FC) finishes on the side stream
        # batch i belongs to stream i % S (consecutive batches of one stream: i, i + S, ...)
        for i, p in enumerate(d_ptrs):
            ctxs[i % S].extract_match_batch_dev(p, B, budgets, THR, 0, 0.6)

    def step_dev_one():
        for p in d_ptrs:
            ctx.extract_match_batch_dev(p, B, budgets, THR, 0, 0.6)

    match_outs = [(pinned_empty((B, c.kp_cap), np.int32), pinned_empty((B, c.kp_cap), np.float32)) for c in ctxs]
    match_out = match_outs[0]
    host_stats = {"kp": 0, "matches": 0}

    host_mode = {"descriptors": True}

    def host_stream(si, res):
        # HFextractor::operator() on host frames (H2D of the u8 frames, D2H of keypoints / descriptors / global
        # descriptors) + the association on the descriptors still resident in HBM (D2H of the match rows) through
        # the synchronous hfb_extract_match_batch, batch after batch of stream si (ctypes releases the GIL in the call)
        kp, idx = 0, None
        for i in range(si, NB, S):
            feats, idx, val = ctxs[si].extract_match_batch(host_batches[i], budgets, THR, 0, 0.6, pinned=True, out=match_outs[si],
                                                           descriptors=host_mode["descriptors"])
            kp += sum(len(f["x"]) for f in feats)
        res[si] = (kp, int((idx >= 0).sum()) if idx is not None else 0)

    def step_host():
        import threading
        res = [None] * S
        if S == 1:
            host_stream(0, res)
        else:
            th = [threading.Thread(target=host_stream, args=(si, res)) for si in range(S)]
            for t in th:
                t.start()
            for t in th:
                t.join()
        host_stats["kp"] = sum(r[0] for r in res)
        host_stats["matches"] = res[(NB - 1) % S][1]

    def timed(fn, steps, warm):
        for _ in range(warm):
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        lc0 = sum(c.launch_count for c in ctxs)
        wall0 = time.perf_counter()
        for s in range(steps):
            flush.fill_(s & 0xFF)                                    # L2 flush between timed iterations (untimed)
            torch.cuda.synchronize(dev)
            ev[s][0].record(stream)
            fn()
            join_streams()                                           # the end event waits for every context's stream
            ev[s][1].record(stream)
        barrier()
        wall1 = time.perf_counter()
        ms = float(sum(a.elapsed_time(b) for a, b in ev))
        return ms, (wall0, wall1), sum(c.launch_count for c in ctxs) - lc0

    warm = max(args.warmup, 3)
    n_host = max(3, min(args.steps, 10))
    for c in ctxs:
        c.reset_stream()
    with ClockSampler(local) as cs:
        time.sleep(0.05)
        ms_dev, (w0, w1), launches = timed(step_dev, args.steps, warm)
    clocks = cs.summary(w0, w1)
    clocks["note"] = "nvidia-smi sampled every 20 ms inside the timed device loop"
    for c in ctxs:
        c.reset_stream()
    ms_host, (h0, h1), _ = timed(step_host, n_host, 2)
    # the same host-buffer call with the 256-d local descriptors left resident in HBM (hfb_features.descriptors = NULL):
    # keypoints, global descriptors and match rows come back; the matcher, the resident windowed search and the keyframe
    # store read the descriptors where the extraction left them
    for c in ctxs:
        c.reset_stream()
    host_mode["descriptors"] = False
    ms_host_res, _, _ = timed(step_host, n_host, 2)
    host_mode["descriptors"] = True
    kp_res = host_stats["kp"]
    ms_one = None
    if S > 1:                                                        # the same step on ONE context / stream, for continuity
        ctx.reset_stream()
        ms_one, _, _ = timed(step_dev_one, max(3, args.steps // 4), 2)
        ms_one /= max(3, args.steps // 4)
    # sanity inside the bench: the device chain produced real keypoints and matches
    f0 = ctx.fetch_features(0)
    midx, _ = ctx.fetch_matches(1 if B > 1 else 0, len(ctx.fetch_features(1 if B > 1 else 0)["x"]))

    def maxred(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_dev, ms_host, ms_host_res = maxred(ms_dev), maxred(ms_host), maxred(ms_host_res)
    if ms_one is not None:
        ms_one = maxred(ms_one)
    value = world * B * NB * args.steps / (ms_dev / 1e3)
    e2e = world * B * NB * n_host / (ms_host / 1e3)
    h2d = B * NB * H * W
    d2h = host_stats["kp"] * (16 + 256 * 4) + B * NB * (4096 * 4 + 32) + 2 * NB * B * NKP * 4

    # ---- roofline of the dominant kernel, measured live with CUDA events on the library's stream (one batch, un-graphed)
    prof = ctx.profile_extract(B, budgets, THR)
    prof = ctx.profile_extract(B, budgets, THR)
    agg = {}
    for r in prof:
        a = agg.setdefault(r["name"], dict(ms=0.0, bytes=0.0, flops=0.0, n=0))
        a["ms"] += r["ms"]; a["bytes"] += r["bytes"]; a["flops"] += r["flops"]; a["n"] += 1
    total_ms = sum(a["ms"] for a in agg.values())
    name, a = max(agg.items(), key=lambda kv: kv[1]["ms"])
    t_s = a["ms"] / 1e3 / a["n"]
    by, fl = a["bytes"] / a["n"], a["flops"] / a["n"]
    frac_h = by / t_s / 1e9 / pk["hbm"] if t_s > 0 else 0.0
    frac_t = fl / t_s / 1e12 / pk["tf"] if t_s > 0 else 0.0
    if frac_t >= frac_h:
        roof = {"bound": "tensor", "achieved": fl / t_s / 1e12, "peak": pk["tf"], "unit": "TFLOP/s", "frac": frac_t}
    else:
        roof = {"bound": "hbm", "achieved": by / t_s / 1e9, "peak": pk["hbm"], "unit": "GB/s", "frac": frac_h}
    traffic = None
    tp = ROOT / "profiles" / "ncu_traffic.json"      # dram__bytes_read+write per launch from the committed ncu --set full capture
    if tp.exists():
        traffic = json.loads(tp.read_text()).get(name, {}).get("dram_bytes_per_launch")
    roof.update({"traffic": traffic, "kernel": name, "ms_per_launch": a["ms"] / a["n"], "share_of_step": a["ms"] / total_ms,
                 "peak_source": pk["src"], "algorithmic_bytes": by, "algorithmic_flops": fl,
                 "launch": f"one launch over a batch of {B} frames"})
    kernels = sorted(({"name": k, "ms": v["ms"], "share": v["ms"] / total_ms,
                       "hbm_frac": (v["bytes"] / (v["ms"] / 1e3) / 1e9 / pk["hbm"]) if v["ms"] > 0 else 0,
                       "tensor_frac": (v["flops"] / (v["ms"] / 1e3) / 1e12 / pk["tf"]) if v["ms"] > 0 else 0}
                      for k, v in agg.items()), key=lambda r: -r["ms"])
    step_flops = 2 * 4.226e9 * B        # 4226 MMAC per 752x480 frame
    extra = {"keypoints_frame0": int(len(f0["x"])), "matches_frame1_vs_frame0": int((midx >= 0).sum()),
             "host_keypoints_per_step": host_stats["kp"], "host_matches_last_batch": host_stats["matches"],
             "ungraphed_batch_ms": total_ms, "kernels": kernels,
             "whole_batch_tensor_frac": step_flops / (ms_dev / 1e3 / (args.steps * NB)) / 1e12 / pk["tf"]}
    extra["e2e_resident_descriptors"] = {
        "value": world * B * NB * n_host / (ms_host_res / 1e3), "unit": "frames/s", "ms_per_step": ms_host_res / n_host,
        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": kp_res * 16 + B * NB * (4096 * 4 + 32) + 2 * NB * B * NKP * 4,
        "note": "the end-to-end arm with hfb_features.descriptors = NULL: frames in, keypoints + global descriptors + match "
                "rows out, the 256-d local descriptors stay in HBM for the device-side consumers (association, resident "
                "windowed search, keyframe store); the headline e2e above moves them to the host as the reference's Frame holds them"}
    if ms_one is not None:
        extra["single_context"] = {"frames_per_s": world * B * NB / (ms_one / 1e3), "ms_per_batch": ms_one / NB,
                                   "note": "the same device-resident step with every batch on ONE context / CUDA stream "
                                           "(the round-1 configuration)"}

    # ---- single-stream latency: a batch of ONE frame per call == the reference's per-frame TrackMonocular loop (every
    # frame is matched against the frame of the previous call)
    ctx.reset_stream()
    one_ptrs = [d_ring[i, 0].data_ptr() for i in range(NB)]
    it = {"i": 0}

    def one_dev():
        ctx.extract_match_batch_dev(one_ptrs[it["i"] % NB], 1, budgets, THR, 0, 0.6)
        it["i"] += 1

    match_out1 = (pinned_empty((1, ctx.kp_cap), np.int32), pinned_empty((1, ctx.kp_cap), np.float32))

    def one_host():
        ctx.extract_match_batch([pinned_ring[it["i"] % NB, 0]], budgets, THR, 0, 0.6, pinned=True, out=match_out1)
        it["i"] += 1

    ms1_dev = event_ms(torch, stream, one_dev, 100)
    ms1_host = event_ms(torch, stream, one_host, 100)
    extra["single_stream"] = {"device_ms_per_frame": ms1_dev, "host_api_ms_per_frame": ms1_host,
                              "frames_per_s_host_api": 1e3 / ms1_host,
                              "note": "one frame per call, streaming association with the previous call's frame"}

    ceil = host_transfer_ceiling(torch, dev, barrier, maxred, d2h // NB, h2d // NB)
    ceil["frames_per_s_bound"] = world * B / (ceil["ms_per_call_transfers"] / 1e3)
    ceil["e2e_fraction_of_bound"] = e2e / ceil["frames_per_s_bound"]
    extra["host_transfer_ceiling"] = ceil

    if not args.skip_extra:
        if world == 1:
            extra["match_c1"] = extra_match_c1(torch, dev, pk, cpu_arms)
            extra["lba"] = extra_lba(ctx, cpu_arms)
        extra["c5"] = c5_stream(torch, dev, args.c5_frames, barrier, maxred, world)
        extra["loopdb"] = extra_loopdb(torch, dist, ctx, dev, pk, world, rank, args.db_rows, barrier, maxred, stream, cpu_arms)
        if world == 1:
            c3 = {"gpu": c3_gpu(torch, dev, args.c3_frames, 2), "gpu_one_thread": c3_gpu(torch, dev, args.c3_frames, 1)}
            if cpu_arms:
                c3["cpu"] = c3_cpu(args.c3_cpu_frames)
                c3["speedup_vs_cpu_arm"] = c3["gpu"]["frames_per_s"] / c3["cpu"]["frames_per_s"]
            c3["vs_published_50fps"] = c3["gpu"]["frames_per_s"] / 50.0
            c3["note"] = ("BASELINE.json configs[2] schedule on synthetic frames through the public host API; 'gpu' = tracking and "
                          "local mapping on two host threads / two contexts as in the reference (src/System.cc), "
                          "'gpu_one_thread' = the same work serialised on one host thread (the CPU arm is serialised too); "
                          "'published' = the reference's own 50 FPS claim (README.md:17, RTX 2070)")
            c3["one_thread_vs_published_50fps"] = c3["gpu_one_thread"]["frames_per_s"] / 50.0
            extra["c3"] = c3
            try:
                sys.path.insert(0, str(ROOT / "tools"))
                import library_baseline
                eager, graph = library_baseline.cudnn_fp16_forward_ms(weights.synthetic(seed=0), pinned_ring[0], dev)
                extra["library_baseline"] = {"cudnn_fp16_channels_last_ms_per_batch": graph if graph is not None else eager,
                                             "eager_ms_per_batch": eager, "frames": B,
                                             "this_library_ms_per_batch": ms_dev / (args.steps * NB),
                                             "note": "network only (no selection / sampling / matching) as torch cuDNN fp16 channels_last, CUDA-graph "
                                                     "replay: the vendor-library stand-in for the reference's TensorRT FP16 engine"}
            except Exception as ex:
                extra["library_baseline"] = {"error": str(ex)[:200]}

    cpu = None
    if cpu_arms:
        fps, cores, dt = run_cpu(args.cpu_frames, 1)
        cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"{args.cpu_frames} frames of the same workload: oracle fp32 forward (torch CPU, {cores} threads) + CPU "
                         f"select/resample + cv2.BFMatcher crossCheck, {dt:.1f} s"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f16 operands / f32 accumulate (network), f32 (select, match recheck)",
                "data": "synthetic", "config": workload_config(B, NB, world, S),
                "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_host / n_host, "steps": n_host},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "extra": extra,
                "ms_per_batch": ms_dev / (args.steps * NB),
                "wall_s": {"device_loop": w1 - w0, "host_loop": h1 - h0}}
        print(json.dumps(line))
    for c in ctxs:
        c.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="frames per call and GPU")
    ap.add_argument("--batches", type=int, default=50, help="calls per step (ring of distinct frame batches)")
    ap.add_argument("--streams", type=int, default=4, help="camera streams (library contexts) served concurrently per GPU")
    ap.add_argument("--db-rows", type=int, default=50000)
    ap.add_argument("--cpu-frames", type=int, default=80)
    ap.add_argument("--ref-frames", type=int, default=8, help="reference arm: frames per step (bounded sample)")
    ap.add_argument("--c3-frames", type=int, default=120)
    ap.add_argument("--c3-cpu-frames", type=int, default=13)
    ap.add_argument("--c5-frames", type=int, default=200)
    ap.add_argument("--skip-extra", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_gpu(args)


if __name__ == "__main__":
    main()
