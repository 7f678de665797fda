# loop-DB end-to-end queries/s at 1 and 2 GPUs (bench.py's loopdb block only)
python - <<'P'
import json, subprocess, sys
for n in (1, 2):
    cmd = ([sys.executable] if n == 1 else [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29533"]) + ["bench.py", "--gpus", str(n), "--steps", "3", "--warmup", "3",
           "--batches", "4", "--skip-cpu", "--c3-frames", "6"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        l = d["extra"]["loopdb"]
        print(n, "gpus: q/s", round(l["queries_per_sec_e2e"]), "scan_ms", round(l["scan_ms"], 4), "us/query", round(1e6 / l["queries_per_sec_e2e"], 1))
    except Exception as ex:
        print(n, "failed", ex, r.stderr[-1500:])
P
