"""Summarise an ncu --set full report here (no GPU): per kernel launch the headline metrics, the stall mix and the
hottest source lines (needs -lineinfo).  Usage: python tools/ncu_hot.py report.ncu-rep [launch_index]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_lsu.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_shared_mem"]
for li, r in enumerate(rows[2:]):
    if which is not None and li != which:
        continue
    print(f"=== launch {li}: {r[idx['Kernel Name']][:60]}")
    for w in want:
        if w in idx:
            print(f"   {w}: {r[idx[w]]} {units[idx[w]]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(li), "--launch-count", "1",
                          "--print-source", "cuda,sass"] , capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    h = None
    for k, rr in enumerate(srows):
        if "Source" in rr and "# Samples" in rr:
            h, start = rr, k + 1
            break
    if not h:
        continue
    isamp, isrc = h.index("# Samples"), h.index("Source")
    stall = collections.Counter()
    lines = []
    for rr in srows[start:]:
        if len(rr) != len(h):
            continue
        try:
            s = int(rr[isamp])
        except ValueError:
            continue
        lines.append((s, rr[isrc].strip()))
        for i, name in enumerate(h):
            if name.startswith("stall_") and "Not Issued" not in name:
                try:
                    stall[name] += int(rr[i])
                except ValueError:
                    pass
    tot = sum(s for s, _ in lines) or 1
    print("   stalls:", ", ".join(f"{k[6:]} {100 * v / sum(stall.values()):.0f}%" for k, v in stall.most_common(7)))
    for s, t in sorted(lines, key=lambda x: -x[0])[:22]:
        print(f"   {100 * s / tot:5.1f}%  {t[:110]}")
