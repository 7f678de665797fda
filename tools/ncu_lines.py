"""Per-CUDA-source-line totals (warp instructions executed, stall samples) of one launch in an ncu --set full report.
Usage: python tools/ncu_lines.py report.ncu-rep launch_index [top_n]"""
import csv
import subprocess
import sys

rep, li = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(li), "--launch-count", "1",
                      "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = None
fname = ""
recs = []
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if "Source" in r and "# Samples" in r:
        h = r
        isamp, iex = h.index("# Samples"), h.index("Instructions Executed")
        continue
    if h and len(r) == len(h) and r[0].isdigit():
        try:
            recs.append((fname, int(r[0]), r[1].strip(), int(r[isamp]), int(r[iex])))
        except ValueError:
            pass
ts = sum(x[3] for x in recs) or 1
te = sum(x[4] for x in recs) or 1
print(f"total samples {ts}, warp instructions {te}")
print("--- by instructions executed")
for f, ln, src, s, e in sorted(recs, key=lambda x: -x[4])[:top]:
    print(f"{100 * e / te:5.1f}% inst {100 * s / ts:5.1f}% samp  {f}:{ln}  {src[:100]}")
print("--- by stall samples")
for f, ln, src, s, e in sorted(recs, key=lambda x: -x[3])[:top]:
    print(f"{100 * s / ts:5.1f}% samp {100 * e / te:5.1f}% inst  {f}:{ln}  {src[:100]}")
