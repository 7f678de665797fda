"""Per-launch DRAM traffic and pipe utilisation from an `ncu --set full` report of one bench step (no GPU needed here).

Writes profiles/ncu_traffic.json ({bench kernel label: {"dram_bytes_per_launch": ...}}, read by bench.py for
roofline.traffic) and prints a markdown table.  Launches are matched to bench.py's labels by kernel name and order:
the n-th fused_block_cpl launch of a step is layer 3 + n, the first gemm_tc_kernel<EpiStore> is the 3x3 head conv.

Usage: python tools/ncu_traffic.py report.ncu-rep [--json profiles/ncu_traffic.json]"""
import csv
import json
import re
import subprocess
import sys

rep = sys.argv[1]
out_json = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, name, scale_units=True):
    if name not in ix:
        return None
    try:
        v = float(r[ix[name]].replace(",", ""))
    except ValueError:
        return None
    u = units[ix[name]]
    if scale_units:
        v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
    return v


labels = {}
counts = {}
table = []
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
    ckey = "fused_block_cpl_kernel" if name.startswith("fused_block_cpl_kernel") else name
    n = counts.get(ckey, 0)
    counts[ckey] = n + 1
    label = None
    if name.startswith("fused_block_cpl_kernel"):
        label = f"l{3 + n}.fused"
    elif name.startswith("gemm_tc_kernel<EpiStore>"):
        label = ["head.conv3x3", "head.desc1x1+l2norm"][n] if n < 2 else None
    elif name.startswith("gemm_tc_kernel<EpiL2Norm>"):
        label = "head.desc1x1+l2norm"
    elif name.startswith("gemm_tc_kernel<EpiSoftmaxD2S>"):
        label = "head.det1x1+softmax+d2s"
    elif name.startswith("gemm_tc_kernel<EpiArgmax>"):
        label = "match.argmax"
    else:
        label = {"nms_kernel": "nms", "conv1_kernel": "conv1", "dw_project_small_kernel<24, 16>": "l2.dw+project",
                 "fc_mma_kernel": "global.fc", "vlad_kernel": "global.netvlad", "select_topk_kernel": "select_topk",
                 "stem_kernel": "stem.conv1+l2"}.get(name, name)
    rd, wr = val(r, "dram__bytes_read.sum") or 0.0, val(r, "dram__bytes_write.sum") or 0.0
    rec = {"kernel": name, "label": label, "ncu_us": val(r, "gpu__time_duration.sum"),
           "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
           "dram_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False),
           "tensor_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", False),
           "issue_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active", False),
           "regs": val(r, "launch__registers_per_thread", False), "grid": val(r, "launch__grid_size", False)}
    table.append(rec)
    if label and label not in labels:
        labels[label] = rec
print("| launch | label | ncu us | DRAM MB (r+w) | DRAM % | tensor % | issue % | regs | grid |\n|---|---|---:|---:|---:|---:|---:|---:|---:|")
f = lambda v, p=1: "-" if v is None else f"{v:.{p}f}"
for t in table:
    print(f"| `{t['kernel'][:44]}` | {t['label']} | {f(t['ncu_us'])} | {f(t['dram_bytes_per_launch'] / 1e6, 2)} | {f(t['dram_pct'])} | "
          f"{f(t['tensor_pct'])} | {f(t['issue_pct'])} | {f(t['regs'], 0)} | {f(t['grid'], 0)} |")
if out_json:
    with open(out_json, "w") as fh:
        json.dump(labels, fh, indent=1)
    print(f"\nwrote {out_json} ({len(labels)} labels)")
