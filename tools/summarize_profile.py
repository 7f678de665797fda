"""Turns an ncu launch list (--metrics gpu__time_duration.sum --csv) and a bench.py JSON line into a markdown summary
for profiles/.  Usage: python tools/summarize_profile.py launches.csv bench.json [steps_in_capture] > profiles/xxx.md"""
import collections
import csv
import json
import re
import sys


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    return agg, n


def main():
    agg, n = launches(sys.argv[1])
    bench = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    tot = sum(a[1] for a in agg.values())
    print(f"## ncu launch list ({n} launches over {steps} step(s); cold-cache, serialised: compare SHARES)\n")
    print("| kernel | launches/step | us/step | share |\n|---|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {a[0] / steps:.0f} | {a[1] / steps:.1f} | {100 * a[1] / tot:.1f}% |")
    print(f"\ntotal {tot / steps:.1f} us/step under ncu\n")
    print("## bench.py line (CUDA events, no profiler)\n")
    print(f"* value {bench['value']:.1f} {bench['unit']} ({bench['ms_per_step']:.3f} ms/step, "
          f"{bench['config']['frames_per_step']} frames/step), e2e {bench['e2e']['value']:.1f} {bench['unit']} "
          f"({bench['e2e'].get('ms_per_step', 0):.3f} ms/step), launches/timed region {bench['gpu_launches']}")
    print(f"* clocks {bench['clocks']}")
    r = bench["roofline"]
    tr = r.get("traffic")
    print(f"* roofline: `{r['kernel']}` {r['bound']}-bound, {r['achieved']:.1f} {r['unit']} of {r['peak']} "
          f"({100 * r['frac']:.1f}%), {r['ms_per_launch'] * 1e3:.1f} us/launch, share of step {100 * r['share_of_step']:.1f}%, "
          f"algorithmic {r['algorithmic_bytes'] / 1e6:.1f} MB, DRAM traffic (ncu) {('%.1f MB' % (tr / 1e6)) if tr else 'n/a'}")
    print("\n| kernel (event-timed, un-graphed step) | ms | share | HBM frac | tensor frac |\n|---|---:|---:|---:|---:|")
    for k in bench["extra"]["kernels"]:
        print(f"| {k['name']} | {k['ms']:.4f} | {100 * k['share']:.1f}% | {100 * k['hbm_frac']:.1f}% | {100 * k['tensor_frac']:.1f}% |")
    for key in ("loopdb", "lba"):
        if key in bench["extra"]:
            print(f"\n* {key}: {bench['extra'][key]}")
    if bench.get("cpu_baseline"):
        print(f"\n* cpu_baseline: {bench['cpu_baseline']}")


if __name__ == "__main__":
    main()
