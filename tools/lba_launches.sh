# ncu launch list of the local-BA / pose-optimisation kernels (tools/ba_time.py), averaged per kernel
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lba_launches.csv python tools/ba_time.py > /dev/null 2>&1
python - <<'P'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/lba_launches.csv')) if len(r) > 5]
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, start = r, i
        break
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[start + 2:]:
    try:
        agg.setdefault(r[ki].split('(')[0], []).append(float(r[vi].replace(',', '')))
    except ValueError:
        pass
for k, v in agg.items():
    print(f"{k:28s} n={len(v):3d} avg {sum(v) / len(v) / 1000:7.1f} us   last {[round(x / 1000, 1) for x in v[-4:]]}")
P
