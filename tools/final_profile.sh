#!/bin/bash
# End-of-round evidence: GPU tests, the bench line, the ncu launch list of a short bench run and one --set full capture
# of a whole un-graphed step (per-launch DRAM traffic / pipe utilisation).  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${R:-r02}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${R}_gputests.txt
python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --batches 4 --skip-extra --skip-cpu > gpurun_out/${R}_bench_under_ncu.json 2>/dev/null
STEPS=2 ncu --set full --clock-control none --import-source on -s 36 -c 36 -o gpurun_out/${R}_full_step \
    python tools/profile_step.py > gpurun_out/${R}_full_step.log 2>&1
cat gpurun_out/${R}_gputests.txt; tail -c 300 gpurun_out/${R}_bench.err; ls -la gpurun_out/${R}_*
