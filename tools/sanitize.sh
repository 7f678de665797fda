#!/bin/bash
# compute-sanitizer passes (memcheck / racecheck / synccheck) over the un-graphed hot path at small sizes.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/sanitizer.txt
: > $out
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  echo "=== compute-sanitizer --tool $tool: tools/cpl_check.py CH=128 CW=160 BATCH=2" >> $out
  HFB_NO_GRAPH=1 CH=128 CW=160 BATCH=2 timeout ${T:-600} compute-sanitizer --tool $tool python tools/cpl_check.py 2>&1 | grep -E "CPL_CHECK|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error" | sort | uniq -c | sort -rn | head -12 >> $out
  echo "=== compute-sanitizer --tool $tool: __graft_entry__.py --smoke" >> $out
  HFB_NO_GRAPH=1 timeout ${T:-600} compute-sanitizer --tool $tool python __graft_entry__.py --smoke 2>&1 | grep -E "smoke ok|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error" | sort | uniq -c | sort -rn | head -12 >> $out
done
cat $out
