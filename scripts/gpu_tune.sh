#!/bin/bash
# Kernel tuning visit: occupancy sweep of the proposal kernel, then ncu of steady-state build / proposal launches.
set -u
TAG=${1:-tune}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
ARGS="--steps 100 --warmup 10 --no_cpu_baseline --no_e2e"
for occ in 5 6 7; do
  SCICONE_OCC=$occ python bench.py $ARGS > "$OUT/occ$occ.json" 2> "$OUT/occ$occ.err"
  python - "$OUT/occ$occ.json" $occ <<'PY'
import json,sys
b=json.load(open(sys.argv[1]))
print("occ",sys.argv[2],"ms_per_step",round(b["ms_per_step"],4),"score_us",round(b["roofline"]["avg_launch_us"],2),"frac",round(b["roofline"]["frac"],3))
PY
done
SHORT="--steps 6 --warmup 3 --no_cpu_baseline --no_e2e"
ncu --set full --clock-control none --import-source on -k regex:build_rows -s 34 -c 6 -o "$OUT/prof_build" python bench.py $SHORT > "$OUT/ncu_build.log" 2>&1
echo "ncu build rc=$?"
ncu --set full --clock-control none --import-source on -k regex:score_proposals -s 70 -c 3 -o "$OUT/prof" python bench.py $SHORT > "$OUT/ncu_full.log" 2>&1
echo "ncu score rc=$?"
ls -la "$OUT"
