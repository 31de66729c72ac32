#!/bin/bash
# Slow host visits of the native host (which move, how long), then a sweep of lanes x min_batch on one prepared input.
set -u
TAG=${1:-sw2}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
SCICONE_DEBUG_VISITS=1 timeout 120 python scripts/host_bench.py --iters 100 --extra="--warmup_iters=100 --profile_kernels=0" > "$OUT/visits.json" 2> "$OUT/visits.err"
grep "^visit" "$OUT/visits.err" | awk '{m[$7]++; s[$7]+=$9; if ($9>mx[$7]) mx[$7]=$9} END{for (k in m) print "move",k,"visits>40us",m[k],"mean",s[k]/m[k],"max_us",mx[k]}'
grep "^visit" "$OUT/visits.err" | sort -k9 -n -r | head -12
timeout 500 python scripts/host_sweep.py --iters 20,300 --reps 3 --variants "${2:-n_streams=2,min_batch=8;n_streams=2,min_batch=4;n_streams=2,min_batch=2;n_streams=2,min_batch=1;n_streams=3,min_batch=4;n_streams=3,min_batch=2;n_streams=3,min_batch=1;n_streams=4,min_batch=2;n_streams=4,min_batch=1}" | tee "$OUT/sweep.jsonl"
