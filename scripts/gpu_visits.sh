#!/bin/bash
set -u
TAG=${1:-visits}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
for rep in 1 2; do
SCICONE_DEBUG_VISITS=1 timeout 120 python scripts/host_bench.py --iters 100 --extra="--warmup_iters=100 --profile_kernels=0 --min_batch=4" > "$OUT/visits_$rep.json" 2> "$OUT/visits_$rep.err"
grep "^visit" "$OUT/visits_$rep.err" | awk '{m[$7]++; s[$7]+=$9; if ($9>mx[$7]) mx[$7]=$9} END{for (k in m) print "move",k,"visits>40us",m[k],"mean",s[k]/m[k],"max_us",mx[k]}'
grep "^visit" "$OUT/visits_$rep.err" | sort -k9 -n -r | head -8
done
