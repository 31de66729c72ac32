#!/bin/bash
# The ncu evidence for profiles/: launch list of the bench command, then full captures (source-level) of the dominant
# kernels.  Each ncu pass runs only after the plain command has exited 0.  Usage (under gpurun): bash scripts/gpu_ncu.sh <tag>
set -u
TAG=${1:-ncu}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
SHORT="--steps 5 --warmup 3 --no_cpu_baseline --no_e2e"
python bench.py $SHORT > "$OUT/plain.json" 2> "$OUT/plain.err" || { echo "plain bench failed"; tail -5 "$OUT/plain.err"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file "$OUT/launches.csv" python bench.py $SHORT > "$OUT/ncu_launches.log" 2>&1
echo "ncu launches rc=$?"
# steady-state batched launches come after the set-up launches (K0 x3, transpose ..., 2 x chains single-chain launches)
ncu --set full --clock-control none --import-source on -k regex:score_proposals -s 70 -c 3 -o "$OUT/prof_score" python bench.py $SHORT > "$OUT/ncu_score.log" 2>&1
echo "ncu score rc=$?"
ncu --set full --clock-control none --import-source on -k regex:build_increments -s 70 -c 3 -o "$OUT/prof_k1b" python bench.py $SHORT > "$OUT/ncu_k1b.log" 2>&1
echo "ncu k1b rc=$?"
ncu --set full --clock-control none --import-source on -k regex:build_tables -s 70 -c 3 -o "$OUT/prof_k1a" python bench.py $SHORT > "$OUT/ncu_k1a.log" 2>&1
echo "ncu k1a rc=$?"
ncu --set full --clock-control none --import-source on -k regex:condense_rows -s 1 -c 2 -o "$OUT/prof_k0" python bench.py $SHORT > "$OUT/ncu_k0.log" 2>&1
echo "ncu k0 rc=$?"
ls -la "$OUT"
