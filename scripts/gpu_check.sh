#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), native-host timing, then the ncu passes of B200_PROFILING.md.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_check.sh <tag> [bench args...]
set -u
TAG=${1:-run}; shift || true
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > "$OUT/gpu.txt" 2>&1
nproc > "$OUT/nproc.txt"
if [ -z "${SKIP_TESTS:-}" ]; then
python -m pytest tests -m gpu -x -q > "$OUT/tests.log" 2>&1; echo "tests rc=$?" | tee -a "$OUT/summary.txt"
tail -3 "$OUT/tests.log"
python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/summary.txt"
fi
python bench.py "$@" > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$OUT/summary.txt"
cat "$OUT/bench.json"; tail -5 "$OUT/bench.err"
if [ -z "${SKIP_REF:-}" ]; then
python bench.py --impl reference "$@" > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; echo "bench_ref rc=$?" | tee -a "$OUT/summary.txt"
cat "$OUT/bench_ref.json"
fi
python scripts/host_bench.py --iters 300 --ref_iters ${REF_ITERS:-0} > "$OUT/host_bench.json" 2> "$OUT/host_bench.err"; echo "host_bench rc=$?" | tee -a "$OUT/summary.txt"
cat "$OUT/host_bench.json"; tail -3 "$OUT/host_bench.err"
if [ -z "${SKIP_NCU:-}" ]; then
SHORT="--steps 5 --warmup 3 --no_cpu_baseline --no_e2e"
python bench.py $SHORT > "$OUT/plain.log" 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches.csv" python bench.py $SHORT > "$OUT/ncu_launches.log" 2>&1
echo "ncu launches rc=$?" | tee -a "$OUT/summary.txt"
# the batched proposal launches come after the 2 x chains single-chain set-up launches: skip those
ncu --set full --clock-control none --import-source on -k regex:score_proposals -s 70 -c 3 -o "$OUT/prof" python bench.py $SHORT > "$OUT/ncu_full.log" 2>&1
echo "ncu full rc=$?" | tee -a "$OUT/summary.txt"
ncu --set full --clock-control none --import-source on -k regex:build_rows -s 36 -c 4 -o "$OUT/prof_build" python bench.py $SHORT > "$OUT/ncu_build.log" 2>&1
echo "ncu build rc=$?" | tee -a "$OUT/summary.txt"
fi
ls -la "$OUT"
