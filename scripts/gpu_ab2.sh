#!/bin/bash
# Parity suite, launch-shape table, and the native host at 20 / 300 iterations (3 repetitions each) on one prepared input.
set -u
TAG=${1:-ab2}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/tests.log" 2>&1; echo "tests rc=$?" | tee -a "$OUT/summary.txt"; tail -4 "$OUT/tests.log"
timeout 200 python scripts/kernel_model.py > "$OUT/model.txt" 2>&1; cat "$OUT/model.txt"
timeout 100 python scripts/nu_probe.py | tee "$OUT/probe.txt"
timeout 300 python scripts/host_sweep.py --iters 20,300 --reps 3 --env SCICONE_THREADS=12 --variants "n_streams=2" | tee "$OUT/sweep.jsonl"
