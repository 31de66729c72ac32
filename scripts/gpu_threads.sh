#!/bin/bash
# Host threads of the native host on one GPU: 20- and 300-iteration windows per thread count (spinning workers vs cores).
set -u
TAG=${1:-threads}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nproc | tee "$OUT/nproc.txt"
for t in ${2:-16 14 12 10 8}; do
  echo "threads $t"
  timeout 300 python scripts/host_sweep.py --iters 20,300 --reps 4 --env SCICONE_THREADS=$t --variants "n_streams=2,min_batch=4" | tee -a "$OUT/threads_$t.jsonl"
done
