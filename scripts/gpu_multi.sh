#!/bin/bash
# One visit of an N-GPU box: the 2-GPU parity test, then bench.py at N GPUs with the dynamic (default) and the static
# schedule of the native host.  Usage (under gpurun --gpus N):  bash scripts/gpu_multi.sh <tag> <N> [bench args...]
set -u
TAG=${1:-multi}; N=${2:-2}; shift; shift || true
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name --format=csv > "$OUT/gpu.txt" 2>&1; nproc > "$OUT/nproc.txt"
if [ -z "${SKIP_TESTS:-}" ]; then
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > "$OUT/tests.log" 2>&1; echo "tests rc=$?" | tee -a "$OUT/summary.txt"
tail -3 "$OUT/tests.log"
fi
run() {  # name, extra bench args
  local name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29571 \
    bench.py --gpus "$N" "$@" > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
  echo "bench $name rc=$?" | tee -a "$OUT/summary.txt"
  cat "$OUT/bench_$name.json"; tail -3 "$OUT/bench_$name.err"
}
run dynamic "$@"
if [ -z "${SKIP_STATIC:-}" ]; then run static --n_groups 2 "$@"; fi
if [ -n "${NCCL_TOO:-}" ]; then SCICONE_EXCHANGE=nccl run dynamic_nccl "$@"; fi
if [ -n "${ONE_LANE:-}" ]; then run dynamic_1lane --host_args=--n_streams=1 "$@"; fi
if [ -n "${MIN_BATCH:-}" ]; then run dynamic_mb$MIN_BATCH --host_args=--min_batch=$MIN_BATCH "$@"; fi
ls -la "$OUT"
