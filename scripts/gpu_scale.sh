#!/bin/bash
# One visit of an N-GPU box: the 2-GPU parity tests, then bench.py at N GPUs as the driver launches it (20 steps + 5 warm-up,
# everything on) and a longer run without the side records.  Usage (under gpurun --gpus N): bash scripts/gpu_scale.sh <tag> <N>
set -u
TAG=${1:-scale}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi --query-gpu=name --format=csv > "$OUT/gpu.txt" 2>&1; nproc > "$OUT/nproc.txt"
if [ -z "${SKIP_TESTS:-}" ]; then
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > "$OUT/tests.log" 2>&1; echo "tests rc=$?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/tests.log"
fi
run() {  # name, bench args
  local name=$1; shift
  ( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29571 \
    bench.py --gpus "$N" "$@" > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err" ) 2> "$OUT/time_$name.txt"
  echo "bench $name rc=$?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/bench_$name.err"; grep real "$OUT/time_$name.txt"
}
run driver --steps 20 --warmup 5
run long --steps 200 --warmup 20 --no_other_configs --host_args=--timeline=1
if [ -n "${MORE:-}" ]; then run long_s3 --steps 200 --warmup 20 --no_other_configs --host_args=--n_streams=3; fi
python - "$OUT" <<'PY'
import json,sys
o=sys.argv[1]
for f in ("bench_driver","bench_long","bench_long_s3"):
    try: j=json.loads(open(f"{o}/{f}.json").read().strip().splitlines()[-1])
    except Exception as e: print(f,"unreadable",e); continue
    e=j["e2e"]
    print("==",f,"n",j["n_gpus"],"value %.3g"%j["value"],"ms/step %.4f"%j["ms_per_step"],"| e2e %.3g"%e["value"],"it/s %.0f"%e["mcmc_iters_per_s"],"ms/step %.4f"%e["ms_per_step"],
          "score %.1f build %.1f us/step"%(e["score_kernel_us_per_step"],e["build_kernel_us_per_step"]),"launches",e.get("launches"),"equal",j.get("sharded_equals_single"))
    for k in ("roofline","roofline_e2e","roofline_k1","roofline_k0"):
        if j.get(k): print("  ",k,"frac %.3f"%j[k]["frac"], {kk:j[k][kk] for kk in ("avg_launch_us","launches","kernel_us") if kk in j[k]})
    if j.get("other_configs"): print("   other", json.dumps(j["other_configs"])[:1800])
PY
