#!/bin/bash
# bench.py at N GPUs exactly as the driver launches it.  Usage (under gpurun --gpus N): bash scripts/gpu_driver_n.sh <tag> <N>
set -u
TAG=${1:-drv}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29571 \
  bench.py --gpus "$N" --steps 20 --warmup 5 > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"; echo "bench rc=$?"
python - "$OUT/bench_n$N.json" <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e=j["e2e"]
print("n",j["n_gpus"],"value %.3g"%j["value"],"ms/step %.4f"%j["ms_per_step"],"| e2e %.3g"%e["value"],"it/s %.0f"%e["mcmc_iters_per_s"],"ms/step %.4f"%e["ms_per_step"],"equal",j.get("sharded_equals_single"))
for k in ("roofline","roofline_e2e","roofline_k1","roofline_k0"):
    if j.get(k): print("  ",k,"frac %.3f"%j[k]["frac"])
print("   other", json.dumps(j.get("other_configs"))[:1500])
PY
