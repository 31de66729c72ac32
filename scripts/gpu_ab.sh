#!/bin/bash
# A/B of the engine switches on one GPU: parity tests first, then the launch-shape table and the native host (real MCMC,
# configs[2]) with each switch off in turn.  Usage (under gpurun): bash scripts/gpu_ab.sh <tag>
set -u
TAG=${1:-ab}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 300 python -m pytest tests/test_gpu_od_hist.py -m gpu -x -q > "$OUT/tests_new.log" 2>&1; echo "new tests rc=$?" | tee -a "$OUT/summary.txt"; tail -4 "$OUT/tests_new.log"
if [ -z "${SKIP_TESTS:-}" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/tests.log" 2>&1; echo "tests rc=$?" | tee -a "$OUT/summary.txt"; tail -4 "$OUT/tests.log"
fi
timeout 200 python scripts/kernel_model.py > "$OUT/model.txt" 2>&1; cat "$OUT/model.txt"
SCICONE_OD_HIST=0 timeout 200 python scripts/kernel_model.py > "$OUT/model_nohist.txt" 2>&1; cat "$OUT/model_nohist.txt"
run() {  # name, env assignments...
  name=$1; shift
  for rep in 1 2; do
    env "$@" timeout 120 python scripts/host_bench.py --iters 300 --extra="--warmup_iters=100 --profile_kernels=0" > "$OUT/host300_${name}_$rep.json" 2> "$OUT/host300_${name}_$rep.err" || echo "host $name failed"
  done
  env "$@" timeout 120 python scripts/host_bench.py --iters 20 --extra="--warmup_iters=100 --profile_kernels=0" > "$OUT/host20_${name}.json" 2> "$OUT/host20_${name}.err" || echo "host20 $name failed"
}
run default SCICONE_X=1
run nopdl SCICONE_PDL=0
run nohist SCICONE_OD_HIST=0
run neither SCICONE_PDL=0 SCICONE_OD_HIST=0
env timeout 120 python scripts/host_bench.py --iters 300 --extra="--warmup_iters=100 --profile_kernels=1" > "$OUT/host300_prof.json" 2> "$OUT/host300_prof.err"
python - "$OUT" <<'PY'
import json,sys,glob
for f in sorted(glob.glob(sys.argv[1]+"/host*.json")):
    try: j=json.loads(open(f).read().splitlines()[0])
    except Exception as e: print(f,"unreadable",e); continue
    print(f.split('/')[-1], "us/step %.1f"%j["us_per_step"], "it/s %.0f"%j["mcmc_iters_per_s"], "score %.1f build %.1f"%(j["score_kernel_us_per_step"],j["build_kernel_us_per_step"]), "launches",j["launches"], "gpu_launches", j["gpu_launches"])
PY
