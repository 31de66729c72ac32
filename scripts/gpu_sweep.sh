#!/bin/bash
# Parity tests, launch-shape table, then a sweep of the native host's schedule parameters (lanes x min_batch).
# Usage (under gpurun): bash scripts/gpu_sweep.sh <tag> ["ns:mb ns:mb ..."]
set -u
TAG=${1:-sweep}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
SPECS=${2:-"2:8 3:8 3:6 4:6 4:4 4:3"}
if [ -z "${SKIP_TESTS:-}" ]; then
python -m pytest tests -m gpu -x -q > "$OUT/tests.log" 2>&1; echo "tests rc=$?" | tee -a "$OUT/summary.txt"; tail -4 "$OUT/tests.log"
fi
python scripts/kernel_model.py > "$OUT/model.txt" 2>&1; cat "$OUT/model.txt"
for spec in $SPECS; do
  ns=${spec%%:*}; mb=${spec##*:}
  python scripts/host_bench.py --iters 300 --extra="--warmup_iters=20 --n_streams=$ns --min_batch=$mb ${HOST_EXTRA:-}" > "$OUT/host300_s${ns}_b${mb}.json" 2> "$OUT/host300_s${ns}_b${mb}.err" || echo "host $spec failed"
done
python scripts/host_bench.py --iters 20 --extra="--warmup_iters=5 ${HOST_EXTRA:-}" > "$OUT/host20.json" 2> "$OUT/host20.err"
python scripts/host_bench.py --iters 300 --extra="--warmup_iters=20 --profile_kernels=0 ${HOST_EXTRA:-}" > "$OUT/host300_noprof.json" 2> "$OUT/host300_noprof.err"
python - "$OUT" <<'PY'
import json,sys,glob
for f in sorted(glob.glob(sys.argv[1]+"/host*.json")):
    try: j=json.loads(open(f).read().splitlines()[0])
    except Exception as e: print(f,"unreadable",e); continue
    print(f.split('/')[-1], "us/step %.1f"%j["us_per_step"], "it/s %.0f"%j["mcmc_iters_per_s"], "score %.1f build %.1f"%(j["score_kernel_us_per_step"],j["build_kernel_us_per_step"]), "launches",j["launches"], "gpu_launches", j["gpu_launches"])
PY
