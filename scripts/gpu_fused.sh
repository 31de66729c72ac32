#!/bin/bash
# One GPU-box visit after a kernel change: parity tests, the launch-shape table, native-host timings per lane count.
# Usage (under gpurun): bash scripts/gpu_fused.sh <tag>
set -u
TAG=${1:-fused}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
python -m pytest tests -m gpu -x -q > "$OUT/tests.log" 2>&1; echo "tests rc=$?" | tee -a "$OUT/summary.txt"; tail -15 "$OUT/tests.log"
python scripts/kernel_model.py > "$OUT/model.txt" 2>&1; cat "$OUT/model.txt"
for ns in 2 3 4; do
  python scripts/host_bench.py --iters 300 --extra="--warmup_iters=20 --n_streams=$ns" > "$OUT/host300_s$ns.json" 2> "$OUT/host300_s$ns.err"
  echo "host300 streams=$ns rc=$?" | tee -a "$OUT/summary.txt"
done
python scripts/host_bench.py --iters 20 --extra="--warmup_iters=5" > "$OUT/host20.json" 2> "$OUT/host20.err"
python - "$OUT" <<'PY'
import json,sys,glob
for f in sorted(glob.glob(sys.argv[1]+"/host*.json")):
    try: j=json.loads(open(f).read().splitlines()[0])
    except Exception as e: print(f,"unreadable",e); continue
    print(f, "us/step %.1f"%j["us_per_step"], "it/s %.0f"%j["mcmc_iters_per_s"], "score %.1f build %.1f"%(j["score_kernel_us_per_step"],j["build_kernel_us_per_step"]), "launches",j["launches"], "gpu_launches", j["gpu_launches"])
PY
