#!/bin/bash
# Parity subset + launch-shape table + ncu captures (full set, source-level) of the proposal kernel in the launch-shape script.
# Usage (under gpurun): bash scripts/gpu_prof.sh <tag>
set -u
TAG=${1:-prof}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
python -m pytest tests/test_gpu_parity.py tests/test_gpu_random.py tests/test_gpu_fullsize.py -m gpu -x -q > "$OUT/tests.log" 2>&1; echo "tests rc=$?" | tee -a "$OUT/summary.txt"; tail -5 "$OUT/tests.log"
python scripts/kernel_model.py > "$OUT/model.txt" 2>&1; cat "$OUT/model.txt"
for ns in 2 3; do
  python scripts/host_bench.py --iters 300 --extra="--warmup_iters=20 --n_streams=$ns" > "$OUT/host300_s$ns.json" 2> "$OUT/host300_s$ns.err"
done
python - "$OUT" <<'PY'
import json,sys,glob
for f in sorted(glob.glob(sys.argv[1]+"/host*.json")):
    try: j=json.loads(open(f).read().splitlines()[0])
    except Exception as e: print(f,"unreadable",e); continue
    print(f, "us/step %.1f"%j["us_per_step"], "it/s %.0f"%j["mcmc_iters_per_s"], "score %.1f build %.1f"%(j["score_kernel_us_per_step"],j["build_kernel_us_per_step"]), "launches",j["launches"], "gpu_launches", j["gpu_launches"])
PY
if [ -z "${SKIP_NCU:-}" ]; then
ncu --set full --clock-control none --import-source on -k regex:score_proposals -s 33 -c 2 -o "$OUT/prof_light32" python scripts/kernel_model.py > "$OUT/ncu1.log" 2>&1; echo "ncu light32 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:score_proposals -s 82 -c 2 -o "$OUT/prof_light1" python scripts/kernel_model.py > "$OUT/ncu2.log" 2>&1; echo "ncu light1 rc=$?"
fi
ls -la "$OUT"
