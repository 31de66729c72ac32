#!/bin/bash
# Cold-start study of the native host: short timed runs after short / long warm-up phases, with the launch timeline.
# Usage (under gpurun): bash scripts/gpu_cold.sh <tag>
set -u
TAG=${1:-cold}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
for spec in "20 5" "20 100" "200 20" "600 20"; do
  set -- $spec
  python scripts/host_bench.py --iters $1 --extra="--warmup_iters=$2 --timeline=1" > "$OUT/host_$1_$2.json" 2> "$OUT/host_$1_$2.err"
  echo "host $1/$2 rc=$?" | tee -a "$OUT/summary.txt"
done
python - "$OUT" <<'PY'
import json,sys,glob
for f in sorted(glob.glob(sys.argv[1]+"/host_*.json")):
    try: j=json.loads(open(f).read().splitlines()[0])
    except Exception as e: print(f,"unreadable",e); continue
    tl=j.pop("timeline",[])
    print(f, "us/step %.1f"%j["us_per_step"], "it/s %.0f"%j["mcmc_iters_per_s"], "score %.1f build %.1f"%(j["score_kernel_us_per_step"],j["build_kernel_us_per_step"]), "launches",j["launches"])
    if j["n_iters"]<=20: print("  timeline", [(round(a),int(b),int(c)) for a,b,c in tl])
PY
