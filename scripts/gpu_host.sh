#!/bin/bash
# host-side visit: native host timing under both schedules and a few batch thresholds
set -u
OUT=gpurun_out/${1:-host}
mkdir -p "$OUT"
for extra in "--schedule=static" "--schedule=dynamic" "--schedule=dynamic --min_batch=8" "--schedule=dynamic --min_batch=16" "--schedule=dynamic --min_batch=4"; do
  python scripts/host_bench.py --iters 400 --extra "$extra --warmup_iters=50" > "$OUT/tmp.json" 2> "$OUT/tmp.err" || tail -3 "$OUT/tmp.err"
  python - "$OUT/tmp.json" "$extra" <<'PY'
import json,sys
s=json.loads(open(sys.argv[1]).readline())
print(sys.argv[2], "| us/step", round(s["us_per_step"],1), "iters/s", round(s["mcmc_iters_per_s"]), "score", round(s["score_kernel_us_per_step"],1), "build", round(s["build_kernel_us_per_step"],1), "launches", s.get("launches"), "avg batch", round(s.get("launched_chains",0)/max(1,s.get("launches",1)),1))
PY
  cat "$OUT/tmp.json" >> "$OUT/host_runs.jsonl"
done
