#!/bin/bash
# host-side visit: native host timing under both schedules, stream counts and a few batch thresholds
set -u
OUT=gpurun_out/${1:-host}
mkdir -p "$OUT"
for extra in "--schedule=static" "--schedule=dynamic --n_streams=1" "--schedule=dynamic --n_streams=2" "--schedule=dynamic --n_streams=2 --min_batch=6" "--schedule=dynamic --n_streams=2 --min_batch=16"; do
  python scripts/host_bench.py --iters 400 "--extra=$extra --warmup_iters=50" > "$OUT/tmp.json" 2> "$OUT/tmp.err" || tail -3 "$OUT/tmp.err"
  python - "$OUT/tmp.json" "$extra" <<'PY'
import json,sys
try:
    s=json.loads(open(sys.argv[1]).readline())
    print(sys.argv[2], "| us/step", round(s["us_per_step"],1), "iters/s", round(s["mcmc_iters_per_s"]), "score", round(s["score_kernel_us_per_step"],1), "build", round(s["build_kernel_us_per_step"],1), "launches", s.get("launches"), "avg batch", round(s.get("launched_chains",0)/max(1,s.get("launches",1)),1))
except Exception as e:
    print(sys.argv[2], "FAILED", open(sys.argv[1]).read()[-600:])
PY
  cat "$OUT/tmp.json" >> "$OUT/host_runs.jsonl"
done
