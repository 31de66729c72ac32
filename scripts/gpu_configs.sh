#!/bin/bash
# native-host timing on the other BASELINE.json shapes (they are parity-test cases, not bench lines): configs[1], configs[4]
set -u
OUT=gpurun_out/${1:-cfg}
mkdir -p "$OUT"
run() {
  python scripts/host_bench.py "$@" > "$OUT/tmp.json" 2> "$OUT/tmp.err" || tail -3 "$OUT/tmp.err"
  python - "$OUT/tmp.json" "$*" <<'PY'
import json,sys
for line in open(sys.argv[1]):
    try: s=json.loads(line)
    except Exception: print("FAILED", line[:300]); continue
    if s.get("impl")=="native_host":
        print(sys.argv[2], "| native: us/step", round(s["us_per_step"],1), "iters/s", round(s["mcmc_iters_per_s"]), "score", round(s["score_kernel_us_per_step"],1), "build", round(s["build_kernel_us_per_step"],1))
    else:
        print("   reference 1 thread: iters/s", round(s["mcmc_iters_per_s"],1))
PY
  cat "$OUT/tmp.json" >> "$OUT/config_runs.jsonl"
}
run --cells 1000 --bins 10000 --regions 50 --nodes 10 --chains 32 --iters 2000 --ref_iters 400 "--extra=--warmup_iters=100"
run --cells 1000 --bins 10000 --regions 50 --nodes 30 --chains 32 --iters 2000 --ref_iters 200 --max_scoring false "--extra=--warmup_iters=100 --move_probs=0,1,0,1,0,1,0,1,0,1,0.01,0,0,1,0.01"
run --cells 100000 --bins 20000 --regions 200 --nodes 50 --chains 16 --iters 100 --ref_iters 4 "--extra=--warmup_iters=10"
run --cells 10000 --bins 18000 --regions 200 --nodes 200 --chains 32 --iters 200 --ref_iters 20 "--extra=--warmup_iters=20"
