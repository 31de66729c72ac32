#!/bin/bash
# Final N=1 record: full GPU tests, smoke, bench both arms at the driver's setting and at the default, ncu evidence.
set -u
TAG=${1:-final1}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/tests.log" 2>&1; echo "tests rc=$?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/tests.log"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/smoke.log"
timeout 900 bash scripts/gpu_bench1.sh $TAG
timeout 900 bash scripts/gpu_ncu.sh $TAG/ncu
