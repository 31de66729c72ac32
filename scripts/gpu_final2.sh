#!/bin/bash
# Final N=1 record: full GPU tests, smoke, ncu evidence (launch list + full captures of the four kernels; the DRAM bytes of
# the proposal kernel become profiles/r2_traffic.json before the bench reads it), then bench both arms at the driver's
# setting and at the default.  Usage (under gpurun): bash scripts/gpu_final2.sh <tag>
set -u
TAG=${1:-final2}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/tests.log" 2>&1; echo "tests rc=$?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/tests.log"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/smoke.log"
timeout 900 bash scripts/gpu_ncu.sh $TAG/ncu
for k in score k1a k1b k0; do
  python scripts/ncu_summary.py "$OUT/ncu/prof_$k.ncu-rep" --json "$OUT/ncu/ncu_$k.json" > "$OUT/ncu/ncu_$k.md" 2>&1
done
python scripts/make_traffic.py "$OUT/ncu/ncu_score.json" "$OUT/r2_traffic.json" "ncu --set full --clock-control none, steady-state launches of 32 proposals of \`bench.py --steps 5 --warmup 3 --no_cpu_baseline --no_e2e\` (profiles/r2_final_ncu_score.json): dram__bytes_read.sum + dram__bytes_write.sum" && cp "$OUT/r2_traffic.json" profiles/r2_traffic.json
rm -f "$OUT/ncu/prof_k1a.ncu-rep" "$OUT/ncu/prof_k1b.ncu-rep"   # the summaries travel; two reports are enough to bring back
timeout 900 bash scripts/gpu_bench1.sh $TAG
