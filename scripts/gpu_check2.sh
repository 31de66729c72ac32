#!/bin/bash
# Full GPU parity suite, smoke, then the driver's N=1 bench command (own arm only).
set -u
TAG=${1:-check2}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/tests.log" 2>&1; echo "tests rc=$?" | tee -a "$OUT/summary.txt"; tail -4 "$OUT/tests.log"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/smoke.log"
( time python bench.py --steps 20 --warmup 5 > "$OUT/bench_20.json" 2> "$OUT/bench_20.err" ) 2> "$OUT/time_20.txt"; echo "bench 20 rc=$?"; tail -3 "$OUT/bench_20.err"; grep real "$OUT/time_20.txt"
python - "$OUT" <<'PY'
import json,sys
o=sys.argv[1]
for f in ("bench_20",):
    try: j=json.loads(open(f"{o}/{f}.json").read().strip().splitlines()[-1])
    except Exception as e: print(f,"unreadable",e); continue
    print("==",f,"value %.3g"%j["value"],"ms/step %.4f"%j["ms_per_step"],"e2e %.3g"%j["e2e"]["value"], "e2e ms/step", j["e2e"].get("ms_per_step"), j["value_how"][-120:])
    for k in ("roofline","roofline_e2e","roofline_k1","roofline_k0"):
        if j.get(k): print("  ",k,"frac %.3f"%j[k]["frac"],"achieved %.0f"%j[k]["achieved"], {kk:j[k][kk] for kk in ("avg_launch_us","launches","kernel_us") if kk in j[k]})
    print("   e2e", {k:v for k,v in j["e2e"].items() if k!="what"})
    if j.get("other_configs"): print("   other", json.dumps(j["other_configs"])[:1500])
PY
