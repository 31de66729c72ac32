#!/bin/bash
# The driver's N=1 commands: bench.py (both arms) at the driver's short setting and at the default one.
set -u
TAG=${1:-bench1}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
( time python bench.py --steps 20 --warmup 5 > "$OUT/bench_20.json" 2> "$OUT/bench_20.err" ) 2> "$OUT/time_20.txt"; echo "bench 20 rc=$?"; tail -3 "$OUT/bench_20.err"; cat "$OUT/time_20.txt" | grep real
python bench.py --impl reference --steps 20 --warmup 5 > "$OUT/ref_20.json" 2> "$OUT/ref_20.err"; echo "ref 20 rc=$?"
( time python bench.py > "$OUT/bench_default.json" 2> "$OUT/bench_default.err" ) 2> "$OUT/time_default.txt"; echo "bench default rc=$?"; tail -3 "$OUT/bench_default.err"; cat "$OUT/time_default.txt" | grep real
python - "$OUT" <<'PY'
import json,sys
o=sys.argv[1]
for f in ("bench_20","bench_default","ref_20"):
    try: j=json.loads(open(f"{o}/{f}.json").read().strip().splitlines()[-1])
    except Exception as e: print(f,"unreadable",e); continue
    print("==",f,"value %.3g"%j["value"],"ms/step %.4f"%j["ms_per_step"],"e2e %.3g"%j["e2e"]["value"], "e2e ms/step", j["e2e"].get("ms_per_step"))
    for k in ("roofline","roofline_e2e","roofline_k1","roofline_k0"):
        if j.get(k): print("  ",k,"frac %.3f"%j[k]["frac"],"achieved %.0f"%j[k]["achieved"], {kk:j[k][kk] for kk in ("avg_launch_us","launches","kernel_us") if kk in j[k]})
    if j.get("other_configs"): print("   other", json.dumps(j["other_configs"])[:1500])
    if j.get("cpu_baseline"): print("   cpu", j["cpu_baseline"]["value"], j["cpu_baseline"]["cores"])
PY
