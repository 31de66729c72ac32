set -u
OUT=gpurun_out/r2_exp1; mkdir -p $OUT
run() { name=$1; shift; env "$@" python scripts/host_bench.py --iters 300 --extra="--warmup_iters=20 --n_streams=3 --min_batch=8" > $OUT/$name.json 2> $OUT/$name.err || echo "$name failed"; }
run tables1 SCICONE_TABLES=1
run tables1_nodirect SCICONE_TABLES=1 SCICONE_DIRECT_RESULTS=0
run auto_nodirect SCICONE_DIRECT_RESULTS=0
run auto_nosplit SCICONE_SPLIT_HEAVY=0
run tables0 SCICONE_TABLES=0
SCICONE_TABLES=1 python scripts/kernel_model.py > $OUT/model_tables1.txt 2>&1; cat $OUT/model_tables1.txt
python - "$OUT" <<'PY'
import json,sys,glob
for f in sorted(glob.glob(sys.argv[1]+"/*.json")):
    try: j=json.loads(open(f).read().splitlines()[0])
    except Exception as e: print(f,"unreadable",e); continue
    print(f.split('/')[-1], "us/step %.1f"%j["us_per_step"], "it/s %.0f"%j["mcmc_iters_per_s"], "score %.1f build %.1f"%(j["score_kernel_us_per_step"],j["build_kernel_us_per_step"]), "launches",j["launches"], "gpu_launches", j["gpu_launches"])
PY
