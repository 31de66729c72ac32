#!/bin/bash
# Launch timeline of short timed windows of the native host (configs[2], one GPU): when is each launch submitted and
# collected, how many proposals does it carry, when does each chain finish its 20 iterations.
# Usage (under gpurun): bash scripts/gpu_timeline.sh <tag> [extra host flags]
set -u
TAG=${1:-tl}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"; shift
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_od_hist.py -m gpu -x -q > "$OUT/tests_some.log" 2>&1; echo "some tests rc=$?"; tail -3 "$OUT/tests_some.log"
for rep in 1 2 3; do
  timeout 120 python scripts/host_bench.py --iters 20 --extra="--warmup_iters=100 --profile_kernels=0 --timeline=1 $*" > "$OUT/tl20_$rep.json" 2> "$OUT/tl20_$rep.err" || echo "run $rep failed"
done
timeout 120 python scripts/host_bench.py --iters 100 --extra="--warmup_iters=100 --profile_kernels=0 --timeline=1 $*" > "$OUT/tl100.json" 2> "$OUT/tl100.err"
python - "$OUT" <<'PY'
import json,sys,glob
import numpy as np
for f in sorted(glob.glob(sys.argv[1]+"/tl*.json")):
    j=json.loads(open(f).read().splitlines()[0])
    tl=np.array(j["timeline"])
    sub=tl[tl[:,1]==0]; col=tl[tl[:,1]==1]; done=tl[tl[:,1]==2]
    print(f.split('/')[-1], "us/step %.1f wall %.0f launches %d"%(j["us_per_step"], j["wall_us"], len(sub)))
    print("  first submits:", [(round(t),int(n)) for t,_,n in sub[:8]])
    print("  last submits :", [(round(t),int(n)) for t,_,n in sub[-10:]])
    print("  last collects:", [(round(t),int(n)) for t,_,n in col[-6:]])
    d=np.sort(done[:,0]); print("  chains done at: min %.0f q25 %.0f med %.0f q75 %.0f max %.0f"%(d[0],d[len(d)//4],d[len(d)//2],d[3*len(d)//4],d[-1]))
    sizes=sub[:,2]; print("  launch sizes: mean %.1f ; in the last 25%% of the window mean %.1f"%(sizes.mean(), sub[sub[:,0]>0.75*j["wall_us"]][:,2].mean()))
PY
