#!/bin/bash
# Heap checking of the host side of libscicone_score.so + the native host on a GPU box: both built with AddressSanitizer
# into a scratch directory, then the given inference command lines are run.  Usage: bash scripts/gpu_asan.sh <tag>
set -u
TAG=${1:-asan}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"; W=/tmp/asan_build; rm -rf $W; mkdir -p $W
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O1 -g -fmad=false -std=c++17 -Xcompiler -fPIC -Xcompiler -fsanitize=address -Xcompiler -fno-omit-frame-pointer \
  -shared -o $W/libscicone_score.so scicone_b200/csrc/scicone_score.cu -ldl 2>&1 | grep -v "warning\|^$\|Remark\|\^\|mapped_total" | tail -5
g++ -std=c++17 -O1 -g -fsanitize=address -fno-omit-frame-pointer -ffp-contract=off -Wno-sign-compare scicone_b200/host/inference_main.cpp -L$W -lscicone_score -Wl,-rpath,$W -o $W/inference 2>&1 | tail -3
python - <<'PY'
import sys
sys.path.insert(0,'.')
import numpy as np
from scicone_b200.synth import make_counts
data = make_counts(5000, 9000, 60, seed=31)
np.savetxt("/tmp/asan_build/d.csv", data["D"], delimiter=",", fmt="%.0f"); np.savetxt("/tmp/asan_build/r.txt", data["region_sizes"], fmt="%d")
PY
cd $W
export ASAN_OPTIONS=protect_shadow_gap=0:detect_leaks=0:abort_on_error=0
for ms in false true; do for sched in dynamic static; do
  ./inference --d_matrix_file=d.csv --region_sizes_file=r.txt --n_cells=5000 --n_regions=60 --n_iters=300 --n_nodes=12 --seed=5 --nu=1.0 --verbosity=2 \
     --max_scoring=$ms --n_chains=6 --schedule=$sched --min_batch=2 --postfix=one_${ms}_$sched > $OLDPWD/$OUT/run_${ms}_$sched.out 2> $OLDPWD/$OUT/run_${ms}_$sched.err
  echo "max_scoring=$ms $sched rc=$?"; head -40 $OLDPWD/$OUT/run_${ms}_$sched.err | cut -c1-220
done; done
