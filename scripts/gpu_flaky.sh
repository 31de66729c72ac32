#!/bin/bash
# Repeat the single-GPU run of tests/test_gpu_multi.py many times to catch a rare crash (gdb backtrace on a failure).
set -u
TAG=${1:-flaky}; N=${2:-30}; OUT=$PWD/gpurun_out/$TAG; mkdir -p "$OUT"; W=/tmp/flaky; rm -rf $W; mkdir -p $W
python - <<'PY'
import sys
sys.path.insert(0,'.')
import numpy as np
from scicone_b200.synth import make_counts
data = make_counts(5000, 9000, 60, seed=31)
np.savetxt("/tmp/flaky/d.csv", data["D"], delimiter=",", fmt="%.0f"); np.savetxt("/tmp/flaky/r.txt", data["region_sizes"], fmt="%d")
PY
EXE=$PWD/scicone_b200/bin/inference
cd $W; ulimit -c unlimited
fails=0
for i in $(seq 1 $N); do for ms in false true; do
  $EXE --d_matrix_file=d.csv --region_sizes_file=r.txt --n_cells=5000 --n_regions=60 --n_iters=300 --n_nodes=12 --seed=5 --nu=1.0 --verbosity=2 \
     --max_scoring=$ms --n_chains=6 --schedule=dynamic --min_batch=2 --postfix=one > run.out 2> run.err; rc=$?
  if [ $rc -ne 0 ]; then fails=$((fails+1)); echo "run $i max_scoring=$ms rc=$rc"; tail -3 run.err; cp run.err "$OUT/fail_${i}_$ms.err";
     if [ $fails -le 2 ] && which gdb > /dev/null; then gdb -batch -ex run -ex bt -ex "thread apply all bt 8" --args $EXE --d_matrix_file=d.csv --region_sizes_file=r.txt --n_cells=5000 --n_regions=60 --n_iters=300 --n_nodes=12 --seed=5 --nu=1.0 --verbosity=2 --max_scoring=$ms --n_chains=6 --schedule=dynamic --min_batch=2 --postfix=one 2>&1 | tail -60 > "$OUT/gdb_${i}_$ms.txt"; fi
  fi
done; done
echo "failures: $fails of $((2*N))"
