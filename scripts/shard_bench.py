#!/usr/bin/env python3
"""Tuning aid: the native host cell-sharded over N GPUs (weak scaling: N x 10 000 cells, 32 chains) outside bench.py, for a list
of schedule variants.  Prints one line per variant: ms per MCMC iteration of all chains (max over ranks), launches.
    python scripts/shard_bench.py <N> [iters] [variant ...]     variant = "threads:n_streams:min_batch" (0 = default)"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scicone_b200.api import comm_unique_id  # noqa: E402
from scicone_b200.synth import SEED, make_counts  # noqa: E402

N = int(sys.argv[1])
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 150
variants = sys.argv[3:] or ["0:2:0"]
cells = 10000 * N
tmp = tempfile.mkdtemp(prefix="sbench_")
d = make_counts(cells, 18000, 200, seed=SEED)
np.ascontiguousarray(d["D"], dtype="<f8").tofile(os.path.join(tmp, "d.f64"))
np.savetxt(os.path.join(tmp, "r.txt"), d["region_sizes"], fmt="%d")
exe = os.path.join(ROOT, "scicone_b200", "bin", "inference")
cores = os.cpu_count() or 1
for v in variants:
    threads, ns, mb = (int(x) for x in v.split(":"))
    threads = threads or (max(2, cores // N - 1) if N > 1 else cores)
    uid = comm_unique_id().hex() if N > 1 else None
    procs = []
    for r in range(N):
        cmd = [exe, "--d_matrix_file=d.f64", "--region_sizes_file=r.txt", f"--n_cells={cells}", "--n_regions=200", f"--n_iters={iters}", "--warmup_iters=25",
               "--n_nodes=49", "--seed=42", "--nu=1.0", "--verbosity=0", "--no_cnv_file=1", "--max_scoring=true", f"--postfix=s{r}", "--n_chains=32",
               f"--device={r}", "--schedule=dynamic", f"--n_streams={ns}", f"--stats_file=stats{r}.json", "--stats_all_ranks=1", "--profile_kernels=0"]
        if mb:
            cmd.append(f"--min_batch={mb}")
        if N > 1:
            cmd += [f"--rank={r}", f"--world={N}", f"--nccl_id={uid}"]
        procs.append(subprocess.Popen(cmd, cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=dict(os.environ, SCICONE_THREADS=str(threads))))
    ok = True
    for r, p in enumerate(procs):
        o, e = p.communicate(timeout=600)
        if p.returncode != 0:
            ok = False
            print("variant", v, "rank", r, "failed:", (o[-200:] + e[-400:]).replace("\n", " | "))
    if not ok:
        continue
    st = [json.load(open(os.path.join(tmp, f"stats{r}.json"))) for r in range(N)]
    wall = max(s["wall_us"] for s in st)
    print(json.dumps({"n_gpus": N, "variant": v, "threads": threads, "n_streams": ns, "min_batch": mb, "iters": iters, "ms_per_step": wall * 1e-3 / iters,
                      "mcmc_iters_per_s": 32 * iters / (wall * 1e-6), "launches": st[0]["launches"], "chains_per_launch": st[0]["launched_chains"] / max(1, st[0]["launches"])}), flush=True)
