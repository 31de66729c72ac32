#!/usr/bin/env python3
"""Sweep schedule parameters of the native MCMC host on ONE prepared configs[2] input (GPU box): the matrix is written once,
every variant is one run of scicone_b200/bin/inference.  Prints one line per run.

    python scripts/host_sweep.py --iters 20,300 --variants "n_streams=2,min_batch=8;n_streams=3,min_batch=4" [--reps 2]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=10000)
    ap.add_argument("--regions", type=int, default=200)
    ap.add_argument("--nodes", type=int, default=50)
    ap.add_argument("--chains", type=int, default=32)
    ap.add_argument("--iters", default="20,300")
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--variants", default="n_streams=2,min_batch=8")
    ap.add_argument("--env", default="", help="NAME=VALUE,... for every run")
    a = ap.parse_args()
    from scicone_b200.synth import SEED, make_counts

    data = make_counts(a.cells, 18000, a.regions, seed=SEED)
    tmp = tempfile.mkdtemp(prefix="hostsweep_")
    np.ascontiguousarray(data["D"], dtype="<f8").tofile(os.path.join(tmp, "d.f64"))
    np.savetxt(os.path.join(tmp, "r.txt"), data["region_sizes"], fmt="%d")
    env = dict(os.environ)
    for kv in filter(None, a.env.split(",")):
        k, v = kv.split("=")
        env[k] = v
    exe = os.path.join(ROOT, "scicone_b200", "bin", "inference")
    for var in a.variants.split(";"):
        flags = [f"--{kv}" for kv in filter(None, var.split(","))]
        for iters in (int(x) for x in a.iters.split(",")):
            us = []
            for rep in range(a.reps):
                stats = os.path.join(tmp, "stats.json")
                cmd = [exe, "--d_matrix_file=d.f64", "--region_sizes_file=r.txt", f"--n_cells={a.cells}", f"--n_regions={a.regions}",
                       f"--n_nodes={a.nodes - 1}", "--seed=42", "--nu=1.0", "--verbosity=1", "--max_scoring=true", f"--n_iters={iters}",
                       f"--n_chains={a.chains}", "--postfix=s", f"--stats_file={stats}", "--no_cnv_file=1", f"--warmup_iters={a.warmup}",
                       "--profile_kernels=0"] + flags
                res = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True, env=env)
                if res.returncode != 0:
                    print("FAILED", var, res.stdout[-500:], res.stderr[-500:])
                    break
                s = json.load(open(stats))
                us.append(s["wall_us"] / iters)
            print(json.dumps({"variant": var, "iters": iters, "us_per_step": [round(x, 1) for x in us], "launches": s["launches"],
                              "props_per_launch": round(s["launched_chains"] / max(1, s["launches"]), 2)}), flush=True)


if __name__ == "__main__":
    main()
