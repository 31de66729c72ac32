#!/usr/bin/env python3
"""Time the native MCMC host (scicone_b200/bin/inference, real moves / accept-reject through the C ABI) on a synthetic
matrix of a BASELINE.json config; optionally the unmodified reference binary beside it.  Prints one JSON line per run.

    python scripts/host_bench.py --cells 10000 --regions 200 --nodes 50 --chains 32 --iters 300 [--ref_iters 100]
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=10000)
    ap.add_argument("--bins", type=int, default=18000)
    ap.add_argument("--regions", type=int, default=200)
    ap.add_argument("--nodes", type=int, default=50)
    ap.add_argument("--chains", type=int, default=32)
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--ref_iters", type=int, default=0)
    ap.add_argument("--max_scoring", default="true")
    ap.add_argument("--extra", default="")
    ap.add_argument("--cluster_rows", type=int, default=0, help="configs[3]: condense the cells into this many cluster means with cluster_sizes")
    a = ap.parse_args()
    from scicone_b200.synth import SEED, make_counts

    data = make_counts(a.cells, a.bins, a.regions, seed=SEED, cluster_rows=a.cluster_rows or None)
    n_rows = data["D"].shape[0]
    tmp = tempfile.mkdtemp(prefix="hostbench_")
    d = os.path.join(tmp, "d.csv")
    np.savetxt(d, data["D"], delimiter=",", fmt="%.0f" if not a.cluster_rows else "%.18e")
    r = os.path.join(tmp, "r.txt")
    np.savetxt(r, data["region_sizes"], fmt="%d")
    common = [f"--d_matrix_file={d}", f"--region_sizes_file={r}", f"--n_cells={n_rows}", f"--n_regions={a.regions}",
              f"--n_nodes={a.nodes - 1}", "--seed=42", "--nu=1.0", "--verbosity=1", f"--max_scoring={a.max_scoring}"]
    if a.cluster_rows:
        w = os.path.join(tmp, "w.txt")
        np.savetxt(w, data["cluster_sizes"], fmt="%d")
        common.append(f"--cluster_sizes_file={w}")
    stats = os.path.join(tmp, "stats.json")
    cmd = [os.path.join(ROOT, "scicone_b200", "bin", "inference")] + common + [f"--n_iters={a.iters}", f"--n_chains={a.chains}",
                                                                                 "--postfix=nat", f"--stats_file={stats}", "--no_cnv_file=1"] + a.extra.split()
    t0 = time.perf_counter()
    res = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True)
    wall = time.perf_counter() - t0
    if res.returncode != 0:
        print(res.stdout[-2000:], res.stderr[-2000:])
        raise SystemExit(1)
    for line in res.stderr.splitlines():  # SCICONE_DEBUG_VISITS=1: the host's notes about slow visits
        if line.startswith("visit"):
            print(line, file=sys.stderr)
    s = json.load(open(stats))
    s["process_wall_s"] = wall
    s["mcmc_iters_per_s"] = s["n_chains"] * s["n_iters"] / (s["wall_us"] * 1e-6)
    s["cell_iters_per_s"] = s["mcmc_iters_per_s"] * a.cells
    s["rows"] = n_rows
    s["us_per_step"] = s["wall_us"] / s["n_iters"]
    s["score_kernel_us_per_step"] = s["score_kernel_us"] / s["n_iters"]
    s["build_kernel_us_per_step"] = s["build_kernel_us"] / s["n_iters"]
    print(json.dumps({"impl": "native_host", **s}))
    if a.ref_iters > 0:
        cmd = [os.path.join(ROOT, "oracle", "_ref", "inference")] + common + [f"--n_iters={a.ref_iters}", "--postfix=ref"]
        res = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS="1"))
        m = re.search(r"Time taken by infer_mcmc function: (\d+) microseconds", res.stdout)
        us = int(m.group(1))
        print(json.dumps({"impl": "reference_1thread", "n_iters": a.ref_iters, "wall_us": us, "mcmc_iters_per_s": a.ref_iters / (us * 1e-6)}))


if __name__ == "__main__":
    main()
