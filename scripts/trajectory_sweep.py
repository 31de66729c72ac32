#!/usr/bin/env python3
"""Randomised trajectory-parity sweep (CPU only): the native MCMC host on the oracle-backed C ABI against the UNMODIFIED
reference `inference` (oracle/_ref), same seed => same accepted-move sequence, same final tree, scores within 1e-9.

    python scripts/trajectory_sweep.py --n 40 [--first_seed 1000]

Every case draws its own configuration: scoring mode, tree size, nu, move probabilities (all 15 moves on / defaults /
random subset), tempering, size limit, cluster sizes, ploidy and copy-number limit.  Divergences are reported with the
command line that reproduces them.  tests/test_host_trajectory.py holds the fixed cases; this is the wide net."""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
from host_util import CPU_BACKEND_DIR, NATIVE, REF, compare_runs, run_inference  # noqa: E402
from scicone_b200.synth import make_counts  # noqa: E402


def draw_case(seed, big=False):
    rng = np.random.default_rng(seed)
    M = int(rng.choice([400, 800] if big else [40, 120, 300]))
    K = int(rng.choice([60, 120] if big else [8, 20, 45]))
    clustered = rng.random() < 0.25
    d = make_counts(M * (20 if clustered else 1), 40 * K, K, seed=seed, cluster_rows=M if clustered else 0)
    max_scoring = bool(rng.random() < 0.5)
    flags = [f"--n_iters={int(rng.choice([400, 900]))}", f"--n_nodes={int(rng.integers(10, 41) if big else rng.integers(2, 13))}", f"--seed={seed}",
             f"--max_scoring={'true' if max_scoring else 'false'}"]
    mode = rng.random()
    if mode < 0.35:  # every move on (sum scoring has no genotype-preserving moves: infer_trees.cpp:198-210)
        probs = [1] * 15
        if not max_scoring:
            probs[11] = probs[12] = 0
        flags.append("--move_probs=" + ",".join(str(p) for p in probs))
    elif mode < 0.6:  # random subset / weights
        probs = [round(float(p), 2) for p in rng.choice([0, 0.01, 0.5, 1.0], size=15)]
        if not max_scoring:
            probs[11] = probs[12] = 0
        if sum(probs) == 0:
            probs[1] = 1.0
        flags.append("--move_probs=" + ",".join(str(p) for p in probs))
    if rng.random() < 0.8:
        flags.append(f"--nu={float(rng.choice([0.3, 1.0, 2.5, 8.0]))}")
    if rng.random() < 0.3:
        flags += [f"--alpha={float(rng.choice([0.1, 0.3]))}", f"--gamma={float(rng.choice([0.5, 2.0]))}"]
    if rng.random() < 0.3:
        flags.append(f"--size_limit={int(rng.integers(6, 20))}")
    if rng.random() < 0.3:
        flags += [f"--ploidy={int(rng.choice([2, 3]))}", f"--copy_number_limit={int(rng.choice([2, 4]))}"]
    if rng.random() < 0.2:
        flags.append(f"--cf={float(rng.choice([0.0, 0.5, 1.0]))}")
    return d["D"], d["region_sizes"], (d["cluster_sizes"] if clustered else None), flags


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=20)
    ap.add_argument("--first_seed", type=int, default=1000)
    ap.add_argument("--big", action="store_true", help="400-800 cells, 60-120 regions, 10-40 node trees")
    a = ap.parse_args()
    env = {"LD_LIBRARY_PATH": CPU_BACKEND_DIR}
    bad = 0
    t0 = time.time()
    for seed in range(a.first_seed, a.first_seed + a.n):
        D, r, w, flags = draw_case(seed, a.big)
        with tempfile.TemporaryDirectory(prefix="sweep_") as tmp:
            try:
                ref = run_inference(REF, tmp, "ref", D, r, flags, w=w)
                new = run_inference(NATIVE, tmp, "new", D, r, flags, w=w, env=env)
                out = compare_runs(ref, new, knife_edge_ok=True)
                acc = int(np.sum(ref["acc"] >= 0))
                print(f"seed {seed}: ok  {D.shape[0]}x{D.shape[1]} accepted {acc}/{ref['acc'].size} {out if isinstance(out, tuple) else f'rel {out:.1e}'}  {' '.join(flags)}",
                      flush=True)
            except AssertionError as e:
                bad += 1
                print(f"seed {seed}: MISMATCH {str(e)[:300]}\n    {D.shape} {' '.join(flags)}", flush=True)
    print(f"{a.n - bad} of {a.n} cases identical ({time.time() - t0:.0f} s)")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
