#!/usr/bin/env python3
"""Stress of the multi-process chain-ownership protocol on the CPU (oracle-backed ABI, --no_shard: every process holds all
cells, so there is no collective and the shared-memory mailbox / launch log carry all the ordering): random world size,
host threads, chains, min_batch and schedule; every owner's traces must equal the single-chain run of the native host.

    python scripts/ownership_stress.py --n 30"""
import argparse
import os
import subprocess
import sys
import tempfile
import uuid

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
from host_util import CPU_BACKEND_DIR, NATIVE  # noqa: E402
from scicone_b200.synth import make_counts  # noqa: E402

FILES = ("acceptance_ratio.csv", "markov_chain.csv", "gamma_values.csv", "tree_inferred.txt")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10)
    ap.add_argument("--first_seed", type=int, default=1)
    a = ap.parse_args()
    bad = 0
    for seed in range(a.first_seed, a.first_seed + a.n):
        rng = np.random.default_rng(seed)
        world = int(rng.choice([2, 3, 4]))
        threads = int(rng.choice([1, 2, 4]))
        chains = int(rng.integers(world, 13))
        schedule = str(rng.choice(["dynamic", "dynamic", "static"]))
        min_batch = int(rng.integers(1, 5))
        iters = int(rng.choice([150, 400]))
        d = make_counts(60, 400, 10, seed=seed)
        with tempfile.TemporaryDirectory(prefix="own_") as tmp:
            np.savetxt(os.path.join(tmp, "d.csv"), d["D"], delimiter=",", fmt="%.0f")
            np.savetxt(os.path.join(tmp, "r.txt"), d["region_sizes"], fmt="%d")
            env = dict(os.environ, LD_LIBRARY_PATH=CPU_BACKEND_DIR, SCICONE_THREADS=str(threads))
            base = [NATIVE, "--d_matrix_file=d.csv", "--region_sizes_file=r.txt", "--n_cells=60", "--n_regions=10", f"--n_iters={iters}",
                    "--n_nodes=5", "--nu=1.0", "--max_scoring=true", "--verbosity=2", "--no_cnv_file=1"]
            shm = "/scicone_stress_" + uuid.uuid4().hex[:10]
            procs = [subprocess.Popen(base + [f"--seed={seed}", f"--n_chains={chains}", f"--schedule={schedule}", f"--min_batch={min_batch}", "--no_shard",
                                              f"--rank={k}", f"--world={world}", f"--shm_name={shm}", f"--postfix=w{k}"],
                                      cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env) for k in range(world)]
            ok = True
            for p in procs:
                try:
                    out, err = p.communicate(timeout=300)
                except subprocess.TimeoutExpired:
                    p.kill()
                    ok = False
                    err = "timeout"
                if p.returncode != 0:
                    ok = False
                    print("   rank failed:", err[-300:])
            if ok:
                one = subprocess.run(base + [f"--seed={seed}", f"--n_chains={chains}", "--schedule=static", "--postfix=one"], cwd=tmp, capture_output=True,
                                     text=True, env=env)
                ok = one.returncode == 0
                for c in range(chains if ok else 0):
                    for name in FILES:
                        x = open(os.path.join(tmp, f"one_c{c}_{name}")).read()
                        y = open(os.path.join(tmp, f"w{c % world}_c{c}_{name}")).read()
                        if x != y:
                            ok = False
                            print(f"   chain {c} {name} differs")
                            break
            bad += 0 if ok else 1
            print(f"seed {seed}: {'ok' if ok else 'FAILED'}  world {world} threads {threads} chains {chains} {schedule} min_batch {min_batch} iters {iters}", flush=True)
    print(f"{a.n - bad} of {a.n} runs identical")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
