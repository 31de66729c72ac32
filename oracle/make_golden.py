#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Runs oracle/_ref/ref_dump (our driver linked against the
reference's Inference/Tree/MathOp classes, built by oracle/Makefile) on

  * the reference's worked example (scicone/tests/validation.h:22-31 and
    Inference::initialize_worked_example, Inference.h:158-181),
  * the bundled tutorial inputs docs/example_data/* with their stored trees,
  * small seeded synthetic matrices (integer counts, per-region neutral states,
    non-integer cluster means with cluster_sizes),

in both likelihood branches and both scoring modes, with move traces that cover
every rescoring move (prune/reattach, swap, add/remove, insert/delete,
condense/split, expand/shrink, common ancestor, nu, delete leaf).

Needs /root/reference and `make -C oracle`; the outputs are committed, so neither
tests nor the GPU box need this script to run.
"""
import gzip
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SCICONE_REF", "/root/reference")
DUMP = os.path.join(HERE, "_ref", "ref_dump")
OUT = os.path.join(HERE, "..", "tests", "golden")
EX = os.path.join(REF, "docs", "example_data")

WORKED_D = [[39, 37, 45, 49, 30], [31, 28, 34, 46, 11], [69, 58, 68, 34, 21], [72, 30, 31, 46, 21], [64, 17, 19, 37, 13]]
WORKED_R = [4, 2, 3, 5, 2]

ALL_MOVES_SUM = "1,1,1,1,1,1,1,1,1,1,0,0,0,{od},0.5"
ALL_MOVES_MAX = "1,1,1,1,1,1,1,1,1,1,0,0.5,0.5,{od},0"


def run(name, args, tmp):
    cmd = [DUMP] + [f"--{k}={v}" for k, v in args.items()]
    res = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True)
    if res.returncode != 0:
        sys.exit(f"{name}: ref_dump failed\n{res.stderr[-2000:]}")
    doc = json.loads(res.stdout)
    doc["generator"] = {"tool": "oracle/_ref/ref_dump", "args": {k: (os.path.basename(v) if isinstance(v, str) and os.sep in v else v) for k, v in args.items()}}
    path = os.path.join(OUT, name + ".json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(doc, separators=(",", ":")).encode())
    n_moves = len(doc["moves"])
    ids = sorted({m["move_id"] for m in doc["moves"]})
    print(f"{name}: cells={doc['n_cells']} regions={doc['n_regions']} moves={n_moves} ids={ids} acc={sum(m['accepted'] for m in doc['moves'])} size={os.path.getsize(path)}")


def write_inputs(tmp, tag, D, r, w=None, neutral=None):
    d = os.path.join(tmp, tag + "_d.csv")
    np.savetxt(d, np.asarray(D, dtype=np.float64), delimiter=",", fmt="%.18e")
    rf = os.path.join(tmp, tag + "_r.txt")
    np.savetxt(rf, np.asarray(r), fmt="%d")
    out = {"d_matrix_file": d, "region_sizes_file": rf, "n_cells": len(D), "n_regions": len(r)}
    if w is not None:
        wf = os.path.join(tmp, tag + "_w.txt")
        np.savetxt(wf, np.asarray(w), fmt="%d")
        out["cluster_sizes_file"] = wf
    if neutral is not None:
        nf = os.path.join(tmp, tag + "_n.txt")
        np.savetxt(nf, np.asarray(neutral), fmt="%d")
        out["region_neutral_states_file"] = nf
    return out


def synth(seed, M, K, neutral=None, clusters=None):
    """Small Dirichlet-multinomial-like count matrix: a few clones with +-1 events."""
    rng = np.random.default_rng(seed)
    r = rng.integers(10, 60, size=K)
    base = np.full(K, 2) if neutral is None else np.asarray(neutral)
    clones = [base.copy()]
    for _ in range(5):
        c = clones[rng.integers(len(clones))].copy()
        k = rng.integers(K)
        c[k] = np.clip(c[k] + rng.choice([-1, 1]), 0, 6)
        clones.append(c)
    D = np.zeros((M, K))
    for j in range(M):
        cn = clones[rng.integers(len(clones))]
        p = rng.dirichlet(10.0 * (cn * r + 1e-3))
        D[j] = rng.multinomial(int(10 * r.sum()), p)
    if clusters:
        # cluster means (non-integer) + integer sizes, as scicone.py:331-335 produces them
        lab = rng.integers(clusters, size=M)
        lab[:clusters] = np.arange(clusters)
        Dc = np.stack([D[lab == c].mean(axis=0) for c in range(clusters)])
        w = np.array([(lab == c).sum() for c in range(clusters)])
        return Dc, r, w
    return D, r, None


def main():
    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        # --- A. the reference's worked example -------------------------------------------
        base = write_inputs(tmp, "worked", WORKED_D, WORKED_R)
        for od in (0, 1):
            for mx in (0, 1):
                a = dict(base, worked_example=1, is_overdispersed=od, nu=2.0 if od else 0.2, max_scoring="true" if mx else "false",
                         seed=42 + od, n_moves=120, cf=0 if mx else 1,
                         move_probs=(ALL_MOVES_MAX if mx else ALL_MOVES_SUM).format(od=1 if od else 0))
                run(f"worked_od{od}_max{mx}", a, tmp)
        # --- B. bundled tutorial inputs + stored trees ---------------------------------------
        tut = {"d_matrix_file": os.path.join(EX, "5nodes_10regions_10000reads_test_d_mat_segmented_counts.txt"),
               "region_sizes_file": os.path.join(EX, "test_segmented_region_sizes.txt"), "n_cells": 200, "n_regions": 8}
        run("tutorial_full_sum", dict(tut, tree_file=os.path.join(EX, "test_tree_inferred.txt"), nu=8.31517949127393, cf=1,
                                      max_scoring="false", n_moves=0), tmp)
        run("tutorial_full_max", dict(tut, tree_file=os.path.join(EX, "test_tree_inferred.txt"), nu=8.31517949127393, cf=0,
                                      max_scoring="true", n_moves=12, seed=7,
                                      move_probs=ALL_MOVES_MAX.format(od=1)), tmp)
        clu = {"d_matrix_file": os.path.join(EX, "5nodes_10regions_10000reads_test_d_mat_segmented_counts_condensed.txt"),
               "cluster_sizes_file": os.path.join(EX, "5nodes_10regions_10000reads_test_d_mat_segmented_counts_cluster_sizes.txt"),
               "region_sizes_file": os.path.join(EX, "test_segmented_region_sizes.txt"), "n_cells": 6, "n_regions": 8}
        run("tutorial_cluster_sum", dict(clu, tree_file=os.path.join(EX, "cluster_tree_inferred.txt"), nu=6.21243258083137, cf=1,
                                         max_scoring="false", n_moves=100, seed=3, move_probs=ALL_MOVES_SUM.format(od=1)), tmp)
        run("tutorial_cluster_max", dict(clu, tree_file=os.path.join(EX, "cluster_tree_inferred.txt"), nu=6.21243258083137, cf=0,
                                         max_scoring="true", n_moves=100, seed=4, move_probs=ALL_MOVES_MAX.format(od=1)), tmp)
        # --- C. seeded synthetic matrices -----------------------------------------------------
        D, r, _ = synth(1, 40, 12)
        base = write_inputs(tmp, "syn1", D, r)
        for mx in (0, 1):
            run(f"synth40x12_od1_max{mx}", dict(base, n_nodes=8, max_regions_per_node=2, is_overdispersed=1, nu=1.0,
                                                max_scoring="true" if mx else "false", seed=11 + mx, n_moves=150, cf=0 if mx else 1,
                                                move_probs=(ALL_MOVES_MAX if mx else ALL_MOVES_SUM).format(od=1)), tmp)
        run("synth40x12_od0_max0", dict(base, n_nodes=8, max_regions_per_node=2, is_overdispersed=0, nu=1.0, max_scoring="false",
                                        seed=13, n_moves=100, cf=1, move_probs=ALL_MOVES_SUM.format(od=0)), tmp)
        neutral = [2] * 9 + [1] * 3  # e.g. a male X chromosome
        D, r, _ = synth(2, 32, 12, neutral=neutral)
        base = write_inputs(tmp, "syn2", D, r, neutral=neutral)
        run("synth32x12_neutral_max1", dict(base, n_nodes=6, is_overdispersed=1, nu=0.7, max_scoring="true", seed=21, n_moves=120,
                                            move_probs=ALL_MOVES_MAX.format(od=1)), tmp)
        D, r, w = synth(3, 300, 10, clusters=9)
        base = write_inputs(tmp, "syn3", D, r, w=w)
        for mx in (0, 1):
            run(f"synth_cluster9x10_max{mx}", dict(base, n_nodes=5, is_overdispersed=1, nu=1.5, max_scoring="true" if mx else "false",
                                                   seed=31 + mx, n_moves=120, cf=0 if mx else 1,
                                                   move_probs=(ALL_MOVES_MAX if mx else ALL_MOVES_SUM).format(od=1)), tmp)


if __name__ == "__main__":
    main()
