"""Helpers for the trajectory tests: run an `inference` binary (the unmodified reference, the reference host forwarded
to the C ABI, or this repo's native host) on a matrix and compare everything it writes."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "inference")                # unmodified reference (oracle/Makefile)
REF_ON_ABI = os.path.join(ROOT, "oracle", "_ref", "inference_b200")    # reference host -> libscicone_score.so
NATIVE = os.path.join(ROOT, "scicone_b200", "bin", "inference")        # this repo's host -> libscicone_score.so
CPU_BACKEND_DIR = os.path.join(ROOT, "oracle", "_ref", "cpu_backend")  # C ABI on the oracle port (tests only)


def run_inference(exe, tmp, tag, D, r, extra, w=None, env=None, timeout=900):
    d = os.path.join(tmp, "d.csv")
    integer = np.allclose(D, np.round(D))
    np.savetxt(d, D, delimiter=",", fmt="%.0f" if integer else "%.18e")
    rf = os.path.join(tmp, "r.txt")
    np.savetxt(rf, r, fmt="%d")
    cmd = [exe, f"--d_matrix_file={d}", f"--region_sizes_file={rf}", f"--n_cells={D.shape[0]}", f"--n_regions={D.shape[1]}",
           "--verbosity=2", f"--postfix={tag}"] + extra
    if w is not None:
        wf = os.path.join(tmp, "w.txt")
        np.savetxt(wf, w, fmt="%d")
        cmd.append(f"--cluster_sizes_file={wf}")
    e = dict(os.environ)
    if env:
        e.update(env)
    res = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True, timeout=timeout, env=e)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    return read_outputs(tmp, tag)


def read_outputs(tmp, tag):
    def csv(name):
        return np.array([float(x) for x in open(os.path.join(tmp, f"{tag}_{name}.csv")).read().strip().split(",")])

    return {"acc": csv("acceptance_ratio"), "chain": csv("markov_chain"), "gamma": csv("gamma_values"),
            "tree": open(os.path.join(tmp, f"{tag}_tree_inferred.txt")).read(),
            "cell_nodes": open(os.path.join(tmp, f"{tag}_cell_node_ids.tsv")).read(),
            "cnvs": open(os.path.join(tmp, f"{tag}_inferred_cnvs.csv")).read(),
            "region_cnvs": open(os.path.join(tmp, f"{tag}_cell_region_cnvs.csv")).read(),
            "scores": np.loadtxt(os.path.join(tmp, f"{tag}_attachment_scores.csv"), delimiter=",", ndmin=2)}


def tree_parts(txt):
    lines = txt.strip().split("\n")
    head = [float(l.split(":")[1]) for l in lines[:7]]
    return head, lines[7:]


def compare_runs(a, b, tol=1e-9, knife_edge_ok=False):
    """Identical accepted-move trajectory and final tree, per-move scores within `tol` relative (north_star).

    knife_edge_ok: the reference's accept rule (Inference.h:497-515) draws a uniform only when log_acc <= 0.  After a
    move that is likelihood-NEUTRAL in exact arithmetic (e.g. swapping two labels no cell is attached to), log_acc is
    +-(a few ulp of a 4e6-sized total), and the SIGN of that rounding noise decides whether the random stream advances.
    No two libm's / summation orders agree on that sign, so from such a move on the two hosts may legitimately walk
    different (equally valid) chains.  With this flag a divergence is tolerated ONLY directly after an accepted move
    whose score change in the reference is below `tol` relative, and everything before it must still agree; the
    function then returns ("knife_edge", iteration)."""
    n = min(a["acc"].size, b["acc"].size)
    diff = np.nonzero(a["acc"][:n] != b["acc"][:n])[0]
    if knife_edge_ok and a["acc"].size == b["acc"].size and diff.size:
        i = int(diff[0])
        assert i >= 2, "trajectories differ from the start"
        assert a["acc"][i - 1] >= 0 and b["acc"][i - 1] >= 0, "divergence is not preceded by an accepted move"
        step = abs(a["chain"][i - 1] - a["chain"][i - 2]) / max(1.0, abs(a["chain"][i - 1]))
        assert step <= tol, f"trajectories diverge at iteration {i} after a non-neutral move (relative step {step})"
        rel = float(np.max(np.abs(a["chain"][:i] - b["chain"][:i]) / np.maximum(1.0, np.abs(a["chain"][:i]))))
        assert rel <= tol, rel
        assert np.allclose(a["gamma"][:i], b["gamma"][:i], rtol=1e-12)
        return ("knife_edge", i)
    assert np.array_equal(a["acc"], b["acc"]), "accepted-move trajectories differ"
    rel = float(np.max(np.abs(a["chain"] - b["chain"]) / np.maximum(1.0, np.abs(a["chain"]))))
    assert rel <= tol, rel
    assert np.allclose(a["gamma"], b["gamma"], rtol=1e-12)
    ha, na = tree_parts(a["tree"])
    hb, nb = tree_parts(b["tree"])
    assert na == nb, "final trees differ"
    assert np.allclose(ha, hb, rtol=tol, atol=tol)
    assert a["cell_nodes"] == b["cell_nodes"]
    assert a["cnvs"] == b["cnvs"]
    assert a["region_cnvs"] == b["region_cnvs"]
    assert a["scores"].shape == b["scores"].shape
    assert np.max(np.abs(a["scores"] - b["scores"]) / np.maximum(1.0, np.abs(a["scores"]))) <= tol
    return rel
