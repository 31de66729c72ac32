"""scicone_b200/bin/learn_tree: restarts spread over the GPUs of a box, best-of-N selection, robustness score and retries
- pyscicone's SCICoNE.learn_tree_parallel and the cluster-tree loop of SCICoNE.learn_tree
(reference: pyscicone/scicone/scicone.py:349-375 and :491-503), which pyscicone runs as a multiprocessing.Pool of
`inference` processes with seeds 42 + i.

The check: the UNMODIFIED reference binary is run once per seed exactly as pyscicone would run it (cluster-tree mode:
non-integer cluster means weighted by cluster_sizes), the selection rule is restated here from the cited lines, and
learn_tree must report the same per-restart scores (1e-9), the same best restart and the same robustness - on the first
try and on the retry that restarts every chain from the best tree of the first try (--tree_file + --nu of that tree).
pyscicone itself cannot be imported here (it needs phenograph); the rule is six lines."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

from host_util import CPU_BACKEND_DIR, NATIVE, REF
from scicone_b200.synth import make_counts

LEARN = os.path.join(os.path.dirname(NATIVE), "learn_tree")
needs_bins = pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(NATIVE) and os.path.exists(LEARN)),
                                reason="oracle/_ref/inference or scicone_b200/bin/{inference,learn_tree} not built")


def _tree_score(path):
    txt = open(path).read()
    return float(re.search(r"Tree score: (\S+)", txt).group(1)), float(re.search(r"Nu: (\S+)", txt).group(1))


def _select(scores):
    """scicone.py:365-373."""
    scores = np.array(scores, dtype=np.float64)
    raw_best = None
    sel = scores.copy()
    sel[sel == 0.0] = -np.inf  # "just in case"
    arg = int(np.argmax(sel))
    raw_best = scores[arg]
    robustness = np.count_nonzero(np.isclose(sel, raw_best, atol=5)) / len(scores)
    return arg, robustness


def _run(tmp_path, env, n_gpus):
    d = make_counts(3000, 10000, 40, seed=16, cluster_rows=24)  # configs[3]: condensed centroids + cluster_sizes
    D, r, w = d["D"], d["region_sizes"], d["cluster_sizes"]
    tmp = str(tmp_path)
    np.savetxt(os.path.join(tmp, "d.csv"), D, delimiter=",", fmt="%.18e")
    np.savetxt(os.path.join(tmp, "r.txt"), r, fmt="%d")
    np.savetxt(os.path.join(tmp, "w.txt"), w, fmt="%d")
    n_reps = 5
    common = ["--d_matrix_file=d.csv", "--region_sizes_file=r.txt", "--cluster_sizes_file=w.txt", f"--n_cells={D.shape[0]}",
              f"--n_regions={D.shape[1]}", "--n_iters=400", "--n_nodes=5", "--max_scoring=true", "--verbosity=1"]

    def reference_try(tag, extra):
        scores, nus = [], []
        for i in range(n_reps):
            res = subprocess.run([REF] + common + extra + [f"--seed={42 + i}", f"--postfix={tag}{i}"], cwd=tmp, capture_output=True, text=True, timeout=600)
            assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
            s, nu = _tree_score(os.path.join(tmp, f"{tag}{i}_tree_inferred.txt"))
            scores.append(s)
            nus.append(nu)
        return scores, nus

    # robustness_thr above 1 can never be met: learn_tree must use both tries
    e = dict(os.environ)
    e.update(env or {})
    res = subprocess.run([LEARN, f"--n_reps={n_reps}", f"--n_gpus={n_gpus}", "--max_tries=2", "--robustness_thr=1.5", "--postfix=lt", "--nu=1.0"] + common,
                         cwd=tmp, capture_output=True, text=True, env=e, timeout=900)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
    out = json.load(open(os.path.join(tmp, "lt_learn_tree.json")))
    assert out["n_tries"] == 2 and len(out["tries"]) == 2

    s0, nu0 = reference_try("ref0_", ["--nu=1.0"])
    arg0, rob0 = _select(s0)
    t0 = out["tries"][0]
    assert np.allclose(t0["scores"], s0, rtol=1e-9, atol=0.0), (t0["scores"], s0)
    assert t0["best_rep"] == arg0 and t0["robustness"] == rob0
    # the retry: every restart from the best tree of try 0, nu of that tree (scicone.py:499, tree.py:330-345)
    best_file = os.path.join(tmp, f"ref0_{arg0}_tree_inferred.txt")
    s1, _ = reference_try("ref1_", [f"--tree_file={best_file}", f"--nu={nu0[arg0]!r}"])
    arg1, rob1 = _select(s1)
    t1 = out["tries"][1]
    assert np.allclose(t1["scores"], s1, rtol=1e-9, atol=0.0), (t1["scores"], s1)
    assert t1["best_rep"] == arg1 and t1["robustness"] == rob1
    assert out["best_rep"] == arg1 and out["robustness"] == rob1
    # the best restart's files under the caller's postfix
    got = open(os.path.join(tmp, "lt_tree_inferred.txt")).read().split("\n")[7:]
    want = open(os.path.join(tmp, f"ref1_{arg1}_tree_inferred.txt")).read().split("\n")[7:]
    assert got == want
    return out


def test_selection_rule():
    assert _select([-10.0, -3.0, -7.9, 0.0]) == (1, 0.5)  # 0.0 is "no score", -7.9 is within 5 of -3
    assert _select([-1e6, -1e6 - 14.0, -1e6 - 16.0])[1] == pytest.approx(2 / 3)  # numpy's rtol: 5 + 1e-5 * 1e6 = 15


@needs_bins
@pytest.mark.skipif(not os.path.exists(os.path.join(CPU_BACKEND_DIR, "libscicone_score.so")), reason="cpu_backend not built")
@pytest.mark.parametrize("n_gpus", [1, 2])
def test_learn_tree_selects_like_pyscicone_on_cpu(tmp_path, n_gpus):
    """n_gpus = 2 here means two `inference` processes (3 + 2 restarts); the oracle-backed ABI ignores --device."""
    _run(tmp_path, {"LD_LIBRARY_PATH": CPU_BACKEND_DIR}, n_gpus)


@needs_bins
@pytest.mark.gpu
def test_learn_tree_selects_like_pyscicone_on_gpu(tmp_path):
    import torch

    _run(tmp_path, None, min(2, torch.cuda.device_count()))
