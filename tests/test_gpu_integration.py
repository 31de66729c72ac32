"""Drop-in check: the reference's own MCMC host (unmodified Tree/Node/Inference move logic, RNG, priors, accept
rule, CLI and writers) linked against libscicone_score.so through oracle/integration/b200_adapter.inl must walk
the SAME accepted-move trajectory and end in the SAME tree as the stock reference binary, with per-move tree
scores within 1e-9 relative (north_star).  Both binaries are built by oracle/Makefile into oracle/_ref/ and
travel to the GPU box; inputs are the bundled tutorial matrix (carried inside tests/golden) and seeded synthetic
matrices."""
import os
import subprocess

import numpy as np
import pytest

from golden_util import load_golden
from host_util import REF, REF_ON_ABI, compare_runs as _compare, run_inference as _run
from scicone_b200.synth import make_counts

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B200 = REF_ON_ABI


needs_bins = pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(B200)),
                                reason="oracle/_ref/{inference,inference_b200} not built (make -C oracle)")


@needs_bins
@pytest.mark.parametrize("max_scoring", ["true", "false"])
def test_tutorial_run_same_trajectory(tmp_path, max_scoring):
    """configs[0]: the docs/tutorial.md inference command on the bundled 200 x 8 matrix, fixed seed."""
    doc = load_golden("tutorial_full_sum")
    D, r = np.array(doc["D"]), np.array(doc["region_sizes"])
    extra = ["--n_iters=4000", "--n_nodes=3", "--seed=42", "--ploidy=2", "--copy_number_limit=2", f"--max_scoring={max_scoring}"]
    a = _run(REF, str(tmp_path), "ref", D, r, extra)
    b = _run(B200, str(tmp_path), "b200", D, r, extra)
    rel = _compare(a, b)
    print("accepted", int(np.sum(a["acc"] >= 0)), "of", a["acc"].size, "max rel diff of per-move scores", rel)


@needs_bins
@pytest.mark.parametrize("max_scoring,n_nodes", [("true", 10), ("false", 30)])
def test_config2_same_trajectory(tmp_path, max_scoring, n_nodes):
    """configs[1]: simulated 1 000 cells x 10 000 bins -> 50 regions, full-tree MCMC, default move weights."""
    data = make_counts(1000, 10000, 50, seed=5)
    extra = ["--n_iters=600", f"--n_nodes={n_nodes}", "--seed=43", "--nu=1.0", f"--max_scoring={max_scoring}"]
    a = _run(REF, str(tmp_path), "ref", data["D"], data["region_sizes"], extra)
    b = _run(B200, str(tmp_path), "b200", data["D"], data["region_sizes"], extra)
    rel = _compare(a, b)
    print("accepted", int(np.sum(a["acc"] >= 0)), "of", a["acc"].size, "max rel diff", rel)


@needs_bins
def test_cluster_tree_mode_same_trajectory(tmp_path):
    """configs[3]: cluster means (non-integer) weighted by cluster_sizes."""
    data = make_counts(3000, 10000, 40, seed=6, cluster_rows=30)
    extra = ["--n_iters=1500", "--n_nodes=5", "--seed=44", "--nu=1.0", "--max_scoring=true"]
    a = _run(REF, str(tmp_path), "ref", data["D"], data["region_sizes"], extra, w=data["cluster_sizes"])
    b = _run(B200, str(tmp_path), "b200", data["D"], data["region_sizes"], extra, w=data["cluster_sizes"])
    _compare(a, b)
