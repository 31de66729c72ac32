"""The identity behind the histogram form of the root score (kernels.cuh, ITEM_OD_HIST), checked on the CPU against the
oracle port of Tree::get_od_root_score / Inference::compute_t_od_scores (reference scicone/src/Tree.h:1774-1810,
Inference.h:1406-1419): the reference only uses the cluster-size-weighted SUM of the per-cell root scores, and in that
sum the cells enter the region part only through how often each count occurs,

    sum_j w_j sum_k (lgamma(D_jk + a_k) - lgamma(a_k)) = sum_k sum_v hist_k[v] (lgamma(v + a_k) - lgamma(a_k)),
    hist_k[v] = sum_j w_j [D_jk == v],   a_k = nu * neutral_k * r_k,

while the cell-level term lgamma(nu z) - lgamma(S_j + nu z) stays per cell.  The GPU side of the same statement is
tests/test_gpu_od_hist.py."""
import numpy as np
import pytest
from scipy.special import gammaln

from golden_util import rel_err
from oracle_ffi import OracleChain
from trace_util import make_case


@pytest.mark.parametrize("n_cells,n_regions,weights", [(500, 12, False), (300, 40, True), (1, 5, True), (129, 3, False)])
def test_root_score_is_a_function_of_the_count_histograms(n_cells, n_regions, weights):
    data, tree = make_case(n_cells, 100 * n_regions, n_regions, 6, seed=900 + n_cells)
    D, r, neutral = data["D"], data["region_sizes"].astype(np.int64), data["neutral_states"].astype(np.int64)
    w = np.random.default_rng(3).integers(1, 5, size=n_cells).astype(np.int32) if weights else np.ones(n_cells, dtype=np.int32)
    cpu = OracleChain(D=D, region_sizes=data["region_sizes"], neutral_states=data["neutral_states"], cluster_sizes=w, eta=1e-4,
                      is_overdispersed=1, max_scoring=1)
    cpu.compute_t_table(tree.flat())
    assert np.array_equal(D, np.round(D))
    z = int((r * neutral).sum())
    S = D.sum(axis=1)
    for nu in (1.0, 0.37, 11.0):
        a = nu * neutral * r
        total = float((w * (gammaln(nu * z) - gammaln(S + nu * z))).sum())  # cell-level term, per cell
        for k in range(n_regions):
            vmin = int(D[:, k].min())
            hist = np.bincount(D[:, k].astype(np.int64) - vmin, weights=w)
            v = vmin + np.arange(hist.size)
            nz = hist != 0
            total += float((hist[nz] * (gammaln(v[nz] + a[k]) - gammaln(a[k]))).sum())
        assert rel_err(total, cpu.compute_t_od_scores(nu)) <= 1e-11, nu
