"""Two engine switches that must not change results:

* the root score (Tree::get_od_root_score, reference scicone/src/Tree.h:1774-1810) from the dataset's count histograms
  (ITEM_OD_HIST, default for integer counts) against the per-cell slices (SCICONE_OD_HIST=0, the path non-integer data
  takes) and against the oracle port;
* programmatic dependent launches of K1a -> K1b -> proposal kernel (default) against ordinary serialised launches
  (SCICONE_PDL=0): same kernels, same arithmetic, so every total must be bit-identical.
Both switches are read when the dataset is created."""
import numpy as np
import pytest

from golden_util import rel_err
from oracle_ffi import OracleChain
from trace_util import make_case, run_trace

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _kw(data, od=1, mx=1):
    return dict(D=data["D"], region_sizes=data["region_sizes"], neutral_states=data["neutral_states"],
                cluster_sizes=data["cluster_sizes"], eta=1e-4, is_overdispersed=od, max_scoring=mx)


@pytest.mark.parametrize("n_cells,n_regions", [(3000, 40), (129, 3), (1, 5), (5000, 200)])
def test_root_score_from_histograms_matches_slices_and_oracle(monkeypatch, n_cells, n_regions):
    from scicone_b200.api import Chain

    data, tree = make_case(n_cells, 100 * n_regions, n_regions, 8, seed=300 + n_cells)
    rng = np.random.default_rng(5)
    data["cluster_sizes"] = rng.integers(1, 4, size=n_cells).astype(np.int32)  # weights enter the histograms
    ft = tree.flat()
    monkeypatch.setenv("SCICONE_OD_HIST", "0")
    slices = Chain(**_kw(data))
    monkeypatch.delenv("SCICONE_OD_HIST")
    hist = Chain(**_kw(data))
    cpu = OracleChain(**_kw(data))
    for ch in (slices, hist, cpu):
        ch.compute_t_table(ft)
    for nu in (ft.nu, 0.37, 2.5, 11.0):
        a, b, c = hist.compute_t_od_scores(nu), slices.compute_t_od_scores(nu), cpu.compute_t_od_scores(nu)
        assert rel_err(a, b) <= 1e-12, (nu, a, b)
        assert rel_err(a, c) <= TOL, (nu, a, c)
    hist.close()
    slices.close()


def test_nu_proposal_with_histograms(monkeypatch):
    """A nu move carries its root score in the same launch (PROP_WITH_OD): both forms against the oracle."""
    data, tree = make_case(2000, 3000, 30, 25, seed=41)
    from scicone_b200.api import Chain

    for env in (None, "0"):
        if env is None:
            monkeypatch.delenv("SCICONE_OD_HIST", raising=False)
        else:
            monkeypatch.setenv("SCICONE_OD_HIST", env)
        gpu, cpu = Chain(**_kw(data)), OracleChain(**_kw(data))
        worst, _ = run_trace(gpu, cpu, tree.copy(), 40, TOL, kinds=("nu", "prune_reattach", "add_remove"))
        assert worst["argmax_mismatch"] == 0
        gpu.close()


def test_programmatic_launches_are_bit_identical(monkeypatch):
    from scicone_b200.api import Chain, Dataset, batch_compute_t_prime_sums

    data, tree0 = make_case(2500, 3000, 30, 20, seed=77)
    kinds = ("prune_reattach", "swap", "add_remove", "insert", "delete", "nu")
    B = 6
    out = {}
    for pdl in ("1", "0"):
        monkeypatch.setenv("SCICONE_PDL", pdl)
        ds = Dataset(data["D"], data["region_sizes"], neutral_states=data["neutral_states"], cluster_sizes=data["cluster_sizes"],
                     max_scoring=1)
        chains = [Chain(dataset=ds) for _ in range(B)]
        trees = []
        for b in range(B):
            t = tree0.copy()
            t.rng = np.random.default_rng(1000 + b)
            trees.append(t)
        totals = [ch.compute_t_table(t.flat()) for ch, t in zip(chains, trees)]
        for step in range(40):
            props = [t.propose(kinds=kinds) for t in trees]
            live = [b for b in range(B) if props[b] is not None]
            fts = [props[b][0].flat() for b in live]
            for b, ft in zip(live, fts):
                for rid in props[b][1]:
                    chains[b].compute_t_prime_scores(ft, ft.index_of(rid))
            sums, ods = batch_compute_t_prime_sums([chains[b] for b in live], fts, with_od=True)
            totals += [float(x) for x in sums] + [float(x) for x in ods if x == x]
            for b in live:
                if (step + b) % 2 == 0:
                    chains[b].accept()
                    trees[b] = props[b][0]
                else:
                    chains[b].reject()
        out[pdl] = totals
        for ch in chains:
            ch.close()
        ds.close()
    assert len(out["1"]) == len(out["0"]) and len(out["1"]) > 200
    assert out["1"] == out["0"]
