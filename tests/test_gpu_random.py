"""CUDA engine vs the oracle port on seeded random proposal traces at sizes the oracle finishes in seconds,
including the edge shapes (one cell, one region, cell counts that are not a multiple of the CTA width,
non-integer cluster means with cluster_sizes)."""
import numpy as np
import pytest

from oracle_ffi import OracleChain
from trace_util import make_case, run_trace

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _pair(data, od, mx):
    from scicone_b200.api import Chain

    kw = dict(D=data["D"], region_sizes=data["region_sizes"], neutral_states=data["neutral_states"],
              cluster_sizes=data["cluster_sizes"], eta=1e-4, is_overdispersed=od, max_scoring=mx)
    return Chain(**kw), OracleChain(**kw)


@pytest.mark.parametrize("od,mx", [(1, 0), (1, 1), (0, 0), (0, 1)])
def test_random_trace_3000x40(od, mx):
    data, tree = make_case(3000, 4000, 40, 30, seed=100 + 2 * od + mx)
    gpu, cpu = _pair(data, od, mx)
    kinds = ("prune_reattach", "swap", "add_remove", "insert", "delete") + (("nu",) if od else ())
    worst, tree = run_trace(gpu, cpu, tree, 80, TOL, kinds=kinds)
    print(worst)
    assert worst["argmax_mismatch"] == 0
    ft = tree.flat()
    gi, gcn = gpu.assign_cells_to_nodes(ft)
    ci, ccn = cpu.assign_cells_to_nodes(ft)
    assert np.array_equal(gi, ci) and np.array_equal(gcn, ccn)
    gpu.close()


@pytest.mark.parametrize("n_cells,n_regions", [(1, 5), (127, 3), (129, 1), (1025, 7)])
def test_ragged_shapes(n_cells, n_regions):
    data, tree = make_case(n_cells, 100 * n_regions, n_regions, 6, seed=7 + n_cells)
    for od, mx in [(1, 0), (1, 1)]:
        gpu, cpu = _pair(data, od, mx)
        worst, _ = run_trace(gpu, cpu, tree.copy(), 25, TOL)
        assert worst["argmax_mismatch"] == 0
        gpu.close()


def test_cluster_means_with_sizes():
    """cluster-tree mode: non-integer D rows weighted by cluster_sizes (reference test_cluster_scoring, validation.h:1237-1282)."""
    data, tree = make_case(2000, 3000, 30, 10, seed=55, cluster_rows=40)
    assert not np.allclose(data["D"], np.round(data["D"]))
    for mx in (0, 1):
        gpu, cpu = _pair(data, 1, mx)
        worst, _ = run_trace(gpu, cpu, tree.copy(), 60, TOL)
        assert worst["argmax_mismatch"] == 0
        gpu.close()


def test_cluster_sizes_equal_duplicated_cells():
    """validation.h:1237-1282: weights {2,1,3} give the same tree score as duplicating the rows."""
    from scicone_b200.api import Chain

    data, tree = make_case(3, 300, 6, 5, seed=9)
    w = np.array([2, 1, 3], dtype=np.int32)
    ft = tree.flat()
    a = Chain(D=data["D"], region_sizes=data["region_sizes"], cluster_sizes=w)
    b = Chain(D=np.repeat(data["D"], w, axis=0), region_sizes=data["region_sizes"])
    ta, tb = a.compute_t_table(ft), b.compute_t_table(ft)
    assert abs(ta - tb) <= 1e-12 * max(1.0, abs(tb))
    oa, ob = a.compute_t_od_scores(ft.nu), b.compute_t_od_scores(ft.nu)
    assert abs(oa - ob) <= 1e-12 * abs(ob)


def test_batched_chains_match_single_launches():
    """scicone_batch_compute_t_prime_sums: B chains in one launch == B separate launches, bit for bit."""
    from scicone_b200.api import Chain, Dataset, batch_compute_t_prime_sums

    data, tree0 = make_case(2500, 3000, 30, 12, seed=77)
    ds = Dataset(data["D"], data["region_sizes"], neutral_states=data["neutral_states"], cluster_sizes=data["cluster_sizes"])
    B = 5
    trees = []
    for b in range(B):
        t = tree0.copy()
        t.rng = np.random.default_rng(1000 + b)
        trees.append(t)
    batch = [Chain(dataset=ds) for _ in range(B)]
    single = [Chain(dataset=ds) for _ in range(B)]
    for b in range(B):
        ft = trees[b].flat()
        assert batch[b].compute_t_table(ft) == single[b].compute_t_table(ft)
    for step in range(20):
        props = [trees[b].propose(kinds=("prune_reattach", "swap", "add_remove", "insert", "delete")) for b in range(B)]
        fts = [p[0].flat() for p in props]
        for b in range(B):
            for rid in props[b][1]:
                batch[b].compute_t_prime_scores(fts[b], fts[b].index_of(rid))
                single[b].compute_t_prime_scores(fts[b], fts[b].index_of(rid))
        totals = batch_compute_t_prime_sums(batch, fts)
        for b in range(B):
            assert totals[b] == single[b].compute_t_prime_sums(fts[b])
            assert np.array_equal(batch[b].read_t_prime_sums(), single[b].read_t_prime_sums())
            if (step + b) % 2 == 0:
                batch[b].accept()
                single[b].accept()
                trees[b] = props[b][0]
            else:
                batch[b].reject()
                single[b].reject()


def test_nu_proposal_root_score_comes_with_the_launch():
    """Inference::apply_overdispersion_change (Inference.h:916-954): od score at the proposed nu + full rescoring;
    the batch call returns both from one launch and they equal the two-call sequence."""
    from scicone_b200.api import Chain, Dataset, batch_compute_t_prime_sums

    data, tree = make_case(1500, 2000, 20, 8, seed=91)
    ds = Dataset(data["D"], data["region_sizes"], neutral_states=data["neutral_states"], cluster_sizes=data["cluster_sizes"])
    a, b = Chain(dataset=ds), Chain(dataset=ds)
    cpu = OracleChain(D=data["D"], region_sizes=data["region_sizes"], neutral_states=data["neutral_states"],
                      cluster_sizes=data["cluster_sizes"])
    ft = tree.flat()
    for ch in (a, b, cpu):
        ch.compute_t_table(ft)
    for step in range(6):
        t2, roots, kind = tree.propose(kinds=("nu", "add_remove"), weights=[0.7, 0.3])
        f2 = t2.flat()
        od_ref = b.compute_t_prime_od_scores(f2.nu) if kind == "nu" else None
        for rid in roots:
            for ch in (a, b, cpu):
                ch.compute_t_prime_scores(f2, f2.index_of(rid))
        totals, ods = batch_compute_t_prime_sums([a], [f2], with_od=True)
        tb, tc = b.compute_t_prime_sums(f2), cpu.compute_t_prime_sums(f2)
        assert totals[0] == tb
        assert abs(tb - tc) <= 1e-9 * max(1.0, abs(tc))
        if kind == "nu":
            assert ods[0] == od_ref
            oc = cpu.compute_t_prime_od_scores(f2.nu)
            assert abs(od_ref - oc) <= 1e-9 * abs(oc)
        else:
            assert np.isnan(ods[0])
        if step % 2 == 0:
            for ch in (a, b):
                ch.accept()
            cpu.accept(f2)
            tree = t2
        else:
            for ch in (a, b, cpu):
                ch.reject()


def test_errors_are_reported_not_thrown():
    from scicone_b200.api import Chain, SciconeError

    data, tree = make_case(10, 300, 4, 4, seed=3)
    ch = Chain(D=data["D"], region_sizes=data["region_sizes"])
    ft = tree.flat()
    with pytest.raises(SciconeError) as e:  # proposal before the table exists
        ch.compute_t_prime_scores(ft, 1)
    assert e.value.code == -3
    ch.compute_t_table(ft)
    with pytest.raises(SciconeError) as e:  # sums without scores
        ch.compute_t_prime_sums(ft)
    assert e.value.code == -3
    with pytest.raises(SciconeError) as e:
        ch.compute_t_prime_scores(ft, 99)
    assert e.value.code == -1
    ch.reject()
    with pytest.raises(SciconeError):  # region sizes / matrix mismatch == reference std::logic_error (Tree.h:170-171)
        Chain(D=data["D"], region_sizes=[1, 2, 3])
