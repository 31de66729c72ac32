"""Size-independent properties at BASELINE.json's full shapes (the oracle cannot replay these sizes in seconds, so it
is only run on a random subset of the cells - every per-cell quantity is independent of the other cells).

configs[2]: 10 000 cells x 200 regions, 50-node tree;  configs[4]: 100 000 cells x 200 regions (one GPU's view of it)."""
import numpy as np
import pytest

from golden_util import rel_err
from oracle_ffi import OracleChain
from scicone_b200.synth import SEED, TreeState, make_counts

pytestmark = pytest.mark.gpu
KINDS = ("prune_reattach", "swap", "add_remove", "insert", "delete", "nu")


def _walk(chains, tree, n_moves, rng_accept):
    """Same random proposal trace on every chain; returns the final tree and the totals of every move."""
    totals = []
    for step in range(n_moves):
        prop = tree.propose(kinds=KINDS)
        if prop is None:
            continue
        t2, roots, kind = prop
        ft = t2.flat()
        row = []
        for ch in chains:
            for rid in roots:
                ch.compute_t_prime_scores(ft, ft.index_of(rid))
            row.append(ch.compute_t_prime_sums(ft))
        totals.append(row)
        if rng_accept.random() < 0.5:
            for ch in chains:
                ch.accept(ft) if isinstance(ch, OracleChain) else ch.accept()
            tree = t2
        else:
            for ch in chains:
                ch.reject()
    return tree, np.array(totals)


@pytest.mark.parametrize("max_scoring", [1, 0])
def test_config2_delta_updates_equal_a_fresh_full_pass_and_the_oracle_on_a_subset(max_scoring):
    from scicone_b200.api import Chain, Dataset

    M, K, N = 10000, 200, 50
    data = make_counts(M, 18000, K, seed=SEED)
    rng = np.random.default_rng(5)
    tree = TreeState(K, data["neutral_states"], rng, nu=1.0).grow(N)
    ds = Dataset(data["D"], data["region_sizes"], neutral_states=data["neutral_states"], cluster_sizes=data["cluster_sizes"],
                 is_overdispersed=1, max_scoring=max_scoring)
    gpu = Chain(dataset=ds)
    subset = np.sort(np.random.default_rng(6).choice(M, 160, replace=False))
    cpu = OracleChain(D=data["D"][subset], region_sizes=data["region_sizes"], neutral_states=data["neutral_states"],
                      cluster_sizes=data["cluster_sizes"][subset], is_overdispersed=1, max_scoring=max_scoring)
    ft = tree.flat()
    gpu.compute_t_table(ft)
    cpu.compute_t_table(ft)
    tree, _ = _walk([gpu, cpu], tree, 40, np.random.default_rng(7))
    # the committed tables after 40 delta updates: per-cell aggregates and scores of the subset against the oracle ...
    assert rel_err(gpu.read_t_sums()[subset], cpu.read_t_sums()) <= 1e-9
    assert list(gpu.t_node_ids()) == list(cpu.t_node_ids())
    assert rel_err(gpu.read_t_scores()[subset], cpu.read_t_scores()) <= 1e-9
    # ... and every cell against a fresh full pass of the final tree (incremental == from scratch)
    fresh = Chain(dataset=ds)
    ft = tree.flat()
    total_fresh = fresh.compute_t_table(ft)
    assert rel_err(gpu.read_t_scores(), fresh.read_t_scores()) <= 1e-9
    assert rel_err(gpu.read_t_sums(), fresh.read_t_sums()) <= 1e-9
    sums = gpu.read_t_sums()
    assert abs(total_fresh - float(np.sum(sums * data["cluster_sizes"]))) <= 1e-9 * abs(total_fresh)
    if max_scoring:  # assignment: first maximum per cell
        ids, _ = gpu.assign_cells_to_nodes(ft, with_cn=False)
        sc = fresh.read_t_scores()
        node_ids = np.array(fresh.t_node_ids())
        assert np.array_equal(np.asarray(ids), node_ids[np.argmax(sc, axis=1)])


def test_config4_shape_weights_scale_the_total_exactly_and_batches_equal_single_launches():
    from scicone_b200.api import Chain, Dataset, batch_compute_t_prime_sums

    M, K = 100000, 200
    base = make_counts(20000, 20000, K, seed=SEED + 4)
    D = np.tile(base["D"], (5, 1))  # 100 000 rows; the engine treats every row as its own cell
    D += np.arange(M)[:, None] % 7  # ... and they are not identical
    w1 = np.ones(M, dtype=np.int32)
    rng = np.random.default_rng(8)
    tree = TreeState(K, base["neutral_states"], rng, nu=1.3).grow(60)
    ds1 = Dataset(D, base["region_sizes"], neutral_states=base["neutral_states"], cluster_sizes=w1, max_scoring=1)
    ds2 = Dataset(D, base["region_sizes"], neutral_states=base["neutral_states"], cluster_sizes=2 * w1, max_scoring=1)
    a, b, c = Chain(dataset=ds1), Chain(dataset=ds2), Chain(dataset=ds1)
    # the oracle on a random subset of the cells (every per-cell quantity is independent of the other cells): it walks the
    # same proposals as chain a and is compared at the end
    subset = np.sort(np.random.default_rng(10).choice(M, 200, replace=False))
    cpu = OracleChain(D=D[subset], region_sizes=base["region_sizes"], neutral_states=base["neutral_states"], cluster_sizes=w1[subset],
                      is_overdispersed=1, max_scoring=1)
    ft = tree.flat()
    ta, tb = a.compute_t_table(ft), b.compute_t_table(ft)
    cpu.compute_t_table(ft)
    assert rel_err(a.read_t_sums()[subset], cpu.read_t_sums()) <= 1e-9
    assert tb == 2.0 * ta  # a power-of-two weight scales every term and every partial sum exactly
    assert b.compute_t_od_scores(ft.nu) == 2.0 * a.compute_t_od_scores(ft.nu)
    c.compute_t_table(ft)
    other = tree.copy()
    other.rng = np.random.default_rng(9)
    for step in range(6):
        t2, roots, kind = tree.propose(kinds=KINDS)
        f2 = t2.flat()
        for ch in (a, b, cpu):
            for rid in roots:
                ch.compute_t_prime_scores(f2, f2.index_of(rid))
        cpu.compute_t_prime_sums(f2)
        # chain c gets a different proposal; a and c share a dataset and go out in ONE launch
        t3, roots3, _ = other.propose(kinds=KINDS)
        f3 = t3.flat()
        for rid in roots3:
            c.compute_t_prime_scores(f3, f3.index_of(rid))
        tot = batch_compute_t_prime_sums([a, c], [f2, f3])
        tb = b.compute_t_prime_sums(f2)
        assert tb == 2.0 * tot[0]
        c.reject()
        for rid in roots3:  # the same proposal alone in a launch gives the same bits
            c.compute_t_prime_scores(f3, f3.index_of(rid))
        assert c.compute_t_prime_sums(f3) == tot[1]
        c.reject()
        # the proposal's per-cell aggregates and argmax ids on the subset against the oracle, before it is settled
        assert rel_err(a.read_t_prime_sums()[subset], cpu.read_t_prime_sums()) <= 1e-9
        assert rel_err(np.asarray(a.read_t_prime_maxs()[1])[subset], cpu.read_t_prime_maxs()[1]) <= 1e-9  # (ids may differ at exact ties)
        for ch in (a, b):
            ch.accept() if step % 2 == 0 else ch.reject()
        cpu.accept(f2) if step % 2 == 0 else cpu.reject()
        if step % 2 == 0:
            tree = t2
    # the committed tables after the walk: scores of the subset against the oracle
    assert list(a.t_node_ids()) == list(cpu.t_node_ids())
    assert rel_err(a.read_t_scores()[subset], cpu.read_t_scores()) <= 1e-9
    assert rel_err(a.read_t_sums()[subset], cpu.read_t_sums()) <= 1e-9


@pytest.mark.parametrize("M", [10000, 2500])
def test_tabulated_lgamma_rows_equal_direct_evaluation_bit_for_bit(M, monkeypatch):
    """K1 evaluates lgamma(count + c) once per possible count of a block of cells when the counts are integers in a
    narrow range (the dataset's count-range table) - the same function of the same argument, so every total and every
    score must be IDENTICAL to a dataset created with SCICONE_LUT=0 (direct evaluation per cell), nu proposals (all rows
    rebuilt, root-score slices) included.  A matrix with some non-integer and some widely spread regions mixes both
    paths inside one launch."""
    from scicone_b200.api import Chain, Dataset

    K, N = 200, 30
    data = make_counts(M, 18000, K, seed=SEED + 3)
    D = data["D"].copy()
    if M == 10000:  # (the other case keeps integer counts everywhere: the row sums S are tabulated too)
        D[:, 5] += 0.5                  # not integers: direct
        D[::7, 11] += 40000.0           # range too wide: direct
    kw = dict(neutral_states=data["neutral_states"], cluster_sizes=data["cluster_sizes"], is_overdispersed=1, max_scoring=1)
    # the count tables are what is compared here: the root score takes the per-cell slices on both sides (its histogram
    # form - which a dataset without count ranges does not have - is held against the slices in tests/test_gpu_od_hist.py)
    monkeypatch.setenv("SCICONE_OD_HIST", "0")
    monkeypatch.setenv("SCICONE_LUT", "1")
    ds_tab = Dataset(D, data["region_sizes"], **kw)
    monkeypatch.setenv("SCICONE_LUT", "0")
    ds_dir = Dataset(D, data["region_sizes"], **kw)
    monkeypatch.delenv("SCICONE_LUT")
    a, b = Chain(dataset=ds_tab), Chain(dataset=ds_dir)
    tree = TreeState(K, data["neutral_states"], np.random.default_rng(15), nu=1.0).grow(N)
    ft = tree.flat()
    assert a.compute_t_table(ft) == b.compute_t_table(ft)
    assert a.compute_t_od_scores(ft.nu) == b.compute_t_od_scores(ft.nu)
    tree, totals = _walk([a, b], tree, 60, np.random.default_rng(16))
    assert len(totals) > 30 and np.array_equal(totals[:, 0], totals[:, 1])
    assert np.array_equal(a.read_t_scores(), b.read_t_scores())
    assert np.array_equal(a.read_t_sums(), b.read_t_sums())
