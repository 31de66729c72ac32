"""Drive two chains (CUDA engine and oracle port) through the same random proposal trace."""
import numpy as np

from golden_util import rel_err, worse
from scicone_b200.synth import TreeState, make_counts


def make_case(n_cells, n_bins, n_regions, n_nodes, seed, cluster_rows=None, nu=1.0):
    data = make_counts(n_cells, n_bins, n_regions, seed=seed, cluster_rows=cluster_rows)
    rng = np.random.default_rng(seed + 1)
    tree = TreeState(n_regions, data["neutral_states"], rng, nu=nu).grow(n_nodes)
    return data, tree


def run_trace(gpu, cpu, tree, n_moves, tol, kinds=None, check_tables_every=10):
    """Returns dict of worst relative errors.  Asserts id sets equal and errors <= tol."""
    worst = {"total": 0.0, "t_prime_scores": 0.0, "t_prime_sums": 0.0, "t_scores": 0.0, "od": 0.0, "argmax_mismatch": 0}
    ft = tree.flat()
    a, b = gpu.compute_t_table(ft), cpu.compute_t_table(ft)
    worst["total"] = worse(worst["total"], rel_err(a, b))
    a, b = gpu.compute_t_od_scores(ft.nu), cpu.compute_t_od_scores(ft.nu)
    worst["od"] = worse(worst["od"], rel_err(a, b))
    worst["t_scores"] = rel_err(gpu.read_t_scores(), cpu.read_t_scores())
    kw = {} if kinds is None else {"kinds": kinds}
    n_done = 0
    for step in range(n_moves):
        prop = tree.propose(**kw)
        if prop is None:
            continue
        t2, roots, kind = prop
        ft = t2.flat()
        if kind == "nu":
            a, b = gpu.compute_t_prime_od_scores(ft.nu), cpu.compute_t_prime_od_scores(ft.nu)
            worst["od"] = worse(worst["od"], rel_err(a, b))
        for rid in roots:
            idx = ft.index_of(rid)
            gpu.compute_t_prime_scores(ft, idx)
            cpu.compute_t_prime_scores(ft, idx)
        a, b = gpu.compute_t_prime_sums(ft), cpu.compute_t_prime_sums(ft)
        worst["total"] = worse(worst["total"], rel_err(a, b))
        gi, gs = gpu.read_t_prime_scores()
        ci, cs = cpu.read_t_prime_scores()
        assert list(gi) == list(ci), (step, kind, gi, ci)
        worst["t_prime_scores"] = worse(worst["t_prime_scores"], rel_err(gs, cs))
        worst["t_prime_sums"] = worse(worst["t_prime_sums"], rel_err(gpu.read_t_prime_sums(), cpu.read_t_prime_sums()))
        if gpu.max_scoring:
            g_ids, _ = gpu.read_t_prime_maxs()
            c_ids, c_vals = cpu.read_t_prime_maxs()
            bad = np.nonzero(g_ids != c_ids)[0]
            if bad.size:
                # a different argmax is only legitimate where two nodes tie within rounding (same genotype reached
                # along different paths): the oracle's own score of the engine's choice must equal the oracle's max
                worst["argmax_ties"] = worst.get("argmax_ties", 0) + int(bad.size)
                ids_t, sc_t = list(cpu.t_node_ids()), cpu.read_t_scores()
                col = {int(i): sc_t[:, n] for n, i in enumerate(ids_t)}
                for n, i in enumerate(ci):
                    col[int(i)] = cs[:, n]
                for j in bad:
                    v = col[int(g_ids[j])][j]
                    if abs(v - c_vals[j]) > tol * max(1.0, abs(c_vals[j])):
                        worst["argmax_mismatch"] += 1
        if tree.rng.random() < 0.5:
            gpu.accept(ft)
            cpu.accept(ft)
            tree = t2
        else:
            gpu.reject()
            cpu.reject()
        n_done += 1
        if n_done % check_tables_every == 0:
            assert list(gpu.t_node_ids()) == list(cpu.t_node_ids())
            worst["t_scores"] = worse(worst["t_scores"], rel_err(gpu.read_t_scores(), cpu.read_t_scores()))
    assert list(gpu.t_node_ids()) == list(cpu.t_node_ids())
    worst["t_scores"] = worse(worst["t_scores"], rel_err(gpu.read_t_scores(), cpu.read_t_scores()))
    for k, v in worst.items():
        if k.startswith("argmax"):
            continue
        # Per-cell aggregates in sum mode go through MathOp::log_replace_sum (MathOp.cpp:307-361), which subtracts
        # the old terms in linear space and keeps the remainder whenever it exceeds 1e-6: the reference's own
        # algorithm amplifies a 1e-13 perturbation of the scores (device vs libm lgamma, a few ulp) by up to 1e6.
        # Scores and the tree totals (the quantity north_star bounds) are held to `tol`.
        bar = 1e-6 if (k == "t_prime_sums" and not gpu.max_scoring) else tol
        assert v <= bar, (k, v)
    worst["n_moves"] = n_done
    return worst, tree
