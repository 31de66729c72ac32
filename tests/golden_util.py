"""Loading and replaying the golden traces under tests/golden (made by oracle/make_golden.py)."""
import glob
import gzip
import json
import os

import numpy as np

from scicone_b200.flat_tree import FlatTree

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(p)[: -len(".json.gz")] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.json.gz")))


def load_golden(name):
    with gzip.open(os.path.join(GOLDEN_DIR, name + ".json.gz"), "rt") as f:
        return json.load(f)


def subtree_roots(tree, rescored_ids):
    """Entries handed to compute_t_prime_scores: rescored nodes whose parent was not rescored."""
    rs = set(rescored_ids)
    roots = []
    for i, nid in enumerate(tree.ids):
        if nid in rs and (tree.parent[i] < 0 or tree.ids[tree.parent[i]] not in rs):
            roots.append(i)
    return roots


def rel_err(a, b):
    """Largest relative error over the FINITE entries.  Non-finite entries (an int-truncated Node::z of 0 gives
    lgamma(0) = inf and inf - inf = NaN on both sides, Tree.h:227-231) must sit at exactly the same positions and be of
    the same kind (NaN / +inf / -inf) on both sides; otherwise the error is infinite, so a NaN that only one side
    produces can never pass a tolerance check."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        a, b = np.broadcast_arrays(a, b)
    if a.size == 0:
        return 0.0
    fa, fb = np.isfinite(a), np.isfinite(b)
    if not np.array_equal(fa, fb):
        return float("inf")
    if not fa.all():
        na, nb = a[~fa], b[~fa]
        if not np.array_equal(np.isnan(na), np.isnan(nb)) or not np.array_equal(na[~np.isnan(na)], nb[~np.isnan(nb)]):
            return float("inf")
        if not fa.any():
            return 0.0
        a, b = a[fa], b[fa]
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))


def worse(worst, err):
    """max() that cannot lose a NaN: Python's max(0.0, nan) is 0.0."""
    return err if not (err <= worst) else worst


def replay(doc, chain, check, od_of=None):
    """Drive `chain` (OracleChain or scicone_b200.api.Chain) through a golden trace.

    `check(what, got, want, step)` is called for every compared quantity.
    """
    t = FlatTree.from_json(doc["init"]["tree"])
    total = chain.compute_t_table(t)
    check("init.total", total, doc["init"]["total"], -1)
    od = chain.compute_t_od_scores(t.nu)
    check("init.od_score", od, doc["init"]["od_score"], -1)
    check("init.ids", list(chain.t_node_ids()), doc["init"]["t_scores"]["ids"], -1)
    check("init.t_scores", chain.read_t_scores(), np.array(doc["init"]["t_scores"]["values"]), -1)
    check("init.t_sums", chain.read_t_sums(), doc["init"]["t_sums"], -1)
    if doc["max_scoring"]:
        ids, vals = chain.read_t_maxs()
        check("init.t_maxs.values", vals, doc["init"]["t_maxs"]["values"], -1)
        check("init.t_maxs.ids", list(ids), doc["init"]["t_maxs"]["ids"], -1)
    for step, mv in enumerate(doc["moves"]):
        tp = FlatTree.from_json(mv["tree"])
        if mv["move_id"] == 13:
            odp = chain.compute_t_prime_od_scores(tp.nu)
            check("move.od_score", odp, mv["od_score"], step)
        for idx in subtree_roots(tp, mv["t_prime_scores"]["ids"]):
            chain.compute_t_prime_scores(tp, idx)
        total = chain.compute_t_prime_sums(tp)
        ids, sc = chain.read_t_prime_scores()
        check("move.rescored_ids", list(ids), mv["t_prime_scores"]["ids"], step)
        check("move.t_prime_scores", sc, np.array(mv["t_prime_scores"]["values"]), step)
        check("move.t_prime_sums", chain.read_t_prime_sums(), mv["t_prime_sums"], step)
        if doc["max_scoring"]:
            ids, vals = chain.read_t_prime_maxs()
            check("move.t_prime_maxs.values", vals, mv["t_prime_maxs"]["values"], step)
            check("move.t_prime_maxs.ids", list(ids), mv["t_prime_maxs"]["ids"], step)
        check("move.total", total, mv["total"], step)
        if mv["accepted"]:
            chain.accept(tp)
        else:
            chain.reject()
    check("final.ids", list(chain.t_node_ids()), doc["final"]["t_scores"]["ids"], len(doc["moves"]))
    check("final.t_scores", chain.read_t_scores(), np.array(doc["final"]["t_scores"]["values"]), len(doc["moves"]))
    check("final.t_sums", chain.read_t_sums(), doc["final"]["t_sums"], len(doc["moves"]))


def chain_kwargs(doc):
    return dict(D=np.array(doc["D"], dtype=np.float64), region_sizes=doc["region_sizes"], neutral_states=doc["neutral_states"],
                cluster_sizes=doc["cluster_sizes"], eta=doc["eta"], is_overdispersed=doc["is_overdispersed"],
                max_scoring=doc["max_scoring"])
