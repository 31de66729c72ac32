"""The CPU restatement (oracle/scicone_oracle.c) against outputs of the unmodified reference.

The fixtures were produced by oracle/make_golden.py from oracle/_ref/ref_dump, which links the
reference's own Inference/Tree/MathOp classes.  Same libm, same operation order => the port must agree
to the last bit on scores and aggregates; we allow 1e-13 relative to stay robust to a libm update.
"""
import numpy as np
import pytest

from golden_util import chain_kwargs, golden_names, load_golden, rel_err, replay, worse
from oracle_ffi import OracleChain

TOL = 1e-13


@pytest.mark.parametrize("name", golden_names())
def test_port_replays_reference_trace(name):
    doc = load_golden(name)
    chain = OracleChain(**chain_kwargs(doc))
    worst = {}

    def check(what, got, want, step):
        if what.endswith("ids"):
            assert list(got) == list(want), (what, step, got, want)
            return
        e = rel_err(got, want)
        worst[what] = worse(worst.get(what, 0.0), e)
        assert e <= TOL, (what, step, e)

    replay(doc, chain, check)
    assert worst, "nothing compared"


def test_known_answers_of_reference_unit_tests():
    """scicone/tests/validation.h: test_tree_attachment (:684-732) -- non-overdispersed worked example."""
    doc = load_golden("worked_od0_max0")
    lse = [4.364, 3.668, 29.222, 8.098, 10.497]  # validation.h:704
    mx = [4.327, 3.520, 29.222, 7.684, 10.464]  # validation.h:722
    assert np.allclose(doc["init"]["t_sums"], lse, atol=1e-3)
    chain = OracleChain(**chain_kwargs(doc))
    from scicone_b200.flat_tree import FlatTree

    t = FlatTree.from_json(doc["init"]["tree"])
    chain.compute_t_table(t)
    assert np.allclose(chain.read_t_sums(), lse, atol=1e-3)
    sc = chain.read_t_scores()
    assert np.allclose(sc.max(axis=1), mx, atol=1e-3)
    # root-score sum of test_prune_reattach (validation.h:734-835): -1510.724
    assert abs(chain.compute_od_scores(t.nu) - (-1510.724)) < 1e-3
    # overdispersed branch, nu = 2 (test_overdispersed_score, validation.h:1082-1162)
    doc = load_golden("worked_od1_max0")
    chain = OracleChain(**chain_kwargs(doc))
    t = FlatTree.from_json(doc["init"]["tree"])
    chain.compute_t_table(t)
    total, per_cell = chain.compute_od_scores(2.0, per_cell=True)
    assert np.allclose(per_cell, [-323.687, -233.387, -391.958, -307.943, -219.915], atol=1e-3)
    assert abs(total - (-1476.89)) < 1e-2


def test_tutorial_stored_tree_scores():
    """docs/example_data/test_tree_inferred.txt: Root score and (re-derived) log likelihood, SURVEY.md App. C."""
    doc = load_golden("tutorial_full_sum")
    assert abs(doc["init"]["od_score"] - (-3840283.93784554)) < 1e-8 * 3840283
    assert abs(doc["init"]["total"] - 22009.4988940134) < 1e-9 * 22009
