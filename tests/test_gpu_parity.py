"""CUDA engine (through the C ABI) against the golden traces of the unmodified reference and the oracle port.

Tolerance: the path is fp64; north_star asks for per-move tree log-scores within 1e-9 relative.  Cell-independent
terms are computed with the same libm as the reference, per-cell terms with the device lgamma/exp/log (a few ulp
from glibc), so per-score agreement is ~1e-12 relative in practice; the bar in the asserts is 1e-9.
Node ids, argmax ids and rescored-id sets must match exactly.
"""
import numpy as np
import pytest

from golden_util import chain_kwargs, golden_names, load_golden, rel_err, replay, worse

pytestmark = pytest.mark.gpu

TOL = 1e-9


def _check_factory(worst):
    def check(what, got, want, step):
        if what.endswith("ids") and "maxs" not in what:
            assert list(got) == list(want), (what, step, got, want)
            return
        if what.endswith("maxs.ids"):
            # an argmax may legitimately differ only where two node scores tie within rounding
            worst.setdefault("argmax_mismatch", 0)
            worst["argmax_mismatch"] += int(np.sum(np.asarray(got) != np.asarray(want)))
            return
        e = rel_err(got, want)
        worst[what] = worse(worst.get(what, 0.0), e)
        assert e <= TOL, (what, step, e)

    return check


@pytest.mark.parametrize("name", golden_names())
def test_engine_replays_reference_trace(name):
    from scicone_b200.api import Chain

    doc = load_golden(name)
    chain = Chain(**chain_kwargs(doc))
    worst = {}
    replay(doc, chain, _check_factory(worst))
    assert worst.get("argmax_mismatch", 0) == 0, worst
    print(name, {k: f"{v:.2e}" if isinstance(v, float) else v for k, v in worst.items()})
    chain.close()
