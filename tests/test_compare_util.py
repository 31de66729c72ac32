"""The comparison helpers of the parity tests themselves: a NaN or an infinity that only ONE side produces must never
pass a tolerance check (Python's max(0.0, nan) is 0.0 and numpy's max propagates NaN only if nothing masks it)."""
import numpy as np

from golden_util import rel_err, worse


def test_rel_err_on_finite_values():
    assert rel_err([1.0, 2.0], [1.0, 2.0]) == 0.0
    assert abs(rel_err([100.0], [101.0]) - 1.0 / 101.0) < 1e-15
    assert rel_err([], []) == 0.0
    assert rel_err(3.0, 3.0) == 0.0


def test_one_sided_non_finite_values_are_an_infinite_error():
    nan, inf = float("nan"), float("inf")
    assert rel_err([1.0, nan], [1.0, 2.0]) == inf
    assert rel_err([1.0, 2.0], [1.0, nan]) == inf
    assert rel_err([inf], [1e308]) == inf
    assert rel_err([inf], [-inf]) == inf
    assert rel_err([nan], [inf]) == inf
    assert rel_err(nan, 1.0) == inf


def test_agreeing_non_finite_values_are_skipped_and_the_rest_is_compared():
    nan, inf = float("nan"), float("inf")
    assert rel_err([nan, 1.0, -inf], [nan, 1.0, -inf]) == 0.0
    assert abs(rel_err([nan, 10.0], [nan, 11.0]) - 1.0 / 11.0) < 1e-15
    assert rel_err([nan, nan], [nan, nan]) == 0.0


def test_worse_keeps_nan_and_inf():
    assert worse(0.0, 1e-12) == 1e-12
    assert worse(1e-9, 1e-12) == 1e-9
    assert worse(0.0, float("inf")) == float("inf")
    assert np.isnan(worse(0.0, float("nan")))
    assert not (worse(0.0, float("nan")) <= 1e-9)  # and a NaN fails the bar
