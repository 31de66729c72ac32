"""K0, bins -> regions (Utils::condense_matrix, reference scicone/src/Utils.cpp:104-138).

CPU: the oracle's restatement against the reference's own tutorial pair (tests/golden/tutorial_bins.npz: bundled bins
matrix and the segmented counts the reference pipeline made of it).  GPU: the TMA streaming kernel through the C ABI,
bit-exact against the golden pair and against the oracle on ragged shapes (odd bin counts, regions wider than a tile,
single-bin regions, unused trailing bins, non-integer values), and a size-independent property at config scale."""
import os

import numpy as np
import pytest

from oracle_ffi import condense_matrix as oracle_condense

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "tutorial_bins.npz"))
    return z["bins"].astype(np.float64), z["regions"], z["region_sizes"]


def test_oracle_matches_reference_pair():
    bins, regions, sizes = _golden()
    assert np.array_equal(oracle_condense(bins, sizes), regions)


def test_oracle_ignores_trailing_bins_and_adds_in_bin_order():
    rng = np.random.default_rng(3)
    bins = rng.normal(size=(5, 40)) * 10.0 ** rng.integers(-8, 8, size=(5, 40))
    sizes = np.array([7, 1, 20, 5], dtype=np.int32)  # 33 of 40 bins used
    got = oracle_condense(bins, sizes)
    off = np.concatenate([[0], np.cumsum(sizes)])
    for i in range(5):
        for r in range(4):
            acc = 0.0
            for b in range(off[r], off[r + 1]):
                acc += bins[i, b]
            assert got[i, r] == acc


def _cases():
    rng = np.random.default_rng(11)
    yield "odd_bins", rng.poisson(20, size=(37, 1001)).astype(np.float64), np.array([1, 500, 3, 297, 200], dtype=np.int32)
    yield "wide_region", rng.poisson(5, size=(9, 5000)).astype(np.float64), np.array([10, 3000, 1024, 966], dtype=np.int32)
    yield "single_row", rng.poisson(5, size=(1, 2049)).astype(np.float64), np.array([2048, 1], dtype=np.int32)
    yield "trailing_unused", rng.poisson(9, size=(20, 777)).astype(np.float64), np.array([100, 200, 300], dtype=np.int32)
    yield "many_small", rng.poisson(9, size=(33, 3000)).astype(np.float64), np.full(3000, 1, dtype=np.int32)
    floats = rng.normal(size=(17, 4099)) * 10.0 ** rng.integers(-6, 6, size=(17, 4099))
    yield "non_integer", floats, np.array([4099 // 7] * 7, dtype=np.int32)


@pytest.mark.gpu
def test_gpu_condense_golden_pair():
    from scicone_b200.api import condense_matrix

    bins, regions, sizes = _golden()
    assert np.array_equal(condense_matrix(bins, sizes), regions)


@pytest.mark.gpu
@pytest.mark.parametrize("name", [c[0] for c in _cases()])
def test_gpu_condense_bit_exact_vs_oracle(name):
    from scicone_b200.api import condense_matrix

    bins, sizes = next((b, s) for n, b, s in _cases() if n == name)
    got = condense_matrix(bins, sizes)
    want = oracle_condense(bins, sizes)
    assert np.array_equal(got, want), np.abs(got - want).max()


@pytest.mark.gpu
def test_gpu_condense_config_scale_properties():
    """configs[2] scale (10 000 x 18 000 -> 200): row totals are preserved exactly (integer counts) and a second collapse of
    the result into one region equals the row total; also reports the kernel's bandwidth."""
    from scicone_b200.api import condense_matrix
    from scicone_b200.synth import SEED, region_sizes

    rng = np.random.default_rng(SEED)
    sizes = region_sizes(rng, 18000, 200)
    bins = rng.poisson(10.0, size=(10000, 18000)).astype(np.float64)
    got, us = condense_matrix(bins, sizes, return_kernel_us=True)
    assert np.array_equal(got.sum(axis=1), bins.sum(axis=1))
    sample = rng.choice(10000, 64, replace=False)
    assert np.array_equal(got[sample], oracle_condense(bins[sample], sizes))
    again = condense_matrix(got, np.array([200], dtype=np.int32))
    assert np.array_equal(again[:, 0], bins.sum(axis=1))
    print("condense 10000 x 18000: kernel %.1f us, %.0f GB/s" % (us, bins.nbytes / us / 1e3))


# ---- cluster condensation (reference scripts/condense_clusters.py:28-38; pyscicone scicone.py:331-335) -----------------
def _cluster_golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "tutorial_clusters.npz"))
    return z["regions"], z["assignments"], z["condensed"], z["cluster_sizes"]


def test_oracle_cluster_means_match_reference_files():
    """tests/golden/tutorial_clusters.npz: the reference's condensed matrix and cluster sizes of the tutorial data."""
    from oracle_ffi import condense_clusters as oracle_clusters

    regions, assign, condensed, sizes = _cluster_golden()
    means, n = oracle_clusters(regions, assign, len(sizes))
    assert np.array_equal(n, sizes)
    assert np.array_equal(means, condensed)


@pytest.mark.gpu
def test_gpu_cluster_means_bit_exact():
    from oracle_ffi import condense_clusters as oracle_clusters
    from scicone_b200.api import condense_clusters

    regions, assign, condensed, sizes = _cluster_golden()
    means, n = condense_clusters(regions, assign)
    assert np.array_equal(n, sizes) and np.array_equal(means, condensed)
    rng = np.random.default_rng(21)  # non-integer values, an empty cluster, ragged sizes
    D = rng.normal(size=(5003, 37)) * 10.0 ** rng.integers(-5, 5, size=(5003, 37))
    a = rng.integers(0, 40, size=5003).astype(np.int32)
    a[a == 17] = 16
    got, gn = condense_clusters(D, a, n_clusters=41)
    want, wn = oracle_clusters(D, a, 41)
    assert np.array_equal(gn, wn) and np.array_equal(got, want)
