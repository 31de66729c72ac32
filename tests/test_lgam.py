"""The device log-gamma (scicone_b200/csrc/lgam_core.h): table-driven logarithm + Stirling series for arguments >= 16.

CPU half: the same header compiled for the host (tests/native/lgam_check.cpp) against glibc's log / lgamma - the routine
the reference calls (Tree.h:214-231) - over 2 M log-spaced arguments: <= 1 ulp for the logarithm, <= 2 ulp for lgamma.
GPU half: the device build through the test hook scicone_test_lgamma against glibc's lgamma."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_build_of_device_lgamma_matches_glibc(tmp_path):
    exe = str(tmp_path / "lgam_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-mfma", "-ffp-contract=off", os.path.join(ROOT, "tests", "native", "lgam_check.cpp"), "-o", exe])
    rep = json.loads(subprocess.check_output([exe, "2000000"]).decode())
    print(rep)
    assert rep["log_max_ulp"] <= 1.0
    assert rep["lgamma_max_ulp"] <= 2.0
    assert rep["lgamma_bit_exact_frac"] >= 0.99


def _ulps(got, want):
    m, e = np.frexp(want)
    return np.abs(got - want) / np.ldexp(1.0, e - 53)


def _libm_lgamma(x):
    """glibc's lgamma, the routine the reference binary calls (CPython's math.lgamma is a different implementation)."""
    import ctypes

    libm = ctypes.CDLL("libm.so.6")
    libm.lgamma.restype = ctypes.c_double
    libm.lgamma.argtypes = [ctypes.c_double]
    return np.array([libm.lgamma(float(v)) for v in x])


@pytest.mark.gpu
def test_device_lgamma_matches_libm():
    from scicone_b200.api import device_lgamma

    rng = np.random.default_rng(7)
    x = np.concatenate([np.exp(rng.uniform(np.log(16.0), np.log(1e9), 400000)), np.floor(rng.uniform(16, 5e5, 200000)),
                        np.floor(rng.uniform(16, 5e4, 100000)) + 0.5, np.array([16.0, 16.000000001, 1e12, 1e15])])
    want = _libm_lgamma(x)
    got = device_lgamma(x)
    u = _ulps(got, want)
    print("large arguments: max ulp", u.max(), "bit-exact", float((got == want).mean()))
    assert u.max() <= 2.0
    # below 16 the CUDA library routine is used (regions without reads: eta-sized arguments, Tree.h:194-195)
    xs = np.concatenate([rng.uniform(1e-4, 16.0, 100000), np.array([1e-4, 2e-4, 1.0, 2.0, 15.999999])])
    wants = _libm_lgamma(xs)
    gots = device_lgamma(xs)
    err = np.abs(gots - wants) / np.maximum(1.0, np.abs(wants))
    print("small arguments: max abs/rel error", err.max())
    assert err.max() <= 1e-14
