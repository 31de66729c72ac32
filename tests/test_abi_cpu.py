"""The C-ABI library: loads, exports every symbol include/scicone_score.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from scicone_b200 import build
    from scicone_b200 import api

    build.build_score_lib()
    return api.lib(), api


def test_header_symbols_are_exported():
    L, api = _lib()
    header = open(os.path.join(ROOT, "include", "scicone_score.h")).read()
    declared = set(re.findall(r"\b(scicone_[a-z_0-9]+)\s*\(", header))
    assert declared == set(api.ABI_SYMBOLS), declared ^ set(api.ABI_SYMBOLS)
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert L.scicone_abi_version() == 2


def test_no_cpu_fallback():
    """Without a CUDA device the dataset cannot be created: there is no CPU path to fall back to."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    _, api = _lib()
    with pytest.raises(api.SciconeError) as e:
        api.Dataset(np.ones((4, 3)), [1, 2, 3])
    assert e.value.code == -2


def test_product_does_not_reference_the_oracle():
    """Nothing under scicone_b200/ may import, link or load oracle/."""
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "scicone_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"scicone_oracle|oracle_ffi|orc_[a-z_]+\(|oracle/_ref", txt):
                    bad.append(f)
    assert not bad, bad
