"""Host-level unit tests of the native tree engine (scicone_b200/host/tree.hpp), compiled as tests/native/tree_check.cpp:
the reference's tree-file validity fixtures (scicone/tests/trees_to_validate, copied by oracle/make_golden_trees.py into
tests/golden/trees) with the verdicts of test_tree_validation (validation.h:999-1063), subtree sizes, copy semantics
(sibling-order reversal) and genotype bookkeeping."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tree_engine_known_answers(tmp_path):
    exe = str(tmp_path / "tree_check")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wno-sign-compare", os.path.join(ROOT, "tests", "native", "tree_check.cpp"), "-o", exe])
    res = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "trees")], capture_output=True, text=True)
    print(res.stdout)
    assert res.returncode == 0, res.stdout + res.stderr
