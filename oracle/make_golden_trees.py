#!/usr/bin/env python3
"""Copies the reference's tree-file validity fixtures (scicone/tests/trees_to_validate/*.txt, used by
test_tree_validation, validation.h:999-1063) into tests/golden/trees/.  TEST INFRASTRUCTURE; needs /root/reference,
the outputs are committed (five text files of a few lines each)."""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(os.environ.get("SCICONE_REF", "/root/reference"), "scicone", "tests", "trees_to_validate")
DST = os.path.join(HERE, "..", "tests", "golden", "trees")
os.makedirs(DST, exist_ok=True)
for f in sorted(os.listdir(SRC)):
    shutil.copy(os.path.join(SRC, f), os.path.join(DST, f))
    print(f)
