#!/usr/bin/env python3
"""Debug helper (GPU box): run one trajectory case through every host/backend pair and print where they diverge."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from host_util import CPU_BACKEND_DIR, NATIVE, REF, REF_ON_ABI, run_inference  # noqa: E402
from test_host_trajectory import CASES  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "alpha_gamma"
data, extra = CASES[case]
D, r, w = data()
extra = extra + ["--print_precision=17"]
tmp = tempfile.mkdtemp()
runs = {"ref": run_inference(REF, tmp, "ref", D, r, extra, w=w),
        "native_cpu": run_inference(NATIVE, tmp, "ncpu", D, r, extra, w=w, env={"LD_LIBRARY_PATH": CPU_BACKEND_DIR})}
try:
    runs["ref_on_abi"] = run_inference(REF_ON_ABI, tmp, "rabi", D, r, extra, w=w)
    runs["native_gpu"] = run_inference(NATIVE, tmp, "ngpu", D, r, extra, w=w)
except AssertionError as e:
    print("gpu runs failed:", str(e)[-500:])
base = runs["ref"]
for name, o in runs.items():
    if name == "ref":
        continue
    n = min(len(base["acc"]), len(o["acc"]))
    d = np.nonzero(base["acc"][:n] != o["acc"][:n])[0]
    rel = np.abs(base["chain"][:n] - o["chain"][:n]) / np.maximum(1.0, np.abs(base["chain"][:n]))
    first_rel = np.nonzero(rel > 1e-9)[0]
    print(name, "first acc divergence:", d[:1], "first score divergence:", first_rel[:1], "max rel before:",
          rel[: (d[0] if len(d) else n)].max() if n else None)
    if len(d):
        i = d[0]
        for k in range(max(0, i - 3), min(n, i + 3)):
            print("   it", k, "acc", base["acc"][k], o["acc"][k], "chain %.17g %.17g" % (base["chain"][k], o["chain"][k]),
                  "gamma", base["gamma"][k], o["gamma"][k])
