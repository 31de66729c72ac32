#!/usr/bin/env python3
"""K0 (bins -> regions on the device) on the shapes of BASELINE.json: kernel time from CUDA events, fraction of the measured
copy bandwidth, result compared with numpy.  Usage: python scripts/k0_bench.py [rows bins regions]..."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scicone_b200.api import condense_matrix  # noqa: E402
from scicone_b200.synth import region_sizes  # noqa: E402

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
shapes = [(10000, 18000, 200), (20000, 20000, 200), (1000, 10000, 50), (3001, 9999, 77)]
if len(sys.argv) > 3:
    a = [int(x) for x in sys.argv[1:]]
    shapes = [tuple(a[i:i + 3]) for i in range(0, len(a), 3)]
for rows, bins, K in shapes:
    rng = np.random.default_rng(rows + bins)
    r = region_sizes(rng, bins, K)
    D = rng.integers(0, 30, size=(rows, bins)).astype(np.float64)
    off = np.concatenate([[0], np.cumsum(r)])[:-1]
    want = np.add.reduceat(D, off, axis=1)
    us = []
    for _ in range(4):
        got, t = condense_matrix(D, r, return_kernel_us=True)
        us.append(t)
    ok = bool(np.array_equal(got, want))
    best = min(us[1:])
    gbs = 8.0 * rows * (bins + K) / (best * 1e-6) / 1e9
    print(json.dumps({"rows": rows, "bins": bins, "regions": K, "kernel_us": best, "all_us": us, "GB/s": gbs, "frac_of_copy_bw": gbs / peak, "equal_to_numpy": ok}))
