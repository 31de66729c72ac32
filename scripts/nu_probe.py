#!/usr/bin/env python3
"""One chain on configs[2]: three launches of one nu proposal (whole-tree rescoring + root score), then three of one
prune/reattach proposal - the two launch shapes whose latency bounds a real MCMC launch.  For ncu (GPU box):

    ncu --set full --import-source on -k regex:score_proposals -s 1 -c 6 -o probe python scripts/nu_probe.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scicone_b200.api import BatchLauncher, Chain, Dataset  # noqa: E402
from scicone_b200.synth import SEED, TreeState, make_counts  # noqa: E402

M, K, N = 10000, 200, 50
data = make_counts(M, 18000, K, seed=SEED)
ds = Dataset(data["D"], data["region_sizes"], neutral_states=data["neutral_states"], cluster_sizes=data["cluster_sizes"],
             is_overdispersed=1, max_scoring=1)
ds.set_profiling(True)
tree = TreeState(K, data["neutral_states"], np.random.default_rng(100), nu=1.0).grow(N)
ch = Chain(dataset=ds)
ch.compute_t_table(tree.flat())
L = BatchLauncher([ch])
for kind in ["nu"] * 3 + ["prune_reattach"] * 3:
    while True:
        prop = tree.propose(kinds=(kind,))
        if prop is not None:
            break
    t2, roots, _ = prop
    ft = t2.flat()
    for rid in roots:
        ch.compute_t_prime_scores(ft, ft.index_of(rid))
    ds.kernel_time_stats(reset=True)
    L.run()
    ks, _ = ds.kernel_time_stats()
    bs, _ = ds.build_time_stats()
    print(f"{kind:16s} K2-K4 {ks:6.1f} us  K1 {bs:6.1f} us  rescored {len(ch.read_t_prime_scores()[0])}")
    ch.reject()
