#!/usr/bin/env python3
"""Cost model of the proposal kernel: device time of launches with controlled composition (GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scicone_b200.api import BatchLauncher, Chain, Dataset  # noqa: E402
from scicone_b200.synth import SEED, TreeState, make_counts  # noqa: E402

M, K, N, B = 10000, 200, 50, 32
data = make_counts(M, 18000, K, seed=SEED)
ds = Dataset(data["D"], data["region_sizes"], neutral_states=data["neutral_states"], cluster_sizes=data["cluster_sizes"],
             is_overdispersed=1, max_scoring=1)
ds.set_profiling(True)
trees, chains = [], []
for b in range(B):
    rng = np.random.default_rng(100 + b)
    t = TreeState(K, data["neutral_states"], rng, nu=1.0).grow(N)
    ch = Chain(dataset=ds)
    ch.compute_t_table(t.flat())
    trees.append(t)
    chains.append(ch)


def launch(kinds_per_chain, reps=6):
    sel = [b for b, k in enumerate(kinds_per_chain) if k]
    L = BatchLauncher([chains[b] for b in sel])
    score, build, tasks = [], [], []
    for r in range(reps):
        fts = []
        for b in sel:
            while True:
                prop = trees[b].propose(kinds=(kinds_per_chain[b],))
                if prop is not None:
                    break
            t2, roots, kind = prop
            ft = t2.flat()
            fts.append(ft)
            for rid in roots:
                chains[b].compute_t_prime_scores(ft, ft.index_of(rid))
        ds.kernel_time_stats(reset=True)
        L.run()
        ks, n = ds.kernel_time_stats()
        bs, nb = ds.build_time_stats()
        score.append(ks)
        build.append(bs)
        tasks.append(sum(len(c.read_t_prime_scores()[0]) for c in L.chains))
        for c in L.chains:
            c.reject()
    return np.median(score), np.median(build), np.median(tasks)


cases = {
    "32 light (prune_reattach)": ["prune_reattach"] * 32,
    "16 light": ["prune_reattach"] * 16 + [None] * 16,
    "32 add_remove": ["add_remove"] * 32,
    "32 nu": ["nu"] * 32,
    "8 nu": ["nu"] * 8 + [None] * 24,
    "1 nu": ["nu"] + [None] * 31,
    "1 nu + 31 light": ["nu"] + ["prune_reattach"] * 31,
    "5 nu + 27 light": ["nu"] * 5 + ["prune_reattach"] * 27,
    "1 light": ["prune_reattach"] + [None] * 31,
}
for name, kinds in cases.items():
    s, b, t = launch(kinds)
    print(f"{name:28s} K2-K4 {s:7.1f} us   K1 {b:7.1f} us   rescored nodes {t}")
