"""Deterministic synthetic inputs for tests and bench.py (no reference data can travel to the GPU box).

* ``make_counts``   region-level Dirichlet-multinomial read counts under a random clone tree, following the
  reference simulator's data model (reference: scicone/simulations/Simulation.h:61-186): region sizes from a
  Dirichlet split of the bins with a minimum width, clones that differ by +-1 single-region events, copy numbers
  clipped to [0, 7], cells assigned uniformly to clones, reads/cell = 10 x bins, overdispersion nu_true = 10.
  Drawing at region level is distributionally identical to drawing per bin and collapsing (Dirichlet and
  multinomial both aggregate), which is what the reference's bins->regions collapse does before `inference`.
* ``TreeState``     a mutable event tree plus a random proposal generator covering the rescoring moves of the
  reference (prune/reattach Tree.h:544, swap labels :878, add/remove events :999, insert/delete node :1226,
  nu change Inference.h:916).  It is NOT the MCMC host (that is scicone_b200/host, C++): it only produces valid
  (tree, rescored subtree roots) pairs to drive the scoring engine and the oracle with identical inputs.
"""
import numpy as np

from .flat_tree import FlatTree

SEED = 20261019


def region_sizes(rng, n_bins, n_regions, min_width=10):
    free = n_bins - n_regions * min_width
    if free < 0:
        raise ValueError("too many regions for the bins")
    p = rng.dirichlet(np.ones(n_regions))
    extra = np.floor(p * free).astype(np.int64)
    extra[: free - extra.sum()] += 1
    return (extra + min_width).astype(np.int32)


def make_counts(n_cells, n_bins, n_regions, seed=SEED, n_clones=20, nu_true=10.0, reads_per_bin=10, ploidy=2,
                max_cn=7, cluster_rows=None):
    """Returns dict(D [cells x regions] float64, region_sizes, neutral_states, cluster_sizes, clone_of_cell)."""
    rng = np.random.default_rng(seed)
    r = region_sizes(rng, n_bins, n_regions)
    clones = [np.full(n_regions, ploidy, dtype=np.int64)]
    for _ in range(n_clones - 1):
        c = clones[rng.integers(len(clones))].copy()
        k = rng.integers(n_regions)
        c[k] = np.clip(c[k] + rng.choice([-1, 1]), 0, max_cn)
        clones.append(c)
    clones = np.stack(clones)
    lab = rng.integers(n_clones, size=n_cells)
    reads = int(reads_per_bin * n_bins)
    alpha = nu_true * (clones * r[None, :] + 1e-3)
    # Dirichlet via gamma draws, vectorised over cells
    g = rng.standard_gamma(alpha[lab])
    p = g / g.sum(axis=1, keepdims=True)
    D = np.empty((n_cells, n_regions), dtype=np.float64)
    for j in range(n_cells):
        D[j] = rng.multinomial(reads, p[j])
    w = np.ones(n_cells, dtype=np.int32)
    if cluster_rows:
        # PhenoGraph-like condensation: cluster means (non-integer) + integer sizes (scicone.py:331-335)
        cl = lab % cluster_rows
        cl[:cluster_rows] = np.arange(cluster_rows)
        D = np.stack([D[cl == c].mean(axis=0) for c in range(cluster_rows)])
        w = np.array([(cl == c).sum() for c in range(cluster_rows)], dtype=np.int32)
        lab = None
    return dict(D=D, region_sizes=r, neutral_states=np.full(n_regions, ploidy, dtype=np.int32), cluster_sizes=w,
                clone_of_cell=lab)


def expand_to_bins(D, region_sizes, seed=SEED):
    """A cells x bins matrix whose bins->regions collapse (Utils::condense_matrix, reference scicone/src/Utils.cpp:104-138)
    is exactly the integer region-level matrix D: every region's count is spread over its bins as evenly as integers allow,
    the remainder going to a run of bins that starts at a per-cell random offset.  (A bin-level multinomial draw would be
    distributionally closer to the simulator but takes half a minute at 10 000 x 18 000; the collapse kernel is a
    bandwidth kernel and only sees 8-byte values.)"""
    D = np.asarray(D)
    r = np.asarray(region_sizes, dtype=np.int64)
    M, K = D.shape
    out = np.empty((M, int(r.sum())), dtype=np.float64)
    shift = np.random.default_rng(seed).integers(0, 1 << 30, size=M)
    o = 0
    for k in range(K):
        c = D[:, k].astype(np.int64)
        base, rem = c // r[k], c % r[k]
        pos = (np.arange(r[k])[None, :] + shift[:, None]) % r[k]
        out[:, o:o + r[k]] = base[:, None] + (pos < rem[:, None])
        o += r[k]
    return out


class TreeState:
    """Event tree: node id -> (parent id, {region: event}).  Root id 0 has no events."""

    def __init__(self, n_regions, neutral, rng, nu=1.0, max_cn=7):
        self.K = n_regions
        self.neutral = np.asarray(neutral, dtype=np.int64)
        self.rng = rng
        self.nu = float(nu)
        self.max_cn = max_cn
        self.parent = {0: -1}
        self.events = {0: {}}
        self.next_id = 1

    def copy(self):
        o = TreeState.__new__(TreeState)
        o.K, o.neutral, o.rng, o.nu, o.max_cn = self.K, self.neutral, self.rng, self.nu, self.max_cn
        o.parent = dict(self.parent)
        o.events = {k: dict(v) for k, v in self.events.items()}
        o.next_id = self.next_id
        return o

    # ---- helpers
    def children(self, u):
        return [v for v, p in self.parent.items() if p == u]

    def subtree(self, u):
        out, stack = [], [u]
        while stack:
            x = stack.pop()
            out.append(x)
            stack.extend(self.children(x))
        return out

    def genotype(self, u):
        g = {}
        path = []
        while u != -1:
            path.append(u)
            u = self.parent[u]
        for x in reversed(path):
            for k, v in self.events[x].items():
                g[k] = g.get(k, 0) + v
        return {k: v for k, v in g.items() if v != 0}

    def valid(self):
        for u in self.parent:
            if u != 0 and not self.events[u]:
                return False
            for k, v in self.genotype(u).items():
                cn = v + self.neutral[k]
                if cn < 0 or cn > self.max_cn:
                    return False
        return True

    def flat(self):
        nodes = [{"id": u, "parent": p, "c_change": self.events[u]} for u, p in self.parent.items()]
        return FlatTree(nodes, self.nu)

    def random_event(self):
        k = int(self.rng.integers(self.K))
        return {k: int(self.rng.choice([-1, 1]))}

    def grow(self, n_nodes):
        """Random initial tree (cf. Inference::random_initialize, Inference.h:99-156)."""
        while len(self.parent) < n_nodes:
            t = self.copy()
            u = t.next_id
            t.next_id += 1
            t.parent[u] = int(self.rng.choice(list(t.parent.keys())))
            t.events[u] = t.random_event()
            if t.valid():
                self.parent, self.events, self.next_id = t.parent, t.events, t.next_id
        return self

    # ---- proposals: return (new TreeState, [ids of subtree roots to rescore], kind) or None
    def propose(self, kinds=("prune_reattach", "swap", "add_remove", "insert", "delete", "nu"), weights=None):
        rng = self.rng
        for _ in range(50):
            kind = str(rng.choice(kinds, p=weights))
            t = self.copy()
            nonroot = [u for u in t.parent if u != 0]
            roots = None
            if kind == "prune_reattach" and len(nonroot) >= 2:
                u = int(rng.choice(nonroot))
                sub = set(t.subtree(u))
                dest = [v for v in t.parent if v not in sub and v != t.parent[u]]
                if not dest:
                    continue
                t.parent[u] = int(rng.choice(dest))
                roots = [u]
            elif kind == "swap" and len(nonroot) >= 2:
                u, v = (int(x) for x in rng.choice(nonroot, size=2, replace=False))
                t.events[u], t.events[v] = t.events[v], t.events[u]
                if v in t.subtree(u):
                    roots = [u]
                elif u in t.subtree(v):
                    roots = [v]
                else:
                    roots = [u, v]
            elif kind == "add_remove" and nonroot:
                u = int(rng.choice(nonroot))
                for k, d in t.random_event().items():
                    nv = t.events[u].get(k, 0) + d
                    if nv == 0:
                        t.events[u].pop(k)
                    else:
                        t.events[u][k] = nv
                roots = [u]
            elif kind == "insert":
                u = t.next_id
                t.next_id += 1
                par = int(rng.choice(list(t.parent.keys())))
                kids = t.children(par)
                t.parent[u] = par
                t.events[u] = t.random_event()
                for c in kids:
                    if rng.random() < 0.5:
                        t.parent[c] = u
                roots = [u]
            elif kind == "delete" and len(nonroot) >= 2:
                u = int(rng.choice(nonroot))
                par = t.parent[u]
                for c in t.children(u):
                    t.parent[c] = par
                del t.parent[u], t.events[u]
                roots = [par]
            elif kind == "nu":
                t.nu = float(np.exp(np.log(t.nu) + rng.normal(0, 0.02)))
                roots = [0]
            if roots is not None and t.valid():
                return t, roots, kind
        return None
