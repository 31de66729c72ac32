"""ctypes binding of oracle/_ref/libscicone_oracle.so (the CPU restatement of the reference path).

Test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from scicone_b200.flat_tree import CTree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "_ref", "libscicone_oracle.so")

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"], stdout=subprocess.DEVNULL)
        L = C.CDLL(LIB_PATH)
        dp, ip, tp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(CTree)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int32, C.c_int32, dp, ip, ip, ip, C.c_double, C.c_int32, C.c_int32]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_compute_t_table.argtypes = [C.c_void_p, tp, dp]
        L.orc_compute_od_scores.argtypes = [C.c_void_p, C.c_double, dp, dp]
        L.orc_compute_t_prime_scores.argtypes = [C.c_void_p, tp, C.c_int32]
        L.orc_compute_t_prime_sums.argtypes = [C.c_void_p, tp, dp]
        L.orc_accept.argtypes = [C.c_void_p, tp]
        L.orc_reject.argtypes = [C.c_void_p]
        L.orc_t_n_nodes.argtypes = [C.c_void_p]
        L.orc_t_node_ids.argtypes = [C.c_void_p, ip]
        L.orc_read_t_scores.argtypes = [C.c_void_p, dp]
        L.orc_read_t_sums.argtypes = [C.c_void_p, dp]
        L.orc_read_t_maxs.argtypes = [C.c_void_p, ip, dp]
        L.orc_read_t_prime_sums.argtypes = [C.c_void_p, dp]
        L.orc_read_t_prime_maxs.argtypes = [C.c_void_p, ip, dp]
        L.orc_t_prime_n_rescored.argtypes = [C.c_void_p]
        L.orc_read_t_prime_scores.argtypes = [C.c_void_p, ip, dp]
        L.orc_assign_cells_to_nodes.argtypes = [C.c_void_p, tp, ip, ip]
        L.orc_log_sum.restype = C.c_double
        L.orc_log_sum.argtypes = [dp, C.c_int]
        L.orc_log_replace_sum.restype = C.c_double
        L.orc_log_replace_sum.argtypes = [C.c_double, dp, C.c_int, dp, C.c_int, dp, C.c_int]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class OracleChain:
    """Same method names as scicone_b200.api.Chain so that parity tests drive both alike."""

    def __init__(self, D, region_sizes, neutral_states=None, cluster_sizes=None, eta=1e-4, is_overdispersed=1,
                 max_scoring=0, ploidy=2):
        D = np.ascontiguousarray(D, dtype=np.float64)
        self.M, self.K = D.shape
        r = np.ascontiguousarray(region_sizes, dtype=np.int32)
        neutral = np.ascontiguousarray(neutral_states if neutral_states is not None else [ploidy] * self.K, dtype=np.int32)
        w = np.ascontiguousarray(cluster_sizes if cluster_sizes is not None else [1] * self.M, dtype=np.int32)
        self._h = lib().orc_create(self.M, self.K, _dp(D), _ip(r), _ip(neutral), _ip(w), eta, int(is_overdispersed), int(max_scoring))
        self.max_scoring = int(max_scoring)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    def compute_t_table(self, tree):
        out = C.c_double()
        rc = lib().orc_compute_t_table(self._h, tree.byref(), C.byref(out))
        assert rc == 0, rc
        return out.value

    def compute_od_scores(self, nu, per_cell=False):
        out = C.c_double()
        pc = np.zeros(self.M) if per_cell else None
        lib().orc_compute_od_scores(self._h, nu, C.byref(out), _dp(pc) if per_cell else None)
        return (out.value, pc) if per_cell else out.value

    compute_t_od_scores = compute_od_scores
    compute_t_prime_od_scores = compute_od_scores

    def compute_t_prime_scores(self, tree, node_idx):
        rc = lib().orc_compute_t_prime_scores(self._h, tree.byref(), node_idx)
        assert rc == 0, rc

    def compute_t_prime_sums(self, tree):
        out = C.c_double()
        rc = lib().orc_compute_t_prime_sums(self._h, tree.byref(), C.byref(out))
        assert rc == 0, rc
        return out.value

    def accept(self, tree=None):
        lib().orc_accept(self._h, tree.byref() if tree is not None else None)

    def reject(self):
        lib().orc_reject(self._h)

    def t_node_ids(self):
        n = lib().orc_t_n_nodes(self._h)
        ids = np.zeros(n, dtype=np.int32)
        lib().orc_t_node_ids(self._h, _ip(ids))
        return ids

    def read_t_scores(self):
        n = lib().orc_t_n_nodes(self._h)
        out = np.zeros((self.M, n))
        lib().orc_read_t_scores(self._h, _dp(out))
        return out

    def read_t_sums(self):
        out = np.zeros(self.M)
        lib().orc_read_t_sums(self._h, _dp(out))
        return out

    def read_t_maxs(self):
        ids, vals = np.zeros(self.M, dtype=np.int32), np.zeros(self.M)
        lib().orc_read_t_maxs(self._h, _ip(ids), _dp(vals))
        return ids, vals

    def read_t_prime_sums(self):
        out = np.zeros(self.M)
        lib().orc_read_t_prime_sums(self._h, _dp(out))
        return out

    def read_t_prime_maxs(self):
        ids, vals = np.zeros(self.M, dtype=np.int32), np.zeros(self.M)
        lib().orc_read_t_prime_maxs(self._h, _ip(ids), _dp(vals))
        return ids, vals

    def read_t_prime_scores(self):
        n = lib().orc_t_prime_n_rescored(self._h)
        ids, out = np.zeros(n, dtype=np.int32), np.zeros((self.M, n))
        lib().orc_read_t_prime_scores(self._h, _ip(ids), _dp(out))
        return ids, out

    def assign_cells_to_nodes(self, tree, with_cn=True):
        ids = np.zeros(self.M, dtype=np.int32)
        cn = np.zeros((self.M, self.K), dtype=np.int32) if with_cn else None
        lib().orc_assign_cells_to_nodes(self._h, tree.byref(), _ip(ids), _ip(cn) if with_cn else None)
        return ids, cn


def log_sum(v):
    v = np.ascontiguousarray(v, dtype=np.float64)
    return lib().orc_log_sum(_dp(v), len(v))


def log_replace_sum(s, sub, add, unch):
    sub, add, unch = (np.ascontiguousarray(x, dtype=np.float64) for x in (sub, add, unch))
    return lib().orc_log_replace_sum(s, _dp(sub), len(sub), _dp(add), len(add), _dp(unch), len(unch))


def condense_matrix(D_bins, region_sizes):
    """orc_condense_matrix: the oracle's restatement of Utils::condense_matrix (Utils.cpp:104-138)."""
    D_bins = np.ascontiguousarray(D_bins, dtype=np.float64)
    r = np.ascontiguousarray(region_sizes, dtype=np.int32)
    out = np.empty((D_bins.shape[0], r.size), dtype=np.float64)
    L = lib()
    L.orc_condense_matrix.argtypes = [C.POINTER(C.c_double), C.c_int64, C.c_int64, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_double)]
    rc = L.orc_condense_matrix(_dp(D_bins), D_bins.shape[0], D_bins.shape[1], _ip(r), int(r.size), _dp(out))
    assert rc == 0
    return out


def condense_clusters(D, assignments, n_clusters):
    """orc_condense_clusters: the oracle's restatement of scripts/condense_clusters.py:28-38."""
    D = np.ascontiguousarray(D, dtype=np.float64)
    a = np.ascontiguousarray(assignments, dtype=np.int32)
    means = np.empty((n_clusters, D.shape[1]), dtype=np.float64)
    sizes = np.empty(n_clusters, dtype=np.int32)
    L = lib()
    L.orc_condense_clusters.argtypes = [C.POINTER(C.c_double), C.c_int64, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int32)]
    assert L.orc_condense_clusters(_dp(D), D.shape[0], D.shape[1], _ip(a), n_clusters, _dp(means), _ip(sizes)) == 0
    return means, sizes
