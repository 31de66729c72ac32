"""ctypes binding of libscicone_score.so (the C ABI in include/scicone_score.h).

The class and method names mirror the members of the reference's ``Inference`` class that own the
scoring tables (reference: scicone/src/Inference.h:224 compute_t_table, :1015 compute_t_prime_scores,
:1069 compute_t_prime_sums, :1406 compute_t_od_scores, :1421 compute_t_prime_od_scores,
:996 update_t_scores via accept, :1311 assign_cells_to_nodes), so that parity tests read like the
reference's own tests (scicone/tests/validation.h).

There is no CPU fallback: loading fails loudly if the CUDA library has not been built
(``python -c 'import __graft_entry__ as g; g.build()'``) and every compute call raises
``SciconeError`` without a CUDA device.
"""
import atexit
import ctypes as C
import os
import weakref

import numpy as np

from .flat_tree import CTree, FlatTree  # noqa: F401  (re-exported)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libscicone_score.so")

# every symbol include/scicone_score.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "scicone_abi_version", "scicone_last_error", "scicone_dataset_create", "scicone_dataset_destroy",
    "scicone_chain_create", "scicone_chain_destroy", "scicone_compute_t_table", "scicone_compute_t_od_scores",
    "scicone_compute_t_prime_od_scores", "scicone_compute_t_prime_scores", "scicone_compute_t_prime_sums",
    "scicone_batch_compute_t_prime_sums", "scicone_batch_submit", "scicone_batch_wait", "scicone_accept", "scicone_reject", "scicone_t_n_nodes",
    "scicone_t_node_ids", "scicone_read_t_scores", "scicone_read_t_sums", "scicone_read_t_maxs",
    "scicone_read_t_prime_sums", "scicone_read_t_prime_maxs", "scicone_t_prime_n_rescored",
    "scicone_read_t_prime_scores", "scicone_assign_cells_to_nodes", "scicone_comm_unique_id",
    "scicone_dataset_attach_comm", "scicone_set_profiling", "scicone_last_kernel_time_us", "scicone_kernel_time_stats", "scicone_build_time_stats", "scicone_build_counters", "scicone_t_prime_cost", "scicone_kernel_launches",
    "scicone_counters", "scicone_region_mark", "scicone_region_elapsed_ms", "scicone_device_bytes",
    "scicone_parallel_for", "scicone_host_threads", "scicone_test_lgamma", "scicone_prepare_t_prime",
    "scicone_batch_compute_t_prime_scores", "scicone_batch_settle", "scicone_condense_matrix", "scicone_batch_poll", "scicone_set_streams", "scicone_condense_clusters",
    "scicone_dataset_reserve", "scicone_max_batch", "scicone_chain_create_follower", "scicone_prepare_t_table", "scicone_export_t_prime",
    "scicone_import_t_prime",
]

ERR_NAMES = {0: "OK", -1: "ERR_ARG", -2: "ERR_CUDA", -3: "ERR_STATE", -4: "ERR_NAN", -5: "ERR_COMM"}


_LIVE = weakref.WeakSet()  # datasets still open: closed at interpreter exit while the CUDA context is alive


@atexit.register
def _close_all():
    for ds in list(_LIVE):
        try:
            ds.close()
        except Exception:
            pass


class SciconeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class DatasetDesc(C.Structure):
    _fields_ = [
        ("n_cells", C.c_int32),
        ("n_regions", C.c_int32),
        ("D", C.POINTER(C.c_double)),
        ("region_sizes", C.POINTER(C.c_int32)),
        ("neutral_states", C.POINTER(C.c_int32)),
        ("cluster_sizes", C.POINTER(C.c_int32)),
        ("eta", C.c_double),
        ("is_overdispersed", C.c_int32),
        ("max_scoring", C.c_int32),
        ("device", C.c_int32),
        ("cell_offset", C.c_int64),
        ("n_cells_global", C.c_int64),
    ]


_lib = None


def lib():
    """Load libscicone_score.so; raises if it was not built (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() (needs nvcc)")
        L = C.CDLL(LIB_PATH)
        dp, ip, tp, vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(CTree), C.c_void_p
        L.scicone_last_error.restype = C.c_char_p
        L.scicone_dataset_create.argtypes = [C.POINTER(vp), C.POINTER(DatasetDesc)]
        L.scicone_dataset_destroy.argtypes = [vp]
        L.scicone_dataset_destroy.restype = None
        L.scicone_chain_create.argtypes = [C.POINTER(vp), vp]
        L.scicone_chain_destroy.argtypes = [vp]
        L.scicone_chain_destroy.restype = None
        L.scicone_compute_t_table.argtypes = [vp, tp, dp]
        L.scicone_compute_t_od_scores.argtypes = [vp, C.c_double, dp]
        L.scicone_compute_t_prime_od_scores.argtypes = [vp, C.c_double, dp]
        L.scicone_compute_t_prime_scores.argtypes = [vp, tp, C.c_int32]
        L.scicone_compute_t_prime_sums.argtypes = [vp, tp, dp]
        L.scicone_batch_compute_t_prime_sums.argtypes = [C.POINTER(vp), C.POINTER(tp), C.c_int32, dp, dp]
        L.scicone_batch_submit.argtypes = [C.POINTER(vp), C.POINTER(tp), C.c_int32, C.POINTER(C.c_uint64)]
        L.scicone_batch_wait.argtypes = [vp, C.c_uint64, dp, dp]
        L.scicone_kernel_time_stats.argtypes = [vp, dp, C.POINTER(C.c_uint64), C.c_int32]
        L.scicone_build_time_stats.argtypes = [vp, dp, C.POINTER(C.c_uint64)]
        L.scicone_build_counters.argtypes = [vp, dp, C.c_int32]
        L.scicone_counters.argtypes = [vp, dp, C.c_int32]
        L.scicone_region_mark.argtypes = [vp, C.c_int32]
        L.scicone_region_elapsed_ms.argtypes = [vp, dp]
        L.scicone_device_bytes.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.scicone_batch_compute_t_prime_scores.argtypes = [C.POINTER(vp), C.POINTER(tp), ip, C.c_int32]
        L.scicone_batch_settle.argtypes = [C.POINTER(vp), ip, C.c_int32]
        L.scicone_set_streams.argtypes = [vp, C.c_int32]
        L.scicone_condense_clusters.argtypes = [C.c_int32, dp, C.c_int64, C.c_int32, ip, C.c_int32, dp, ip]
        L.scicone_condense_matrix.argtypes = [C.c_int32, dp, C.c_int64, C.c_int64, ip, C.c_int32, dp, dp]
        L.scicone_test_lgamma.argtypes = [C.c_int32, dp, C.c_int32, dp]
        L.scicone_accept.argtypes = [vp]
        L.scicone_reject.argtypes = [vp]
        L.scicone_t_n_nodes.argtypes = [vp, ip]
        L.scicone_t_node_ids.argtypes = [vp, ip]
        L.scicone_read_t_scores.argtypes = [vp, dp]
        L.scicone_read_t_sums.argtypes = [vp, dp]
        L.scicone_read_t_maxs.argtypes = [vp, ip, dp]
        L.scicone_read_t_prime_sums.argtypes = [vp, dp]
        L.scicone_read_t_prime_maxs.argtypes = [vp, ip, dp]
        L.scicone_t_prime_n_rescored.argtypes = [vp, ip]
        L.scicone_read_t_prime_scores.argtypes = [vp, ip, dp]
        L.scicone_assign_cells_to_nodes.argtypes = [vp, tp, ip, ip]
        L.scicone_comm_unique_id.argtypes = [C.POINTER(C.c_uint8)]
        L.scicone_dataset_attach_comm.argtypes = [vp, C.POINTER(C.c_uint8), C.c_int32, C.c_int32]
        L.scicone_set_profiling.argtypes = [vp, C.c_int32]
        L.scicone_last_kernel_time_us.argtypes = [vp, dp]
        L.scicone_kernel_launches.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.scicone_dataset_reserve.argtypes = [vp, C.c_int32, C.c_int64]
        L.scicone_max_batch.argtypes = [vp, ip]
        L.scicone_chain_create_follower.argtypes = [C.POINTER(vp), vp]
        L.scicone_prepare_t_table.argtypes = [vp, tp, C.c_int32]
        L.scicone_prepare_t_prime.argtypes = [vp]
        L.scicone_export_t_prime.argtypes = [vp, vp, C.c_uint64, C.POINTER(C.c_uint64)]
        L.scicone_import_t_prime.argtypes = [vp, vp, C.c_uint64]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise SciconeError(rc, lib().scicone_last_error().decode())


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 creates it, the host program broadcasts it)."""
    buf = (C.c_uint8 * 128)()
    _check(lib().scicone_comm_unique_id(buf))
    return bytes(buf)


class Dataset:
    """Device-resident count matrix of one rank's cells (immutable; shared by any number of chains)."""

    def __init__(self, D, region_sizes, neutral_states=None, cluster_sizes=None, eta=1e-4, is_overdispersed=1,
                 max_scoring=0, ploidy=2, device=0, cell_offset=0, n_cells_global=None):
        D = np.ascontiguousarray(D, dtype=np.float64)
        if D.ndim != 2:
            raise ValueError("D must be cells x regions")
        self.M, self.K = D.shape
        r = np.ascontiguousarray(region_sizes, dtype=np.int32)
        if r.shape != (self.K,):
            # the reference throws std::logic_error when D and r disagree (scicone/src/Tree.h:170-171)
            raise SciconeError(-1, f"region_sizes has {r.size} entries for {self.K} regions")
        neutral = np.ascontiguousarray(neutral_states if neutral_states is not None else [ploidy] * self.K, dtype=np.int32)
        w = np.ascontiguousarray(cluster_sizes if cluster_sizes is not None else [1] * self.M, dtype=np.int32)
        if neutral.shape != (self.K,) or w.shape != (self.M,):
            raise SciconeError(-1, "neutral_states / cluster_sizes have the wrong length")
        desc = DatasetDesc(self.M, self.K, _dp(D), _ip(r), _ip(neutral), _ip(w), float(eta), int(is_overdispersed),
                           int(max_scoring), int(device), int(cell_offset),
                           int(n_cells_global if n_cells_global is not None else self.M))
        self.max_scoring = int(max_scoring)
        self.is_overdispersed = int(is_overdispersed)
        self._h = C.c_void_p()
        self._chains = weakref.WeakSet()  # chains are destroyed before their dataset, whatever the finalisation order
        _check(lib().scicone_dataset_create(C.byref(self._h), C.byref(desc)))
        _LIVE.add(self)

    def close(self):
        if getattr(self, "_h", None) and not getattr(self, "_closing", False):
            self._closing = True  # a chain that owns this dataset calls back into close()
            for ch in list(getattr(self, "_chains", ())):
                ch.close()
            lib().scicone_dataset_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def attach_comm(self, unique_id, rank, world_size):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _check(lib().scicone_dataset_attach_comm(self._h, buf, rank, world_size))

    def reserve(self, max_batch=0, rows=0):
        """scicone_dataset_reserve: size the reduction scratch and back the row pool with memory before the first launch."""
        _check(lib().scicone_dataset_reserve(self._h, int(max_batch), int(rows)))

    def max_batch(self):
        out = C.c_int32()
        _check(lib().scicone_max_batch(self._h, C.byref(out)))
        return out.value

    def set_profiling(self, on=True):
        _check(lib().scicone_set_profiling(self._h, int(on)))

    def set_streams(self, n):
        """scicone_set_streams: 2 = consecutive launches alternate between two streams (collect a launch before settling it)."""
        _check(lib().scicone_set_streams(self._h, int(n)))

    def last_kernel_time_us(self):
        out = C.c_double()
        _check(lib().scicone_last_kernel_time_us(self._h, C.byref(out)))
        return out.value

    def kernel_time_stats(self, reset=False):
        """(sum of proposal-kernel device time in us, number of timed launches) since the last reset."""
        us, n = C.c_double(), C.c_uint64()
        _check(lib().scicone_kernel_time_stats(self._h, C.byref(us), C.byref(n), int(reset)))
        return us.value, n.value

    def build_time_stats(self):
        """(sum of build_rows_kernel device time in us, number of timed launches) since the last reset."""
        us, n = C.c_double(), C.c_uint64()
        _check(lib().scicone_build_time_stats(self._h, C.byref(us), C.byref(n)))
        return us.value, n.value

    def counters(self, reset=False):
        out = np.zeros(4)
        _check(lib().scicone_counters(self._h, _dp(out), int(reset)))
        return dict(h2d_bytes=out[0], d2h_bytes=out[1], alg_bytes=out[2], n_scores=out[3])

    def build_counters(self, reset=False):
        out = np.zeros(2)
        _check(lib().scicone_build_counters(self._h, _dp(out), int(reset)))
        return dict(alg_bytes=out[0], n_lgamma=out[1])

    def region_mark(self, which):
        _check(lib().scicone_region_mark(self._h, int(which)))

    def region_elapsed_ms(self):
        out = C.c_double()
        _check(lib().scicone_region_elapsed_ms(self._h, C.byref(out)))
        return out.value

    def device_bytes(self):
        out = C.c_uint64()
        _check(lib().scicone_device_bytes(self._h, C.byref(out)))
        return out.value

    def kernel_launches(self):
        out = C.c_uint64()
        _check(lib().scicone_kernel_launches(self._h, C.byref(out)))
        return out.value


class Chain:
    """Score tables of one MCMC chain == one reference ``Inference`` object (Inference.h:39-44)."""

    def __init__(self, D=None, region_sizes=None, dataset=None, follower=False, **kw):
        self._own = dataset is None
        self.ds = dataset if dataset is not None else Dataset(D, region_sizes, **kw)
        self.M, self.K = self.ds.M, self.ds.K
        self.max_scoring = self.ds.max_scoring
        self._h = C.c_void_p()
        # follower: the handle another rank of a cell-sharded run holds for a chain it does not own (import + submit only)
        _check((lib().scicone_chain_create_follower if follower else lib().scicone_chain_create)(C.byref(self._h), self.ds._h))
        self.ds._chains.add(self)

    def prepare_t_table(self, tree, with_od=False):
        """Full pass as a queued proposal (scicone_prepare_t_table): launched by a batch submit, committed by accept()."""
        _check(lib().scicone_prepare_t_table(self._h, tree.byref(), int(with_od)))

    def prepare_t_prime(self):
        _check(lib().scicone_prepare_t_prime(self._h))

    def export_t_prime(self):
        """The compiled proposal as bytes (valid on every rank of a cell-sharded run)."""
        need = C.c_uint64()
        lib().scicone_export_t_prime(self._h, None, 0, C.byref(need))
        buf = (C.c_uint8 * need.value)()
        _check(lib().scicone_export_t_prime(self._h, buf, need.value, C.byref(need)))
        return bytes(buf)

    def import_t_prime(self, data):
        buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
        _check(lib().scicone_import_t_prime(self._h, buf, len(data)))

    def close(self):
        if getattr(self, "_h", None):
            if getattr(self.ds, "_h", None):  # a dataset that is already gone took its chains with it
                lib().scicone_chain_destroy(self._h)
            self._h = None
        if self._own and self.ds is not None:
            self.ds.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- Inference::compute_t_table, Inference.h:224-279
    def compute_t_table(self, tree):
        out = C.c_double()
        _check(lib().scicone_compute_t_table(self._h, tree.byref(), C.byref(out)))
        return out.value

    # ---- Inference::compute_t_od_scores / compute_t_prime_od_scores, Inference.h:1406-1433
    def compute_t_od_scores(self, nu):
        out = C.c_double()
        _check(lib().scicone_compute_t_od_scores(self._h, float(nu), C.byref(out)))
        return out.value

    def compute_t_prime_od_scores(self, nu):
        out = C.c_double()
        _check(lib().scicone_compute_t_prime_od_scores(self._h, float(nu), C.byref(out)))
        return out.value

    # ---- Inference::compute_t_prime_scores, Inference.h:1015-1052
    def compute_t_prime_scores(self, tree, node_idx):
        _check(lib().scicone_compute_t_prime_scores(self._h, tree.byref(), int(node_idx)))

    # ---- Inference::compute_t_prime_sums (Inference.h:1069-1196) + the sum of comparison (:292-294)
    def compute_t_prime_sums(self, tree):
        out = C.c_double()
        _check(lib().scicone_compute_t_prime_sums(self._h, tree.byref(), C.byref(out)))
        return out.value

    def accept(self, tree=None):
        _check(lib().scicone_accept(self._h))

    def reject(self):
        _check(lib().scicone_reject(self._h))

    def t_node_ids(self):
        n = C.c_int32()
        _check(lib().scicone_t_n_nodes(self._h, C.byref(n)))
        ids = np.zeros(n.value, dtype=np.int32)
        _check(lib().scicone_t_node_ids(self._h, _ip(ids)))
        return ids

    def read_t_scores(self):
        n = len(self.t_node_ids())
        out = np.zeros((self.M, n))
        _check(lib().scicone_read_t_scores(self._h, _dp(out)))
        return out

    def read_t_sums(self):
        out = np.zeros(self.M)
        _check(lib().scicone_read_t_sums(self._h, _dp(out)))
        return out

    def read_t_maxs(self):
        ids, vals = np.zeros(self.M, dtype=np.int32), np.zeros(self.M)
        _check(lib().scicone_read_t_maxs(self._h, _ip(ids), _dp(vals)))
        return ids, vals

    def read_t_prime_sums(self):
        out = np.zeros(self.M)
        _check(lib().scicone_read_t_prime_sums(self._h, _dp(out)))
        return out

    def read_t_prime_maxs(self):
        ids, vals = np.zeros(self.M, dtype=np.int32), np.zeros(self.M)
        _check(lib().scicone_read_t_prime_maxs(self._h, _ip(ids), _dp(vals)))
        return ids, vals

    def read_t_prime_scores(self):
        n = C.c_int32()
        _check(lib().scicone_t_prime_n_rescored(self._h, C.byref(n)))
        ids, out = np.zeros(n.value, dtype=np.int32), np.zeros((self.M, n.value))
        _check(lib().scicone_read_t_prime_scores(self._h, _ip(ids), _dp(out)))
        return ids, out

    # ---- Inference::assign_cells_to_nodes, Inference.h:1346-1362
    def assign_cells_to_nodes(self, tree, with_cn=True):
        ids = np.zeros(self.M, dtype=np.int32)
        cn = np.zeros((self.M, self.K), dtype=np.int32) if with_cn else None
        _check(lib().scicone_assign_cells_to_nodes(self._h, tree.byref(), _ip(ids), _ip(cn) if with_cn else None))
        return ids, cn


def batch_compute_t_prime_sums(chains, trees, with_od=False):
    """One kernel launch for the queued proposals of several chains on one dataset."""
    n = len(chains)
    hs = (C.c_void_p * n)(*[c._h for c in chains])
    ts = (C.POINTER(CTree) * n)(*[C.pointer(t.ctree) for t in trees])
    out, od = np.zeros(n), np.zeros(n)
    _check(lib().scicone_batch_compute_t_prime_sums(hs, ts, n, _dp(out), _dp(od)))
    return (out, od) if with_od else out


class BatchLauncher:
    """Pre-marshalled handles for repeated batch launches of the same chains (keeps ctypes overhead out of loops)."""

    def __init__(self, chains):
        self.chains = list(chains)
        self.n = len(self.chains)
        self.ds = self.chains[0].ds
        self._hs = (C.c_void_p * self.n)(*[c._h for c in self.chains])
        self._out = np.zeros(self.n)
        self._od = np.zeros(self.n)
        self._ticket = C.c_uint64()

    def submit(self):
        _check(lib().scicone_batch_submit(self._hs, None, self.n, C.byref(self._ticket)))
        return self._ticket.value

    def wait(self, ticket):
        _check(lib().scicone_batch_wait(self.ds._h, ticket, _dp(self._out), _dp(self._od)))
        return self._out

    def run(self):
        return self.wait(self.submit())

    def marshal_queue(self, trees, node_idx):
        """Pre-marshalled arguments of queue() (keeps the ctypes array building out of timed loops)."""
        tp = C.POINTER(CTree)
        arr = (tp * self.n)(*[C.pointer(t.ctree) for t in trees])
        idx = (C.c_int32 * self.n)(*[int(i) for i in node_idx])
        return (arr, idx, trees)  # the CTree buffers must outlive the call

    def marshal_settle(self, accept):
        return (C.c_int32 * self.n)(*[1 if a else 0 for a in accept])

    def queue(self, trees=None, node_idx=None, marshalled=None):
        """scicone_batch_compute_t_prime_scores: subtree root node_idx[i] of trees[i] on chain i (one call for all chains)."""
        arr, idx, self._keep = marshalled if marshalled is not None else self.marshal_queue(trees, node_idx)
        _check(lib().scicone_batch_compute_t_prime_scores(self._hs, arr, idx, self.n))

    def settle(self, accept=None, marshalled=None):
        """scicone_batch_settle: accept[i] ? accept : reject for every chain."""
        flags = marshalled if marshalled is not None else self.marshal_settle(accept)
        _check(lib().scicone_batch_settle(self._hs, flags, self.n))


def device_lgamma(x, device=0):
    """The engine's device log-gamma on an array of positive arguments (test hook, scicone_test_lgamma)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    L = lib()
    rc = L.scicone_test_lgamma(int(device), x.ctypes.data_as(C.POINTER(C.c_double)), int(x.size), out.ctypes.data_as(C.POINTER(C.c_double)))
    if rc != 0:
        raise SciconeError(rc, L.scicone_last_error().decode())
    return out


def condense_matrix(D_bins, region_sizes, device=0, return_kernel_us=False):
    """Utils::condense_matrix (Utils.cpp:104-138) on the device: cells x bins -> cells x regions, bins added in order."""
    D_bins = np.ascontiguousarray(D_bins, dtype=np.float64)
    r = np.ascontiguousarray(region_sizes, dtype=np.int32)
    out = np.empty((D_bins.shape[0], r.size), dtype=np.float64)
    us = C.c_double(0.0)
    L = lib()
    rc = L.scicone_condense_matrix(int(device), D_bins.ctypes.data_as(C.POINTER(C.c_double)), D_bins.shape[0], D_bins.shape[1],
                                   r.ctypes.data_as(C.POINTER(C.c_int32)), int(r.size), out.ctypes.data_as(C.POINTER(C.c_double)), C.byref(us))
    if rc != 0:
        raise SciconeError(rc, L.scicone_last_error().decode())
    return (out, us.value) if return_kernel_us else out


def condense_clusters(D, assignments, n_clusters=None, device=0):
    """Cluster means and sizes (reference scripts/condense_clusters.py:28-38) on the device: rows added in row order."""
    D = np.ascontiguousarray(D, dtype=np.float64)
    a = np.ascontiguousarray(assignments, dtype=np.int32)
    n = int(a.max()) + 1 if n_clusters is None else int(n_clusters)
    means = np.empty((n, D.shape[1]), dtype=np.float64)
    sizes = np.empty(n, dtype=np.int32)
    L = lib()
    rc = L.scicone_condense_clusters(int(device), D.ctypes.data_as(C.POINTER(C.c_double)), D.shape[0], D.shape[1],
                                     a.ctypes.data_as(C.POINTER(C.c_int32)), n, means.ctypes.data_as(C.POINTER(C.c_double)),
                                     sizes.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        raise SciconeError(rc, L.scicone_last_error().decode())
    return means, sizes
