"""NumPy model of the cell-shard reduction - test infrastructure for tests/test_sharding_gloo.py, not
part of the product: the product's reduction is finish_reduction / gather_totals_kernel in scicone_b200/csrc/kernels.cuh.

Every per-cell quantity of the scoring path is independent across cells; the only cross-cell step is the weighted
sum over cells (reference: scicone/src/Inference.h:292-294).  Cells are split into contiguous shards whose
boundaries fall on multiples of CHUNK cells; every rank reduces its cells to one partial per chunk (CHUNK = the
8 x 128-cell CTAs of kernels.cuh: kCtasPerChunk x kCells), all ranks exchange the chunk partials (peer-memory stores or an
NCCL all-gather inside libscicone_score.so; `gloo` in the CPU tests) and add them in global chunk order.  Because the order of additions is
fixed by the chunk grid and not by the rank layout, the tree score is bit-identical for 1, 2, 4 or 8 GPUs.
"""
import numpy as np

CHUNK = 1024  # kCells * kCtasPerChunk in scicone_b200/csrc/kernels.cuh


def shard_bounds(n_cells, world_size, chunk=CHUNK):
    """[(start, stop)] per rank: contiguous, chunk-aligned, sizes differ by at most one chunk; trailing ranks may be empty."""
    n_chunks = (n_cells + chunk - 1) // chunk
    base, extra = divmod(n_chunks, world_size)
    out, c0 = [], 0
    for r in range(world_size):
        c1 = c0 + base + (1 if r < extra else 0)
        out.append((min(c0 * chunk, n_cells), min(c1 * chunk, n_cells)))
        c0 = c1
    return out


def max_chunks_per_rank(n_cells, world_size, chunk=CHUNK):
    return max((b - a + chunk - 1) // chunk for a, b in shard_bounds(n_cells, world_size, chunk))


def chunk_partials(values, weights, chunk=CHUNK, cta=128):
    """Per-chunk partial sums of values*weights with the kernel's reduction shape: xor-shuffle tree inside a warp,
    warps added in order inside a CTA, CTAs added in order inside a chunk (kernels.cuh block_sum/finish_reduction)."""
    v = np.asarray(values, dtype=np.float64) * np.asarray(weights, dtype=np.float64)
    n = v.size
    pad = (-n) % cta
    v = np.concatenate([v, np.zeros(pad)])
    w = v.reshape(-1, 32).copy()
    for o in (16, 8, 4, 2, 1):  # lane i ends with the butterfly sum; all lanes equal, take lane 0
        w = w + w[:, np.arange(32) ^ o]
    warp = w[:, 0].reshape(-1, cta // 32)
    ctas = np.zeros(warp.shape[0])
    for i in range(cta // 32):
        ctas = ctas + warp[:, i]
    per_chunk = chunk // cta
    n_chunks = (ctas.size + per_chunk - 1) // per_chunk
    out = np.zeros(n_chunks)
    for c in range(n_chunks):
        s = 0.0
        for b in ctas[c * per_chunk:(c + 1) * per_chunk]:
            s = s + b
        out[c] = s
    return out


def combine(gathered):
    """gathered[rank] = zero-padded chunk partials of that rank; serial sum in rank-then-chunk (= global chunk) order."""
    s = 0.0
    for part in gathered:
        for x in np.asarray(part, dtype=np.float64):
            s = s + x
    return s
