"""world_size-2 (and 4) `gloo` test of the cell-sharded total: the N>1 host logic on CPU.

Each rank scores ITS cells with the oracle port (test infrastructure), reduces them to per-chunk partials with the
engine's reduction shape, all-gathers the partials and adds them in global chunk order.  The result must be
bit-identical to the single-rank chunk-ordered total, for every world size, and equal the reference's serial sum
(Inference.h:292-294) to rounding."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import sharding_model as sharding  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case():
    from trace_util import make_case

    data, tree = make_case(5000, 4000, 24, 9, seed=321)
    w = np.random.default_rng(5).integers(1, 4, size=5000).astype(np.int32)
    return data, tree, w


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle_ffi import OracleChain

    data, tree, w = _case()
    a, b = sharding.shard_bounds(data["D"].shape[0], world)[rank]
    width = sharding.max_chunks_per_rank(data["D"].shape[0], world)
    mine = np.zeros(width)
    if b > a:
        ch = OracleChain(D=data["D"][a:b], region_sizes=data["region_sizes"], cluster_sizes=w[a:b])
        ch.compute_t_table(tree.flat())
        part = sharding.chunk_partials(ch.read_t_sums(), w[a:b])
        mine[: part.size] = part
    gathered = [torch.zeros(width, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(mine))
    total = sharding.combine([g.numpy() for g in gathered])
    if rank == 0:
        np.save(out_path, np.array([total]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_total_is_independent_of_world_size(tmp_path, world):
    from oracle_ffi import OracleChain

    out = str(tmp_path / "total.npy")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    sharded = float(np.load(out)[0])
    data, tree, w = _case()
    ch = OracleChain(D=data["D"], region_sizes=data["region_sizes"], cluster_sizes=w)
    serial = ch.compute_t_table(tree.flat())  # reference order: j ascending
    single = sharding.combine([sharding.chunk_partials(ch.read_t_sums(), w)])
    assert sharded == single  # bit-identical for any number of ranks
    assert abs(sharded - serial) <= 1e-12 * abs(serial)


def test_shard_bounds_cover_cells_and_are_chunk_aligned():
    for n in (1, 1023, 1024, 1025, 10000, 100000):
        for world in (1, 2, 4, 8):
            b = sharding.shard_bounds(n, world)
            assert b[0][0] == 0 and b[-1][1] == n
            for (a0, a1), (b0, b1) in zip(b, b[1:]):
                assert a1 == b0
            for a0, a1 in b:
                assert a0 % sharding.CHUNK == 0 or a0 == n
