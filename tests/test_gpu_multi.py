"""Cell-sharded runs on real GPUs: the same chain on 1 GPU and sharded over 2 GPUs must write
byte-identical traces - the tree score is assembled from per-1024-cell chunk sums added in global order, so it does not
depend on the number of GPUs, and neither does any accept decision.  Skipped on boxes with fewer than 2 GPUs (the CPU
side of the sharding logic is covered by tests/test_sharding_gloo.py)."""
import os
import subprocess

import numpy as np
import pytest

from host_util import NATIVE
from scicone_b200.synth import make_counts

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("max_scoring,schedule,n_chains,exchange", [("true", "static", 2, "peer"), ("false", "static", 2, "nccl"),
                                                                     ("true", "dynamic", 6, "peer"), ("false", "dynamic", 6, "nccl")])
def test_two_gpu_run_equals_one_gpu_run(tmp_path, max_scoring, schedule, n_chains, exchange):
    from scicone_b200.api import comm_unique_id

    data = make_counts(5000, 9000, 60, seed=31)  # 5 chunks of 1024 cells: shards of 3 and 2 chunks
    d = str(tmp_path / "d.csv")
    r = str(tmp_path / "r.txt")
    np.savetxt(d, data["D"], delimiter=",", fmt="%.0f")
    np.savetxt(r, data["region_sizes"], fmt="%d")
    common = [NATIVE, f"--d_matrix_file={d}", f"--region_sizes_file={r}", "--n_cells=5000", "--n_regions=60", "--n_iters=300",
              "--n_nodes=12", "--seed=5", "--nu=1.0", "--verbosity=2", f"--max_scoring={max_scoring}", f"--n_chains={n_chains}", f"--schedule={schedule}", "--min_batch=2"]
    one = subprocess.run(common + ["--postfix=one", "--device=0"], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert one.returncode == 0, one.stderr[-2000:]
    uid = comm_unique_id().hex()
    # the chunk sums travel through peer memory (the proposal kernel stores them into every rank's buffer over NVLink) or,
    # with SCICONE_EXCHANGE=nccl, through an all-gather
    env = dict(os.environ, SCICONE_EXCHANGE=exchange)
    procs = [subprocess.Popen(common + [f"--postfix=two{k}", f"--device={k}", f"--rank={k}", "--world=2", f"--nccl_id={uid}"],
                              cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env) for k in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=600)
        assert p.returncode == 0, err[-2000:]
        if exchange == "peer":
            assert "peer-memory exchange unavailable" not in err, err[-500:]
    # chain c is searched by rank c % 2 (the other rank follows it through the mailbox): its owner wrote the traces
    # (dynamic schedule: rank 0 composes the launches on two lanes, rank 1 replays its launch log)
    for c in range(n_chains):
        for name in ("acceptance_ratio.csv", "markov_chain.csv", "gamma_values.csv", "tree_inferred.txt"):
            a = open(os.path.join(str(tmp_path), f"one_c{c}_{name}")).read()
            assert a == open(os.path.join(str(tmp_path), f"two{c % 2}_c{c}_{name}")).read(), (c, name)
