"""bench.py's contract, checked where no GPU exists: the reference arm prints exactly ONE JSON line on stdout with the
keys the driver reads, and the engine arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "inference")
SMALL = ["--steps", "2", "--warmup", "0", "--cells", "200", "--regions", "8", "--bins", "1000", "--nodes", "5", "--chains", "2"]


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/inference not built")
def test_reference_arm_prints_one_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"] + SMALL, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    out = json.loads(lines[0])
    assert out["impl"] == "reference" and out["metric"] == "mcmc_cell_iters_per_s" and out["unit"] == "cell_iters/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert key in out, key
    assert out["value"] > 0 and out["higher_is_better"] is True and out["vs_baseline"] is None
    assert out["cpu_baseline"]["kind"] == "reference" and out["cpu_baseline"]["cores"] >= 1 and out["cpu_baseline"]["value"] == out["value"]
    assert out["e2e"] == {"value": out["value"], "unit": out["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in out["config"]


def test_engine_arm_needs_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + SMALL, capture_output=True, text=True, timeout=600)
    assert res.returncode != 0
    assert "no CPU fallback" in (res.stderr + res.stdout)
