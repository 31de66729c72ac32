"""The parallel count-matrix reader (scicone_b200/host/csv_reader.hpp) against the serial
getline + strtod loop with the reference's Utils::read_counts semantics (Utils.cpp:74-102): bit-identical matrices on
integer counts, decimals, exponents, negative values, CRLF, blank lines, ragged rows, a missing final newline, and the same
error for a wrong row count; then the native host reading cells x bins (--bins_matrix_file, condensed through the ABI)
must give the run it gives on the pre-condensed matrix."""
import os
import subprocess

import numpy as np
import pytest

from host_util import CPU_BACKEND_DIR, NATIVE, read_outputs

CPU_ENV = {"LD_LIBRARY_PATH": CPU_BACKEND_DIR}

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def csv_check(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("csv") / "csv_check")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-pthread", os.path.join(ROOT, "tests", "native", "csv_check.cpp"), "-o", exe])
    return exe


def _run(exe, path, rows, cols, threads=4):
    res = subprocess.run([exe, path, str(rows), str(cols), str(threads)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    return res.stdout


@pytest.mark.parametrize("fmt", ["%.0f", "%.1f", "%.6f", "%.18e", "%.17g", "%d"])
def test_formats(csv_check, tmp_path, fmt):
    rng = np.random.default_rng(3)
    X = rng.poisson(120, size=(700, 333)).astype(float)
    if fmt not in ("%.0f", "%d"):
        X = X * rng.normal(0, 3.7, size=X.shape)
    p = str(tmp_path / "m.csv")
    np.savetxt(p, X.astype(int) if fmt == "%d" else X, delimiter=",", fmt=fmt)
    assert "identical" in _run(csv_check, p, 700, 333)
    assert "identical" in _run(csv_check, p, 700, 333, threads=1)


def test_edge_cases(csv_check, tmp_path):
    p = str(tmp_path / "e.csv")
    lines = ["1,2,3", " 4, 5 ,6\r", "", "7,8", "9,10,11,12,13", "-0,+3,.5", "5.,1e3,-2.5E-3", "0x10,inf,nan", "12abc,4,5", "1,,3",
             "123456789012345678901234567890,0.1234567890123456789012345,9007199254740993", "4.35,0.1,100.3"]
    open(p, "w").write("\n".join(lines))  # no final newline
    assert "identical" in _run(csv_check, p, len(lines), 3)
    open(p, "w").write("\n".join(lines) + "\n")
    assert "identical" in _run(csv_check, p, len(lines), 3)
    assert "identical" in _run(csv_check, p, len(lines), 5, threads=3)
    # wrong row counts: the same error either way
    assert "both failed" in _run(csv_check, p, len(lines) - 1, 3)
    assert "both failed" in _run(csv_check, p, len(lines) + 1, 3)
    open(p, "w").write("")
    assert "both failed" in _run(csv_check, p, 1, 3)


def test_many_pieces(csv_check, tmp_path):
    """Files of several MB are cut into pieces at line boundaries; rows of very different lengths move the cuts."""
    rng = np.random.default_rng(5)
    p = str(tmp_path / "big.csv")
    with open(p, "w") as f:
        for i in range(4000):
            n = 600 if i % 7 else 3
            f.write(",".join(str(v) for v in rng.integers(0, 100000, size=n)) + "\n")
    assert "identical" in _run(csv_check, p, 4000, 600, threads=8)
    assert "identical" in _run(csv_check, p, 4000, 600, threads=2)


@pytest.mark.skipif(not os.path.exists(os.path.join(CPU_BACKEND_DIR, "libscicone_score.so")), reason="cpu_backend not built")
def test_host_reads_bins(tmp_path):
    _host_reads_bins(tmp_path, dict(os.environ, **CPU_ENV))


@pytest.mark.gpu
def test_host_reads_bins_on_gpu(tmp_path):
    _host_reads_bins(tmp_path, dict(os.environ))  # K0 on the device


def _host_reads_bins(tmp_path, env):
    rng = np.random.default_rng(11)
    sizes = rng.integers(3, 12, size=9)
    bins = rng.poisson(30, size=(40, int(sizes.sum()) + 2)).astype(float)  # two trailing bins beyond the regions are ignored
    edges = np.concatenate([[0], np.cumsum(sizes)])
    D = np.stack([bins[:, edges[k]:edges[k + 1]].sum(axis=1) for k in range(len(sizes))], axis=1)
    np.savetxt(str(tmp_path / "bins.csv"), bins, delimiter=",", fmt="%.0f")
    np.savetxt(str(tmp_path / "d.csv"), D, delimiter=",", fmt="%.0f")
    np.savetxt(str(tmp_path / "r.txt"), sizes, fmt="%d")
    common = [NATIVE, "--region_sizes_file=r.txt", "--n_cells=40", f"--n_regions={len(sizes)}", "--n_iters=200", "--n_nodes=4", "--seed=3", "--nu=1.0",
              "--verbosity=2", "--max_scoring=true"]
    a = subprocess.run(common + ["--d_matrix_file=d.csv", "--postfix=regions"], cwd=str(tmp_path), capture_output=True, text=True, env=env)
    assert a.returncode == 0, a.stderr[-2000:]
    b = subprocess.run(common + ["--bins_matrix_file=bins.csv", f"--n_bins={bins.shape[1]}", "--postfix=bins"], cwd=str(tmp_path), capture_output=True,
                       text=True, env=env)
    assert b.returncode == 0, b.stderr[-2000:]
    ra, rb = read_outputs(str(tmp_path), "regions"), read_outputs(str(tmp_path), "bins")
    assert np.array_equal(ra["acc"], rb["acc"]) and np.array_equal(ra["chain"], rb["chain"]) and ra["tree"] == rb["tree"]
