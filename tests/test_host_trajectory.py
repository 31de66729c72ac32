"""The native MCMC host (scicone_b200/host -> scicone_b200/bin/inference: move set, priors, RNG consumption, CLI and
writers re-implemented on top of the C ABI) against the unmodified reference `inference` (oracle/_ref/inference):
identical accepted-move trajectory, final tree and per-cell outputs; per-move scores within 1e-9 relative.

CPU half (not gpu): the host binary runs with oracle/_ref/cpu_backend on LD_LIBRARY_PATH, i.e. the C ABI served by the
oracle port - this checks the HOST logic (moves, validity rules, neighbourhood corrections, retries, file formats)
where no GPU exists.  GPU half: the same binary with its normal rpath, i.e. the CUDA library."""
import os

import numpy as np
import pytest

from golden_util import load_golden
from host_util import CPU_BACKEND_DIR, NATIVE, REF, compare_runs, run_inference
from scicone_b200.synth import make_counts

needs_bins = pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(NATIVE)),
                                reason="oracle/_ref/inference or scicone_b200/bin/inference not built")
CPU_ENV = {"LD_LIBRARY_PATH": CPU_BACKEND_DIR}


def _tutorial():
    doc = load_golden("tutorial_full_sum")
    return np.array(doc["D"]), np.array(doc["region_sizes"])


CASES = {
    # configs[0]: docs/tutorial.md command, both scoring modes
    "tutorial_max": (lambda: _tutorial() + (None,), ["--n_iters=3000", "--n_nodes=3", "--seed=42", "--ploidy=2", "--copy_number_limit=2", "--max_scoring=true"]),
    "tutorial_sum": (lambda: _tutorial() + (None,), ["--n_iters=3000", "--n_nodes=3", "--seed=42", "--ploidy=2", "--copy_number_limit=2", "--max_scoring=false"]),
    # every move enabled (genotype-preserving prune-reattach, expand/shrink, common ancestor, delete leaf)
    "all_moves_max": (lambda: _tutorial() + (None,), ["--n_iters=2500", "--n_nodes=6", "--seed=7", "--nu=2.0", "--max_scoring=true",
                                                       "--move_probs=1,1,1,1,1,1,1,1,1,1,1,1,1,1,1"]),
    "all_moves_sum": (lambda: _tutorial() + (None,), ["--n_iters=2500", "--n_nodes=6", "--seed=8", "--nu=2.0", "--max_scoring=false",
                                                       "--move_probs=1,1,1,1,1,1,1,1,1,1,1,0,0,1,1"]),
    # plain multinomial branch: no nu given and the nu move disabled
    "multinomial": (lambda: _tutorial() + (None,), ["--n_iters=1500", "--n_nodes=4", "--seed=9", "--max_scoring=true",
                                                     "--move_probs=0,1,0,1,0,1,0,1,0,1,0.01,0.1,0.01,0,0"]),
    # learning rate / tempering and a size limit
    "alpha_gamma": (lambda: _tutorial() + (None,), ["--n_iters=1500", "--n_nodes=5", "--seed=10", "--nu=1.0", "--alpha=0.2", "--gamma=2.0",
                                                     "--size_limit=8", "--max_scoring=true"]),
}


def _config2():
    d = make_counts(1000, 10000, 50, seed=5)
    return d["D"], d["region_sizes"], None


def _cluster():
    d = make_counts(3000, 10000, 40, seed=6, cluster_rows=30)
    return d["D"], d["region_sizes"], d["cluster_sizes"]


CASES["config2_max"] = (_config2, ["--n_iters=300", "--n_nodes=10", "--seed=43", "--nu=1.0", "--max_scoring=true"])
CASES["config2_sum"] = (_config2, ["--n_iters=200", "--n_nodes=20", "--seed=44", "--nu=1.0", "--max_scoring=false"])
CASES["cluster_max"] = (_cluster, ["--n_iters=1500", "--n_nodes=5", "--seed=44", "--nu=1.0", "--max_scoring=true"])


# the BASELINE.json test configs must be identical end to end; the stress cases (every move enabled, tempering) may end
# at a likelihood-neutral knife edge on the GPU (see host_util.compare_runs)
STRICT = {"tutorial_max", "tutorial_sum", "config2_max", "config2_sum", "cluster_max"}


def _both(tmp_path, case, env):
    data, extra = CASES[case]
    D, r, w = data()
    a = run_inference(REF, str(tmp_path), "ref", D, r, extra, w=w)
    b = run_inference(NATIVE, str(tmp_path), "new", D, r, extra, w=w, env=env)
    rel = compare_runs(a, b, knife_edge_ok=(env is None and case not in STRICT))
    print(case, "accepted", int(np.sum(a["acc"] >= 0)), "of", a["acc"].size, "max rel diff", rel)
    return a, b


@needs_bins
@pytest.mark.skipif(not os.path.exists(os.path.join(CPU_BACKEND_DIR, "libscicone_score.so")), reason="cpu_backend not built")
@pytest.mark.parametrize("case", sorted(CASES))
def test_native_host_logic_on_cpu(tmp_path, case):
    _both(tmp_path, case, CPU_ENV)


@needs_bins
@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CASES))
def test_native_host_on_gpu(tmp_path, case):
    _both(tmp_path, case, None)


def _config2_full():
    # configs[2] at full size for the scoring path: 10 000 cells x 200 regions, 50-node trees.  Only the bin count is
    # reduced (4 000 instead of 18 000): bins enter nothing but the width of the inferred_cnvs.csv text, which both
    # hosts write and the comparison reads in full (40 MB instead of 360 MB per run).
    d = make_counts(10000, 4000, 200, seed=12)
    return d["D"], d["region_sizes"], None


FULL_CASES = {
    "config2_full_max": (_config2_full, ["--n_iters=150", "--n_nodes=49", "--seed=42", "--nu=1.0", "--max_scoring=true"]),
    "config2_full_sum": (_config2_full, ["--n_iters=150", "--n_nodes=49", "--seed=43", "--nu=1.0", "--max_scoring=false"]),
}


@needs_bins
@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(FULL_CASES))
def test_native_host_on_gpu_at_config2_size(tmp_path, case):
    """Inference::infer_mcmc (Inference.h:544-913) at BASELINE configs[2] size against the unmodified reference binary:
    identical accepted-move sequence, final tree and per-cell outputs, per-move scores within 1e-9 relative - strict
    comparison, no knife-edge allowance."""
    data, extra = FULL_CASES[case]
    D, r, w = data()
    a = run_inference(REF, str(tmp_path), "ref", D, r, extra, w=w, env={"OMP_NUM_THREADS": str(os.cpu_count() or 1)}, timeout=1500)
    b = run_inference(NATIVE, str(tmp_path), "new", D, r, extra, w=w, timeout=600)
    rel = compare_runs(a, b)
    assert int(np.sum(a["acc"] >= 0)) > 5, "the run must accept some moves to mean anything"
    print(case, "accepted", int(np.sum(a["acc"] >= 0)), "of", a["acc"].size, "max rel diff", rel)


@needs_bins
@pytest.mark.skipif(not os.path.exists(os.path.join(CPU_BACKEND_DIR, "libscicone_score.so")), reason="cpu_backend not built")
def test_tree_file_start_on_cpu(tmp_path):
    """--tree_file: continue from a stored tree (format of Tree::load_from_file, Tree.h:1565-1660)."""
    D, r = _tutorial()
    first = ["--n_iters=400", "--n_nodes=4", "--seed=3", "--nu=1.5", "--max_scoring=true"]
    run_inference(REF, str(tmp_path), "seed", D, r, first)
    tf = os.path.join(str(tmp_path), "seed_tree_inferred.txt")
    extra = ["--n_iters=600", "--seed=4", f"--tree_file={tf}", "--max_scoring=true"]
    a = run_inference(REF, str(tmp_path), "ref", D, r, extra)
    b = run_inference(NATIVE, str(tmp_path), "new", D, r, extra, env=CPU_ENV)
    compare_runs(a, b)


def _multi_chain(tmp_path, env, n_chains, extra_flags, check):
    """--n_chains=B: restarts advanced in groups (parallel host, pipelined launches) must each equal the single-chain
    run of the unmodified reference with seed+c, whatever the grouping and the number of host threads."""
    import subprocess

    from host_util import read_outputs

    D, r = _tutorial()
    base = ["--n_iters=800", "--n_nodes=4", "--nu=1.0", "--max_scoring=true"]
    d = os.path.join(str(tmp_path), "d.csv")
    np.savetxt(d, D, delimiter=",", fmt="%.0f")
    rf = os.path.join(str(tmp_path), "r.txt")
    np.savetxt(rf, r, fmt="%d")
    e = dict(os.environ)
    e.update(env or {})
    cmd = [NATIVE, f"--d_matrix_file={d}", f"--region_sizes_file={rf}", f"--n_cells={D.shape[0]}", f"--n_regions={D.shape[1]}",
           "--verbosity=2", "--postfix=multi", "--seed=20", f"--n_chains={n_chains}"] + base + extra_flags
    res = subprocess.run(cmd, cwd=str(tmp_path), capture_output=True, text=True, env=e, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    for c in check:
        single = run_inference(REF, str(tmp_path), f"single{c}", D, r, base + [f"--seed={20 + c}"])
        compare_runs(single, read_outputs(str(tmp_path), f"multi_c{c}"))


@needs_bins
@pytest.mark.skipif(not os.path.exists(os.path.join(CPU_BACKEND_DIR, "libscicone_score.so")), reason="cpu_backend not built")
@pytest.mark.parametrize("n_chains,flags,threads", [(3, [], "1"), (9, [], "4"), (10, ["--n_groups=3"], "3")])
def test_multi_chain_equals_single_runs_on_cpu(tmp_path, n_chains, flags, threads):
    _multi_chain(tmp_path, dict(CPU_ENV, SCICONE_THREADS=threads), n_chains, flags, [0, n_chains // 2, n_chains - 1])


@needs_bins
@pytest.mark.skipif(not os.path.exists(os.path.join(CPU_BACKEND_DIR, "libscicone_score.so")), reason="cpu_backend not built")
def test_timeline_records_every_launch_and_every_chain(tmp_path):
    """--timeline=1 (timing runs): the stats file lists every submit (kind 0) and collect (1) of the measured phase with
    the number of proposals, and the moment each chain finished it (2, chain) - what profiles/r2b_timeline_* was made from.
    The warm-up phase leaves no entries, and the trajectories are those of a run without it."""
    import json
    import subprocess

    from host_util import read_outputs

    D, r = _tutorial()
    base = ["--n_nodes=4", "--nu=1.0", "--max_scoring=true", "--seed=20", "--n_chains=6"]
    d = os.path.join(str(tmp_path), "d.csv")
    np.savetxt(d, D, delimiter=",", fmt="%.0f")
    rf = os.path.join(str(tmp_path), "r.txt")
    np.savetxt(rf, r, fmt="%d")
    e = dict(os.environ, **dict(CPU_ENV, SCICONE_THREADS="3"))
    common = [NATIVE, f"--d_matrix_file={d}", f"--region_sizes_file={rf}", f"--n_cells={D.shape[0]}", f"--n_regions={D.shape[1]}", "--verbosity=2"] + base
    stats = os.path.join(str(tmp_path), "stats.json")
    res = subprocess.run(common + ["--postfix=tl", "--n_iters=60", "--warmup_iters=40", "--timeline=1", f"--stats_file={stats}"],
                         cwd=str(tmp_path), capture_output=True, text=True, env=e, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    st = json.load(open(stats))
    tl = np.array(st["timeline"])
    sub, col, done = tl[tl[:, 1] == 0], tl[tl[:, 1] == 1], tl[tl[:, 1] == 2]
    assert len(sub) == len(col) == st["launches"] > 0
    assert sub[:, 2].sum() == col[:, 2].sum() == st["launched_chains"] <= 6 * 60
    assert sorted(done[:, 2].astype(int)) == list(range(6))  # every chain finished the measured phase exactly once
    assert tl[:, 0].min() >= 0 and done[:, 0].max() <= st["wall_us"] + 1000
    # the same 100 iterations in one phase walk the same trajectories
    res = subprocess.run(common + ["--postfix=plain", "--n_iters=100"], cwd=str(tmp_path), capture_output=True, text=True, env=e, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    for c in (0, 5):
        compare_runs(read_outputs(str(tmp_path), f"plain_c{c}"), read_outputs(str(tmp_path), f"tl_c{c}"))


@needs_bins
@pytest.mark.gpu
def test_multi_chain_equals_single_runs_on_gpu(tmp_path):
    _multi_chain(tmp_path, None, 12, [], [0, 5, 11])


@needs_bins
@pytest.mark.gpu
def test_more_chains_than_the_initial_slot_count_on_gpu(tmp_path):
    """70 chains under the dynamic schedule: the library's reduction scratch starts with 64 slots and cannot grow while a
    launch is outstanding, so the driver must reserve it up front (scicone_dataset_reserve) - launches of 65+ proposals
    while another launch is in flight used to abort the run."""
    _multi_chain(tmp_path, None, 70, ["--min_batch=66"], [0, 37, 69])


@needs_bins
@pytest.mark.skipif(not os.path.exists(os.path.join(CPU_BACKEND_DIR, "libscicone_score.so")), reason="cpu_backend not built")
@pytest.mark.parametrize("schedule,world,threads,warmup", [("static", 2, "2", 0), ("dynamic", 2, "1", 0), ("dynamic", 3, "3", 0), ("dynamic", 2, "2", 150)])
def test_chain_ownership_protocol_on_cpu(tmp_path, schedule, world, threads, warmup):
    """Multi-process runs: every chain is searched by ONE owner process, the others follow it through the shared-memory
    mailbox (proposal + decision messages).  Two processes with --no_shard (each holds all cells, oracle-backed ABI):
    the owner's traces of every chain equal the unmodified reference's single-chain run, and the followers' device state
    stays in step (a follower that missed a commit would make the owner's next totals differ on the GPU; here it is
    checked through the final tables of both processes being readable and the run finishing).  With the dynamic schedule
    rank 0 composes the launches and the other ranks replay its launch log.  `warmup`: --warmup_iters splits the run into an
    untimed and a timed phase (bench.py); the chains continue across the mark, so the trajectory is that of one run."""
    import subprocess
    import uuid

    from host_util import read_outputs

    D, r = _tutorial()
    base = ["--n_iters=600", "--n_nodes=4", "--nu=1.0", "--max_scoring=true"]
    split = [f"--n_iters={600 - warmup}", f"--warmup_iters={warmup}"] if warmup else []
    d = os.path.join(str(tmp_path), "d.csv")
    np.savetxt(d, D, delimiter=",", fmt="%.0f")
    rf = os.path.join(str(tmp_path), "r.txt")
    np.savetxt(rf, r, fmt="%d")
    env = dict(os.environ, **CPU_ENV, SCICONE_THREADS=threads)
    shm = "/scicone_test_" + uuid.uuid4().hex[:12]
    procs = []
    for k in range(world):
        cmd = [NATIVE, f"--d_matrix_file={d}", f"--region_sizes_file={rf}", f"--n_cells={D.shape[0]}", f"--n_regions={D.shape[1]}",
               "--verbosity=2", f"--postfix=own{k}", "--seed=30", "--n_chains=5", f"--schedule={schedule}", "--min_batch=2", "--no_shard", f"--rank={k}", f"--world={world}",
               f"--shm_name={shm}"] + base + split
        procs.append(subprocess.Popen(cmd, cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env))
    for p in procs:
        out, err = p.communicate(timeout=600)
        assert p.returncode == 0, err[-2000:]
    for c in range(5):
        single = run_inference(REF, str(tmp_path), f"single{c}", D, r, base + [f"--seed={30 + c}"])
        owner = c % world
        got = {}
        for name in ("acceptance_ratio", "markov_chain", "gamma_values"):
            got[name] = np.array([float(x) for x in open(os.path.join(str(tmp_path), f"own{owner}_c{c}_{name}.csv")).read().strip().split(",")])
        assert np.array_equal(got["acceptance_ratio"], single["acc"]), c
        assert np.max(np.abs(got["markov_chain"] - single["chain"]) / np.maximum(1.0, np.abs(single["chain"]))) <= 1e-9
        assert open(os.path.join(str(tmp_path), f"own{owner}_c{c}_tree_inferred.txt")).read().split("\n")[7:] == single["tree"].split("\n")[7:]


@needs_bins
@pytest.mark.skipif(not os.path.exists(os.path.join(CPU_BACKEND_DIR, "libscicone_score.so")), reason="cpu_backend not built")
@pytest.mark.parametrize("seed", list(range(9000, 9012)))
def test_random_configurations_on_cpu(tmp_path, seed):
    """A slice of scripts/trajectory_sweep.py: random scoring mode, tree size, nu, move weights, tempering, limits, cluster
    sizes and shapes; the native host must reproduce the unmodified reference run for run."""
    import importlib.util

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("trajectory_sweep", os.path.join(root, "scripts", "trajectory_sweep.py"))
    sweep = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sweep)
    D, r, w, flags = sweep.draw_case(seed)
    a = run_inference(REF, str(tmp_path), "ref", D, r, flags, w=w)
    b = run_inference(NATIVE, str(tmp_path), "new", D, r, flags, w=w, env=CPU_ENV)
    compare_runs(a, b)


@needs_bins
@pytest.mark.skipif(not os.path.exists(os.path.join(CPU_BACKEND_DIR, "libscicone_score.so")), reason="cpu_backend not built")
def test_cli_errors_match_the_reference(tmp_path):
    """Same messages as the reference's CLI for missing arguments (infer_trees.cpp:157-176) and for a region-size vector of
    the wrong length (Tree.h:170-171); the reference aborts on the uncaught exception, the native host exits with 1."""
    import subprocess

    rng = np.random.default_rng(0)
    np.savetxt(str(tmp_path / "d.csv"), rng.poisson(50, size=(30, 6)).astype(float), delimiter=",", fmt="%.0f")
    np.savetxt(str(tmp_path / "r.txt"), rng.integers(3, 9, size=6), fmt="%d")
    np.savetxt(str(tmp_path / "r5.txt"), rng.integers(3, 9, size=5), fmt="%d")
    env = dict(os.environ, **CPU_ENV)
    cases = {
        "the D matrix file is not provided.": ["--region_sizes_file=r.txt", "--n_cells=30", "--n_regions=6"],
        "the number of regions is not provided.": ["--d_matrix_file=d.csv", "--region_sizes_file=r.txt", "--n_cells=30"],
        "the number of cells is not provided.": ["--d_matrix_file=d.csv", "--region_sizes_file=r.txt", "--n_regions=6"],
        "Size of the counts per cell needs to be equal to the number of regions!":
            ["--d_matrix_file=d.csv", "--region_sizes_file=r5.txt", "--n_cells=30", "--n_regions=6", "--n_iters=10", "--verbosity=0"],
    }
    for message, args in cases.items():
        ref = subprocess.run([REF] + args, cwd=str(tmp_path), capture_output=True, text=True, timeout=120)
        new = subprocess.run([NATIVE] + args, cwd=str(tmp_path), capture_output=True, text=True, timeout=120, env=env)
        assert ref.returncode != 0 and new.returncode == 1, (message, ref.returncode, new.returncode)
        assert message in ref.stdout + ref.stderr, message
        assert message in new.stdout + new.stderr, message
    # files that do not exist: the reference dies on an assertion, the native host says which file
    new = subprocess.run([NATIVE, "--d_matrix_file=nope.csv", "--region_sizes_file=r.txt", "--n_cells=30", "--n_regions=6", "--n_iters=10"],
                         cwd=str(tmp_path), capture_output=True, text=True, timeout=120, env=env)
    assert new.returncode == 1 and "cannot open nope.csv" in new.stderr


def _score_tool(tmp_path, env, exact):
    """The reference's `score` executable (score_tree.cpp) and scicone_b200/bin/score on a tree inferred from the tutorial
    matrix: same printed tree; the seven score lines identical (CPU, oracle-backed ABI) or within 1e-9 relative (GPU)."""
    import subprocess

    D, r = _tutorial()
    first = run_inference(REF, str(tmp_path), "seed", D, r, ["--n_iters=600", "--n_nodes=5", "--seed=3", "--nu=1.5", "--max_scoring=false"])
    assert first["tree"]
    tree_file = os.path.join(str(tmp_path), "seed_tree_inferred.txt")
    ref_score = os.path.join(os.path.dirname(REF), "score")
    native_score = os.path.join(os.path.dirname(NATIVE), "score")
    common = [f"--d_matrix_file={os.path.join(str(tmp_path), 'seed_d.csv')}", f"--region_sizes_file={os.path.join(str(tmp_path), 'seed_r.txt')}",
              f"--n_cells={D.shape[0]}", f"--n_regions={D.shape[1]}", f"--tree_file={tree_file}"]
    np.savetxt(os.path.join(str(tmp_path), "seed_d.csv"), D, delimiter=",", fmt="%.0f")
    np.savetxt(os.path.join(str(tmp_path), "seed_r.txt"), r, fmt="%d")
    for extra in (["--nu=1.5"], ["--nu=0.4", "--cf=0.5"], ["--is_overdispersed=0"]):
        a = subprocess.run([ref_score] + common + extra, capture_output=True, text=True, timeout=300)
        b = subprocess.run([native_score] + common + extra, capture_output=True, text=True, timeout=300,
                           env=dict(os.environ, **env) if env else None)
        assert a.returncode == 0 and b.returncode == 0, b.stderr[-1000:]
        if exact:
            assert a.stdout == b.stdout, extra
            continue
        la, lb = a.stdout.strip().split("\n"), b.stdout.strip().split("\n")
        assert len(la) == len(lb)
        for x, y in zip(la, lb):
            if ": " in x and x.split(": ")[0] in ("Tree posterior", "Tree prior", "Event prior", "Log likelihood", "Root score", "Tree score", "Nu"):
                vx, vy = float(x.split(": ")[1]), float(y.split(": ")[1])
                assert x.split(": ")[0] == y.split(": ")[0] and abs(vx - vy) <= 1e-9 * max(1.0, abs(vx)), (x, y)
            else:
                assert x == y


@needs_bins
@pytest.mark.skipif(not os.path.exists(os.path.join(CPU_BACKEND_DIR, "libscicone_score.so")), reason="cpu_backend not built")
def test_score_tool_on_cpu(tmp_path):
    _score_tool(tmp_path, CPU_ENV, exact=True)


@needs_bins
@pytest.mark.gpu
def test_score_tool_on_gpu(tmp_path):
    _score_tool(tmp_path, None, exact=False)
