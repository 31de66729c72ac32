#!/usr/bin/env python3
"""bench.py - MCMC tree-scoring throughput of the B200 engine (and of the reference's CPU path beside it).

    python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU `inference` on the host cores

Workload (BASELINE.json configs[2], the "10k cells" headline): synthetic 10 000 cells x 18 000 bins collapsed to
200 regions PER GPU, B independent MCMC chains (restarts, as pyscicone's learn_tree_parallel runs them) on trees of ~50
nodes.  One STEP = one MCMC iteration of every chain.  The run starts from the cells x BINS matrix: the bins -> regions
collapse (K0, Utils::condense_matrix) runs on the device first and its output - checked against the region-level
matrix it was expanded from - is the count matrix of everything that follows.

Metric: cell x MCMC-iterations per second ("cell_iters/s" = chain iterations/s x cells), which aggregates over
cell-sharded GPUs (weak scaling: every rank holds its own 10 000 cells) and over chains, and which the
reference's `inference` timer yields directly (iterations/s x cells).  Chain iterations/s and cell x node
attachment scores/s are reported next to it.

  e2e    THE HEADLINE.  The real MCMC: this repo's native host (scicone_b200/bin/inference - the reference's CLI, move
         set, priors and accept rule on top of the C ABI) runs the same B restarts with the same flags and seeds as the
         reference arm; every step stages its proposals from host memory, launches, and reads the B totals back before
         the chains can decide.  Timer = the host's own clock around infer_mcmc, exactly what the reference arm reports.
         `roofline_e2e` / `roofline_k1` are the kernels' numbers INSIDE this run.
  value  device pipeline with everything resident in HBM: pre-generated valid proposals (scicone_b200.synth.TreeState,
         six move kinds at the weights of the reference's default mix, pre-drawn accept coins) driven through the ctypes
         mirror of the C ABI, region timed with CUDA events on the engine's stream.  `roofline` comes from its first pass
         (one launch of all chains per step on one stream, where the proposal kernel runs alone).
  also   `other_configs`: the other named shapes of BASELINE.json through the same native host - configs[1] (1 000 x 50),
         configs[3] (cluster-tree restarts over the GPUs, scicone_b200/bin/learn_tree), configs[4] (100 000 cells, cell-
         sharded over the N GPUs) - and for N > 1 `sharded_equals_single`: a 5 000-cell run sharded over the N GPUs must
         write byte-identical traces to the same run on one GPU.
"""
import argparse
import json
import os
import re
import shutil
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L2_BYTES = 126e6


_RESULT_FD = None


def protect_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner, for one), so the real
    stdout is kept aside for the result and fd 1 points at stderr for everything else."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, line)


def workload_config(args, world):
    """The workload both arms run - the same dict on both lines."""
    B, M = args.chains, args.cells
    resident = (B * (args.nodes + 4) * M * 8 + M * args.regions * 10) / 1e6
    return {"workload": f"configs[2]: {M} cells x {args.bins} bins -> {args.regions} regions per GPU, {B} chains x "
                        f"~{args.nodes}-node trees, overdispersed, {'max' if args.max_scoring else 'sum'} scoring, default move mix",
            "cells_per_gpu": M, "bins": args.bins, "regions": args.regions, "chains": B, "tree_nodes": args.nodes,
            "max_scoring": int(args.max_scoring), "n_gpus": world,
            "parallelism": f"cells sharded x{world}" if world > 1 else "1 GPU",
            "l2": f"inputs larger than L2: ~{resident:.0f} MB of score rows + counts resident per GPU vs 126 MB L2"}


def threads_per_rank(world):
    """Host threads of one rank's process.  They spin while they wait (an MCMC iteration is ~100 us), so the ranks together
    must leave a core or two per rank to everything else on the box - a spinning thread that loses its core to the scheduler
    stalls every rank of a cell-sharded run for a time slice."""
    cores = os.cpu_count() or 1
    # one GPU: 16 spinning threads on a 16-core box gave 20-step windows of 136-214 us per step (a worker that loses its core
    # in the middle of a visit holds its chain for a time slice), 12 threads 134-139; 300-step runs are the same for 12-16
    # (the host work of 32 chains keeps ~6 cores busy; more than a dozen spinning workers only add contention on a bigger box)
    return max(2, cores // world - 1) if world > 1 else max(2, min(12, cores - 4))


def load_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of this command (profiles/)."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(p):
        p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(p):
        t = json.load(open(p))
        return t.get("dram_bytes_per_launch"), t.get("source")
    return None, None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc = None
        self.path = None
        if shutil.which("nvidia-smi"):
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def build_chain_traces(data, n_chains, n_nodes, n_steps, seed, nu):
    """Pre-generated proposal traces: per chain a start tree and per step (FlatTree, root indices, accept coin)."""
    from scicone_b200.synth import TreeState

    kinds = ("prune_reattach", "swap", "add_remove", "insert", "delete", "nu")
    # default move weights of the reference CLI (infer_trees.cpp:120): five weighted structural moves at 1.0
    # (prune/reattach, swap, add/remove, insert/delete, condense/split) and the nu move at 1.0
    weights = np.array([1.0, 1.0, 1.0, 1.0, 1.0, 1.0])
    weights = weights / weights.sum()
    traces = []
    for c in range(n_chains):
        rng = np.random.default_rng(seed + 42 + c)
        tree = TreeState(data["D"].shape[1], data["neutral_states"], rng, nu=nu).grow(n_nodes)
        start = tree.flat()
        steps = []
        for _ in range(n_steps):
            t2, roots, kind = tree.propose(kinds=kinds, weights=weights)
            ft = t2.flat()
            acc = bool(rng.random() < 0.3)
            steps.append((ft, [ft.index_of(r) for r in roots], acc, kind))
            if acc:
                tree = t2
        traces.append((start, steps))
    return traces


def run_engine(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    import torch.distributed as dist

    if world > 1:  # the ranks share the host cores: split them between the libraries' worker pools
        os.environ.setdefault("SCICONE_THREADS", str(threads_per_rank(world)))

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from scicone_b200.api import BatchLauncher, Chain, Dataset, comm_unique_id, condense_matrix
    from scicone_b200.synth import SEED, expand_to_bins, make_counts

    K, W, B = args.steps, args.warmup, args.chains
    M = args.cells
    # weak scaling: every rank draws its own 10k-cell shard of one (world x M)-cell matrix; region sizes are shared
    data = make_counts(M, args.bins, args.regions, seed=SEED)
    if world > 1:
        shard = make_counts(M, args.bins, args.regions, seed=SEED + 1000 * rank)
        shard["region_sizes"] = data["region_sizes"]
        data = shard if rank > 0 else data
    # K0: the run starts from cells x bins; the device collapses the bins of every region (Utils::condense_matrix) and the
    # result - which must equal the region-level matrix the bins were expanded from - is the count matrix of the run
    k0 = None
    if not args.no_k0:
        bins_mat = expand_to_bins(data["D"], data["region_sizes"], seed=SEED + rank)
        k0_us = []
        for _ in range(3):
            D_k0, us = condense_matrix(bins_mat, data["region_sizes"], device=local, return_kernel_us=True)
            k0_us.append(us)
        if not np.array_equal(D_k0, data["D"]):
            raise SystemExit("bench.py: the device bins->regions collapse does not reproduce the region-level matrix")
        k0_bytes = 8.0 * bins_mat.shape[0] * (bins_mat.shape[1] + args.regions)
        k0 = {"us": float(np.median(k0_us)), "alg_bytes": k0_bytes, "rows": int(bins_mat.shape[0]), "bins": int(bins_mat.shape[1])}
        data["D"] = D_k0
        del bins_mat
    ds = Dataset(data["D"], data["region_sizes"], neutral_states=data["neutral_states"], cluster_sizes=data["cluster_sizes"],
                 is_overdispersed=1, max_scoring=args.max_scoring, device=local, cell_offset=rank * M, n_cells_global=world * M)
    if world > 1:
        uid = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ds.attach_comm(uid[0], rank, world)
    # identical proposal streams on every rank (the tree is replicated, only cells are sharded)
    ref_data = make_counts(64, args.bins, args.regions, seed=SEED)  # only region count / neutral states matter here
    traces = build_chain_traces(ref_data, B, args.nodes, 2 * (W + K), SEED, args.nu)
    ds.reserve(B, B * (3 * (args.nodes + 24) + 48))  # reduction scratch and row pool are sized before the first launch
    chains = [Chain(dataset=ds) for _ in range(B)]
    for ch, (start, _) in zip(chains, traces):
        ch.compute_t_table(start)
        ch.compute_t_od_scores(start.nu)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device pipeline, timed twice over K steps each (CUDA events on the engine's stream):
    #   pass 1  one launch of all B proposals per step on ONE stream, totals collected with a lag (accept / reject are
    #           host-side pointer swaps issued before the wait): the proposal kernel runs alone on the device, so its
    #           CUDA-event durations are clean -> `roofline`
    #   pass 2  two groups of chains that take turns on the dataset's two launch lanes (a launch is collected before its
    #           chains are settled), so K1 of one launch overlaps the proposal kernel of the other -> `value`
    # ctypes arguments of every step are built before the timed regions: the loops are three library calls per launch.
    import gc

    n_steps = 2 * (W + K)

    def marshal(groups, launchers):
        q_args, extras, s_args = [], [], []
        for s in range(n_steps):
            q_args.append([L.marshal_queue([traces[b][1][s][0] for b in g], [traces[b][1][s][1][0] for b in g]) for L, g in zip(launchers, groups)])
            extras.append([[(chains[b], traces[b][1][s][0], extra) for b in g for extra in traces[b][1][s][1][1:]] for g in groups])
            s_args.append([L.marshal_settle([traces[b][1][s][2] for b in g]) for L, g in zip(launchers, groups)])
        return q_args, extras, s_args

    def queue_step(launchers, q_args, extras, s, gi):  # a nu proposal gets its root score from the same launch
        launchers[gi].queue(marshalled=q_args[s][gi])
        for ch, ft, extra in extras[s][gi]:  # label swaps rescore two subtrees
            ch.compute_t_prime_scores(ft, extra)

    # pass 1
    one = [BatchLauncher(chains)]
    q1, x1, s1 = marshal([list(range(B))], one)

    def run_one_lane(first, last):
        tickets = []
        for s in range(first, last):
            queue_step(one, q1, x1, s, 0)
            tickets.append(one[0].submit())
            one[0].settle(marshalled=s1[s][0])
            if len(tickets) >= 6:
                one[0].wait(tickets.pop(0))
        for tk in tickets:
            one[0].wait(tk)

    ds.set_profiling(True)
    run_one_lane(0, W)
    ds.kernel_time_stats(reset=True)
    ds.counters(reset=True)
    launches1 = ds.kernel_launches()
    gc.collect()
    gc.disable()  # the traces are ~10^6 small Python objects: no collector pauses inside the timed loops
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ds.region_mark(0)
    run_one_lane(W, W + K)
    ds.region_mark(1)
    barrier()
    one_lane_ms = ds.region_elapsed_ms()
    kern_us, n_timed = ds.kernel_time_stats()
    cnt_dev = ds.counters()
    launches_dev = ds.kernel_launches() - launches1
    ds.set_profiling(False)

    # pass 2
    two_lanes = B >= 2  # cell-sharded: each lane has a communicator of its own, all ranks launch in the same order
    dev_ms = one_lane_ms
    two_lane_ms = None
    value_pass = "one launch of all chains per step on one stream, totals collected with a lag"
    if two_lanes:
        groups = [list(range(0, B // 2)), list(range(B // 2, B))]
        launchers = [BatchLauncher([chains[b] for b in g]) for g in groups]
        gc.enable()
        q2, x2, s2 = marshal(groups, launchers)
        ds.set_streams(2)

        def run_two_lanes(first, last):
            tickets = [None] * len(groups)
            for s in range(first, last):
                for gi in range(len(groups)):
                    if tickets[gi] is not None:
                        launchers[gi].wait(tickets[gi])
                        launchers[gi].settle(marshalled=s2[s - 1][gi])
                    queue_step(launchers, q2, x2, s, gi)
                    tickets[gi] = launchers[gi].submit()
            for gi in range(len(groups)):
                launchers[gi].wait(tickets[gi])
                launchers[gi].settle(marshalled=s2[last - 1][gi])

        base = W + K
        run_two_lanes(base, base + W)
        launches2 = ds.kernel_launches()
        gc.collect()
        gc.disable()
        barrier()
        ds.region_mark(0)
        run_two_lanes(base + W, base + W + K)
        ds.region_mark(1)
        barrier()
        two_lane_ms = ds.region_elapsed_ms()
        if world > 1:  # the choice between the passes is made on the max over ranks, so every rank makes the same one
            t = torch.tensor([one_lane_ms, two_lane_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            one_lane_ms, two_lane_ms = (float(x) for x in t.cpu())
            dev_ms = one_lane_ms
        if two_lane_ms < one_lane_ms:  # `value` is the faster of the two ways to drive the device pipeline
            dev_ms = two_lane_ms
            launches_dev = ds.kernel_launches() - launches2
            value_pass = "two groups of chains on two launch lanes"
        ds.set_streams(1)
    gc.enable()

    # ---------------- e2e: the native MCMC host on the same matrix (all ranks: cell-sharded run of ONE set of chains)
    dd = dist if world > 1 else None
    host = None if args.no_e2e else run_native_host(args, data, rank, world, local, dd, comm_unique_id, profile_kernels=True)
    timed = None if args.no_e2e else run_native_host(args, data, rank, world, local, dd, comm_unique_id, profile_kernels=False)
    e2e_s = timed["wall_us"] * 1e-6 if timed else float("nan")
    if host and timed:
        host["wall_us_with_kernel_events"] = host["wall_us"]
        for key in ("wall_us", "h2d_bytes", "d2h_bytes", "accepted", "rejected", "gpu_launches", "launches", "launched_chains", "host_threads"):
            host[key] = timed[key]
    clocks = sampler.stop() if sampler else None
    others = None if (args.no_e2e or args.no_other_configs) else other_configs(args, rank, world, local, dd, comm_unique_id)
    equal = sharded_equals_single(rank, world, local, dist, comm_unique_id) if (world > 1 and not args.no_e2e) else None

    # max over ranks
    if world > 1:
        t = torch.tensor([dev_ms, e2e_s, kern_us, k0["us"] if k0 else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, kern_us, k0_us_max = (float(x) for x in t.cpu())
        if k0:
            k0["us"] = k0_us_max
        if host:  # the kernels' times inside the e2e run: max over ranks as well
            t = torch.tensor([host["score_kernel_us"], host["build_kernel_us"]], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            host["score_kernel_us"], host["build_kernel_us"] = (float(x) for x in t.cpu())

    if rank == 0:
        peak, peak_src = load_peaks()
        traffic, traffic_src = load_traffic()
        cells_total = world * M
        iters = B * K
        value = iters * cells_total / (dev_ms * 1e-3)
        e2e = iters * cells_total / e2e_s
        if host is None:  # profiling runs (--no_e2e): placeholders, never a bench line
            host = {"h2d_bytes": 0, "d2h_bytes": 0, "host_threads": 0, "accepted": 0, "rejected": 0, "gpu_launches": 0,
                    "score_kernel_us": 0.0, "build_kernel_us": 0.0, "alg_bytes": 0.0, "build_alg_bytes": 0.0, "build_lgamma": 0.0,
                    "score_kernel_launches": 0, "build_kernel_launches": 0, "launches": 0, "launched_chains": 0}
        achieved = cnt_dev["alg_bytes"] / (kern_us * 1e-6) / 1e9 if kern_us > 0 else 0.0

        def roof(alg_bytes, us, **more):
            gbs = alg_bytes / (us * 1e-6) / 1e9 if us > 0 else 0.0
            return {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, **more}

        out = {
            "metric": "mcmc_cell_iters_per_s", "value": value, "unit": "cell_iters/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(args, world),
            "value_how": "device pipeline, inputs resident in HBM: pre-generated valid proposals (scicone_b200.synth.TreeState: "
                         "prune/reattach, swap, add/remove, insert, delete, nu at equal weights - the reference's default mix has "
                         "condense/split in place of a second insert/delete), accept prob 0.3; " + value_pass
                         + f" (one stream: {one_lane_ms / K:.4f} ms/step"
                         + (f", two lanes: {two_lane_ms / K:.4f} ms/step" if two_lane_ms is not None else "") + ")",
            "mcmc_iters_per_s": iters / (dev_ms * 1e-3),
            "scores_per_s": cnt_dev["n_scores"] * world / (dev_ms * 1e-3),
            "e2e": {"value": e2e, "unit": "cell_iters/s", "mcmc_iters_per_s": iters / e2e_s, "mcmc_iters_per_s_per_chain": K / e2e_s,
                    "ms_per_step": 1e3 * e2e_s / K, "warmup_steps": e2e_warmup(args),
                    "h2d_bytes_per_step": host["h2d_bytes"] / K, "d2h_bytes_per_step": host["d2h_bytes"] / K,
                    "what": "scicone_b200/bin/inference (native MCMC host on the C ABI): real moves + Metropolis-Hastings, "
                            f"{B} restarts (seeds 42..), {host['host_threads']} host threads per rank, timer around infer_mcmc; the warm-up "
                            "iterations are a phase of the same call (all chains stop at the mark, nothing in flight, then the clock starts)",
                    "accepted": host["accepted"], "rejected": host["rejected"], "gpu_launches": host["gpu_launches"],
                    "launches": host["launches"], "proposals_per_launch": host["launched_chains"] / max(1, host["launches"]),
                    "score_kernel_us_per_step": host["score_kernel_us"] / K, "build_kernel_us_per_step": host["build_kernel_us"] / K,
                    "kernel_times_from": "a second identical run (same seeds, same trajectories) with CUDA events around every kernel; that run took "
                                         f"{host.get('wall_us_with_kernel_events', 0) * 1e-3 / K:.4f} ms per step"},
            "gpu_launches": int(launches_dev),
            "roofline": roof(cnt_dev["alg_bytes"], kern_us, kernel="score_proposals_kernel", traffic=traffic, traffic_source=traffic_src,
                             peak_source=peak_src, alg_bytes_per_launch=cnt_dev["alg_bytes"] / max(1, n_timed),
                             avg_launch_us=kern_us / max(1, n_timed), kernel_share_of_step=kern_us * 1e-3 / one_lane_ms,
                             bytes_model="SURVEY.md 8d: 16 + 8 E_n per rescored cell x node, 8 N + 12 per cell aggregate; "
                                         "nu move / full pass: 8 K + 16 N + 12 per cell",
                             measured_in="value pass 1: one launch of all chains per step on one stream (the kernel runs alone); "
                                         f"{one_lane_ms / K:.4f} ms per step there",
                             note="algorithmic bytes are what the reference's algorithm touches per unit (SURVEY.md 8d); this kernel reads "
                                  "only the rows an aggregate needs and takes the data terms from 2-byte count indices and L2-resident tables, "
                                  "so its DRAM traffic (`traffic`) is a fraction of them and frac can exceed 1: the kernel is bound by the latency "
                                  "of dependent row gathers and by issue slots (ncu: issue active 42 %, DRAM 14 %), not by HBM bandwidth; "
                                  "roofline_e2e is the same kernel inside the real MCMC",
                             traffic_over_algorithmic=(traffic / (cnt_dev["alg_bytes"] / max(1, n_timed))) if traffic else None),
            # the same kernel inside the REAL MCMC (launches of ~10 proposals, two to three in flight, other kernels beside it)
            "roofline_e2e": roof(host["alg_bytes"], host["score_kernel_us"], kernel="score_proposals_kernel",
                                 launches=host["score_kernel_launches"],
                                 avg_launch_us=host["score_kernel_us"] / max(1, host["score_kernel_launches"]),
                                 measured_in="e2e run: CUDA events around every launch of the proposal kernel, summed (max over ranks)"),
            # K1a + K1b (value tables and increments of every rescored node, the root score of nu proposals)
            "roofline_k1": roof(host["build_alg_bytes"], host["build_kernel_us"], kernel="build_tables_kernel + build_increments_kernel",
                                launches=host["build_kernel_launches"], lgamma_per_s=host["build_lgamma"] / (host["build_kernel_us"] * 1e-6)
                                if host["build_kernel_us"] > 0 else 0.0,
                                bytes_model="per rescored node of a tabulated proposal and cell: 8 written + 2 (count index) or 8 (count) read per "
                                            "event; root score: 8 per histogram entry (datasets with count histograms), else per slice and cell "
                                            "8 written + 2 or 8 per region; the kernels are launch-floor bound, the byte fraction only documents that",
                                measured_in="e2e run: CUDA events around K1a + K1b of every launch that has them"),
            "roofline_k0": roof(k0["alg_bytes"], k0["us"], kernel="condense_rows_kernel", rows=k0["rows"], bins=k0["bins"], kernel_us=k0["us"],
                                bytes_model="8 n_bins read + 8 K written per row",
                                measured_in="dataset set-up of this run: median of 3 collapses of the rank's cells x bins matrix, CUDA events") if k0 else None,
            "other_configs": others,
            "sharded_equals_single": equal,
            "clocks": clocks,
        }
        out["cpu_baseline"] = cpu_baseline(args, data) if world == 1 and not args.no_cpu_baseline else None
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def host_run(tag, D_full, region_sizes, rank, world, local, dist, comm_unique_id, n_chains, n_nodes, steps, warmup, max_scoring, nu,
             extra=(), cluster_sizes=None, ranks=None):
    """One run of scicone_b200/bin/inference (the native MCMC host on the C ABI) over `world` GPUs - every rank starts its own
    process on its own device, cells sharded, tree replicated - or, with ranks=[0], on rank 0's GPU alone while the others
    wait.  D_full is only read on rank 0 (it writes the matrix to a directory all ranks of the box see).  Returns the
    stats JSON of this rank's process (None on ranks that did not run)."""
    exe = os.path.join(ROOT, "scicone_b200", "bin", "inference")
    if not os.path.exists(exe):
        raise SystemExit("bench.py: scicone_b200/bin/inference is missing - run __graft_entry__.build()")
    solo = ranks is not None
    w_run = 1 if solo else world
    if world > 1:
        box = [tempfile.mkdtemp(prefix=f"scicone_{tag}_") if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        tmp = box[0]
        uid = [comm_unique_id().hex() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        shape = [tuple(D_full.shape) if rank == 0 else None]
        dist.broadcast_object_list(shape, src=0)
        n_rows, Kr = shape[0]
    else:
        tmp, uid = tempfile.mkdtemp(prefix=f"scicone_{tag}_"), [None]
        n_rows, Kr = D_full.shape
    try:
        d = os.path.join(tmp, "d.f64")
        r = os.path.join(tmp, "r.txt")
        wf = os.path.join(tmp, "w.txt")
        if rank == 0:
            np.ascontiguousarray(D_full, dtype="<f8").tofile(d)
            np.savetxt(r, region_sizes, fmt="%d")
            if cluster_sizes is not None:
                np.savetxt(wf, cluster_sizes, fmt="%d")
        if world > 1:
            dist.barrier()
        out = None
        if not solo or rank in ranks:
            stats = os.path.join(tmp, f"stats{rank}.json")
            cmd = [exe, f"--d_matrix_file={d}", f"--region_sizes_file={r}", f"--n_cells={n_rows}", f"--n_regions={Kr}",
                   f"--n_iters={steps}", f"--warmup_iters={warmup}", f"--n_nodes={n_nodes}", "--seed=42", f"--nu={nu}",
                   "--verbosity=0", "--no_cnv_file=1", f"--max_scoring={'true' if max_scoring else 'false'}", f"--postfix={tag}{rank}",
                   f"--n_chains={n_chains}", f"--device={local}", f"--stats_file={stats}", "--stats_all_ranks=1"]
            if cluster_sizes is not None:
                cmd.append(f"--cluster_sizes_file={wf}")
            if w_run > 1:  # rank 0 composes the launches, the other ranks replay its launch log
                cmd += ["--schedule=dynamic", f"--rank={rank}", f"--world={world}", f"--nccl_id={uid[0]}"]
            cmd += list(extra)
            env = dict(os.environ)
            env["SCICONE_THREADS"] = str(threads_per_rank(w_run))  # a run on one GPU has the box's cores to itself (the other ranks sleep)
            res = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True, env=env)
            if res.returncode != 0:
                raise SystemExit(f"bench.py: native host failed ({tag}):\n" + res.stdout[-2000:] + res.stderr[-2000:])
            out = json.load(open(stats))
            keep = os.environ.get("SCICONE_BENCH_KEEP_STATS")  # tuning runs: keep every rank's stats file (launch timeline etc.)
            if keep:
                os.makedirs(keep, exist_ok=True)
                shutil.copy(stats, os.path.join(keep, f"{tag}_stats{rank}_{int(time.time() * 1000) % 100000}.json"))
            if solo and world > 1:
                open(os.path.join(tmp, f"done{rank}"), "w").close()
        elif world > 1:  # a rank that sits this run out waits for a file, not in an NCCL barrier (which spins on its GPU and core)
            while not all(os.path.exists(os.path.join(tmp, f"done{r}")) for r in ranks):
                time.sleep(0.02)
        if world > 1:
            dist.barrier()
        return out
    finally:
        if rank == 0:
            shutil.rmtree(tmp, ignore_errors=True)


def run_native_host(args, data, rank, world, local, dist, comm_unique_id, profile_kernels):
    """The e2e leg: real MCMC on the bench workload.  With world > 1 the ranks' matrices are stacked into one (world x
    cells)-cell matrix and every rank's host takes its chunk-aligned block of rows.  profile_kernels: CUDA events around
    every kernel (three event records per launch, and the collect waits for the kernel to retire instead of for its
    results) - the timed run goes without, a second identical run (same seeds, same trajectories) yields the per-kernel times."""
    full = data["D"]
    if world > 1:
        import torch

        mine = torch.from_numpy(np.ascontiguousarray(data["D"])).cuda()
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        full = torch.cat(parts).cpu().numpy() if rank == 0 else None
    extra = args.host_args.split() + [f"--profile_kernels={1 if profile_kernels else 0}"]
    if args.n_groups:
        extra += [f"--n_groups={args.n_groups}", "--schedule=static"]
    return host_run("e2e", full, data["region_sizes"], rank, world, local, dist, comm_unique_id, args.chains, args.nodes - 1, args.steps,
                    e2e_warmup(args), args.max_scoring, args.nu, extra=extra)


def e2e_warmup(args):
    """Untimed iterations in front of the K timed ones of the e2e leg: at least W, and at least 100 (about 12 ms) - the host
    side of an MCMC iteration (tree moves on a dozen threads, their allocator arenas and caches) is not warm after 5
    iterations of 32 chains, and a real run has thousands of iterations.  The number used is reported as e2e.warmup_steps."""
    return max(args.warmup, args.e2e_min_warmup)


def max_over_ranks(dist, world, x):
    if world == 1:
        return x
    import torch

    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.cpu()[0])


def other_configs(args, rank, world, local, dist, comm_unique_id):
    """The other named shapes of BASELINE.json through the same native host, each a short real-MCMC run (same K and W):
    configs[1] 1 000 cells x 10 000 bins -> 50 regions (N = 1 only), configs[3] cluster-tree restarts spread over the N
    GPUs by scicone_b200/bin/learn_tree (replicas only, no collective), configs[4] 100 000 cells x 20 000 bins -> 200 regions,
    cell-sharded over the N GPUs (strong scaling: the same matrix on 1 GPU is timed beside it)."""
    from scicone_b200.synth import SEED, make_counts

    K, W = args.steps, e2e_warmup(args)  # the same number of untimed iterations as the e2e leg
    out = {}
    noprof = ["--profile_kernels=0"]  # timed runs: no CUDA events around the kernels

    def record(st, wall_s, n_chains, n_cells, n_gpus, **more):
        rec = {"mcmc_iters_per_s": n_chains * K / wall_s, "cell_iters_per_s": n_chains * K * n_cells / wall_s, "ms_per_step": 1e3 * wall_s / K,
               "chains": n_chains, "cells": n_cells, "n_gpus": n_gpus, "steps": K, "warmup": W}
        if st is not None:
            rec.update({"accepted": st["accepted"], "rejected": st["rejected"], "launches": st["launches"], "alg_bytes_per_step": st["alg_bytes"] / K})
        rec.update(more)
        return rec

    if world == 1:
        d1 = make_counts(1000, 10000, 50, seed=SEED + 11)
        st = host_run("c1", d1["D"], d1["region_sizes"], rank, world, local, dist, comm_unique_id, args.chains, 49, K, W, args.max_scoring, args.nu, extra=noprof)
        out["configs[1]"] = record(st, st["wall_us"] * 1e-6, args.chains, 1000, 1, what="1 000 cells x 10 000 bins -> 50 regions, 50-node trees")
        # Head room of the bench workload itself: twice the restarts on the same GPU (restarts are what pyscicone's
        # learn_tree_parallel multiplies).  A launch of 18 proposals costs little more than one of 9, so the extra restarts
        # are close to free; the headline stays at the 32 restarts of earlier rounds.
        d2 = make_counts(args.cells, args.bins, args.regions, seed=SEED)
        st = host_run("c2x2", d2["D"], d2["region_sizes"], rank, world, local, dist, comm_unique_id, 2 * args.chains, args.nodes - 1, K, W, args.max_scoring,
                      args.nu, extra=noprof)
        out["configs[2] with twice the restarts"] = record(st, st["wall_us"] * 1e-6, 2 * args.chains, args.cells, 1,
                                                           what=f"the bench workload with {2 * args.chains} restarts instead of {args.chains} (informational)")
    # configs[4]
    n4, c4 = 100000, 16
    d4 = make_counts(n4, 20000, 200, seed=SEED + 14) if rank == 0 else None
    D4 = d4["D"] if rank == 0 else None
    r4 = d4["region_sizes"] if rank == 0 else None
    if world > 1:
        box = [r4]
        dist.broadcast_object_list(box, src=0)
        r4 = box[0]
    st1 = host_run("c4one", D4, r4, rank, world, local, dist, comm_unique_id, c4, 49, K, W, args.max_scoring, args.nu, ranks=[0], extra=noprof)
    one_s = max_over_ranks(dist, world, st1["wall_us"] * 1e-6 if st1 else 0.0)
    rec = {"what": "100 000 cells x 20 000 bins -> 200 regions, 16 chains, 50-node trees; strong scaling of ONE matrix",
           "one_gpu": record(st1, one_s, c4, n4, 1) if rank == 0 else None}
    if world > 1:
        stn = host_run("c4", D4, r4, rank, world, local, dist, comm_unique_id, c4, 49, K, W, args.max_scoring, args.nu, extra=noprof)
        n_s = max_over_ranks(dist, world, stn["wall_us"] * 1e-6)
        rec["sharded"] = record(stn, n_s, c4, n4, world)
        rec["speedup_vs_one_gpu"] = one_s / n_s
        rec["strong_scaling_efficiency"] = one_s / n_s / world
    out["configs[4]"] = rec
    # configs[3]: cluster-tree mode - condensed centroids with cluster_sizes, many restarts in parallel across the GPUs.
    # learn_tree runs one inference process per GPU; the other ranks of this bench must leave their GPUs alone meanwhile (a
    # rank waiting in an NCCL barrier keeps a spinning kernel on its GPU), so they wait for a file instead.
    sentinel = None
    if world > 1:
        box = [tempfile.mktemp(prefix="scicone_c3_done_") if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        sentinel = box[0]
        import torch

        torch.cuda.synchronize()
    if rank == 0:
        d3 = make_counts(10000, 18000, 200, seed=SEED + 13, cluster_rows=40)
        tmp = tempfile.mkdtemp(prefix="scicone_c3_")
        try:
            np.ascontiguousarray(d3["D"], dtype="<f8").tofile(os.path.join(tmp, "d.f64"))
            np.savetxt(os.path.join(tmp, "r.txt"), d3["region_sizes"], fmt="%d")
            np.savetxt(os.path.join(tmp, "w.txt"), d3["cluster_sizes"], fmt="%d")
            n_reps, iters = 32 * world, 2000
            cmd = [os.path.join(ROOT, "scicone_b200", "bin", "learn_tree"), f"--n_reps={n_reps}", f"--n_gpus={world}", "--max_tries=1",
                   "--postfix=c3", "--d_matrix_file=d.f64", "--region_sizes_file=r.txt", "--cluster_sizes_file=w.txt",
                   f"--n_cells={d3['D'].shape[0]}", "--n_regions=200", f"--n_iters={iters}", "--n_nodes=10", f"--nu={args.nu}", "--verbosity=1",
                   "--no_cnv_file=1", f"--max_scoring={'true' if args.max_scoring else 'false'}"]
            t0 = time.perf_counter()
            res = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True)
            wall = time.perf_counter() - t0
            if res.returncode != 0:
                raise SystemExit("bench.py: learn_tree failed:\n" + res.stdout[-2000:] + res.stderr[-2000:])
            times = [int(x) * 1e-6 for x in re.findall(r"Time taken by infer_mcmc function: (\d+) microseconds", res.stdout)]
            lt = json.load(open(os.path.join(tmp, "c3_learn_tree.json")))
            out["configs[3]"] = {"what": f"cluster-tree mode: 10 000 cells condensed to {d3['D'].shape[0]} centroids with cluster_sizes, {n_reps} restarts x "
                                         f"{iters} iterations over {world} GPU(s) by bin/learn_tree (replicas only); best-of-N + robustness as pyscicone",
                                 "mcmc_iters_per_s": n_reps * iters / max(times) if times else None, "process_wall_s": wall, "restarts": n_reps,
                                 "iters_per_restart": iters, "n_gpus": world, "robustness": lt["robustness"], "best_score": lt["best_score"],
                                 "best_restart": lt["best_rep"]}
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
            if sentinel:
                open(sentinel, "w").close()
    elif sentinel:
        while not os.path.exists(sentinel):
            time.sleep(0.05)
    if world > 1:
        dist.barrier()
        if rank == 0:
            os.unlink(sentinel)
    return out


def sharded_equals_single(rank, world, local, dist, comm_unique_id):
    """The N-GPU parity check of tests/test_gpu_multi.py inside the bench (the GPU test box has one GPU): the same 6 chains on
    5 000 cells on ONE GPU and cell-sharded over the N GPUs must write byte-identical traces and final trees."""
    from scicone_b200.synth import make_counts

    exe = os.path.join(ROOT, "scicone_b200", "bin", "inference")
    n_cells = 1024 * max(5, world) - 120  # every rank gets at least one 1024-cell chunk, the last one ragged
    box = [tempfile.mkdtemp(prefix="scicone_eq_") if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    tmp = box[0]
    uid = [comm_unique_id().hex() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    n_chains = 6
    common = [exe, "--d_matrix_file=d.f64", "--region_sizes_file=r.txt", f"--n_cells={n_cells}", "--n_regions=60", "--n_iters=300", "--n_nodes=12",
              "--seed=5", "--nu=1.0", "--verbosity=2", "--max_scoring=true", f"--n_chains={n_chains}", "--schedule=dynamic", "--min_batch=2", "--no_cnv_file=1"]
    ok = True
    try:
        if rank == 0:
            data = make_counts(n_cells, 9000, 60, seed=31)
            np.ascontiguousarray(data["D"], dtype="<f8").tofile(os.path.join(tmp, "d.f64"))
            np.savetxt(os.path.join(tmp, "r.txt"), data["region_sizes"], fmt="%d")
            one = subprocess.run(common + ["--postfix=one", f"--device={local}"], cwd=tmp, capture_output=True, text=True)
            ok = one.returncode == 0
        dist.barrier()
        env = dict(os.environ)
        env.setdefault("SCICONE_THREADS", str(threads_per_rank(world)))
        res = subprocess.run(common + [f"--postfix=shard{rank}", f"--device={local}", f"--rank={rank}", f"--world={world}", f"--nccl_id={uid[0]}"],
                             cwd=tmp, capture_output=True, text=True, env=env)
        flag = [res.returncode == 0]
        gathered = [None] * world
        dist.all_gather_object(gathered, flag)
        ok = ok and all(g[0] for g in gathered)
        if rank == 0 and ok:
            for c in range(n_chains):  # chain c is searched (and its traces written) by rank c % world
                for name in ("acceptance_ratio.csv", "markov_chain.csv", "gamma_values.csv", "tree_inferred.txt"):
                    a = open(os.path.join(tmp, f"one_c{c}_{name}")).read()
                    b = open(os.path.join(tmp, f"shard{c % world}_c{c}_{name}")).read()
                    ok = ok and a == b
        out = [ok]
        dist.broadcast_object_list(out, src=0)
        return bool(out[0])
    finally:
        dist.barrier()
        if rank == 0:
            shutil.rmtree(tmp, ignore_errors=True)


# ------------------------------------------------------------------------------------------------
def write_inputs(tmp, data):
    d = os.path.join(tmp, "d.csv")
    np.savetxt(d, data["D"], delimiter=",", fmt="%.0f" if np.allclose(data["D"], np.round(data["D"])) else "%.18e")
    r = os.path.join(tmp, "r.txt")
    np.savetxt(r, data["region_sizes"], fmt="%d")
    return d, r


def run_reference_inference(args, data, n_iters, n_procs, omp_threads):
    """The UNMODIFIED reference `inference` (oracle/_ref, built by oracle/Makefile) on this box's host cores."""
    exe = os.path.join(ROOT, "oracle", "_ref", "inference")
    if not os.path.exists(exe):
        return None
    tmp = tempfile.mkdtemp(prefix="scicone_ref_")
    try:
        d, r = write_inputs(tmp, data)
        M, K = data["D"].shape
        procs = []
        env = dict(os.environ, OMP_NUM_THREADS=str(omp_threads))
        t0 = time.perf_counter()
        for p in range(n_procs):
            cmd = [exe, f"--d_matrix_file={d}", f"--region_sizes_file={r}", f"--n_cells={M}", f"--n_regions={K}",
                   f"--n_iters={n_iters}", f"--n_nodes={args.nodes - 1}", f"--seed={42 + p}", f"--nu={args.nu}", "--verbosity=1",
                   f"--max_scoring={'true' if args.max_scoring else 'false'}", f"--postfix=ref{p}"]
            procs.append(subprocess.Popen(cmd, cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env))
        times = []
        for pr in procs:
            out, _ = pr.communicate()
            m = re.search(r"Time taken by infer_mcmc function: (\d+) microseconds", out)
            if pr.returncode != 0 or not m:
                raise RuntimeError("reference inference failed:\n" + out[-2000:])
            times.append(int(m.group(1)) * 1e-6)
        wall = time.perf_counter() - t0
        return {"mcmc_s_max": max(times), "wall_s": wall, "iters_per_s": n_procs * n_iters / max(times)}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def cpu_baseline(args, data):
    """Bounded sample of the same workload on the host CPU: the reference binary if it travelled, else the oracle port."""
    M = data["D"].shape[0]
    n_iters = args.cpu_iters
    res = run_reference_inference(args, data, n_iters, 1, os.cpu_count() or 1)
    if res is not None:
        return {"value": res["iters_per_s"] * M, "unit": "cell_iters/s", "mcmc_iters_per_s": res["iters_per_s"],
                "cores": os.cpu_count(), "kind": "reference",
                "sample": f"1 chain x {n_iters} iterations of oracle/_ref/inference (unmodified reference, -O3 -fopenmp, no ASan), "
                          f"OMP_NUM_THREADS={os.cpu_count()}, same matrix, random {args.nodes}-node start tree, infer_mcmc timer"}
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_ffi import OracleChain

    traces = build_chain_traces(data, 1, args.nodes, n_iters, 7, args.nu)
    ch = OracleChain(D=data["D"], region_sizes=data["region_sizes"], neutral_states=data["neutral_states"],
                     cluster_sizes=data["cluster_sizes"], is_overdispersed=1, max_scoring=args.max_scoring)
    start, steps = traces[0]
    ch.compute_t_table(start)
    t0 = time.perf_counter()
    for ft, roots, acc, kind in steps:
        if kind == "nu":
            ch.compute_t_prime_od_scores(ft.nu)
        for idx in roots:
            ch.compute_t_prime_scores(ft, idx)
        ch.compute_t_prime_sums(ft)
        ch.accept(ft) if acc else ch.reject()
    dt = time.perf_counter() - t0
    return {"value": n_iters * M / dt, "unit": "cell_iters/s", "mcmc_iters_per_s": n_iters / dt, "cores": 1, "kind": "port",
            "sample": f"1 chain x {n_iters} proposals through oracle/scicone_oracle.c (scoring only, no move logic)"}


def run_reference(args):
    """The reference arm: the UNMODIFIED reference `inference` (oracle/_ref) on the box's host cores, on OUR arm's config -
    the same matrix shape, the same number of chains (one single-threaded process per chain, all running at once: the
    reference is one chain per process and pyscicone runs restarts exactly so), the same seeds, flags and metric."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from scicone_b200.synth import SEED, make_counts

    M = args.cells * world  # our arm's config at N GPUs: every GPU holds args.cells cells
    data = make_counts(M, args.bins, args.regions, seed=SEED)
    cores = os.cpu_count() or 1
    n_procs = max(1, args.chains)
    K, W = args.steps, args.warmup
    # one "step" = a bounded sample: every process advances its chain by `iters_per_step` iterations
    iters_per_step = max(1, args.ref_iters_per_step)
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "inference")):
        emit({"impl": "reference", "unavailable": "oracle/_ref/inference was not built (needs /root/reference at build time)"})
        return
    for _ in range(min(W, 1)):
        run_reference_inference(args, data, 2, n_procs, 1)
    t0 = time.perf_counter()
    res = run_reference_inference(args, data, iters_per_step * K, n_procs, 1)
    wall = time.perf_counter() - t0
    v = res["iters_per_s"] * M
    sample = (f"{n_procs} concurrent restarts (seeds 42..) of the unmodified reference `inference` (oracle/_ref, -O3 -fopenmp, no ASan), "
              f"OMP_NUM_THREADS=1 each on {cores} cores, {iters_per_step * K} iterations each on {M} cells x {args.regions} regions, "
              f"random {args.nodes}-node start trees; rate = chains x iterations / max infer_mcmc time")
    out = {"impl": "reference", "metric": "mcmc_cell_iters_per_s", "value": v, "unit": "cell_iters/s", "n_gpus": world, "steps": K,
           "warmup": W, "ms_per_step": 1e3 * res["mcmc_s_max"] / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": workload_config(args, world),
           "mcmc_iters_per_s": res["iters_per_s"], "mcmc_iters_per_s_per_chain": res["iters_per_s"] / n_procs, "wall_s": wall,
           "cpu_baseline": {"value": v, "unit": "cell_iters/s", "cores": cores, "kind": "reference", "sample": sample},
           "e2e": {"value": v, "unit": "cell_iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=10000, help="cells per GPU")
    ap.add_argument("--bins", type=int, default=18000)
    ap.add_argument("--regions", type=int, default=200)
    ap.add_argument("--nodes", type=int, default=50)
    ap.add_argument("--chains", type=int, default=32)
    ap.add_argument("--nu", type=float, default=1.0)
    ap.add_argument("--max_scoring", type=int, default=1)
    ap.add_argument("--cpu_iters", type=int, default=150)
    ap.add_argument("--ref_iters_per_step", type=int, default=1)
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--no_e2e", action="store_true", help="skip the native-host run (profiling passes only)")
    ap.add_argument("--e2e_min_warmup", type=int, default=100, help="the e2e leg runs at least this many untimed iterations first")
    ap.add_argument("--no_k0", action="store_true", help="start from the region-level matrix (profiling passes only)")
    ap.add_argument("--no_other_configs", action="store_true", help="skip the configs[1]/[3]/[4] sub-records")
    ap.add_argument("--host_args", default="", help="extra flags for the native host, space separated (tuning runs)")
    ap.add_argument("--n_groups", type=int, default=0, help="chain groups of the native host's static schedule (cell-sharded runs)")
    args = ap.parse_args()
    protect_stdout()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
