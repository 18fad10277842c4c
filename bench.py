#!/usr/bin/env python
"""bench.py - IPM iterations/s and time-to-LP-optimum of the Mehrotra hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one LP solve (starting point + predictor-corrector loop to mu <= 1e-4, max_iter 100,
eta 0.95 - SURVEY.md 8d) of one synthetic set-covering LP relaxation.  Default workload =
BASELINE.json configs[1]: scpnrh-shaped (1000 x 10000, 5% density, costs 1..100), five rotating
instances (like scpnrh1-5) so a step never finds its inputs in L2.  Prints ONE JSON line.

  value : whole-job IPM iterations/s, models resident in HBM (CSR/CSC/symbolic structure loaded)
  e2e   : the same metric through the reference-facing call with HOST buffers: every step uploads
          the CSR model (sb200_load_model: H2D + CSC + symbolic), solves and reads x, y, s back
  roofline / cpu_baseline : see DESIGN.md "Measurement"

--impl reference times the CPU restatement of the reference's own solver (oracle/, NumPy+SciPy with
all host BLAS threads) on the same workload; /root/reference does not exist on the GPU box.
N > 1: one process per GPU (torchrun), independent LPs per rank, no data-path collective (the path
does not shard inside an LP: "weak" scaling, replicas of the LP workload).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

# OR-Library workloads: the real instances, shipped as compact archives (tests/golden/<name>.npz, written from the
# reference's data/ files by tests/golden/make_golden.py).  name: (instances, description)
ORLIB = {
    "scpnrh": ([f"scpnrh{i}" for i in range(1, 6)],
               "OR-Library scpnrh1-5 LP relaxations (1000x10000, 5% density), BASELINE.json configs[1]"),
    "scp4": (["scp41", "scp42", "scp46", "scp48", "scp49", "scp410"],
             "OR-Library scp41/42/46/48/49/410 LP relaxations (200x1000, 2% density), configs[0]"),
    "scpnre": ([f"scpnre{i}" for i in range(1, 6)], "OR-Library scpnre1-5 LP relaxations (500x5000, 10% density)"),
    "scpnrg": ([f"scpnrg{i}" for i in range(1, 6)], "OR-Library scpnrg1-5 LP relaxations (1000x10000, 2% density)"),
    "scpnrf": (["scpnrf1"], "OR-Library scpnrf1 LP relaxation (500x5000, 20% density), configs[2]"),
    "scpclr13": (["scpclr13"], "OR-Library scpclr13 LP relaxation (4095x715, 511 entries per column), configs[2]"),
}
WORKLOADS = {
    # synthetic look-alikes: name: (m, n_orig, density, description)
    "scpnrh-synth": (1000, 10000, 0.05, "scpnrh-shaped synthetic SCP LP relaxation 1000x10000, 5% density"),
    "scp4-synth": (200, 1000, 0.02, "scp4x-shaped synthetic SCP LP relaxation 200x1000, 2% density"),
    "scpnrf-synth": (500, 5000, 0.20, "scpnrf-shaped synthetic SCP LP relaxation 500x5000, 20% density"),
    "scpclr13-synth": (4095, 715, 0.125, "scpclr13-shaped synthetic SCP LP relaxation 4095x715, 12.5% density"),
    "synth5k": (5000, 100000, 0.001, "synthetic SCP 5000x100000, 0.1% density (configs[3] ladder rung)"),
    "synth50k": (50000, 1000000, 0.001, "synthetic SCP 50000x1000000, 0.1% density (configs[3])"),
}
N_INSTANCES = 5
MAX_ITER = 100
# FP64 denominators: MEASURED_PEAKS.json carries no FP64 entry.  bench.py measures a cuBLAS DGEMM 8192^3 live
# (SURVEY.md 8d names it as the FP64 tensor denominator) and quotes scripts/fp64_peak.cu's committed numbers
# (profiles/r2_a_fp64_peaks.json: DMMA issue rate 37.2, DGEMM 35.9 TFLOP/s) as the fallback
FP64_FALLBACK_TFLOPS = 35.9


def load_models(workload, rank=0):
    """-> (models, description, data) for a workload name."""
    if workload in ORLIB:
        from sypha_b200.instances import load_npz
        names, desc = ORLIB[workload]
        return [load_npz(REPO / "tests" / "golden" / f"{nm}.npz") for nm in names], desc, "orlib"
    from sypha_b200.instances import gen_scp
    m, n0, dens, desc = WORKLOADS[workload]
    k = 1 if workload == "synth50k" else N_INSTANCES     # 1.3 GB of structure per instance: larger than L2 on its own
    return [gen_scp(m, n0, dens, 1000 * rank + i + 1) for i in range(k)], desc, "synthetic"


def base_config(workload, models, desc):
    """The part of `config` both arms share."""
    mdl = models[0]
    return {"workload": desc, "instances": len(models), "m": int(mdl.m), "n_orig": int(mdl.n_orig),
            "nnz": int(mdl.nnz), "max_iter": MAX_ITER, "eta": 0.95, "mu_tol": 1e-4}


def measure_fp64_dgemm(n=8192, reps=4):
    """cuBLAS DGEMM n^3 through torch.matmul(float64) on the current device, best of `reps` (CUDA events)."""
    import torch
    try:
        a = torch.rand((n, n), dtype=torch.float64, device="cuda") - 0.5
        b = torch.rand((n, n), dtype=torch.float64, device="cuda") - 0.5
        torch.matmul(a, b)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        del a, b
        torch.cuda.empty_cache()
        return 2.0 * n ** 3 / best / 1e9, f"cuBLAS DGEMM {n}^3 measured live by bench.py (torch.matmul float64, best of {reps})"
    except Exception as e:           # never lose the line to the denominator
        return FP64_FALLBACK_TFLOPS, f"fallback: profiles/r2_a_fp64_peaks.json cuBLAS DGEMM 8192^3 ({e!r})"


def read_peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.stop_evt = threading.Event()
        self.ready = threading.Event()          # NVML is initialised (it takes driver locks: keep it out of the timed region)
        self.sm, self.maxsm, self.reasons = [], [], set()

    def run(self):
        """NVML in-process (no nvidia-smi process per sample: spawning it every 200 ms takes driver locks and
        slowed the host-side-heavy e2e steps several-fold); nvidia-smi only if NVML cannot be imported."""
        if os.environ.get("SB200_NO_SAMPLER"):
            self.ready.set()
            return
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            self.maxsm.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
            self.ready.set()
            while not self.stop_evt.is_set():
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = int(get_reasons(h))
                for nm, b in bits.items():
                    if r & b:
                        self.reasons.add(nm)
                self.stop_evt.wait(0.25)
            return
        except Exception:
            pass
        self.ready.set()
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                f = [t.strip() for t in out.strip().split(",")]
                self.sm.append(float(f[0]))
                self.maxsm.append(float(f[1]))
                for nm, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self.stop_evt.wait(1.0)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": float(max(self.maxsm)),
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def read_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch group from the committed ncu --set full captures."""
    p = REPO / "profiles" / "traffic.json"
    if p.exists():
        return {k: v for k, v in json.load(open(p)).items() if not k.startswith("_")}
    return {}


def pcg_block(lib, sb, local_rank, peak, full_solve=True):
    """Secondary measurement: one CG iteration of the matrix-free normal-equations solve on the
    50k x 1M synthetic instance (BASELINE.json configs[3]) - CUDA events, model resident."""
    import ctypes as C
    from sypha_b200.instances import gen_scp
    m, n0, dens, desc = WORKLOADS["synth50k"]
    mdl = gen_scp(m, n0, dens, 1)
    env = sb.SyphaEnvironment(cudaDeviceId=local_rank, linearSolverStrategy="pcg")
    node = sb.SyphaNodeSparse.from_csr(mdl.m, mdl.n, mdl.n_orig, mdl.offs, mdl.inds, mdl.vals, mdl.c, mdl.b, env)
    ws = sb.IpmWorkspace()
    sb.initializeIpmWorkspace(ws, device=local_rank)
    try:
        t0 = time.perf_counter()
        node.copyModelOnDevice(ws)
        load_s = time.perf_counter() - t0
        info = (C.c_longlong * 20)()
        lib.sb200_model_info(ws.handle, info, 20)
        out = {"workload": desc, "m": mdl.m, "n": mdl.n, "nnz": int(mdl.nnz), "load_model_s": load_s,
               "representation": ("pattern-only 2-byte entries in 16-byte chunks, vector blocks staged in shared memory"
                                  if info[12] else "value-carrying CSR/CSC")}
        ms = C.c_double()
        for pid, nm in ((6, "cg_iteration"), (7, "At_p"), (8, "A_q")):
            if lib.sb200_time_phase(ws.handle, pid, 20, C.byref(ms)) == 0:
                out[nm + "_us"] = 1e3 * ms.value
        if "cg_iteration_us" in out:
            nseg_r, nseg_c = info[13] * mdl.m, info[14] * mdl.n
            # bytes our representation moves per CG iteration: entry chunks + segment bounds of both copies,
            # the per-block partial sums of A q (written + read), q written + read, d, and 10 m-vector passes
            ours = (16 * (info[15] + info[16]) + 4 * (nseg_r + nseg_c) + 16 * nseg_r + 8 * 3 * mdl.n + 8 * 10 * mdl.m
                    if info[12] else 2 * 12 * int(mdl.nnz) + 8 * (3 * mdl.n + 10 * mdl.m))
            survey = 2 * 12 * int(mdl.nnz) + 8 * (3 * mdl.n + 10 * mdl.m)      # 12 B per stored entry
            t = out["cg_iteration_us"] * 1e-6
            out["roofline"] = {"bound": "hbm", "kernel": "one CG iteration (A'p, A q, 2 fused vector kernels)",
                               "algorithmic_bytes": ours, "achieved": ours / t / 1e9, "peak": peak, "unit": "GB/s",
                               "frac": ours / t / 1e9 / peak,
                               "csr12_equivalent_bytes": survey, "csr12_equivalent_gbs": survey / t / 1e9,
                               "traffic": read_traffic().get("cg_iteration")}
            out["cg_iterations_per_sec"] = 1.0 / t
        if full_solve:
            # time-to-LP-optimum of configs[3]: one whole solve (the reference cannot run this size: its
            # start point needs 840 GB of host memory and its ratio test bails at n > 262144 - SURVEY F6, F7)
            env2 = sb.SyphaEnvironment(cudaDeviceId=local_rank, linearSolverStrategy="pcg", krylovMaxCgIter=200000,
                                       krylovCgTolInitial=1e-8, krylovCgTolFinal=1e-8, krylovCgTolDecayRate=1.0)
            node.env = env2
            res = sb.SolverExecutionResult()
            t0 = time.perf_counter()
            st = sb.solver_sparse_mehrotra_run(node, sb.SolverExecutionConfig(maxIterations=MAX_ITER), res, ws)
            out["lp_solve"] = {"status": int(st), "reason": int(res.terminationReason), "iterations": int(res.iterations),
                               "primal": res.primalObj, "dual": res.dualObj, "mu": res.mu,
                               "cg_iterations": int(res.cgIterations), "cg_tol": 1e-8,
                               "time_to_lp_opt_s": (res.msStart + res.msSetup + res.msLoop) / 1e3,
                               "wall_s": time.perf_counter() - t0, "ipm_iterations_per_sec":
                               res.iterations / max((res.msStart + res.msSetup + res.msLoop) / 1e3, 1e-9),
                               "kernels_launched": int(res.kernelsLaunched)}
        return out
    finally:
        sb.releaseIpmWorkspace(ws)


def bnb_measure(args, rank, world, local_rank, dist, instance, steps, warmup, slots, stream_factor=0):
    """B&B nodes/s (BASELINE.json configs[4]) on one OR-Library instance: the reference's prelude (greedy
    incumbent, columns dearer than it dropped), then batched node LPs on every GPU with the reference's node LP
    configuration; ranks work on disjoint parts of the frontier, NCCL carries only the incumbent bound (an
    asynchronous 24-byte all_gather per round, collected two rounds later) and, when the frontier sizes drift
    apart, donated nodes.  Returns this rank-0 view of the measurement (None on the other ranks)."""
    import torch
    from sypha_b200 import bnb, bnb_exchange
    from sypha_b200.instances import load_npz

    mdl = load_npz(REPO / "tests" / "golden" / f"{instance}.npz")      # same instance on every rank
    ax = bnb_exchange.AsyncBoundExchange(lag=2) if dist is not None else None

    def rebalance(nodes):
        return bnb_exchange.rebalance_frontier(nodes, max_depth=64, min_imbalance=slots // 2)

    drv = bnb.BatchedBnb.with_reference_presolve(
        mdl, slots=slots, device=local_rank, device_heuristics=not args.host_heuristics, share_gpu=not args.no_share,
        poll_every=1, node_lp=args.node_lp, async_exchange=ax, warm_start=args.warm_start,
        pipeline=1 if stream_factor > 0 else args.bnb_pipeline,
        rebalance=rebalance if (dist is not None and not args.no_donation) else None)
    red = drv.base
    # every rank expands the same first levels (deterministic, no exchange yet), then keeps its round-robin share
    drv.async_exchange = None
    while len(drv.frontier) < world * slots and drv.frontier:
        drv.round()
    mine = bnb_exchange.partition_round_robin(list(drv.frontier), rank, world)
    drv.frontier.clear()
    drv.frontier.extend(mine)
    for _ in range(warmup):                        # untimed full windows: every slot has solved a node before t0
        drv.round()
    drv.drain()                                    # nothing in flight when the clock starts
    drv.async_exchange = ax
    warm_nodes = drv.stats.processed
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
        torch.cuda.synchronize()
    st = drv.stats
    before = (st.processed, st.lp_iterations, st.lp_device_ms, st.kernels_launched, st.delta_rows, st.rounds)
    t0 = time.perf_counter()
    # a step = one round: a window of K node LPs, or (stream_factor F) F*K nodes through the continuous batcher
    drv.run(max_nodes=10 ** 9, rounds=steps, stream_nodes=stream_factor * slots)
    drv.drain()                                    # ... and nothing when it stops: every node counted was solved inside
    torch.cuda.synchronize()
    busy = time.perf_counter() - t0               # this rank's own time for its K rounds
    if dist:
        dist.barrier()
        torch.cuda.synchronize()
    elapsed = time.perf_counter() - t0
    nodes, iters, dev_ms, launches, drows, rounds = (st.processed - before[0], st.lp_iterations - before[1],
                                                     st.lp_device_ms - before[2], st.kernels_launched - before[3],
                                                     st.delta_rows - before[4], st.rounds - before[5])
    idle_ms, wait_ms = 1e3 * (elapsed - busy), st.exchange_wait_ms
    per_rank = None
    if dist:
        drv.finish_exchange()
        g = torch.tensor([nodes, 1e3 * busy, wait_ms, len(drv.frontier)], dtype=torch.float64, device="cuda")
        allg = [torch.zeros_like(g) for _ in range(world)]
        dist.all_gather(allg, g)
        per_rank = [[round(float(v), 2) for v in t.tolist()] for t in allg]
        elapsed, (nodes, iters, dev_ms, launches, drows) = bnb_exchange.reduce_counters(elapsed, [nodes, iters, dev_ms, launches, drows])
    out = None
    if rank == 0:
        out = {
            "instance": instance, "value": nodes / elapsed, "unit": "nodes/s", "n_gpus": world, "rounds": steps,
            "warmup_rounds": warmup, "warmup_nodes": warm_nodes, "ms_per_round": 1e3 * elapsed / steps,
            "nodes": int(nodes), "lp_iterations": int(iters), "lp_iterations_per_node": iters / max(nodes, 1),
            "lp_device_ms_per_node": dev_ms / max(nodes, 1), "gpu_launches": int(launches),
            "model": {"m": int(red.m), "n_orig": int(red.n_orig), "n_orig_input": int(mdl.n_orig), "nnz": int(red.nnz),
                      "greedy_incumbent": drv.stats.greedy_incumbent},
            "incumbent": drv.incumbent, "root_bound": drv.stats.root_bound,
            "slots_per_gpu": slots, "node_lp": args.node_lp, "warm_start": bool(args.warm_start),
            "batching": (f"continuous (sb200_solve_stream), {stream_factor} x slots nodes per round"
                         if stream_factor > 0 else "windows of K nodes (sb200_solve_batch)" if drv.pipeline == 1 else
                         f"windows of K nodes, {drv.pipeline} in flight over alternating workspace sets "
                         "(sb200_window_begin / sb200_window_finish)"),
            "exchange": ({"kind": "asynchronous all_gather of (incumbent objective, open nodes, processed nodes) per "
                                  "round, collected 2 rounds later; incumbent vector fetched once at the end; node "
                                  "donation (blocking) only when the gathered frontier sizes differ by > slots/2",
                          "bytes_per_round_per_rank": 24, "posted": ax.posted,
                          "per_rank_[nodes, busy_ms, exchange_wait_ms, open_nodes]": per_rank,
                          "rank0_idle_at_final_barrier_ms": idle_ms,
                          "nodes_sent_by_rank0": st.nodes_sent, "nodes_received_by_rank0": st.nodes_received}
                         if dist else None),
            "rank0": {"round_ms": st.round_ms[-steps:], "solve_ms": st.solve_ms[-steps:], "window_device_ms": st.window_ms[-steps:],
                      "heuristics_ms": st.heur_ms[-steps:], "lp_at_iteration_cap": st.maxiter_nodes,
                      "lp_gap_stalled": st.gap_stalled_nodes, "lp_failed": st.infeasible, "pruned": st.pruned_by_bound,
                      "integral": st.integral},
            "h2d_bytes_per_round": int(20 * drows / max(steps, 1)),
            "d2h_bytes_per_round": int(slots * world * (8 * (2 * red.n + red.m) if args.host_heuristics else 40 + 96)),
        }
    drv.close()
    return out


def run_bnb(args, rank, world, local_rank):
    """--workload bnb: the B&B measurement as the headline line."""
    import torch
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.ready.wait(10)
    blk = bnb_measure(args, rank, world, local_rank, dist, args.bnb_instance, args.steps, args.warmup, args.slots,
                      args.stream_factor)
    sampler.stop_evt.set()
    sampler.join(timeout=2)
    out = None
    if rank == 0:
        out = {
            "metric": "bnb_nodes_per_sec", "value": blk["value"], "unit": "nodes/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": blk["ms_per_round"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "orlib",
            "config": {"workload": f"branch-and-bound on OR-Library {args.bnb_instance} (configs[4]); a step = one round "
                                   "of K batched node LPs per GPU", "slots_per_gpu": args.slots, "node_lp": args.node_lp},
            "timing": "wall clock between device synchronisations around the K rounds, max over ranks "
                      "(node deltas, LP solves, heuristics kernel, host loop, exchange)",
            "e2e": {"value": blk["value"], "unit": "nodes/s", "h2d_bytes_per_step": blk["h2d_bytes_per_round"],
                    "d2h_bytes_per_step": blk["d2h_bytes_per_round"],
                    "note": "the base model is resident; every round sends the K decision lists (20 B per branch row) and "
                            "reads back, per node, the LP result scalars and the 40-byte branching / incumbent record "
                            "inside the timed region: value IS end to end"},
            "gpu_launches": blk["gpu_launches"], "clocks": sampler.summary(), "bnb": blk,
        }
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return out


def algorithmic_bytes(info, phase):
    """Algorithmic bytes (or flops) of one launch group, DESIGN.md 'Kernels'."""
    m, n, nnz, mpad, n_pairs, n_terms, general = (info["m"], info["n"], info["nnz"], info["mpad"],
                                                  info["n_pairs"], info["n_terms"], info["general"])
    if phase == "assemble":      # term stream + entry pointers + M written once
        if info.get("pat_chunks"):                       # compact form: 16-byte chunks of 2-byte column ids
            return info["pat_chunks"] * 16 + n_pairs * 4 + n_pairs * 8
        return n_terms * (12 if general else 4) + n_pairs * 4 + n_pairs * 8
    if phase == "potrf":         # M read once + L written once (lower triangles)
        return 2 * 8 * m * (m + 1) // 2
    if phase == "potrs":         # L read twice (forward + backward)
        return 2 * 8 * m * (m + 1) // 2
    if phase == "spmv_csr":
        return 12 * nnz + 8 * (n + 2 * m)
    if phase == "spmv_csc_recover":
        return 12 * nnz + 8 * (m + 6 * n)
    if phase == "vector":
        return 8 * 4 * n
    raise KeyError(phase)


def run_ours(args, rank, world, local_rank):
    import ctypes as C
    import torch
    import sypha_b200 as sb
    from sypha_b200 import _lib as L

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = L.load()
    strategy = args.strategy
    global N_INSTANCES
    models, desc, data_kind = load_models(args.workload, rank)
    N_INSTANCES = len(models)
    fp64_peak, fp64_src = measure_fp64_dgemm() if rank == 0 else (FP64_FALLBACK_TFLOPS, "")

    env = sb.SyphaEnvironment(cudaDeviceId=local_rank, linearSolverStrategy=strategy,
                              pollEvery=args.poll_every, useGraph=not args.no_graph,
                              krylovMaxCgIter=args.cg_max_iter, krylovCgTolInitial=args.cg_tol,
                              krylovCgTolFinal=args.cg_tol, krylovCgTolDecayRate=1.0)
    cfg = sb.SolverExecutionConfig(maxIterations=MAX_ITER)
    wss, nodes = [], []
    for mdl in models:
        ws = sb.IpmWorkspace()
        sb.initializeIpmWorkspace(ws, device=local_rank)
        node = sb.SyphaNodeSparse.from_csr(mdl.m, mdl.n, mdl.n_orig, mdl.offs, mdl.inds, mdl.vals, mdl.c, mdl.b, env)
        node.copyModelOnDevice(ws)
        wss.append(ws)
        nodes.append(node)

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    def step(i):
        res = sb.SolverExecutionResult()
        st = sb.solver_sparse_mehrotra_run(nodes[i % N_INSTANCES], cfg, res, wss[i % N_INSTANCES])
        if st != sb.CODE_SUCCESSFUL:
            raise RuntimeError(f"LP {i} failed: reason {res.terminationReason}")
        return res

    # one LP at a time (latency form): the headline with --headline single, the `single_lp` block otherwise
    s_steps, s_warm = (1, 0) if args.quick_single else (args.steps, args.warmup)
    for i in range(N_INSTANCES):                   # set-up: every instance's iteration graph is built once
        step(i)
    for i in range(s_warm):
        step(i)
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.ready.wait(10)
    barrier()
    t0 = time.perf_counter()
    results = [step(i) for i in range(s_steps)]
    barrier()
    elapsed = time.perf_counter() - t0
    wall = elapsed
    iters = sum(r.iterations for r in results)
    iters_rank0 = iters
    launches = sum(r.kernelsLaunched for r in results)
    # timed on the device: CUDA events on the workspace stream around every LP (start point, initial
    # residuals, loop); LPs run back to back, so the sum is this rank's device time for the K steps
    dev_ms = sum(r.msStart + r.msSetup + r.msLoop for r in results)
    loop_ms = sum(r.msLoop for r in results)
    elapsed = dev_ms / 1e3

    # ---- e2e: host buffers in, host results out, every step ---------------------------------
    # one persistent workspace per instance, as a caller that re-solves a set of models keeps them (the reference's B&B
    # keeps one IpmWorkspace per base model): every step still uploads the whole model from pinned host memory; the
    # library recognises "the same matrix as the resident one" by fingerprint and keeps its CSC / symbolic structure
    e2e_wss = []
    for _ in range(N_INSTANCES):
        w = sb.IpmWorkspace()
        sb.initializeIpmWorkspace(w, device=local_rank)
        e2e_wss.append(w)
    pinned = []
    for mdl in models:
        arrs = {}
        for k in ("offs", "inds", "vals", "c", "b"):
            t = torch.from_numpy(getattr(mdl, k)).pin_memory()
            arrs[k] = t
        pinned.append(arrs)

    def e2e_step(i):
        mdl, a = models[i % N_INSTANCES], pinned[i % N_INSTANCES]
        node = sb.SyphaNodeSparse(env)
        node.nrows, node.ncols, node.ncolsOriginal, node.nnz = mdl.m, mdl.n, mdl.n_orig, mdl.nnz
        node.hCsrMatOffs, node.hCsrMatInds, node.hCsrMatVals = a["offs"].numpy(), a["inds"].numpy(), a["vals"].numpy()
        node.hObjDns, node.hRhsDns = a["c"].numpy(), a["b"].numpy()
        res = sb.SolverExecutionResult()
        sb.solver_sparse_mehrotra_run(node, cfg, res, e2e_wss[i % N_INSTANCES])      # uploads the model, solves, D2H
        return res

    e2e_steps = 0 if args.no_e2e else max(2, min(s_steps, 5))
    for i in range(max(3, N_INSTANCES) if e2e_steps else 0):   # warm-up: the pools, the driver's allocation paths and the
                                                               # grow-only buffers (every instance seen once: scpclr13's
                                                               # 400 MB structure otherwise grows inside timed steps 4-5)
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    e2e_res, e2e_each = [], []
    for i in range(e2e_steps):
        t1 = time.perf_counter()
        e2e_res.append(e2e_step(i))
        e2e_each.append(round(1e3 * (time.perf_counter() - t1), 2))
    barrier()
    e2e_elapsed = time.perf_counter() - t0
    print(f"[bench] e2e step times (ms): {e2e_each}", file=sys.stderr)
    e2e_iters = sum(r.iterations for r in e2e_res)
    mdl0 = models[0]
    h2d = int(mdl0.offs.nbytes + mdl0.inds.nbytes + mdl0.vals.nbytes + mdl0.c.nbytes + mdl0.b.nbytes)
    d2h = int(8 * (2 * mdl0.n + mdl0.m))
    sampler.stop_evt.set()
    sampler.join(timeout=2)

    # ---- THROUGHPUT form (the headline when the shape fits one thread block, m <= 2048): a step = ONE WINDOW of K LPs of
    #      the workload (the instances in rotation), each solved whole by one thread block, the window ONE launch
    #      (csrc/sb200_cta.cu; sb200_solve_batch over K workspaces).  `value`: CUDA events on the launching stream around
    #      the window kernel (sb200_last_window), models resident.  `e2e`: every step uploads all K models from pinned
    #      host memory (sb200_load_model), solves the window and reads every LP's x, y, s back into pinned memory ----
    tp = None
    if args.headline != "single" and not args.no_batch_block and models[0].m <= 2048 and strategy in ("auto", "cholesky"):
        from sypha_b200.solver import set_solver_form, solve_batch, last_window
        K = args.tp_slots
        tp_ws, tp_nodes, tp_bufs = [], [], []
        for i in range(K):
            w = sb.IpmWorkspace()
            sb.initializeIpmWorkspace(w, device=local_rank)
            mdl = models[i % N_INSTANCES]
            nd = sb.SyphaNodeSparse.from_csr(mdl.m, mdl.n, mdl.n_orig, mdl.offs, mdl.inds, mdl.vals, mdl.c, mdl.b, env)
            nd.copyModelOnDevice(w)
            set_solver_form(w, "throughput")
            tp_ws.append(w)
            tp_nodes.append(nd)
            hb = torch.empty(2 * mdl.n + mdl.m, dtype=torch.float64).pin_memory().numpy()
            tp_bufs.append((hb[:mdl.n], hb[mdl.n:mdl.n + mdl.m], hb[mdl.n + mdl.m:]))

        def window():
            rs = solve_batch(tp_nodes, cfg, tp_ws, host_bufs=tp_bufs)
            ms, k = last_window(tp_ws[0])
            if k != K:
                raise RuntimeError(f"the window was not one launch of {K} thread blocks (got {k})")
            bad = [r.terminationReason for r in rs if r.terminationReason != 0]
            if bad:
                raise RuntimeError(f"{len(bad)} LPs of the window did not converge: {bad[:4]}")
            return rs, ms

        try:
            window()
        except RuntimeError as e:
            if args.headline == "throughput":
                raise
            print(f"[bench] throughput form not available for this workload ({e}); headline = one LP at a time", file=sys.stderr)
            K = 0
        for _ in range(max(args.warmup, 3) - 1 if K else 0):
            window()
        sampler2 = ClockSampler(local_rank)
        sampler2.start()
        sampler2.ready.wait(10)
        barrier()
        t0 = time.perf_counter()
        tp_ms, tp_iters, tp_launch, tp_flops, tp_each = 0.0, 0, 0, 0.0, []
        for _ in range(args.steps if K else 0):
            rs, ms = window()
            tp_ms += ms
            tp_each.append(round(ms, 2))
            tp_iters += sum(r.iterations for r in rs)
            tp_launch += sum(r.kernelsLaunched for r in rs)
            # factorisations per LP: one per iteration + the starting point's (M = A A')
            tp_flops += sum((r.iterations + 1) * (tp_nodes[j].nrows ** 3) / 3.0 for j, r in enumerate(rs))
        barrier()
        tp_wall = time.perf_counter() - t0
        ms = C.c_double()
        ph = []
        for q in range(8):
            lib.sb200_time_phase(tp_ws[K // 2].handle, 100 + q, 1, C.byref(ms))
            ph.append(ms.value)
        it_mid = rs[K // 2].iterations if K else 1

        # e2e of the same step
        def tp_e2e_step():
            for j in range(K):
                mdl, a = models[j % N_INSTANCES], pinned[j % N_INSTANCES]
                node = sb.SyphaNodeSparse(env)
                node.nrows, node.ncols, node.ncolsOriginal, node.nnz = mdl.m, mdl.n, mdl.n_orig, mdl.nnz
                node.hCsrMatOffs, node.hCsrMatInds, node.hCsrMatVals = a["offs"].numpy(), a["inds"].numpy(), a["vals"].numpy()
                node.hObjDns, node.hRhsDns = a["c"].numpy(), a["b"].numpy()
                node.copyModelOnDevice(tp_ws[j])                   # H2D of the whole model, every LP, every step
                tp_nodes[j] = node
            return window()[0]

        tp_e2e = None
        if e2e_steps and K:
            tp_e2e_step()
            barrier()
            t0 = time.perf_counter()
            n_e2e, it_e2e = max(2, min(args.steps, 3)), 0
            for _ in range(n_e2e):
                it_e2e += sum(r.iterations for r in tp_e2e_step())
            barrier()
            el = time.perf_counter() - t0
            tp_e2e = {"iters": it_e2e, "elapsed": el, "steps": n_e2e,
                      "h2d": int(sum(sum(pinned[j % N_INSTANCES][k].numel() * pinned[j % N_INSTANCES][k].element_size()
                                         for k in ("offs", "inds", "vals", "c", "b")) for j in range(K))),
                      "d2h": int(sum(8 * (2 * models[j % N_INSTANCES].n + models[j % N_INSTANCES].m) for j in range(K)))}
        sampler2.stop_evt.set()
        sampler2.join(timeout=2)
        mm = models[0].m
        fl = mm ** 3 / 3.0
        per_it = {nm: 1e3 * v / max(it_mid, 1) for nm, v in zip(("assembly", "factorisation", "solves", "A_v", "At_v", "vector"), ph[:6])}
        if K:
            tp = dict(K=K, ms=tp_ms, iters=tp_iters, launches=tp_launch, flops=tp_flops, each=tp_each, wall=tp_wall,
                      per_it=per_it, whole_lp_ms=ph[7], e2e=tp_e2e, clocks=sampler2.summary(), fl=fl)
        for w in tp_ws:
            sb.releaseIpmWorkspace(w)

    # ---- secondary: B&B nodes/s on scpnre1 and scpnrg1 at THIS N (configs[4]; every rank takes part) ----
    bnb_block = None
    if not args.no_bnb_block and args.workload in ORLIB:
        bnb_block = {}
        for inst_name in ("scpnre1", "scpnrg1"):
            try:
                bnb_block[inst_name] = bnb_measure(args, rank, world, local_rank, dist, inst_name, args.bnb_rounds, 3,
                                                   args.slots, 0)
            except Exception as e:                  # never lose the headline line to a secondary measurement
                if world > 1:
                    raise
                bnb_block[inst_name] = {"error": repr(e)}

    if bnb_block and rank == 0:
        # B&B against the roofline of its dominant kernel (SURVEY 8d iii): the FP64 tensor flops of the node LPs' Cholesky
        # factorisations (m^3/3 each, one per iteration + the starting point's; m = base rows, the branch rows are not
        # counted) over the WALL time of the search rounds - host work, node rules and exchange included - per GPU
        for v in bnb_block.values():
            if "error" in v:
                continue
            fl = (v["lp_iterations"] + v["nodes"]) * (v["model"]["m"] ** 3) / 3.0
            tf = fl / (v["ms_per_round"] * v["rounds"] / 1e3) / world / 1e12
            v["roofline"] = {"bound": "tensor", "kernel": "k_ipm_cta (node LPs)", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s",
                             "frac": tf / fp64_peak, "per": "GPU, wall time of the rounds"}

    # ---- max over ranks / sums ----------------------------------------------------------------
    if dist:
        t = torch.tensor([elapsed, e2e_elapsed, wall], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed, e2e_elapsed, wall = float(t[0]), float(t[1]), float(t[2])
        c = torch.tensor([iters, launches, e2e_iters], device="cuda", dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        iters, launches, e2e_iters = int(c[0]), int(c[1]), int(c[2])
        if tp:
            t = torch.tensor([tp["ms"], tp["wall"], tp["e2e"]["elapsed"] if tp["e2e"] else 0.0], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tp["ms_max"], tp["wall_max"], e2e_max = float(t[0]), float(t[1]), float(t[2])
            c = torch.tensor([tp["iters"], tp["launches"], tp["flops"], tp["e2e"]["iters"] if tp["e2e"] else 0,
                              tp["e2e"]["h2d"] if tp["e2e"] else 0, tp["e2e"]["d2h"] if tp["e2e"] else 0],
                             device="cuda", dtype=torch.float64)
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
            tp["iters_all"], tp["launches_all"], tp["flops_all"] = int(c[0]), int(c[1]), float(c[2])
            if tp["e2e"]:
                tp["e2e"].update(elapsed_max=e2e_max, iters_all=int(c[3]), h2d_all=int(c[4]), d2h_all=int(c[5]))
    elif tp:
        tp.update(ms_max=tp["ms"], wall_max=tp["wall"], iters_all=tp["iters"], launches_all=tp["launches"], flops_all=tp["flops"])
        if tp["e2e"]:
            tp["e2e"].update(elapsed_max=tp["e2e"]["elapsed"], iters_all=tp["e2e"]["iters"], h2d_all=tp["e2e"]["h2d"],
                             d2h_all=tp["e2e"]["d2h"])

    out = None
    if rank == 0:
        # ---- per-phase device timing (CUDA events on the workspace stream) and the roofline ---
        info_arr = (C.c_longlong * 20)()
        lib.sb200_model_info(wss[0].handle, info_arr, 20)
        info = dict(m=info_arr[0], n=info_arr[1], nnz=info_arr[3], mpad=info_arr[4], strategy=info_arr[5],
                    n_pairs=info_arr[6], n_terms=info_arr[7], general=info_arr[8], pat_chunks=info_arr[17])
        peak, peak_src = read_peaks()
        phases = {}
        names = {0: "assemble", 1: "potrf", 2: "potrs", 3: "spmv_csr", 4: "spmv_csc_recover", 5: "vector"}
        per_iter = {"assemble": 1, "potrf": 1, "potrs": 2, "spmv_csr": 2, "spmv_csc_recover": 2, "vector": 2}
        for pid, nm in ({} if args.no_phases else names).items():
            ms = C.c_double()
            if lib.sb200_time_phase(wss[0].handle, pid, 20, C.byref(ms)) == 0:
                by = algorithmic_bytes(info, nm)
                phases[nm] = {"ms": ms.value, "algorithmic_bytes": by, "gbs": by / ms.value / 1e6,
                              "frac_hbm": by / ms.value / 1e6 / peak, "per_iteration": per_iter[nm]}
                if nm == "potrf" or (nm == "assemble" and info["strategy"] == 2):
                    # FP64 tensor-pipe kernels: m^3/3 (Cholesky), m^2 n (SYRK on the lower tiles)
                    fl = info["m"] ** 3 / 3.0 if nm == "potrf" else float(info["m"]) ** 2 * info["n"]
                    phases[nm].update({"flops": fl, "tflops": fl / ms.value / 1e9,
                                       "frac_fp64_tensor": fl / ms.value / 1e9 / fp64_peak})
        roof = None
        if phases:
            dom = max(phases, key=lambda k: phases[k]["ms"] * phases[k]["per_iteration"])
            ph = phases[dom]
            roof = {"kernel": dom, "bound": "hbm", "achieved": ph["gbs"], "peak": peak, "unit": "GB/s",
                    "frac": ph["frac_hbm"], "traffic": None, "peak_source": peak_src,
                    "ms_per_launch_group": ph["ms"], "algorithmic_bytes": ph["algorithmic_bytes"]}
            if dom in ("potrf",):
                # the factorisation runs on the FP64 tensor pipe (DMMA): m^3/3 flops per launch
                flops = info["m"] ** 3 / 3.0
                tf = flops / ph["ms"] / 1e9
                roof.update({"bound": "tensor", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s",
                             "frac": tf / fp64_peak, "flops": flops,
                             "peak_source": fp64_src + "; MEASURED_PEAKS.json has no FP64 entry; scripts/fp64_peak.cu on "
                                            "this pool: DMMA issue rate 37.2, DGEMM 8192^3 35.9 TFLOP/s "
                                            "(profiles/r2_a_fp64_peaks.json); nominal 40",
                             "hbm_gbs": ph["gbs"], "hbm_frac": ph["frac_hbm"],
                             "note": "latency-bound at m = 1000: 16 dependent 64-column steps (pivot chain + one "
                                     "hand-off each); see profiles/ for the per-step timeline. For m <= 1280 the same launch "
                                     "also forms Z = L^-1 for the solves (another m^3/3 flops, NOT counted in `flops`)"})
            tr = read_traffic().get(dom)
            if tr is not None:
                roof["traffic"] = tr
            # the whole iteration against SURVEY.md 8(d)'s algorithmic bytes of the direct path:
            # 5 * 12 * nnz + 2 * 8 * m^2 + 4 * 8 * m^2 / 2 + 24 * 8 * n, over the measured loop time per iteration
            mm, nn, zz = info["m"], info["n"], info["nnz"]
            it_bytes = 5 * 12 * zz + 16 * mm * mm + 16 * mm * mm + 24 * 8 * nn
            it_us = 1e3 * loop_ms / max(iters_rank0, 1)
            roof["whole_iteration"] = {"algorithmic_bytes": it_bytes, "us_per_iteration": it_us,
                                       "achieved_gbs": it_bytes / it_us / 1e3, "frac_hbm": it_bytes / it_us / 1e3 / peak,
                                       "us_at_hbm_peak": it_bytes / peak / 1e3,
                                       "note": "SURVEY.md 8(d) per-iteration bytes / measured loop time per iteration "
                                               "(CUDA events); launches and dependent latency, not bandwidth, bound this shape"}
        out = {
            "metric": "ipm_iterations_per_sec", "value": iters / elapsed, "unit": "iter/s", "n_gpus": world,
            "steps": s_steps, "warmup": s_warm, "ms_per_step": 1e3 * elapsed / s_steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": data_kind,
            "config": dict(base_config(args.workload, models, desc),
                           strategy={1: "cholesky", 2: "syrk", 3: "pcg"}.get(info["strategy"], "?"),
                           l2=(f"{N_INSTANCES} rotating instances, ~75 MB of resident structure each: working set > 126 MB L2 "
                               "(no flush needed)" if N_INSTANCES > 1 else
                               "one instance; its resident structure exceeds the 126 MB L2" if models[0].nnz > 10 ** 7 else
                               "one instance, L2-resident between steps (stated, not flushed)"),
                           poll_every=args.poll_every, graph=not args.no_graph),
            "time_to_lp_opt_ms": 1e3 * elapsed / s_steps,
            "timing": "CUDA events on the workspace stream around every LP, summed over the K steps, max over ranks",
            "wall_ms_per_step": 1e3 * wall / s_steps,
            "iterations_per_lp": iters / (s_steps * world),
            "device_ms_per_lp": dev_ms / s_steps, "loop_ms_per_lp": loop_ms / s_steps,
            "e2e": ({"value": e2e_iters / e2e_elapsed, "unit": "iter/s", "h2d_bytes_per_step": h2d,
                     "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_elapsed / e2e_steps, "steps": e2e_steps}
                    if e2e_steps else None),
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "roofline": roof,
            "phases": phases,
        }
        if tp:
            # the throughput form is the headline; the one-LP-at-a-time (latency form) line above moves into `single_lp`
            single = {k: out[k] for k in ("value", "unit", "steps", "warmup", "ms_per_step", "time_to_lp_opt_ms", "timing",
                                          "wall_ms_per_step", "iterations_per_lp", "device_ms_per_lp", "loop_ms_per_lp",
                                          "e2e", "gpu_launches", "clocks", "roofline", "phases")}
            single["form"] = ("one LP at a time over the whole GPU, ~12 kernels per iteration as one CUDA graph (latency form, "
                              "sb200_set_solver_form LATENCY): the time-to-optimal-LP-relaxation number")
            K = tp["K"]
            win_ms = tp["ms_max"] / args.steps
            tfl = tp["flops_all"] / world / args.steps / win_ms / 1e9          # per launch (one GPU's window)
            ktr = read_traffic().get("k_ipm_cta_per_lp_iteration")
            f_us = tp["per_it"]["factorisation"]
            roof_tp = {"kernel": "k_ipm_cta (one launch per window, one thread block per LP: assembly, left-looking DMMA Cholesky, "
                                 "solves, products, vector steps of every iteration)",
                       "bound": "tensor", "achieved": tfl, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tfl / fp64_peak,
                       "flops": tp["flops_all"] / world / args.steps,
                       "flops_note": "m^3/3 per factorisation, one per iteration + the starting point's, summed over the window's LPs; "
                                     "ALL of the kernel's time is in the denominator (the factorisation is ~40% of it)",
                       "ms_per_launch": win_ms,
                       "traffic": (ktr * tp["iters_all"] / world / args.steps) if ktr else None,
                       "traffic_note": "dram bytes per LP iteration of a K-block window (ncu --set full, profiles/traffic.json) x this "
                                       "window's iterations" if ktr else "no ncu capture of a full window committed",
                       "peak_source": fp64_src + "; MEASURED_PEAKS.json has no FP64 entry; scripts/fp64_peak.cu on this pool: DMMA issue "
                                      "rate 37.2, DGEMM 8192^3 35.9 TFLOP/s (profiles/r2_a_fp64_peaks.json); nominal 40",
                       "factorisation_phase_per_sm": {"achieved_gflops": tp["fl"] / f_us / 1e3 if f_us else None,
                                                      "peak_gflops": 1e3 * fp64_peak / 148.0,
                                                      "frac": (tp["fl"] / f_us / 1e3) / (1e3 * fp64_peak / 148.0) if f_us else None,
                                                      "note": "in-kernel %globaltimer of one block of the window (sb200_time_phase 101)"},
                       "us_per_iteration_inside_one_block": tp["per_it"], "whole_lp_ms_inside_one_block": tp["whole_lp_ms"]}
            e2 = tp["e2e"]
            out.update({
                "value": tp["iters_all"] / (tp["ms_max"] / 1e3), "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": win_ms,
                "timing": "CUDA events on the launching stream around each window's kernel (sb200_last_window), summed over the K "
                          "steps, max over ranks",
                "wall_ms_per_step": 1e3 * tp["wall_max"] / args.steps,
                "iterations_per_lp": tp["iters_all"] / (args.steps * world * K),
                "lps_per_sec": args.steps * world * K / (tp["ms_max"] / 1e3),
                "window_ms": tp["each"],
                "e2e": ({"value": e2["iters_all"] / e2["elapsed_max"], "unit": "iter/s",
                         "h2d_bytes_per_step": e2["h2d_all"], "d2h_bytes_per_step": e2["d2h_all"],
                         "ms_per_step": 1e3 * e2["elapsed_max"] / e2["steps"], "steps": e2["steps"],
                         "note": "every step: all K models host->device from pinned memory (sb200_load_model), the window, every "
                                 "LP's x, y, s device->host into pinned memory; host clock between barriers"} if e2 else None),
                "gpu_launches": tp["launches_all"], "clocks": tp["clocks"], "roofline": roof_tp,
            })
            out["config"] = dict(out["config"], lps_per_step=K, form="throughput: one thread block per LP, one launch per step",
                                 l2=f"{K} LPs x ~75 MB of resident structure per GPU: working set >> 126 MB L2 (no flush needed)")
            out["time_to_lp_opt_ms"] = single["time_to_lp_opt_ms"]
            for k in ("device_ms_per_lp", "loop_ms_per_lp", "phases"):
                out.pop(k, None)
            out["single_lp"] = single
        if bnb_block:
            out["bnb"] = bnb_block
        if world == 1 and not args.no_pcg_block and args.workload != "synth50k":
            try:
                out["pcg_50kx1M"] = pcg_block(lib, sb, local_rank, peak, full_solve=not args.no_pcg_solve)
            except Exception as e:           # never lose the headline line to the secondary measurement
                out["pcg_50kx1M"] = {"error": repr(e)}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(models[0], budget_s=args.cpu_budget)
    for ws in wss + e2e_wss:
        sb.releaseIpmWorkspace(ws)
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return out


def to_oracle_instance(mdl):
    from oracle import scp_io
    return scp_io.ScpInstance(mdl.m, mdl.n, mdl.n_orig, mdl.offs, mdl.inds, mdl.vals, mdl.c, mdl.b, mdl.name)


def cpu_baseline(mdl, budget_s=25.0, solver="ne"):
    """The oracle (CPU port of the reference's solver) on a bounded sample of the same workload."""
    from oracle import mehrotra as mo
    inst = to_oracle_instance(mdl)
    t0 = time.perf_counter()
    probe = mo.solve_instance(inst, mo.Params(max_iter=MAX_ITER), solver, stop_after=2)
    t_probe = time.perf_counter() - t0
    per_iter = max(probe.loop_seconds / max(probe.iterations, 1), 1e-6)
    cap = int(max(2, min(MAX_ITER, (budget_s - probe.start_seconds) / per_iter)))
    r = mo.solve_instance(inst, mo.Params(max_iter=MAX_ITER), solver, stop_after=cap)
    full = r.reason == mo.TERM_CONVERGED
    secs = r.start_seconds + r.loop_seconds
    return {"value": r.iterations / secs, "unit": "iter/s", "cores": os.cpu_count(), "kind": "port",
            "sample": (f"one {'full' if full else 'truncated'} LP solve of instance 0 ({r.iterations} iterations, "
                       f"start point {r.start_seconds:.2f} s + loop {r.loop_seconds:.2f} s), NumPy/SciPy oracle "
                       f"(normal equations + LAPACK Cholesky), BLAS threads = all cores"),
            "seconds": secs, "probe_seconds": t_probe}


def run_reference_port(args, models, desc, data_kind, world):
    """CPU arm of last resort: the oracle restatement of the reference's solver (oracle/, NumPy + SciPy)."""
    budget = max(2.0, 170.0 / (args.steps + args.warmup))
    samples = []
    for i in range(args.warmup):
        cpu_baseline(models[i % len(models)], budget)
    t0 = time.perf_counter()
    for i in range(args.steps):
        samples.append(cpu_baseline(models[i % len(models)], budget))
    elapsed = time.perf_counter() - t0
    iters_s = sum(s["value"] * s["seconds"] for s in samples) / sum(s["seconds"] for s in samples)
    return {
        "impl": "reference", "metric": "ipm_iterations_per_sec", "value": iters_s, "unit": "iter/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": data_kind,
        "config": base_config(args.workload, models, desc),
        "cpu_baseline": {"value": iters_s, "unit": "iter/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": samples[0]["sample"] + f"; {args.steps} such steps, {budget:.0f} s budget each"},
        "e2e": {"value": iters_s, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU restatement (oracle/) of the reference's solver: oracle/_ref/sypha_ref (the reference's own CUDA "
                "build, `make -C oracle`) is not present in this checkout",
    }


def run_reference(args, rank, world):
    """The reference arm, rank 0 only: the reference's OWN solver, unmodified - its CLI (src/main.cpp) over its
    CUDA implementation (dense LU of the KKT matrix with cuSOLVER, cuSPARSE, cuBLAS), built for sm_100a by
    oracle/Makefile into oracle/_ref/sypha_ref - on the same instances, one process per step, on this box's GPU 0.
    A step is a bounded sample: the first `--ref-iters` iterations of the LP (every iteration of the reference does
    the same work: one LU of the (2n+m)^2 KKT matrix and two solves).  value = iterations / the reference's own loop
    timer (`solver` in its "Time (s)" line = node.timeSolver*, src/sypha_solver.cpp:487,821 - the same span our
    ms_loop covers), summed over the K steps."""
    if rank != 0:
        return None
    import re
    import tempfile
    models, desc, data_kind = load_models(args.workload, 0)
    binary = REPO / "oracle" / "_ref" / "sypha_ref"
    if not binary.exists() or models[0].m * 3 + 2 * models[0].n_orig > 110000:     # dense KKT must fit
        return run_reference_port(args, models, desc, data_kind, world)
    from sypha_b200.instances import write_scp
    tmp = Path(tempfile.mkdtemp(prefix="sb200_ref_"))
    files = []
    for i, mdl in enumerate(models):
        files.append(tmp / f"inst{i}.txt")
        write_scp(mdl, files[-1])

    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)      # torchrun pins it to 1: the host-side start point (the GSL
                                                           # stand-in, outside the metric) would take 18 s per step
    def one(i):
        t0 = time.perf_counter()
        r = subprocess.run([str(binary), "--model", "scp", "--input-file", str(files[i % len(files)]),
                            "--mehrotra-max-iter", str(args.ref_iters), "--disable-bnb", "--verbosity", "5"],
                           capture_output=True, text=True, timeout=900, env=env)
        out = r.stdout + r.stderr
        if r.returncode != 0:
            raise RuntimeError(f"reference binary failed ({r.returncode}): {out[-800:]}")
        it = int(re.search(r"Iterations:\s+(\d+)", out).group(1))
        tm = re.search(r"start\s+([0-9.]+)\s+pre\s+([0-9.]+)\s+solver\s+([0-9.]+)\s+total\s+([0-9.]+)", out)
        return it, float(tm.group(3)), float(tm.group(1)), float(tm.group(2)), time.perf_counter() - t0

    for i in range(args.warmup):
        one(i)
    t0 = time.perf_counter()
    runs = [one(i) for i in range(args.steps)]
    elapsed = time.perf_counter() - t0
    iters, loop_s = sum(r[0] for r in runs), sum(r[1] for r in runs)
    val = iters / max(loop_s, 1e-9)
    sample = (f"first {args.ref_iters} iterations of each LP, one process per step ({args.steps} steps, instances in "
              f"rotation); loop timer {loop_s:.2f} s for {iters} iterations; per step also start point "
              f"{np.mean([r[2] for r in runs]):.2f} s (host, the GSL stand-in with OpenMP) + set-up "
              f"{np.mean([r[3] for r in runs]):.2f} s + process start, {np.mean([r[4] for r in runs]):.1f} s wall")
    return {
        "impl": "reference", "metric": "ipm_iterations_per_sec", "value": val, "unit": "iter/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": data_kind,
        "config": base_config(args.workload, models, desc),
        "cpu_baseline": {"value": val, "unit": "iter/s", "cores": 1, "kind": "reference", "sample": sample,
                         "what": "the reference's own CUDA solver (unmodified sources, oracle/Makefile, release flags) "
                                 "driven through its CLI by one host thread, GPU 0 of this box; not a CPU run"},
        "e2e": {"value": val, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "gpu_launches counts OUR kernels: none run in this arm (cuSOLVER/cuSPARSE/cuBLAS + the reference's own)",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="scpnrh", choices=sorted(ORLIB) + sorted(WORKLOADS) + ["bnb"])
    ap.add_argument("--ref-iters", type=int, default=6,
                    help="--impl reference: iterations of each LP the reference's CUDA solver is sampled over")
    ap.add_argument("--slots", type=int, default=148, help="bnb: node LPs per window and GPU (one thread block each, one launch per window; 148 = the SMs of a B200)")
    ap.add_argument("--tp-slots", type=int, default=148, help="LPs in flight in the throughput block of the default line")
    ap.add_argument("--bnb-instance", default="scpnre1", help="bnb: OR-Library instance (tests/golden/<name>.npz)")
    ap.add_argument("--node-lp", default="reference", choices=["reference", "converged"],
                    help="bnb: node LP configuration - the reference's (gap-stagnation exit, window 5, 1 %%) or to mu <= 1e-4")
    ap.add_argument("--warm-start", action="store_true", help="bnb: children start from their parent's iterate (sb200_node_delta.warm_start)")
    ap.add_argument("--bnb-pipeline", type=int, default=2,
                    help="bnb: windows in flight per GPU (2: the next window is queued while the host processes the previous one)")
    ap.add_argument("--no-bnb-block", action="store_true", help="skip the B&B block of the default line")
    ap.add_argument("--bnb-rounds", type=int, default=8, help="rounds per instance in the B&B block of the default line")
    ap.add_argument("--host-heuristics", action="store_true",
                    help="bnb: branching rule and rounding/repair heuristic on the host (NumPy) instead of the device kernel")
    ap.add_argument("--stream-factor", type=int, default=0,
                    help="bnb: > 0 = continuous batching (sb200_solve_stream), a step starts F x slots nodes; 0 = windows of K nodes")
    ap.add_argument("--no-share", action="store_true", help="bnb: keep the single-LP launch geometry (one CTA per task) in every slot")
    ap.add_argument("--no-donation", action="store_true", help="bnb, N > 1: keep the initial round-robin split (no node donation)")
    ap.add_argument("--strategy", default="auto")
    ap.add_argument("--poll-every", type=int, default=None,
                    help="iterations enqueued per read-back of the device scalar block (default: 2 for the LP workloads - "
                         "measured 2955 vs 2909 iter/s at 1, 2929 at 4 - and 1 for bnb, where extra launches cost throughput)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-phases", action="store_true")
    ap.add_argument("--no-pcg-block", action="store_true")
    ap.add_argument("--headline", default="auto", choices=["auto", "throughput", "single"],
                    help="throughput (default where m <= 2048): a step = one window of --tp-slots LPs, one thread block each, one "
                         "launch; single: a step = one LP over the whole GPU (latency form)")
    ap.add_argument("--no-batch-block", action="store_true", help="same as --headline single")
    ap.add_argument("--quick-single", action="store_true", help="profiling runs: the one-LP-at-a-time block does a single step")
    ap.add_argument("--no-pcg-solve", action="store_true", help="skip the whole 50k x 1M LP solve (about 15 s)")
    ap.add_argument("--cpu-budget", type=float, default=25.0)
    ap.add_argument("--cg-max-iter", type=int, default=50000)
    ap.add_argument("--cg-tol", type=float, default=1e-8)
    args = ap.parse_args()
    if args.poll_every is None:
        args.poll_every = 1 if args.workload == "bnb" else 2
    args.warmup = max(args.warmup, 0)
    # stdout carries exactly ONE JSON line: libraries (NCCL banner, ...) are diverted to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        out = run_reference(args, rank, world)
    else:
        if args.warmup < 3:
            args.warmup = 3
        if args.workload == "bnb":
            out = run_bnb(args, rank, world, local_rank)
        else:
            out = run_ours(args, rank, world, local_rank)
    sys.stdout.flush()
    if out is not None:
        os.write(real_stdout, (json.dumps(out) + "\n").encode())


if __name__ == "__main__":
    main()
