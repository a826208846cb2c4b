#!/usr/bin/env python
"""bench.py — headline measurement of the B200 batched CLDDP engine (contract in the task brief).

Metric (BASELINE.json): "DDP iterations/sec, quadrotor n=13 m=4 N=100, batch sweep at 1/2/4/8 GPU".
Unit of work: one INSTANCE-ITERATION = one backward Riccati sweep + one forward line search for one
problem instance.  One bench "step" = one batched DDP iteration = the three kernels
(linearise, backward sweep, forward rollout/line search) over the rank's whole batch.

  value        = instance-iterations/s, whole job, inputs resident in HBM when the timed region starts
  e2e.value    = same metric through the C-ABI the way a user calls it: pinned HOST buffers in,
                 `iters_per_call` DDP iterations, HOST buffers out — H2D + D2H inside the timed region
  roofline     = backward-sweep kernel: algorithmic HBM bytes 8(n^2+2nm+n+3m)*N*B per launch divided
                 by its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline = the CPU oracle (restatement of the reference algorithm; the reference itself cannot
                 be built here: Eigen/autodiff unavailable) on a bounded sample, all host threads

Convergence exits are disabled in the throughput legs (tolerance = acceptable_tolerance = 0) so every
instance performs every iteration; converge-to-tolerance parity is covered by tests/ (pytest -m gpu).

`--impl reference` times the CPU oracle on the same workload/metric (rank 0 only).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIG5_CLDDP, CONFIG5_IPDDP = "manip7_user", "manip7_user_ipddp"  # BASELINE config #5 workloads (user-model plugin)
METRIC = "DDP iterations/sec (instance-iterations: backward sweep + forward line search per problem instance)"
UNIT = "instance-iterations/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="quadrotor")
    ap.add_argument("--batch", type=int, default=0, help="instances per GPU (default: the config's batch)")
    ap.add_argument("--iters-per-call", type=int, default=10, help="DDP iterations per e2e solve call")
    ap.add_argument("--e2e-calls", type=int, default=24)
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip the secondary workloads (configs 1, 2, 4, 5)")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.samples:
            if t0 is not None and not (t0 - 0.15 <= ts <= t1 + 0.15):
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def default_batch(config):
    return {"quadrotor": 4096, "cartpole": 1024, "pendulum": 1, "unicycle_obstacle_teq": 2048}.get(config, 1024)


def bind_to_gpu_numa_node(index):
    """Pin this process (and therefore the pinned host buffers it first-touches) to the CPUs NVML reports as local to
    GPU `index`, so that every rank's host<->device copies stay on its GPU's NUMA node."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def throughput_options(cfg, max_iterations):
    o = dict(cfg["options"])
    o.update(tolerance=0.0, acceptable_tolerance=0.0, max_iterations=max_iterations)
    return o


def run_reference(args, rank, world):
    """CPU arm: the oracle restatement of the reference algorithm, all host threads, rank 0 only."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_binding as ob
    problems = importlib.import_module("cddp-cpp_b200.problems")
    threads = ob.hardware_threads()
    # the SAME workload as the GPU arm's rank 0: its per-GPU batch (4096 instances at config 3), not a small sample.  One
    # step = iters_per_call DDP iterations of that whole batch (~2-3 s of host work on 16 cores).
    sample = args.cpu_sample or args.batch or default_batch(args.config)
    iters = args.iters_per_call
    cfg = problems.make_config(args.config, batch=sample)
    if cfg.get("solver") == "ipddp" or cfg["spec"]["model"] == "user":
        raise SystemExit("bench.py --impl reference: the CPU arm times the CLDDP headline configs")
    P = ob.OracleProblem(cfg["spec"])
    oo = ob.make_options(**throughput_options(cfg, iters))
    times, done = [], 0
    for step in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = ob.solve_batch(P, oo, cfg["x0"], cfg["xref"], cfg["X0"], cfg["U0"], cfg["ref_traj"], nthreads=threads)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt)
            done += int(r["iterations"].sum())
    total = sum(times)
    value = done / total
    spec = cfg["spec"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["notes"], "n": spec["n"], "m": spec["m"], "horizon": spec["horizon"],
                   "batch_per_gpu": sample, "global_batch": sample, "sample_instances": sample, "iterations_per_step": iters,
                   "convergence_exits": "disabled (tolerance=0) so every instance does every iteration",
                   "note": "CPU restatement of the reference algorithm (oracle/); the reference itself cannot be built "
                           "here (Eigen 3.4 / autodiff are network FetchContent deps)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} instances x {iters} DDP iterations per step, {args.steps} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def max_over_ranks(ms, world):
    if world == 1:
        return ms
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(v, world):
    if world == 1:
        return v
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(v)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def rank_barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def make_solver(cddp, cfg, opts, B, device):
    if cfg.get("solver") == "ipddp":
        return cddp.BatchedIPDDP(cfg["spec"], opts, cddp.default_ipddp_options(**cfg.get("ipddp_options", {})), cfg["constraints"], B,
                                 device=device)
    return cddp.BatchedCLDDP(cfg["spec"], opts, B, device=device)


def load_instances(s, cfg):
    ip = cfg.get("solver") == "ipddp"
    s.set_instances(cfg["x0"], cfg["xref"], None if ip else cfg["X0"], cfg["U0"], cfg["ref_traj"])
    s.initialize()


def measure_iterations(cddp, problems, device, name, global_batch, rank, world, with_cpu, K=20, W=3, peak=None):
    """Secondary workload, fixed number of batched iterations with the convergence exits disabled (the headline protocol):
    `global_batch` instances split contiguously over `world` ranks (strong scaling), device-resident inputs, time = max
    over ranks of the CUDA-event time of K iterations; per-kernel CUDA-event times from a second pass; for IPDDP the
    backward sweep's algorithmic-HBM roofline; the CPU oracle beside it on rank 0 (single-GPU runs only)."""
    import torch
    sharding = importlib.import_module("cddp-cpp_b200.sharding")
    lo, hi = sharding.shard_bounds(global_batch, rank, world)
    B = hi - lo
    full = problems.make_config(name, batch=global_batch)
    cfg = dict(full)
    for k in ("x0", "xref", "X0", "U0", "ref_traj"):
        cfg[k] = None if full[k] is None else full[k][lo:hi]
    ip = cfg.get("solver") == "ipddp"
    opts = cddp.default_options(**dict(cfg["options"], tolerance=0.0, acceptable_tolerance=0.0, max_iterations=W + K))
    s = make_solver(cddp, cfg, opts, B, device)
    s.set_stream(torch.cuda.current_stream().cuda_stream)
    load_instances(s, cfg)
    s.iterate(W)
    it0 = int(s.get_scalars()["iterations"].sum())
    rank_barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s.iterate(K)
    e1.record()
    rank_barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    sc = s.get_scalars()
    done = sum_over_ranks(int(sc["iterations"].sum()) - it0, world)  # instance-iterations actually performed while timed
    running = sum_over_ranks(int((sc["status"] == 0).sum()), world)
    load_instances(s, cfg)
    s.iterate(W)
    s.enable_timing(True)
    s.reset_timing()
    s.iterate(K)
    t = s.get_timing()
    spec = cfg["spec"]
    n, m, N = spec["n"], spec["m"], spec["horizon"]
    kms = {"linearize": t.linearize_ms / max(t.linearize_launches, 1), "backward": t.backward_ms / max(t.backward_launches, 1),
           "forward": t.forward_ms / max(t.forward_launches, 1)}
    out = {"workload": cfg["notes"], "solver": "IPDDP" if ip else "CLDDP", "global_batch": global_batch, "n_gpus": world,
           "batch_per_gpu": B, "horizon": N, "scaling": "strong" if world > 1 else "single GPU",
           "value": done / (ms * 1e-3), "unit": UNIT, "ms_per_iteration": ms / K, "us_per_iteration": 1e3 * ms / K,
           "instance_iterations_timed": int(done), "instances_running_all_iterations": int(running),
           "kernel_ms_per_iteration": kms}
    # backward-sweep roofline in the north star's accounting: CLDDP rows 8(n^2+2nm+n+3m) B per step; IPDDP adds
    # the reads of y, s, g (3d) and the writes of k_y, K_y, k_s, K_s (2(d + dn))
    per_step = 8.0 * (n * n + 2 * n * m + n + 3 * m)
    if ip:
        d = s.d
        out["dual_dim"] = d
        per_step += 8.0 * (3 * d + 2 * (d + d * n))
    if peak:
        alg = per_step * N * B
        ach = alg / (kms["backward"] * 1e-3) / 1e9
        out["roofline"] = {"kernel": "ip_backward" if ip else "backward_sweep", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                           "frac": ach / peak, "algorithmic_bytes_per_step": per_step, "algorithmic_bytes_per_launch": alg,
                           "ms_per_launch": kms["backward"], "traffic": None}
        tf = os.path.join(ROOT, "profiles", "ip_backward_traffic.json" if ip else "backward_traffic.json")
        try:  # dram__bytes_read + dram__bytes_write of the backward pass from the committed ncu capture of this workload
            with open(tf) as f:
                tj = json.load(f)
            if tj.get("batch") == B and tj.get("config") == name:
                out["roofline"]["traffic"] = tj.get("dram_bytes_per_launch")
                out["roofline"]["traffic_source"] = tj.get("source")
        except Exception:
            pass
    s.close()
    if with_cpu and rank == 0 and world == 1:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_binding as ob
        threads = ob.hardware_threads()
        P = ob.OracleProblem(full["spec"])
        oo = ob.make_options(**dict(full["options"], tolerance=0.0, acceptable_tolerance=0.0, max_iterations=W + K))
        t0 = time.perf_counter()
        if ip:
            r = ob.ipddp_solve_batch(P, oo, ob.make_ipddp_options(**full["ipddp_options"]), ob.ConstraintSet(full["constraints"]), full["x0"],
                                     full["xref"], full["U0"], None, nthreads=threads)
        else:
            r = ob.solve_batch(P, oo, full["x0"], full["xref"], full["X0"], full["U0"], full["ref_traj"], nthreads=threads)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": float(r["iterations"].sum()) / dt, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": f"{global_batch} instances x {W + K} iterations, {dt:.2f}s wall"}
    return out


def measure_solve(cddp, problems, device, name, global_batch, rank, world, cpu_sample, with_cpu):
    """Secondary workload measured as what a caller runs: ONE solve() of the whole batch to the workload's own tolerances
    (instances stop when they converge), `global_batch` instances split over `world` ranks; instance-iterations actually
    performed / max-over-ranks device time; the CPU oracle runs the same solves on a sample (rank 0, single-GPU runs)."""
    import torch
    sharding = importlib.import_module("cddp-cpp_b200.sharding")
    lo, hi = sharding.shard_bounds(global_batch, rank, world)
    B = hi - lo
    full = problems.make_config(name, batch=global_batch)
    cfg = dict(full)
    for k in ("x0", "xref", "X0", "U0", "ref_traj"):
        cfg[k] = None if full[k] is None else full[k][lo:hi]
    ip = cfg.get("solver") == "ipddp"
    s = make_solver(cddp, cfg, cddp.default_options(**cfg["options"]), B, device)
    s.set_stream(torch.cuda.current_stream().cuda_stream)
    load_instances(s, cfg)
    s.solve()  # untimed: first launches (of the JIT-compiled module for a user model)
    torch.cuda.synchronize()
    load_instances(s, cfg)
    rank_barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s.solve()
    e1.record()
    rank_barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    sc = s.get_scalars()
    done = sum_over_ranks(int(sc["iterations"].sum()), world)
    load_instances(s, cfg)
    s.enable_timing(True)
    s.reset_timing()
    s.solve()
    t = s.get_timing()
    out = {"workload": cfg["notes"], "solver": "IPDDP" if ip else "CLDDP", "global_batch": global_batch, "n_gpus": world,
           "batch_per_gpu": B, "horizon": cfg["spec"]["horizon"], "scaling": "strong" if world > 1 else "single GPU",
           "value": done / (ms * 1e-3), "unit": UNIT, "solve_ms": ms, "instance_iterations": int(done),
           "mean_iterations": done / global_batch,
           "status_counts_rank0": {int(k): int(v) for k, v in zip(*np.unique(sc["status"], return_counts=True))},
           "mean_cost_rank0": float(np.mean(sc["cost"])),
           "kernel_ms_total_rank0": {"linearize": t.linearize_ms, "backward": t.backward_ms, "forward": t.forward_ms},
           "kernel_launches_rank0": {"linearize": t.linearize_launches, "backward": t.backward_launches, "forward": t.forward_launches,
                                     "other": t.other_launches}}
    if ip:
        out["dual_dim"] = s.d
    s.close()
    if with_cpu and rank == 0 and world == 1:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_binding as ob
        threads = ob.hardware_threads()
        ccfg = problems.make_config(name, batch=cpu_sample)
        P = ob.OracleProblem(ccfg["spec"])
        oo = ob.make_options(**ccfg["options"])
        t0 = time.perf_counter()
        if ip:
            r = ob.ipddp_solve_batch(P, oo, ob.make_ipddp_options(**ccfg.get("ipddp_options", {})), ob.ConstraintSet(ccfg["constraints"]),
                                     ccfg["x0"], ccfg["xref"], ccfg["U0"], None, nthreads=threads)
        else:
            r = ob.solve_batch(P, oo, ccfg["x0"], ccfg["xref"], ccfg["X0"], ccfg["U0"], ccfg["ref_traj"], nthreads=threads)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": float(r["iterations"].sum()) / dt, "unit": UNIT, "cores": min(threads, cpu_sample), "kind": "port",
                               "sample": f"{cpu_sample} instance(s) solved to tolerance, {dt:.3f}s wall"}
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa_cpus = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cddp = importlib.import_module("cddp-cpp_b200")
    problems = importlib.import_module("cddp-cpp_b200.problems")
    peaks, peak_src = measured_peaks()

    # ---- workload: BASELINE config #3 per GPU (weak scaling: the batch shards with no exchange) ----
    base = problems.make_config(args.config, batch=1)
    per_gpu = args.batch or default_batch(args.config)
    cfg = problems.make_config(args.config, batch=per_gpu, seed_offset=1000 * rank)
    spec = cfg["spec"]
    n, m, N, B = spec["n"], spec["m"], spec["horizon"], per_gpu
    K, W = args.steps, args.warmup
    opts = cddp.default_options(**throughput_options(cfg, max(W + K, args.iters_per_call)))
    solver = cddp.BatchedCLDDP(spec, opts, B, device=local_rank)
    stream = torch.cuda.current_stream()
    solver.set_stream(stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- leg 1: device-resident throughput ----
    dev = {k: torch.from_numpy(np.ascontiguousarray(cfg[k])).cuda() for k in ("x0", "xref", "X0", "U0")}
    rt = torch.from_numpy(np.ascontiguousarray(cfg["ref_traj"])).cuda() if cfg["ref_traj"] is not None else None
    solver.set_instances_device(dev["x0"].data_ptr(), dev["xref"].data_ptr(), dev["X0"].data_ptr(), dev["U0"].data_ptr(),
                                rt.data_ptr() if rt is not None else None)
    solver.initialize()
    for _ in range(W):
        solver.iterate(1)
    barrier()
    solver.reset_timing()  # launch counters only (event timing stays off in this leg)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record(stream)
    for _ in range(K):
        solver.iterate(1)
    e1.record(stream)
    barrier()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    tl = solver.get_timing()
    gpu_launches = int(tl.linearize_launches + tl.backward_launches + tl.forward_launches + tl.other_launches)
    clocks = sampler.stop(t_wall0, t_wall1)
    sc = solver.get_scalars()
    iters_done = int(sc["iterations"].sum())
    assert iters_done == B * (W + K), f"every instance must perform every iteration ({iters_done} != {B * (W + K)})"
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # the path's single collective: all-gather of per-instance results (cost, iterations, status) over NCCL
    sharding = importlib.import_module("cddp-cpp_b200.sharding")
    all_cost, all_iters, _ = sharding.gather_results(sc["cost"], sc["iterations"], sc["status"], B * world,
                                                     device=torch.device("cuda", local_rank) if world > 1 else None)
    finite = bool(np.isfinite(all_cost).all())
    assert int(all_iters.sum()) == world * B * (W + K)
    value = world * B * K / (ms * 1e-3)

    # ---- leg 2: per-kernel CUDA-event timing of the same iterations (roofline for the backward sweep) ----
    solver.set_instances_device(dev["x0"].data_ptr(), dev["xref"].data_ptr(), dev["X0"].data_ptr(), dev["U0"].data_ptr(),
                                rt.data_ptr() if rt is not None else None)
    solver.initialize()
    for _ in range(W):
        solver.iterate(1)
    solver.enable_timing(True)
    solver.reset_timing()
    for _ in range(K):
        solver.iterate(1)
    tm = solver.get_timing()
    solver.enable_timing(False)
    bw_ms = tm.backward_ms / max(tm.backward_launches, 1)
    alg_bytes = solver.backward_algorithmic_bytes()
    achieved = alg_bytes / (bw_ms * 1e-3) / 1e9
    peak = float(peaks["hbm_gbs"])
    roofline = {"kernel": "backward_sweep", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": bw_ms,
                "kernel_ms_per_iteration": {"linearize": tm.linearize_ms / max(tm.linearize_launches, 1), "backward": bw_ms,
                                            "forward": tm.forward_ms / max(tm.forward_launches, 1)}}
    traffic_file = os.path.join(ROOT, "profiles", "backward_traffic.json")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as f:
                tj = json.load(f)
            if tj.get("batch") == B and tj.get("config") == args.config:
                roofline["traffic"] = tj.get("dram_bytes_per_launch")
                roofline["traffic_source"] = tj.get("source")
        except Exception:
            pass

    # ---- leg 3: end to end through the C ABI with pinned HOST buffers, the way a serving loop calls it ----
    # Every call uploads its inputs (x0, xref, X0, U0) host->device and downloads its full solution (X, U, K, cost,
    # iterations, status) device->host inside the timed region.  Three solver handles on three CUDA streams are used in
    # rotation (call k+1's upload and call k-1's download overlap call k's kernels), each with its own pinned buffers;
    # "serial" is the same measurement on ONE handle with a blocking call sequence.
    e2e = None
    if not args.no_e2e:
        ipc = args.iters_per_call
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
        lib = solver.lib

        class Lane:
            with_gains = False  # payload of the download: trajectories + scalars, or the full CDDPSolution incl. K_u_

            def __init__(self, slv, strm):
                self.s, self.stream = slv, strm
                self.solved = torch.cuda.Event()  # recorded after the last kernel of this lane's solve
                slv.set_stream(strm.cuda_stream)
                slv.set_options(cddp.default_options(**throughput_options(cfg, ipc)))
                self.hin = {k: pin(cfg[k]) for k in ("x0", "xref", "X0", "U0")}
                self.hrt = pin(cfg["ref_traj"]) if cfg["ref_traj"] is not None else None
                self.hout = {"X": torch.empty((B, N + 1, n), dtype=torch.float64).pin_memory(),
                             "U": torch.empty((B, N, m), dtype=torch.float64).pin_memory(),
                             "K": torch.empty((B, N, m, n), dtype=torch.float64).pin_memory(),
                             "cost": torch.empty(B, dtype=torch.float64).pin_memory(),
                             "iters": torch.empty(B, dtype=torch.int32).pin_memory(),
                             "status": torch.empty(B, dtype=torch.int32).pin_memory()}

            def enqueue(self, blocking, after=None):
                h, hin, hout = self.s.handle, self.hin, self.hout
                cddp._check(lib.cddp_b200_set_instances(h, hin["x0"].data_ptr(), hin["xref"].data_ptr(),
                                                        self.hrt.data_ptr() if self.hrt is not None else None,
                                                        hin["X0"].data_ptr(), hin["U0"].data_ptr()))
                # The two lanes overlap COPIES with kernels, not kernels with kernels: this lane's solve starts when the
                # other lane's solve has finished.  Interleaved, every sweep launch (one 255-register CTA per SM) has to
                # wait for whole SMs to drain of the other lane's CTAs, which cost 10 % of the pipelined throughput.
                if after is not None:
                    self.stream.wait_event(after)
                cddp._check(lib.cddp_b200_solve(h))
                self.solved.record(self.stream)
                get = lib.cddp_b200_get_solution if blocking else lib.cddp_b200_get_solution_async
                cddp._check(get(h, hout["X"].data_ptr(), hout["U"].data_ptr(), hout["K"].data_ptr() if Lane.with_gains else None,
                                hout["cost"].data_ptr(), hout["iters"].data_ptr(), hout["status"].data_ptr(), None, None, None))

            def wait(self):
                self.s.synchronize()
                assert int(self.hout["iters"].sum().item()) == B * ipc
                assert bool(torch.isfinite(self.hout["cost"]).all().item())

        lane0 = Lane(solver, torch.cuda.Stream())
        h2d = sum(v.numel() * v.element_size() for v in lane0.hin.values()) + (lane0.hrt.numel() * 8 if lane0.hrt is not None else 0)
        d2h_full = sum(v.numel() * v.element_size() for v in lane0.hout.values())
        d2h = d2h_full - lane0.hout["K"].numel() * 8
        calls = max(args.e2e_calls, 2)
        extra = [cddp.BatchedCLDDP(spec, cddp.default_options(**throughput_options(cfg, ipc)), B, device=local_rank) for _ in range(2)]
        lanes = [lane0] + [Lane(sv, torch.cuda.Stream()) for sv in extra]
        NL = len(lanes)

        def serial():  # one handle, blocking calls
            lane0.s.set_poll_interval(-1)
            lane0.enqueue(True)
            barrier()
            t0 = time.perf_counter()
            for _ in range(calls):
                lane0.enqueue(True)
            barrier()
            return time.perf_counter() - t0

        def pipelined():
            # three handles, asynchronous calls: call k+1's upload runs against call k-1's download (full duplex, two copy
            # engines) and both against call k's kernels
            for ln in lanes:
                ln.s.set_poll_interval(0)
                ln.enqueue(False)
                ln.wait()
            barrier()
            t0 = time.perf_counter()
            inflight = [False] * NL
            for k in range(calls):
                i = k % NL
                if inflight[i]:
                    lanes[i].wait()
                prev = (i - 1) % NL
                lanes[i].enqueue(False, after=lanes[prev].solved if inflight[prev] else None)
                inflight[i] = True
            for j in range(NL):
                i = (calls + j) % NL  # oldest first
                if inflight[i]:
                    lanes[i].wait()
            barrier()
            return time.perf_counter() - t0

        def over_ranks(*ts):
            if world == 1:
                return ts
            t = torch.tensor(list(ts), device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return tuple(float(v) for v in t)

        # payload 1 (e2e.value): what a serving / MPC caller reads back every call — the optimised trajectories X, U and the
        # per-instance cost / iterations / status.  payload 2 (with_gains_*): the whole CDDPSolution of the reference
        # including the feedback gains K_u_ (cddp_core.hpp:54-103), 3x the bytes; both through the same entry points.
        Lane.with_gains = False
        dt_serial = serial()
        dt = pipelined()
        Lane.with_gains = True
        dt_full = pipelined()
        Lane.with_gains = False
        dt, dt_serial, dt_full = over_ranks(dt, dt_serial, dt_full)
        e2e = {"value": world * B * ipc * calls / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "iterations_per_call": ipc, "calls": calls, "ms_per_call": 1e3 * dt / calls,
               "serial_value": world * B * ipc * calls / dt_serial, "serial_ms_per_call": 1e3 * dt_serial / calls,
               "with_gains_value": world * B * ipc * calls / dt_full, "with_gains_ms_per_call": 1e3 * dt_full / calls,
               "with_gains_d2h_bytes_per_step": int(d2h_full),
               "host_link_GBps_per_gpu": {"payload": (h2d + d2h) * calls / dt / 1e9, "with_gains": (h2d + d2h_full) * calls / dt_full / 1e9},
               "numa_local_cpus": numa_cpus,
               "api": "cddp_b200_set_instances + cddp_b200_solve + cddp_b200_get_solution[_async] on pinned host buffers allocated "
                      "on the GPU's NUMA node; value = X, U, cost, iterations, status downloaded every call (K pointer NULL), "
                      "with_gains_value = the same calls downloading the feedback gains K too; three handles used in rotation "
                      "on three CUDA streams (copies overlap the other lanes' kernels, kernels of different lanes are serialised "
                      "by an event); serial_value = one handle, blocking calls"}
        for sv in extra:
            sv.close()
        solver.set_stream(stream.cuda_stream)

    # ---- CPU baseline (rank 0, N=1 only): the oracle on a bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_binding as ob
        threads = ob.hardware_threads()
        sample = args.cpu_sample or B  # the same batch the GPU ran (~10 s of host work at cfg 3)
        ccfg = problems.make_config(args.config, batch=sample)
        P = ob.OracleProblem(ccfg["spec"])
        it_cpu = W + K
        oo = ob.make_options(**throughput_options(ccfg, it_cpu))
        t0 = time.perf_counter()
        r = ob.solve_batch(P, oo, ccfg["x0"], ccfg["xref"], ccfg["X0"], ccfg["U0"], ccfg["ref_traj"], nthreads=threads)
        dt = time.perf_counter() - t0
        cpu = {"value": float(r["iterations"].sum()) / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{sample} instances x {it_cpu} DDP iterations of the same workload, {dt:.1f}s wall, std::thread static partition",
               "note": "CPU restatement of the reference algorithm (Eigen/autodiff unavailable, reference not buildable here)"}

    # ---- the other BASELINE.json configs, each on the GPU count it is stated for (all ranks take part: strong scaling) ----
    other = None
    if args.config == "quadrotor" and not args.no_other:
        peak_gbs = float(peaks["hbm_gbs"])
        with_cpu = not args.no_cpu_baseline
        jobs = []
        if world == 1:
            jobs += [("clddp_config1_pendulum_single_trajectory", lambda: measure_solve(cddp, problems, local_rank, "pendulum", 1, rank, world, 1, with_cpu)),
                     ("clddp_config2_cartpole_batch1024", lambda: measure_iterations(cddp, problems, local_rank, "cartpole", 1024, rank, world, with_cpu, peak=peak_gbs)),
                     ("clddp_config2_cartpole_batch1024_solve", lambda: measure_solve(cddp, problems, local_rank, "cartpole", 1024, rank, world, 1024, with_cpu))]
        if world in (1, 4):
            jobs += [("ipddp_config4_unicycle_obstacle_teq_batch2048", lambda: measure_iterations(cddp, problems, local_rank, "unicycle_obstacle_teq", 2048, rank, world, with_cpu, peak=peak_gbs))]
        if world in (1, 8):
            jobs += [("clddp_config5_7dof_user_model_batch8192", lambda: measure_solve(cddp, problems, local_rank, CONFIG5_CLDDP, 8192, rank, world, 1024, with_cpu)),
                     ("ipddp_config5_7dof_user_model_batch8192", lambda: measure_solve(cddp, problems, local_rank, CONFIG5_IPDDP, 8192, rank, world, 256, with_cpu))]
        other = {}
        for key, fn in jobs:
            try:
                other[key] = fn()
            except Exception as e:  # secondary measurement: never take the bench line down with it
                other[key] = {"error": repr(e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["notes"], "n": n, "m": m, "horizon": N, "batch_per_gpu": B, "global_batch": B * world,
                       "record_layout": "%s (%d B/record)" % solver.get_record_layout(),
                       "parallelism": f"batch sharded over {world} GPU(s), no data-path collective; one all-gather of per-instance results",
                       "step": "one batched DDP iteration (linearise + backward sweep + forward line search, all alphas in parallel)",
                       "convergence_exits": "disabled (tolerance=0) so every instance does every iteration",
                       "l2": "no flush needed: per-iteration working set (linearisation records %.0f MB + gains %.0f MB + "
                             "trajectories %.0f MB) exceeds the 126 MB L2"
                             % (1e-6 * B * N * solver.get_record_layout()[1], 8e-6 * B * N * m * n, 2 * 8e-6 * B * N * (n + m)),
                       "line_search_alphas": solver.num_alphas},
            "batched_iterations_per_s": K / (ms * 1e-3),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": gpu_launches,
            "clocks": clocks, "all_costs_finite": finite, "other_workloads": other,
        }
        print(json.dumps(line), flush=True)
    solver.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
