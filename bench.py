#!/usr/bin/env python
"""bench.py -- system-steps/sec of the thread-per-system ensemble integrator on B200.

    python bench.py --gpus N --steps K --warmup W            # product arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...   # reference arm: the reference's own
                                                              # kernel source on the host cores

Workload at every N: BASELINE.json configs[1] -- T2_Lorenz, RKCK45 adaptive, 2^20 rho-sweep
instances per GPU, FP64 (``--precision f32`` / ``--workload impact|kellermiksis`` select the other
configurations).  One "step" = one integration phase (one Solve()) of all instances from the
tutorial's initial condition.  Metric: system-steps/sec, a system-step being one stepper
invocation (accepted or rejected, 6 right-hand-side evaluations) of one instance.

Timed regions (all after W >= 3 warm-up steps):
  value : inputs resident in HBM; each step restores the initial condition device-to-device and
          launches the phase; CUDA events on the solver's stream, barrier + synchronize on both
          sides, max over ranks.
  e2e   : the same K steps through the public API with HOST buffers: every step uploads the
          inputs from pinned host memory, integrates, and downloads state + time; host wall
          clock between synchronisation points, max over ranks.
Multi-GPU: one process per GPU (torchrun); instances are independent, so each rank integrates its
own 2^20-instance tile-cyclic shard of a (N x 2^20)-instance sweep with no collective on the data
path (weak scaling); torch.distributed/NCCL only carries the barrier and the reduction of the
timings.
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

FLOPS_NOTE = "algorithmic flops per system-step, SURVEY.md 8(d): add/sub/mul/div/sqrt = 1, FMA = 2, transcendentals 0"


def build_workload(W, name: str, precision: str, NT: int, rank: int, world: int):
    """This rank's shard: tile-cyclic slices (4096 instances) of the global (world x NT) sweep."""
    dtype = np.float64 if precision == "f64" else np.float32
    gen = {"lorenz": lambda n: W.lorenz(NT=n, algorithm=W.RKCK45, dtype=dtype),
           "lorenz_rk4": lambda n: W.lorenz(NT=n, algorithm=W.RK4, dtype=dtype),
           "impact": lambda n: W.impact_t6(NT=n),
           "kellermiksis": lambda n: W.keller_miksis_t5(NT=n, ndo=0),
           "kellermiksis_dense": lambda n: W.keller_miksis_t5(NT=n, ndo=200),
           "duffing": lambda n: W.duffing_t1(NT=n, algorithm=W.RK4)}[name]
    sharding = importlib.import_module("massively-parallel-gpu-ode-solver_b200.sharding")
    return sharding.shard_workload(gen(NT * world), NT, rank, world)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 7:
                continue
            try:
                mx = float(p[1])
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(p[0]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(W, name: str, precision: str, sample: int, threads: int):
    """The reference's own kernel source compiled for the host (oracle/_ref, kind "reference"), or
    our CPU restatement of it when that library is absent (kind "port"), over every k-th instance
    of the sweep, on `threads` host threads.  Returns (system-steps, seconds, kind)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py
    from cases import subsample
    full = build_workload(W, name, precision, 1 << 20 if "keller" not in name else 1 << 18, 0, 1)
    wl = subsample(full, sample)
    if oracle_py.ref_available(wl, fast=True):
        run, kind = (lambda: oracle_py.run_ref(wl, phases=1, host_threads=threads, fast=True)), "reference"
    else:
        if not os.path.exists(oracle_py.oracle_library_path(fast=True)):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle_perthread_fast.so"], check=True, capture_output=True)
        run, kind = (lambda: oracle_py.run_oracle(wl, phases=1, host_threads=threads, fast=True, double_literals=True)), "port"
    t0 = time.perf_counter()
    _, att, _ = run()
    return int(att.sum()), time.perf_counter() - t0, kind, wl.NT


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="lorenz")
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--instances", type=int, default=0, help="instances per GPU (default: the BASELINE size)")
    ap.add_argument("--lane-refill", type=int, default=-1)
    ap.add_argument("--threads-per-block", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    pkg = importlib.import_module("massively-parallel-gpu-ode-solver_b200")
    W = pkg.workloads
    NT = args.instances or ((1 << 18) if "keller" in args.workload else (1 << 20))
    flops_per_step = {"lorenz": 236, "lorenz_rk4": 79, "impact": 230, "kellermiksis": 413, "kellermiksis_dense": 413, "duffing": 58}[args.workload]
    metric = "system-steps/sec (%s %s)" % ("FP64" if args.precision == "f64" else "FP32", "RK4" if args.workload in ("lorenz_rk4", "duffing") else "RKCK45")
    config = {"workload": f"{args.workload}: {NT} instances per GPU x {world} GPU(s), one integration phase per step "
                          f"(BASELINE configs[1] T2_Lorenz rho-sweep)" if args.workload == "lorenz" else f"{args.workload}: {NT} instances per GPU x {world}",
              "instances_per_gpu": NT, "sharding": "tile-cyclic (4096), no collective on the data path",
              "l2": "inputs + outputs per step exceed nothing cached: state is restored device-to-device from a snapshot before every step; kernel is FMA-pipe bound, HBM traffic < 1% of time",
              "flops_per_system_step": flops_per_step, "flops_convention": FLOPS_NOTE}

    # ------------------------------------------------------------------ reference arm (CPU) ------
    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        sample = args.cpu_sample or (1 << 16)
        for _ in range(args.warmup):
            cpu_reference_run(W, args.workload, args.precision, max(sample // 8, 256), threads)
        steps_total, secs, kind = 0, 0.0, "reference"
        for _ in range(args.steps):
            n, s, kind, nt = cpu_reference_run(W, args.workload, args.precision, sample, threads)
            steps_total += n
            secs += s
        v = steps_total / secs
        print(json.dumps({"impl": "reference", "metric": metric, "value": v, "unit": "system-steps/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "system-steps/s", "cores": threads, "kind": kind,
                                           "sample": f"every {max(NT // sample, 1)}-th instance of the sweep ({sample} instances) per step, {threads} host threads"},
                          "e2e": {"value": v, "unit": "system-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return

    # ------------------------------------------------------------------ product arm (CUDA) -------
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    wl = build_workload(W, args.workload, args.precision, NT, rank, world)
    sh = wl.shape
    ps = pkg.ProblemSolver(wl.name, NT, NDO=sh.NDO, Algorithm=sh.Algorithm, Precision=sh.dtype, Device=local_rank)
    pkg.apply_workload(ps, wl)
    if args.lane_refill >= 0:
        ps.SolverOption(pkg.Options.LaneRefill, args.lane_refill)
    if args.threads_per_block:
        ps.SolverOption(pkg.Options.ThreadsPerBlock, args.threads_per_block)
    ps.SynchroniseFromHostToDevice(pkg.Variables.All)
    ps.SaveState()
    stream = torch.cuda.ExternalStream(ps.stream, device=local_rank)

    def step_device():
        ps.RestoreState()
        ps.Solve()

    for _ in range(args.warmup):
        step_device()
    ps.SynchroniseDevice()
    ps.ResetStepCounters()

    # ---- value: device-resident ---------------------------------------------------------------
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 2)]
    sampler = ClockSampler(local_rank)
    time.sleep(0.25)
    barrier(); torch.cuda.synchronize()
    t_host0 = time.time()
    ev[0].record(stream)
    for k in range(args.steps):
        ps.RestoreState()
        ev[2 + 2 * k].record(stream)
        ps.Solve()
        ev[3 + 2 * k].record(stream)
    ev[1].record(stream)
    torch.cuda.synchronize(); barrier()
    t_host1 = time.time()
    clocks = sampler.stop(t_host0, t_host1)
    region_ms = ev[0].elapsed_time(ev[1])
    kernel_ms = [ev[2 + 2 * k].elapsed_time(ev[3 + 2 * k]) for k in range(args.steps)]
    attempted, accepted = ps.TotalSteps()

    # ---- e2e: host buffers through the public API ------------------------------------------------
    V = pkg.Variables
    in_vars = [V.TimeDomain, V.ActualState, V.ActualTime, V.ControlParameters, V.SharedParameters, V.Accessories, V.IntegerAccessories, V.DenseIndex]
    out_vars = [V.ActualState, V.ActualTime, V.Accessories, V.IntegerAccessories]
    pinned_in, pinned_out = {}, {}
    for v in in_vars:
        a = ps.host(v)
        if a.size:
            t = torch.from_numpy(np.ascontiguousarray(a).copy()).pin_memory()
            pinned_in[v] = t
    for v in out_vars:
        a = ps.host(v)
        if a.size:
            pinned_out[v] = torch.empty_like(torch.from_numpy(np.ascontiguousarray(a).copy())).pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in pinned_in.values())
    d2h = sum(t.numel() * t.element_size() for t in pinned_out.values())

    def step_e2e():
        for v, t in pinned_in.items():
            ps.Upload(v, t.data_ptr(), t.numel() * t.element_size())
        ps.Solve()
        for v, t in pinned_out.items():
            ps.Download(v, t.data_ptr(), t.numel() * t.element_size())
        ps.InsertSynchronisationPoint()
        ps.SynchroniseSolver()

    for _ in range(2):
        step_e2e()
    ps.ResetStepCounters()
    barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize(); barrier()
    e2e_s = time.perf_counter() - t0
    e2e_attempted, _ = ps.TotalSteps()
    final_state_checksum = float(pinned_out[V.ActualState].double().sum())

    # ---- reduce over ranks -------------------------------------------------------------------------
    sharding = importlib.import_module("massively-parallel-gpu-ode-solver_b200.sharding")
    (region_ms, e2e_s, kmean), counts = sharding.reduce_over_ranks([region_ms, e2e_s, float(np.mean(kernel_ms))],
                                                                   [attempted, accepted, e2e_attempted], device="cuda")
    attempted, accepted, e2e_attempted = [int(x) for x in counts]

    if rank == 0:
        value = attempted / (region_ms * 1e-3)
        precision_bytes = 8 if args.precision == "f64" else 4
        peak_tf = pkg.probe_fma_peak(local_rank, precision_bytes)
        per_rank_attempted_per_launch = attempted / world / args.steps
        achieved_tf = per_rank_attempted_per_launch * flops_per_step / (kmean * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(f"{args.workload}_{args.precision}")
        out = {"metric": metric, "value": value, "unit": "system-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": region_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": args.precision, "data": "synthetic", "config": config,
               "accepted_steps_per_s": accepted / (region_ms * 1e-3),
               "system_steps_per_step": attempted / args.steps,
               "roofline": {"bound": "fp64_fma" if precision_bytes == 8 else "fp32_fma", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                            "frac": achieved_tf / peak_tf, "traffic": traffic,
                            "peak_source": "measured on this GPU by mpgos_probe_fma_peak (register-resident FMA chains, CUDA events); MEASURED_PEAKS.json carries no FP64/FP32 FMA figure",
                            "kernel_ms": kmean, "algorithmic_flops_per_launch": per_rank_attempted_per_launch * flops_per_step},
               "e2e": {"value": e2e_attempted / e2e_s, "unit": "system-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": e2e_s / args.steps * 1e3, "state_checksum": final_state_checksum},
               "gpu_launches": args.steps, "clocks": clocks}
        if world == 1 and not args.no_reference_gpu:
            # reported baseline, not a target: the reference MPGOS compiled for sm_100a, same inputs, same event bracket
            try:
                sys.path.insert(0, os.path.join(ROOT, "oracle"))
                import oracle_py
                if oracle_py.ref_gpu_available(wl):
                    _, ref_ms = oracle_py.run_ref_gpu(wl, phases=1, threads_per_block=64, repeats=3)
                    out["reference_gpu"] = {"kernel_ms": ref_ms, "value": per_rank_attempted_per_launch / (ref_ms * 1e-3), "unit": "system-steps/s",
                                            "speedup_kernel": ref_ms / kmean,
                                            "what": "reference MPGOS SingleSystem_PerThread kernel compiled unmodified by nvcc for sm_100a (oracle/_ref), 64 threads/block, best of 3, same inputs; step count taken from our counters"}
            except Exception as e:   # the baseline must never break the bench line
                out["reference_gpu"] = {"unavailable": str(e)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sample = args.cpu_sample or (1 << 19 if threads >= 16 else 1 << 17)
            n, s, kind, nt = cpu_reference_run(W, args.workload, args.precision, sample, threads)
            out["cpu_baseline"] = {"value": n / s, "unit": "system-steps/s", "cores": threads, "kind": kind,
                                   "sample": f"every {max(NT // nt, 1)}-th instance of the sweep ({nt} instances), one phase, {threads} host threads, {s:.1f} s"}
        print(json.dumps(out))
    ps.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
