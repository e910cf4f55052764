#!/usr/bin/env python
"""bench.py -- system-steps/sec of the B200-native ensemble ODE integrator.

    python bench.py --gpus N --steps K --warmup W            # product arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...   # reference arm: the reference's own
                                                              # kernel source on the host cores

Workload at every N (default): BASELINE.json configs[1] -- T2_Lorenz, RKCK45 adaptive, 2^20 rho-sweep
instances per GPU, FP64.  ``--workload`` selects the other configurations:
    duffing (C1), lorenz / lorenz_rk4 (C2, --precision f32 for the FP32 run), impact (C3),
    kellermiksis / kellermiksis_dense (C4), coupled (C5: 8192 systems of 64 coupled bubbles per GPU).
One "step" = ``--phases`` integration phases (Solve() calls, default 1) of all instances from the
tutorial's initial condition.  Metric: system-steps/sec, a system-step being one stepper invocation
(accepted or rejected; 6 right-hand-side evaluations for RKCK45, 4 for RK4) of one instance (for the
coupled module: of one whole system of 64 units).

Timed regions (all after W >= 3 warm-up steps):
  value : inputs resident in HBM; each step restores the initial condition device-to-device and
          launches the phase(s); L2 is flushed before every step (256 MiB overwritten on the same
          stream); the step time is the sum of the CUDA-event brackets around (restore + phases) on
          the solver's stream, barrier + synchronize on both sides of the K steps, max over ranks.
  e2e   : the same K steps through the public API with HOST buffers: every step uploads the inputs
          from pinned host memory, integrates, and downloads state + time.  The sweep is driven as
          ``--e2e-chunks`` solver objects with their own streams so that copies overlap kernels --
          the double-buffering pattern of the reference's own tutorial
          (Tutorials_SingleSystem_PerThread/T7_DoubleBuffering/DoubleBuffering.cu:89-144): an object is
          waited for, refilled and relaunched while the others compute; every step uploads and downloads
          every chunk, the region ends when the last result is in host memory; host wall clock, max over ranks.
Multi-GPU: one process per GPU (torchrun); instances are independent, so each rank integrates its own
tile-cyclic shard of a (N x per-GPU)-instance sweep with no collective on the data path (weak
scaling); torch.distributed/NCCL only carries the barrier and the reduction of the timings.
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

# name -> (module kind, BASELINE config, instances (systems) per GPU, flops per system-step, algorithm label)
WORKLOADS = {
    "lorenz":             ("perthread", "configs[1] T2_Lorenz rho-sweep", 1 << 20, 236, "RKCK45"),
    "lorenz_rk4":         ("perthread", "T2_Lorenz as shipped (RK4, dt=1e-3)", 1 << 20, 79, "RK4"),
    "impact":             ("perthread", "configs[2] T6_ImpactDynamics", 1 << 20, 230, "RKCK45"),
    "kellermiksis":       ("perthread", "configs[3] T5 Keller-Miksis without dense output", 1 << 18, 413, "RKCK45"),
    "kellermiksis_dense": ("perthread", "configs[3] T5 Keller-Miksis + dense output (200 points)", 1 << 18, 413, "RKCK45"),
    "duffing":            ("perthread", "configs[0] T1_Reference Duffing, RK4", 10000, 58, "RK4"),
    "coupled":            ("coupled", "configs[4] T2_BubbleEnsemble, 64 units, dense coupling matrix", 8192, 78857, "RKCK45"),
    # the same two configurations with the opt-in transcendentals of include/mpgos_b200/fast_math.cuh (pow / sincospi / sin / cos <= 3 ulp)
    "kellermiksis_dense_fm": ("perthread", "configs[3] T5 Keller-Miksis + dense output, opt-in fast_math.cuh transcendentals", 1 << 18, 413, "RKCK45"),
    "coupled_fm":         ("coupled", "configs[4] T2_BubbleEnsemble, 64 units, opt-in fast_math.cuh transcendentals", 8192, 78857, "RKCK45"),
}
COUPLED_TILE = 64


def build_workload(pkg, name: str, precision: str, n: int, rank: int, world: int):
    """This rank's shard: tile-cyclic slices of the global (world x n)-instance sweep."""
    W = pkg.workloads
    sharding = importlib.import_module(pkg.__name__ + ".sharding")
    dtype = np.float64 if precision == "f64" else np.float32
    if name in ("coupled", "coupled_fm"):
        full = pkg.workloads_coupled.bubble_ensemble(NS=n * world, UPS=64, dtype=dtype, fast_math=name.endswith("_fm"))
        if world > 1:
            import copy
            wl = copy.copy(full)
            wl.buffers = full.buffers.take(sharding.shard_indices(n, rank, world, COUPLED_TILE if n % COUPLED_TILE == 0 else 1))
            return wl
        return full
    single = (lambda wl: W.single_precision(wl)) if precision == "f32" else (lambda wl: wl)     # our FP32 settings: tolerances 1e-5
    gen = {"lorenz": lambda m: W.lorenz(NT=m, algorithm=W.RKCK45, dtype=dtype),
           "lorenz_rk4": lambda m: W.lorenz(NT=m, algorithm=W.RK4, dtype=dtype),
           "impact": lambda m: single(W.impact_t6(NT=m)),
           "kellermiksis": lambda m: single(W.keller_miksis_t5(NT=m, ndo=0)),
           "kellermiksis_dense": lambda m: single(W.keller_miksis_t5(NT=m, ndo=200)),
           "kellermiksis_dense_fm": lambda m: W.keller_miksis_t5(NT=m, ndo=200, fast_math=True),
           "duffing": lambda m: single(W.duffing_t1(NT=m, algorithm=W.RK4))}[name]
    tile = sharding.TILE if n % sharding.TILE == 0 else 1
    return sharding.shard_workload(gen(n * world), n, rank, world, tile)


class quiet_stdout:
    """The reference's host object narrates its construction on stdout (C++ iostream); the contract is ONE line there."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *exc):
        os.dup2(self.saved, 1)
        os.close(self.null)
        os.close(self.saved)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
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


def bind_to_gpu_numa_node(local_rank: int):
    """Multi-GPU runs: keep this rank's host thread (and, by first touch, its pinned buffers) on the NUMA node its GPU hangs off, so that
    eight ranks' host<->device copies do not all cross one memory controller.  Returns a description for the bench line.  On a box
    whose GPUs all sit on one node (nvidia-smi topo: the pool's 8 x B200 boxes) this changes nothing -- and says so."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]
        if node < 0 or len(nodes) < 2:
            return {"numa_node": node, "numa_nodes": len(nodes), "bound": False}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "numa_nodes": len(nodes), "bound": bool(cpus), "cpus": len(cpus)}
    except Exception as e:
        return {"bound": False, "error": str(e)[:80]}


def cpu_reference_run(pkg, name: str, precision: str, sample: int, threads: int, phases: int):
    """The reference's own kernel source compiled for the host (oracle/_ref, kind "reference"), or our
    CPU restatement of it when that library is absent / for the coupled module (kind "port"), over every
    k-th instance of the sweep, on `threads` host threads.  Returns (system-steps, seconds, kind, instances)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py
    if name.endswith("_fm"):       # the baselines run the system as the reference defines it (CUDA / C math library)
        name = name[:-3]
    n_full = WORKLOADS[name][2]
    full = build_workload(pkg, name, precision, n_full, 0, 1)
    if name == "coupled":
        import copy
        idx = np.unique(np.linspace(0, n_full - 1, sample).round().astype(np.int64))
        wl = copy.copy(full)
        wl.buffers = full.buffers.take(idx)
        if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle_coupled_fast.so")):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle_coupled_fast.so"], check=True, capture_output=True)
        t0 = time.perf_counter()
        _, att, _ = oracle_py.run_coupled_oracle(wl, phases=phases, host_threads=threads, fast=True)
        return int(att.sum()), time.perf_counter() - t0, "port", int(idx.size)
    from cases import subsample
    wl = subsample(full, sample)
    if oracle_py.ref_available(wl, fast=True):
        run, kind = (lambda: oracle_py.run_ref(wl, phases=phases, host_threads=threads, fast=True)), "reference"
    else:
        if not os.path.exists(oracle_py.oracle_library_path(fast=True)):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle_perthread_fast.so"], check=True, capture_output=True)
        run, kind = (lambda: oracle_py.run_oracle(wl, phases=phases, host_threads=threads, fast=True, double_literals=True)), "port"
    t0 = time.perf_counter()
    _, att, _ = run()
    return int(att.sum()), time.perf_counter() - t0, kind, wl.NT


class Arm:
    """One solver object of either module behind a common face (what bench needs of the public API)."""

    def __init__(self, pkg, wl, kind: str, device: int, args):
        self.pkg, self.wl, self.kind = pkg, wl, kind
        sh = wl.shape
        if kind == "coupled":
            self.V = pkg.CoupledVariables
            self.ps = pkg.CoupledProblemSolver(wl.name, wl.buffers.NS, sh.UPS, CBW=sh.CBW, CCI=sh.CCI, NDO=sh.NDO, Algorithm=sh.Algorithm, Device=device,
                                               TPB=sh.TPB, SPB=sh.SPB, Precision=sh.dtype)
            pkg.apply_coupled_workload(self.ps, wl)
            if args.tensor_core_coupling >= 0:
                self.ps.SolverOption(pkg.CoupledOptions.TensorCoreCoupling, args.tensor_core_coupling)
            self.in_vars = [self.V.TimeDomain, self.V.ActualState, self.V.ActualTime, self.V.UnitParameters, self.V.CouplingStrength]
            self.out_vars = [self.V.ActualState, self.V.ActualTime]
        else:
            self.V = pkg.Variables
            self.ps = pkg.ProblemSolver(wl.name, wl.NT, NDO=sh.NDO, Algorithm=sh.Algorithm, Precision=sh.dtype, Device=device)
            pkg.apply_workload(self.ps, wl)
            if args.lane_refill >= 0:
                self.ps.SolverOption(pkg.Options.LaneRefill, args.lane_refill)
            if args.threads_per_block:
                self.ps.SolverOption(pkg.Options.ThreadsPerBlock, args.threads_per_block)
            V = self.V
            self.in_vars = [V.TimeDomain, V.ActualState, V.ActualTime, V.ControlParameters, V.SharedParameters, V.Accessories, V.IntegerAccessories, V.DenseIndex]
            self.out_vars = [V.ActualState, V.ActualTime, V.Accessories, V.IntegerAccessories]
        self.ps.SynchroniseFromHostToDevice(self.V.All)
        self.ps.SynchroniseDevice()

    def total_steps(self):
        att, acc = self.ps.StepCounters()
        return int(att.sum()), int(acc.sum())

    # ---- e2e: caller-owned pinned host arrays -> device -> caller-owned pinned host arrays ------------
    def make_host_buffers(self, torch):
        self.user_in = {v: torch.from_numpy(np.ascontiguousarray(self.ps.host(v)).copy()).pin_memory() for v in self.in_vars if self.ps.host(v).size}
        self.user_out = {v: torch.empty_like(torch.from_numpy(np.ascontiguousarray(self.ps.host(v)).copy())).pin_memory() for v in self.out_vars if self.ps.host(v).size}
        h2d = sum(t.numel() * t.element_size() for t in self.user_in.values())
        d2h = sum(t.numel() * t.element_size() for t in self.user_out.values())
        return h2d, d2h

    def enqueue_e2e(self, phases: int):
        ps = self.ps
        for v, t in self.user_in.items():
            ps.Upload(v, t.data_ptr(), t.numel() * t.element_size())
        for _ in range(phases):
            ps.Solve()
        for v, t in self.user_out.items():
            ps.Download(v, t.data_ptr(), t.numel() * t.element_size())
        ps.InsertSynchronisationPoint()

    def finish_e2e(self):
        self.ps.SynchroniseSolver()

    def close(self):
        self.ps.close()


EXTRA_CONFIGS = (      # key, workload, precision, phases per step  -- run after the headline, same timing rules, fewer steps
    ("C1_duffing_rk4_f64", "duffing", "f64", 8),
    ("C2_lorenz_rkck45_f32", "lorenz", "f32", 1),
    ("C3_impact_rkck45_f64", "impact", "f64", 8),
    ("C4_kellermiksis_dense_rkck45_f64", "kellermiksis_dense", "f64", 2),
    ("C5_coupled_bubbles_rkck45_f64", "coupled", "f64", 1),
    ("C4_kellermiksis_dense_fastmath", "kellermiksis_dense_fm", "f64", 2),
    ("C5_coupled_bubbles_fastmath", "coupled_fm", "f64", 1),
)


def reference_gpu_baseline(kind, wl, N, precision, phases, per_rank_attempted_per_launch, kmean):
    """Reported baseline, not a target: the reference MPGOS compiled unmodified for sm_100a (oracle/_ref), same inputs, same
    event bracket -- both without a register cap and at the -maxrregcount of the reference's own tutorial makefile ("stock"),
    over 64 / 128 / 256 threads per block; the best setting is the baseline, every setting tried is listed."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    if wl.name.endswith("_fm"):      # the reference has one form of the system: its own, on the CUDA math library
        import copy
        wl = copy.copy(wl)
        wl.name = wl.name[:-3]
    tried = {}
    if kind == "coupled":
        if not (N == 8192 and precision == "f64"):
            return None
        for tag in ("bubble_ensemble_u64_s1", "bubble_ensemble_u64_s2", "bubble_ensemble_u64_s1_stock", "bubble_ensemble_u64_s2_stock", "bubble_ensemble_u64_s4_stock"):
            if not os.path.exists(oracle_py.coupled_ref_library_path(tag, gpu=True)):
                continue
            for shared in (0, 1):
                try:
                    with quiet_stdout():
                        _, ms = oracle_py.run_coupled_ref_gpu(tag, wl, phases=phases, shared_matrices=shared, repeats=2)
                    tried[f"{tag.replace('bubble_ensemble_u64_', '')},SharedCouplingMatrices={shared}"] = ms / phases
                except Exception as e:
                    tried[f"{tag},SharedCouplingMatrices={shared}"] = None
        what = ("reference MPGOS CoupledSystems_PerBlock kernel compiled unmodified by nvcc for sm_100a; settings: s1/s2/s4 = SPB 1/2/4 with TPB = 64 SPB, "
                "'stock' = -maxrregcount=128 as in the reference's T2_BubbleEnsemble/makefile, SharedCouplingMatrices 0/1")
    else:
        for variant in ("", "stock"):
            if not oracle_py.ref_gpu_available(wl, variant):
                continue
            for tpb in (64, 128, 256):
                try:
                    with quiet_stdout():
                        _, ms = oracle_py.run_ref_gpu(wl, phases=phases, threads_per_block=tpb, repeats=2, variant=variant)
                    tried[f"{variant or 'nocap'},tpb={tpb}"] = ms / phases
                except Exception as e:
                    tried[f"{variant or 'nocap'},tpb={tpb}"] = None
        what = ("reference MPGOS SingleSystem_PerThread kernel compiled unmodified by nvcc for sm_100a; settings: 'nocap' = no register cap, "
                "'stock' = the -maxrregcount of the reference's own tutorial makefile (T2_Lorenz 64, T1_Reference / T6_ImpactDynamics 80, T5 128), ThreadsPerBlock 64/128/256")
    ok = {k: v for k, v in tried.items() if v}
    if not ok:
        return None
    best = min(ok, key=ok.get)
    ref_ms = ok[best]
    return {"kernel_ms": ref_ms, "value": per_rank_attempted_per_launch / (ref_ms * 1e-3), "unit": "system-steps/s", "speedup_kernel": ref_ms / kmean,
            "best_setting": best, "kernel_ms_by_setting": tried,
            "what": what + " (oracle/_ref); best of 2 per setting, the fastest setting is the baseline; same inputs; step count taken from our counters"}


_PEAKS = {}


def measured_peak(pkg, device, precision_bytes):
    """FMA peak of this GPU in TFLOP/s, measured in this process: FP64 = max(DFMA probe, DMMA probe), FP32 = FFMA probe."""
    key = (device, precision_bytes)
    if key not in _PEAKS:
        fma = pkg.probe_fma_peak(device, precision_bytes)
        if precision_bytes == 8:
            mma = pkg.probe_dmma_peak(device)
            _PEAKS[key] = (max(fma, mma), {"dfma_probe": fma, "dmma_probe": mma})
        else:
            _PEAKS[key] = (fma, {"ffma_probe": fma})
    return _PEAKS[key]


def measure(pkg, args, workload, precision, N, phases, steps, warmup, rank, local_rank, world, full):
    """One configuration through the product: device-resident `value`, `e2e` through the public API with host buffers, roofline,
    clocks, and (N = 1) the two reported baselines.  Returns the result dict on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()

    kind, cfg_name, n_default, flops_per_step, alg = WORKLOADS[workload]
    unit = "systems" if kind == "coupled" else "instances"
    metric = "system-steps/sec (%s %s)" % ("FP64" if precision == "f64" else "FP32", alg)
    config = {"workload": f"{workload}: {N} {unit} per GPU x {world} GPU(s), {phases} integration phase(s) per step (BASELINE {cfg_name})",
              f"{unit}_per_gpu": N, "phases_per_step": phases,
              "sharding": "tile-cyclic, no collective on the data path",
              "l2": "flushed between timed iterations (a 256 MiB buffer is overwritten on the solver's stream, outside the summed event brackets); state is restored device-to-device from a snapshot before every step; the kernel is bound by instruction issue (an FP64 instruction takes two issue slots), its HBM traffic takes < 1% of its time (profiles/traffic.json), so cache residency between steps does not enter the timing",
              "flops_per_system_step": flops_per_step, "flops_convention": FLOPS_NOTE}

    wl = build_workload(pkg, workload, precision, N, rank, world)
    arm = Arm(pkg, wl, kind, local_rank, args)
    ps = arm.ps
    ps.SaveState()
    stream = torch.cuda.ExternalStream(ps.stream, device=local_rank)

    for _ in range(warmup):
        ps.RestoreState()
        for _ in range(phases):
            ps.Solve()
    ps.SynchroniseDevice()
    ps.ResetStepCounters()

    # ---- value: device-resident ---------------------------------------------------------------
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps + 2)]
    ev_step = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    # L2 (126 MB) is flushed between timed iterations by overwriting a 256 MiB buffer on the solver's stream; the flush itself is
    # outside the event brackets that are summed into the step time
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")
    sampler = ClockSampler(local_rank)
    time.sleep(0.25)
    barrier(); torch.cuda.synchronize()
    t_host0 = time.time()
    ev[0].record(stream)
    for k in range(steps):
        with torch.cuda.stream(stream):
            flush.fill_(k & 0xff)
        ev_step[k].record(stream)
        ps.RestoreState()
        ev[2 + 2 * k].record(stream)
        for _ in range(phases):
            ps.Solve()
        ev[3 + 2 * k].record(stream)
    ev[1].record(stream)
    torch.cuda.synchronize(); barrier()
    t_host1 = time.time()
    clocks = sampler.stop(t_host0, t_host1)
    region_ms = sum(ev_step[k].elapsed_time(ev[3 + 2 * k]) for k in range(steps))      # K x (restore + phases), flushes excluded
    kernel_ms = [ev[2 + 2 * k].elapsed_time(ev[3 + 2 * k]) / phases for k in range(steps)]
    attempted, accepted = arm.total_steps()
    dense_samples = None
    if kind == "perthread" and wl.shape.NDO > 0:
        # points stored by the last phase (the built-in dense systems restart the index every phase) x phases per step
        ps.SynchroniseFromDeviceToHost(arm.V.DenseIndex)
        ps.SynchroniseDevice()
        dense_samples = int(np.minimum(ps.host(arm.V.DenseIndex), wl.shape.NDO).sum()) * phases
    del flush

    # ---- e2e: host buffers through the public API, double-buffered over `chunks` solver objects ----
    chunks = max(1, args.e2e_chunks)
    while N % chunks or (kind == "perthread" and N // chunks < 1):
        chunks -= 1
    if chunks == 1:
        e2e_arms = [arm]
    else:
        import copy
        sharding = importlib.import_module(pkg.__name__ + ".sharding")
        e2e_arms = []
        for c in range(chunks):
            w = copy.copy(wl)
            w.buffers = wl.buffers.take(sharding.shard_indices(N // chunks, c, chunks, 1))
            e2e_arms.append(Arm(pkg, w, kind, local_rank, args))
    h2d = d2h = 0
    for a in e2e_arms:
        hi, do = a.make_host_buffers(torch)
        h2d += hi
        d2h += do

    def step_e2e():
        # T7's loop: wait for a solver object (its previous sweep is then in the caller's arrays), refill it and launch it again,
        # while the other objects compute -- so the copies of one step also overlap the kernels of its neighbours
        for a in e2e_arms:
            a.finish_e2e()
            a.enqueue_e2e(phases)

    def drain_e2e():
        for a in e2e_arms:
            a.finish_e2e()

    for _ in range(2):
        step_e2e()
    drain_e2e()
    # the copies alone, all ranks at once: what the host side (memory controllers, PCIe root) gives each rank for this step's bytes
    barrier(); torch.cuda.synchronize()
    tc0 = time.perf_counter()
    for _ in range(3):
        for a in e2e_arms:
            for v, t in a.user_in.items():
                a.ps.Upload(v, t.data_ptr(), t.numel() * t.element_size())
            for v, t in a.user_out.items():
                a.ps.Download(v, t.data_ptr(), t.numel() * t.element_size())
            a.ps.InsertSynchronisationPoint()
        drain_e2e()
    torch.cuda.synchronize(); barrier()
    copy_gbs = 3 * (h2d + d2h) / (time.perf_counter() - tc0) / 1e9
    for a in e2e_arms:
        a.ps.ResetStepCounters()
    barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_e2e()
    drain_e2e()
    torch.cuda.synchronize(); barrier()
    e2e_s = time.perf_counter() - t0
    e2e_attempted = sum(a.total_steps()[0] for a in e2e_arms)
    final_state_checksum = float(sum(a.user_out[a.V.ActualState].double().sum() for a in e2e_arms))

    # ---- reduce over ranks -------------------------------------------------------------------------
    sharding = importlib.import_module(pkg.__name__ + ".sharding")
    (region_ms, e2e_s, kmean, neg_copy), counts = sharding.reduce_over_ranks([region_ms, e2e_s, float(np.mean(kernel_ms)), -copy_gbs],
                                                                             [attempted, accepted, e2e_attempted], device="cuda")
    attempted, accepted, e2e_attempted = [int(x) for x in counts]
    copy_gbs_min = -neg_copy       # the slowest rank's copy rate

    out = None
    if rank == 0:
        value = attempted / (region_ms * 1e-3)
        precision_bytes = 8 if precision == "f64" else 4
        peak_tf, peak_probes = measured_peak(pkg, local_rank, precision_bytes)
        per_rank_attempted_per_launch = attempted / world / steps / phases
        achieved_tf = per_rank_attempted_per_launch * flops_per_step / (kmean * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(f"{workload}_{precision}")
        # solver kernels of the timed region, plus the three ordering kernels in front of every phase that has a cost history
        # (second and later phases of a step; per-thread module: only launches of more than 16 blocks per SM, coupled module: queues
        # longer than two rounds of slots, and there the first phase too: pilot slice + ordered continuation)
        tpb = (args.threads_per_block or getattr(wl.options, "ThreadsPerBlock", None) or 128) if kind == "perthread" else 1
        tiles = (N + tpb - 1) // tpb
        ordered = (tiles > 16 * 148) if kind == "perthread" else (N > 2 * 148 * 8)
        if kind == "coupled":      # every phase is cost-ordered (3 kernels); the first one of a step after a pilot slice (one more solver launch)
            ordering_kernels = (3 * phases + 1) if ordered else 0
        else:
            ordering_kernels = 3 * (phases - 1) if (ordered and args.lane_refill <= 0) else 0
        out = {"metric": metric, "value": value, "unit": "system-steps/s", "n_gpus": world, "steps": steps, "warmup": warmup,
               "ms_per_step": region_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": precision, "data": "synthetic", "config": config,
               "accepted_steps_per_s": accepted / (region_ms * 1e-3),
               "system_steps_per_step": attempted / steps,
               "roofline": {"bound": "fp64_fma" if precision_bytes == 8 else "fp32_fma", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                            "frac": achieved_tf / peak_tf, "traffic": traffic,
                            "peak_source": "measured on this GPU in this process (~10 ms register-resident chains, 256 FMAs per loop trip, CUDA events): FP64 = max of the DFMA probe and the FP64 tensor-core (DMMA) probe, FP32 = FFMA probe; MEASURED_PEAKS.json carries no FP64/FP32 FMA figure. The issue-rate bound is one DFMA per 2 cycles per SM sub-partition = 32 flop/clk x 592 x SM clock (37.2 TFLOP/s at 1965 MHz)",
                            "peak_probes": peak_probes,
                            "kernel_ms": kmean, "algorithmic_flops_per_launch": per_rank_attempted_per_launch * flops_per_step,
                            "fp64_pipe_utilisation_ncu": NCU_PIPE.get(f"{workload}_{precision}")},
               "e2e": {"value": e2e_attempted / e2e_s, "unit": "system-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": e2e_s / steps * 1e3, "chunks": len(e2e_arms), "state_checksum": final_state_checksum,
                       "copies_alone_gbs_per_rank": copy_gbs_min,
                       "copies_alone_note": "this step's H2D + D2H bytes without any kernel, all ranks at the same time (slowest rank): the host side's share of the e2e time"},
               "gpu_launches": steps * (phases + ordering_kernels), "clocks": clocks}
        if dense_samples is not None:
            # the only HBM stream that matters on this path: (SD + 1) words per stored point (24 B for Keller-Miksis)
            bytes_per_sample = (wl.shape.SD + 1) * precision_bytes
            hbm_peak = None
            mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
            if os.path.exists(mp):
                hbm_peak = json.load(open(mp)).get("hbm_gbs")
            gbs = dense_samples * bytes_per_sample / (kmean * phases * 1e-3) / 1e9
            out["dense_output"] = {"stored_points_per_step_per_gpu": dense_samples, "bytes_per_point": bytes_per_sample, "achieved_gbs": gbs,
                                   "hbm_peak_gbs": hbm_peak, "frac": (gbs / hbm_peak) if hbm_peak else None,
                                   "note": "rank 0; algorithmic bytes of the dense-output stream over the kernel time -- the kernel is issue-bound (FP64), the stream is far from the HBM roof"}
        if world == 1 and not args.no_reference_gpu:
            try:
                r = reference_gpu_baseline(kind, wl, N, precision, phases, per_rank_attempted_per_launch, kmean)
                if r is not None:
                    out["reference_gpu"] = r
            except Exception as e:   # the baseline must never break the bench line
                out["reference_gpu"] = {"unavailable": str(e)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            big = threads >= 16
            if full:
                sample = args.cpu_sample or {"coupled": 512 if big else 128, "kellermiksis": 1 << 15, "kellermiksis_dense": 1 << 15,
                                             "duffing": 10000, "impact": 1 << 17 if big else 1 << 15}.get(workload, 1 << 19 if big else 1 << 17)
            else:       # the extra configurations: a couple of seconds each
                sample = {"coupled": 256 if big else 64, "coupled_fm": 256 if big else 64, "kellermiksis_dense": 1 << 14, "kellermiksis_dense_fm": 1 << 14,
                          "duffing": 10000, "impact": 1 << 15}.get(workload, 1 << 17)
            sample = min(sample, N)
            n, s, kind_cpu, nt = cpu_reference_run(pkg, workload, precision, sample, threads, phases)
            out["cpu_baseline"] = {"value": n / s, "unit": "system-steps/s", "cores": threads, "kind": kind_cpu,
                                   "sample": f"every {max(N // nt, 1)}-th {unit[:-1]} of the sweep ({nt} {unit}), {phases} phase(s), {threads} host threads, {s:.1f} s"}
    for a in e2e_arms:
        if a is not arm:
            a.close()
    arm.close()
    return out


# FP64-pipe activity of the solver kernel as ncu reports it (sm__inst_executed_pipe_fp64 / pipe_fp64_cycles_active, from the
# committed captures under profiles/; not measured live -- a number taken under a profiler is not a bench value): the honest
# utilisation figure for right-hand sides whose transcendentals count 0 algorithmic flops.
NCU_PIPE = {}
_np = os.path.join(ROOT, "profiles", "ncu_pipe_utilisation.json")
if os.path.exists(_np):
    try:
        NCU_PIPE = json.load(open(_np))
    except Exception:
        NCU_PIPE = {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="lorenz", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--instances", type=int, default=0, help="instances (systems) per GPU (default: the BASELINE size)")
    ap.add_argument("--phases", type=int, default=1, help="integration phases (Solve() calls) per step")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="solver objects the e2e sweep is double-buffered over")
    ap.add_argument("--lane-refill", type=int, default=-1)
    ap.add_argument("--threads-per-block", type=int, default=0)
    ap.add_argument("--tensor-core-coupling", type=int, default=-1)
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true")
    ap.add_argument("--configs", default="auto", help="'auto': after the default headline workload also run the other BASELINE configurations (C1, C2-FP32, C3, C4, C5) "
                                                      "and report them in the line's `configs` object; 'none'; or a comma list of their keys")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    pkg = importlib.import_module("massively-parallel-gpu-ode-solver_b200")
    kind, cfg_name, n_default, flops_per_step, alg = WORKLOADS[args.workload]
    N = args.instances or n_default
    unit = "systems" if kind == "coupled" else "instances"
    metric = "system-steps/sec (%s %s)" % ("FP64" if args.precision == "f64" else "FP32", alg)

    # ------------------------------------------------------------------ reference arm (CPU) ------
    if args.impl == "reference":
        if rank != 0:
            return
        config = {"workload": f"{args.workload}: {N} {unit} per GPU x {world} GPU(s), {args.phases} integration phase(s) per step (BASELINE {cfg_name})",
                  f"{unit}_per_gpu": N, "phases_per_step": args.phases, "flops_per_system_step": flops_per_step, "flops_convention": FLOPS_NOTE}
        threads = os.cpu_count() or 1
        sample = args.cpu_sample or {"coupled": 256, "kellermiksis": 1 << 14, "kellermiksis_dense": 1 << 14, "duffing": 10000}.get(args.workload, 1 << 16)
        sample = min(sample, N)
        for _ in range(args.warmup):
            cpu_reference_run(pkg, args.workload, args.precision, max(sample // 8, 32), threads, args.phases)
        steps_total, secs, kind_cpu, nt = 0, 0.0, "reference", sample
        for _ in range(args.steps):
            n, s, kind_cpu, nt = cpu_reference_run(pkg, args.workload, args.precision, sample, threads, args.phases)
            steps_total += n
            secs += s
        v = steps_total / secs
        print(json.dumps({"impl": "reference", "metric": metric, "value": v, "unit": "system-steps/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "system-steps/s", "cores": threads, "kind": kind_cpu,
                                           "sample": f"every {max(N // nt, 1)}-th {unit[:-1]} of the sweep ({nt} {unit}) per step, {threads} host threads"},
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

    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    out = measure(pkg, args, args.workload, args.precision, N, args.phases, args.steps, args.warmup, rank, local_rank, world, full=True)
    if rank == 0 and numa is not None:
        out["host_binding"] = numa

    # ---- the other BASELINE configurations, after the headline (whose timing they cannot disturb) ----
    default_headline = (args.workload == "lorenz" and args.precision == "f64" and not args.instances and args.phases == 1
                        and args.lane_refill < 0 and not args.threads_per_block)
    wanted = [] if args.configs == "none" else ([c for c in EXTRA_CONFIGS] if args.configs == "auto" else [c for c in EXTRA_CONFIGS if c[0] in args.configs.split(",")])
    if args.configs == "auto" and not default_headline:
        wanted = []
    configs = {}
    if wanted and world > 1:
        # strong scaling of the headline sweep: its 2^20 instances split over the GPUs (short kernels: launch latency and tails show)
        wanted = wanted + [("C2_lorenz_rkck45_f64_strong_scaling", "lorenz", "f64", 1)]
    for key, wname, prec, phases in wanted:
        strong = key.endswith("_strong_scaling")
        try:
            r = measure(pkg, args, wname, prec, WORKLOADS[wname][2] // (world if strong else 1), phases, min(args.steps, 5), 3, rank, local_rank, world, full=False)
            if strong and r is not None:
                r["scaling"] = "strong"
        except Exception as e:      # an extra configuration must never break the headline line
            r = {"error": f"{type(e).__name__}: {e}"[:300]}
            if world > 1:
                raise
        if rank == 0 and r is not None:
            configs[key] = r
    if rank == 0:
        if wanted:
            out["configs"] = configs
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
