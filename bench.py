#!/usr/bin/env python
"""bench.py -- headline benchmark of the ScaML-GP hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the fused LML+grad kernel over one batch of synthetic meta-tasks
(config 3 of BASELINE.json: M=4096 tasks x n=256 x d=6 per GPU, R=6 hyper-parameter rows per
task = 1 warm start + 5 prior samples, reference scamlgp/utils.py:173-203).  `value` is
evaluations/s with inputs resident in HBM, `e2e` is the same metric through the public
engine call with pinned HOST buffers (H2D of X, Y, theta and D2H of lml, grad inside the
timed region).  Multi-GPU: tasks are independent -> every rank owns its own 4096 tasks
(weak scaling), no data-path collective; timing is the max over ranks of device time.

`--impl reference` times the reference's CPU path (the oracle restatement of its
botorch/gpytorch arithmetic -- those packages cannot be installed here) on the
box's host cores in the reference's own parallel style: one single-threaded process per
core over disjoint task slices (scamlgp/benchmarking/local_runner.py:107-108,174-181).
"""
from __future__ import annotations

import argparse
import math
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

M_TASKS, R_ROWS, N_PTS, DIM = 4096, 6, 256, 6
WORKLOAD = "config3: 4096 meta-tasks x n=256 x d=6, R=6 hyper-parameter rows/task, LML+grad (per GPU)"


def algorithmic_flops(n: int, d: int) -> float:
    """F_lml = n^3 + n^2 (5d/2 + 10) flop per evaluation."""
    return float(n) ** 3 + float(n) ** 2 * (2.5 * d + 10.0)


# --------------------------------------------------------------------------------------- #
# CPU side (oracle) -- used for cpu_baseline and for --impl reference
# --------------------------------------------------------------------------------------- #
_CPU_JOB = {}


def _cpu_worker(args):
    lo, hi = args
    import torch

    from oracle import scaml_oracle as O

    torch.set_num_threads(1)
    X, Yt, th, spec = _CPU_JOB["X"], _CPU_JOB["Yt"], _CPU_JOB["th"], _CPU_JOB["spec"]
    R = th.shape[1]
    t0 = time.perf_counter()
    acc = 0.0
    for e in range(lo, hi):
        m, r = divmod(e, R)
        v, g = O.lml_and_grad_autograd(X[m], Yt[m], th[m, r], spec, mode="expansion")
        acc += float(v)
    return hi - lo, time.perf_counter() - t0, acc


def _cpu_post_worker(args):
    """Reference-style posterior of the source GPs: Python loop over tasks (scamlgp/model.py:128-134), each a
    kernel row block, an alpha dot and a triangular solve for the variance."""
    lo, hi = args
    import torch

    from oracle import scaml_oracle as O

    torch.set_num_threads(1)
    states, Xc = _CPU_JOB["states"], _CPU_JOB["Xc"]
    t0 = time.perf_counter()
    acc = 0.0
    for m in range(lo, hi):
        mu, var = O.posterior(states[m], Xc)
        acc += float(mu[0]) + float(var[0])
    return (hi - lo) * Xc.shape[0], time.perf_counter() - t0, acc


class CpuReference:
    """Reference-style CPU evaluation of LML+grad on a bounded sample of the workload."""

    def __init__(self, tasks: int, cores: int):
        import torch

        from oracle import scaml_oracle as O

        torch.set_num_threads(1)
        self.cores = cores
        self.tasks = tasks
        X, Y = O.synthetic_tasks(tasks, N_PTS, DIM, seed=0)
        spec = O.HyperSpec.source()
        Yt = torch.stack([O.standardize(Y[m])[0] for m in range(tasks)])
        th = O.sample_theta_raw(tasks, R_ROWS, DIM, spec, seed=0)
        g = torch.Generator().manual_seed(5)
        _CPU_JOB.update(X=X, Yt=Yt, th=th, spec=spec,
                        states=[O.factorize(X[m], Y[m], th[m, 0], spec) for m in range(tasks)],
                        Xc=torch.rand(2048, DIM, dtype=torch.float64, generator=g))
        import multiprocessing as mp

        self.pool = mp.get_context("fork").Pool(cores)
        self.evals = tasks * R_ROWS

    def step(self) -> float:
        """Evaluates the whole sample once across the pool; returns wall seconds."""
        per = (self.evals + self.cores - 1) // self.cores
        chunks = [(i, min(i + per, self.evals)) for i in range(0, self.evals, per)]
        t0 = time.perf_counter()
        self.pool.map(_cpu_worker, chunks)
        return time.perf_counter() - t0

    def posterior_points_per_s(self, candidates: int = 2048, passes: int = 2) -> dict:
        """points/s of the per-task posterior (mean + variance) on the same sample of tasks."""
        per = (self.tasks + self.cores - 1) // self.cores
        chunks = [(i, min(i + per, self.tasks)) for i in range(0, self.tasks, per)]
        self.pool.map(_cpu_post_worker, chunks)  # warm-up
        t0 = time.perf_counter()
        for _ in range(passes):
            self.pool.map(_cpu_post_worker, chunks)
        dt = time.perf_counter() - t0
        return {"value": self.tasks * candidates * passes / dt, "unit": "points/s", "cores": self.cores, "kind": "port",
                "sample": f"{self.tasks} fitted tasks (n={N_PTS}, d={DIM}) x {candidates} candidates, oracle posterior "
                          f"mean+variance per task, {self.cores} single-threaded processes, {passes} timed passes"}

    def close(self):
        self.pool.close()
        self.pool.join()


def host_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_sample_tasks(cores: int) -> int:
    # ~6 evaluations/task; aim at roughly 2-3 s of wall per step with ~5 ms per evaluation
    return max(8, min(M_TASKS, cores * 128 // R_ROWS))


# --------------------------------------------------------------------------------------- #
# clocks
# --------------------------------------------------------------------------------------- #
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def rd():
            for line in self.proc.stdout:
                self.rows.append(line.strip())

        self.thread = threading.Thread(target=rd, daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------- #
def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    tasks = cpu_sample_tasks(cores)
    ref = CpuReference(tasks, cores)
    for _ in range(args.warmup):
        ref.step()
    times = [ref.step() for _ in range(args.steps)]
    ref.close()
    total = sum(times)
    value = ref.evals * args.steps / total
    sample = (f"{tasks} of {M_TASKS} tasks x R={R_ROWS} rows per step, oracle LML + autograd gradient "
              f"(gpytorch-style expansion distances), {cores} single-threaded processes")
    line = {
        "impl": "reference", "metric": "meta-task LML+grad evals/s (n=256,d=6)", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def ncu_traffic(kernel: str):
    """DRAM bytes (read + write) per launch of `kernel` from the committed ncu capture (profiles/ncu_traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)[kernel]
        return float(t["dram_bytes_read"]) + float(t["dram_bytes_write"])
    except Exception:
        return None


def measure_fp64_peak(torch, lib_path: str, device) -> dict:
    """Live FP64 denominators: register-resident DFMA chains and DMMA chains."""
    import ctypes as C

    L = C.CDLL(lib_path)
    L.scaml_microbench_fp64.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_void_p]
    blocks = 148 * 8
    out = torch.empty(blocks * 256, dtype=torch.float64, device=device)
    res = {}
    stream = torch.cuda.current_stream(device).cuda_stream
    for mode, name, iters in ((0, "dfma", 4000), (1, "dmma_m8n8k4", 4000), (2, "dmma_m16n8k16", 2000)):
        flops = C.c_double(0.0)
        best = 0.0
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = L.scaml_microbench_fp64(mode, blocks, iters, out.data_ptr(), C.byref(flops), stream)
            e1.record()
            torch.cuda.synchronize(device)
            if rc != 0:
                break
            ms = e0.elapsed_time(e1)
            if rep > 0:
                best = max(best, flops.value / (ms * 1e-3) / 1e12)
        res[name] = best
    # library DGEMM for reference ("quote roofline fractions of measured DGEMM / DFMA")
    try:
        n = 8192
        a = torch.randn(n, n, dtype=torch.float64, device=device)
        b = torch.randn(n, n, dtype=torch.float64, device=device)
        torch.matmul(a, b)
        best = 0.0
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize(device)
            best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        res["cublas_dgemm_8192"] = best
        del a, b
    except Exception:
        res["cublas_dgemm_8192"] = None
    return res


def run_ours(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = host_cores()

    # CPU baseline first (fork-based pool must be created before CUDA is initialised)
    cpu_baseline = cpu_posterior = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        tasks = cpu_sample_tasks(cores)
        ref = CpuReference(tasks, cores)
        ref.step()  # warm-up
        ts = [ref.step() for _ in range(3)]
        cpu_posterior = None if args.no_posterior else ref.posterior_points_per_s()
        ref.close()
        cpu_baseline = {
            "value": ref.evals * len(ts) / sum(ts), "unit": "evals/s", "cores": cores, "kind": "port",
            "sample": f"{tasks} of {M_TASKS} tasks x R={R_ROWS}, oracle LML+autograd grad (expansion distances), "
                      f"{cores} single-threaded processes, {len(ts)} timed passes",
        }

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device; there is no CPU fallback. "
                         "Use --impl reference for the CPU baseline.")
    torch.cuda.set_device(local_rank)
    device = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    from oracle import scaml_oracle as O  # synthetic data generator only (inputs), not the measured path
    from scamlgp_b200 import HyperSpec
    from scamlgp_b200._capi import CUDA_LIB_PATH
    from scamlgp_b200.engine import Engine, SourceBatch

    eng = Engine(device)
    spec = HyperSpec.source()
    M, R, n, d = M_TASKS, R_ROWS, N_PTS, DIM
    X, Y = O.synthetic_tasks(M, n, d, seed=rank)  # every rank owns its own 4096 tasks
    th = O.sample_theta_raw(M, R, d, O.HyperSpec.source(), seed=rank)
    hX, hY, hT = X.pin_memory(), Y.pin_memory(), th.contiguous().pin_memory()
    batch = SourceBatch.from_padded(hX.to(device), hY.to(device))
    thd = hT.to(device).contiguous()
    lml = torch.empty(M, R, dtype=torch.float64, device=device)
    grad = torch.empty(M, R, d + 2, dtype=torch.float64, device=device)
    info = torch.empty(M, R, dtype=torch.int32, device=device)
    h_lml = torch.empty(M, R, dtype=torch.float64).pin_memory()
    h_grad = torch.empty(M, R, d + 2, dtype=torch.float64).pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=device)  # 256 MB > 126 MB L2

    peaks = measure_fp64_peak(torch, CUDA_LIB_PATH, device) if rank == 0 else {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def step_resident():
        eng.lml_grad_raw(batch, thd, spec, out=(lml, grad, info))

    def step_e2e():
        b = SourceBatch.from_padded(hX.to(device, non_blocking=True), hY.to(device, non_blocking=True))
        t = hT.to(device, non_blocking=True)
        l, g, i = eng.lml_grad_raw(b, t, spec, out=(lml, grad, info))
        h_lml.copy_(l, non_blocking=True)
        h_grad.copy_(g, non_blocking=True)
        torch.cuda.current_stream(device).synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        launches0 = eng.launches
        evs = []
        for _ in range(steps):
            flush.zero_()  # evict L2 between timed iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), eng.launches - launches0

    # ---- second half of BASELINE.json's metric: posterior points/s over all tasks (config 5 shape) ------ #
    def posterior_bench():
        """Weighted ScaML-GP prior prediction (mean + variance, reduced over the rank's 4096 fitted base GPs
        inside the kernel) at B candidates; one "point" = one (task, candidate) pair.  Config 5 asks for
        B = 1 Mi; a step here is a bounded slice of it (B_STEP candidates = 2 waves of 148 x 64-candidate
        tiles) so the default run stays short -- points/s does not depend on B beyond one wave."""
        B_STEP = 2 * 148 * 64
        fs = eng.factorize(batch, thd[:, 0].contiguous(), spec)
        w = torch.full((M,), 1.0 / M, dtype=torch.float64, device=device)
        g = torch.Generator().manual_seed(100 + rank)
        hXc = torch.rand(B_STEP, d, dtype=torch.float64, generator=g).pin_memory()
        Xc = hXc.to(device)
        pm = torch.empty(B_STEP, dtype=torch.float64, device=device)
        pv = torch.empty(B_STEP, dtype=torch.float64, device=device)
        part = torch.empty(2, B_STEP, dtype=torch.float64, device=device)
        h_out = torch.empty(2, B_STEP, dtype=torch.float64).pin_memory()

        def step_res():
            eng.predict_weighted(fs, w, Xc, out=(pm, pv))

        def step_e2e():
            xc = hXc.to(device, non_blocking=True)
            eng.predict_weighted(fs, w, xc, out=(part[0], part[1]))
            if world > 1:  # sum of weighted predictions over the task shards (north_star item 5)
                dist.all_reduce(part, op=dist.ReduceOp.SUM)
            h_out.copy_(part, non_blocking=True)
            torch.cuda.current_stream(device).synchronize()

        psteps = max(2, min(args.steps, 4))
        ms_r, nl = timed(step_res, psteps, 3)
        ms_e, _ = timed(step_e2e, psteps, 1)
        # conditioned variant (n_t = 32 target points): prior mean / variance + cross-covariance with the target
        # inputs in one fused prediction launch; A_m is prepared once per set of target inputs
        n_t = 32
        Xt = torch.rand(n_t, d, dtype=torch.float64, generator=g).to(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.cond_prepare(fs, Xt)
        e0.record()
        A = eng.cond_prepare(fs, Xt)
        e1.record()
        torch.cuda.synchronize(device)
        ms_prep = e0.elapsed_time(e1)
        ms_c, _ = timed(lambda: eng.predict_conditioned(fs, w, Xc, Xt, A), 2, 1)
        # acquisition-gradient leg: value + analytic d mean/dx, d var/dx of the conditioned posterior at
        # 64 candidates (one L-BFGS-B function evaluation of the acquisition optimiser over 64 restarts)
        Bg = 64
        Yt = torch.sin(3.0 * Xt).sum(1)
        Yall = torch.cat([batch.Y_raw.reshape(-1), Yt])
        mu_a, s_a = float(Yall.mean()), float(Yall.std())
        smn, scv = eng.cond_caches(fs, Xt, A)
        from scamlgp_b200 import HyperSpec as _HS
        tsp = _HS.target()

        def _raw(v, lo, hi):  # inverse of the Interval (sigmoid) constraint at the reference's initial values
            q = (v - lo) / (hi - lo)
            return math.log(q / (1.0 - q))

        tht = torch.tensor([_raw(tsp.ls_init, *tsp.ls_bounds)] * d + [_raw(tsp.os_init, *tsp.os_bounds),
                                                                      _raw(tsp.noise_init, *tsp.noise_bounds)],
                           dtype=torch.float64, device=device)
        tstate = eng.target_factorize(smn, scv, Xt, ((Yt - mu_a) / s_a).contiguous(), w, tht, mu_a, s_a, tsp)
        del smn, scv
        Xg = Xc[:Bg].contiguous()

        def step_grad():
            U = eng.cond_prepare(fs, Xg)
            a_, b_, c_ = eng.values_from_u(fs, w, Xg, U, Xt, A)
            _, _, beta = eng.target_posterior_beta(tstate, a_, b_, c_, Xg)
            return eng.posterior_grad(fs, w, Xg, U, tstate, A, beta)

        ms_g, nl_g = timed(step_grad, 3, 2)
        gdm, gdv = step_grad()
        grad_finite = bool(torch.isfinite(gdm).all() and torch.isfinite(gdv).all()) and tstate.info == 0
        del A
        ok_info = int((fs.info == 0).all())
        finite = bool(torch.isfinite(pm).all() and torch.isfinite(pv).all() and (pv > 0).all())
        del fs
        pts = float(M) * B_STEP * world
        Fp = float(n) ** 2 + float(n) * (3.0 * d + 12.0)  # flop per (task, candidate) point
        ach = float(M) * B_STEP * Fp / (ms_r / psteps * 1e-3) / 1e12
        return {"metric": "posterior points/s over all tasks", "value": pts * psteps / (ms_r * 1e-3),
                "unit": "points/s", "ms_per_step": ms_r / psteps, "steps": psteps,
                "config": {"workload": f"config5 slice: {M} fitted base GPs (n={n}, d={d}) per GPU x {B_STEP} "
                                       "candidates per step, weighted mean+variance (q=1)",
                           "candidates_per_step": B_STEP, "factor_info_zero": bool(ok_info), "finite": finite},
                "e2e": {"value": pts * psteps / (ms_e * 1e-3), "unit": "points/s",
                        "h2d_bytes_per_step": int(hXc.numel()) * 8, "d2h_bytes_per_step": int(h_out.numel()) * 8,
                        "collective": "all_reduce(sum) [2,B] fp64 over ranks" if world > 1 else None},
                "gpu_launches": nl,
                "conditioned": {"n_t": n_t, "value": pts * 2 / (ms_c * 1e-3), "unit": "points/s",
                                "ms_per_step": ms_c / 2, "prepare_ms": ms_prep,
                                "what": "weighted prior mean/variance + cross-covariance with n_t target inputs "
                                        "(fused), all ranks"},
                "acq_grad": {"n_t": n_t, "candidates": Bg, "ms_per_step": ms_g / 3, "gpu_launches": nl_g // 3,
                             "value": float(M) * Bg * world * 3 / (ms_g * 1e-3), "unit": "(task, candidate) gradients/s",
                             "finite": grad_finite,
                             "what": "posterior value + analytic d mean/dx, d var/dx (conditioned on n_t target points): "
                                     "cond_prepare at the candidates, values from U, beta, DMMA mix, gradient contraction"},
                "roofline": {"bound": "tensor", "pipe": "FP64 tensor cores (DMMA)", "achieved": ach, "unit": "TFLOP/s", "flops_per_point": Fp,
                             "traffic": ncu_traffic("scaml_predict_kernel<RBF>"),
                             "traffic_unit": "DRAM bytes per launch (ncu)",
                             "algorithmic_bytes": 8.0 * (M * (n * n / 2 + n * d + n) * 2 + B_STEP * (d + 2)),
                             "kernel": "scaml_predict_kernel<RBF>"}}

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_res, launches = timed(step_resident, args.steps, max(3, args.warmup))
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, _ = timed(step_e2e, max(2, min(args.steps, 5)), 2)
    e2e_steps = max(2, min(args.steps, 5))
    ok = int((info == 0).all())
    posterior = None if args.no_posterior else posterior_bench()
    if world > 1:
        okt = torch.tensor([ok], device=device)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        ok = int(okt.item())

    if rank == 0:
        evals_per_step = M * R * world
        value = evals_per_step * args.steps / (ms_res * 1e-3)
        e2e_value = evals_per_step * e2e_steps / (ms_e2e * 1e-3)
        F = algorithmic_flops(n, d)
        kernel_ms = ms_res / args.steps  # one launch per step: the step IS the dominant kernel
        achieved = M * R * F / (kernel_ms * 1e-3) / 1e12  # per GPU
        peak = peaks.get("dfma", 0.0) or None
        line = {
            "metric": "meta-task LML+grad evals/s (n=256,d=6)", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_res / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "tasks_per_gpu": M, "rows_per_task": R, "n": n, "d": d,
                       "l2": "256 MB flush buffer written between timed iterations",
                       "parallelism": f"task-sharded x{world}, no data-path collective", "all_info_zero": bool(ok)},
            "e2e": {"value": e2e_value, "unit": "evals/s",
                    "h2d_bytes_per_step": int(hX.numel() + hY.numel() + hT.numel()) * 8,
                    "d2h_bytes_per_step": int(h_lml.numel() + h_grad.numel()) * 8, "steps": e2e_steps},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "pipe": "FP64 tensor cores (mma.sync f64 -> DMMA; shares its pipe with DFMA)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": (achieved / peak) if peak else None,
                         "traffic": ncu_traffic("scaml_fit_kernel<RBF>"), "traffic_unit": "DRAM bytes per launch (ncu)",
                         "algorithmic_bytes": 8.0 * M * (n * d + n + R * (2 * (d + 2) + 1)),
                         "kernel": "scaml_fit_kernel<RBF>", "flops_per_eval": F,
                         "peak_source": "live register-resident DFMA microbench in this run "
                                        "(MEASURED_PEAKS.json has no FP64 entry)",
                         "fp64_peaks_tflops": peaks},
            "clocks": clocks,
        }
        if posterior is not None:
            pk = peak
            posterior["roofline"]["peak"] = pk
            posterior["roofline"]["frac"] = (posterior["roofline"]["achieved"] / pk) if pk else None
            if cpu_posterior is not None:
                posterior["cpu_baseline"] = cpu_posterior
            line["posterior"] = posterior
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-posterior", action="store_true", help="skip the posterior points/s measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
