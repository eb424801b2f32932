#!/usr/bin/env python
"""bench.py -- the reference's headline workload on B200: batched log-marginal-likelihood + gradient
evaluations per second (BASELINE.json metric, config 3: theta batch x M=512, 2-D SE kernel with first-derivative
observations, synthetic data of SURVEY.md section 8d).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of B = 4096 hyperparameter vectors per GPU (assembly ->
Cholesky -> alpha / ll -> explicit inverse -> fused gradient), i.e. one launch of the persistent batched kernel.
  value  whole-job evals/s with the theta batch already resident in HBM, timed with CUDA events on the
         launching stream, max over ranks; for N > 1 the timed region also contains the single NCCL all-gather
         of the ll / gradient scalars.  Weak scaling: every rank owns its own 4096-theta batch.
  e2e    the same metric through the public API (GaussianProcess.update_hyperparameters_batch) with HOST
         buffers: pinned theta -> device and ll / grad / status -> host inside the timed region.
  roofline  FP64 tensor pipe: B * M^3 algorithmic flop per launch / measured launch time, against the cuBLAS
         Dgemm throughput measured in this very run (MEASURED_PEAKS.json carries no FP64 figure).
  cpu_baseline  the numpy/scipy restatement of the reference's algorithm (oracle/gp_oracle.py, "port") timed on
         the host cores on a bounded sample of the same workload.
--impl reference times that CPU path alone, in the reference's own parallel mode (one theta per worker
process, gaussian_process.py:723-735), with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_OBS = 512
B_PER_GPU = 4096
METRIC = "log-ML+grad evals/sec (batched theta, N=512)"
UNIT = "evals/s"
FLOP_PER_EVAL = float(M_OBS) ** 3  # potrf M^3/3 + explicit inverse 2M^3/3


def c3_problem():
    """Config-3 synthetic data, exact generation order of SURVEY.md section 8(d)."""
    from numpy.random import RandomState
    rs = RandomState(0)
    Xv = rs.rand(256, 2)
    Xd1 = rs.rand(128, 2)
    Xd2 = rs.rand(128, 2)

    def f(x):
        return np.sin(3 * x[:, 0]) * np.cos(2 * x[:, 1])
    yv = f(Xv) + 0.05 * rs.randn(256)
    y1 = 3 * np.cos(3 * Xd1[:, 0]) * np.cos(2 * Xd1[:, 1]) + 0.05 * rs.randn(128)
    y2 = -2 * np.sin(3 * Xd2[:, 0]) * np.sin(2 * Xd2[:, 1]) + 0.05 * rs.randn(128)
    X = np.vstack([Xv, Xd1, Xd2])
    n = np.vstack([np.zeros((256, 2), dtype=int), np.tile([1, 0], (128, 1)), np.tile([0, 1], (128, 1))])
    y = np.concatenate([yv, y1, y2])
    err = np.full(512, 0.05)
    return X, n, y, err


def theta_batch(B, seed=1):
    from numpy.random import RandomState
    return np.array([1.0, 0.3, 0.4]) * np.exp(0.1 * RandomState(seed).randn(B, 3))


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, gpu_index):
        self.rows = []
        self.stop = threading.Event()
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())
            if self.stop.is_set():
                break

    def finish(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.stop.set()
        try:
            self.proc.terminate()
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            busy = sorted(sm)[len(sm) // 2:]          # median of the upper half = clocks under load
            out["sm_mhz"] = float(np.median(busy))
            out["sm_max_mhz"] = float(max(mx))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------------------------------------
# CPU legs (oracle = test infrastructure; only timed here, never part of the product path)
# ----------------------------------------------------------------------------------------------------------
def _cpu_eval(args):
    from oracle import gp_oracle as orc
    X, n, y, err, th = args
    r = orc.compute_K_L_alpha_ll(orc.KERNEL_SE, th, X, n, y, err, grad_idx=[0, 1, 2])
    return r["ll"]


def cpu_baseline_single_process(nsample):
    """Oracle port in ONE process (numpy/scipy threads = all cores), nsample thetas after one warm-up."""
    X, n, y, err = c3_problem()
    th = theta_batch(nsample + 1)
    _cpu_eval((X, n, y, err, th[0]))
    t0 = time.perf_counter()
    for b in range(1, nsample + 1):
        _cpu_eval((X, n, y, err, th[b]))
    dt = time.perf_counter() - t0
    return nsample / dt, dt


def run_reference_arm(args, rank):
    """--impl reference: the reference's CPU algorithm (oracle port) in its own parallel mode, all host threads."""
    if rank != 0:
        return
    import multiprocessing as mp
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    cores = host_threads()
    X, n, y, err = c3_problem()
    per_step = cores                      # bounded sample: one theta per worker per step
    th = theta_batch(per_step * (args.steps + args.warmup))
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pos = 0
        for _ in range(args.warmup):
            pool.map(_cpu_eval, [(X, n, y, err, t) for t in th[pos:pos + per_step]])
            pos += per_step
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_eval, [(X, n, y, err, t) for t in th[pos:pos + per_step]])
            pos += per_step
        dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = "%d thetas per step (one per worker process) of the %d-theta config-3 batch, M=512, ll+grad" % (
        per_step, B_PER_GPU)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "config 3: batched ll+grad, SE 2-D kernel, M=512 obs (256 values + 2x128 first "
                               "derivatives), P=3 free params; CPU sample of the theta batch",
                   "per_step_thetas": per_step, "parallelism": "%d worker processes" % cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import warnings
    import torch
    import torch.distributed as dist
    import gptools_b200 as g

    warnings.simplefilter("ignore")
    torch.cuda.set_device(local_rank)
    dev_t = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=dev_t)

    X, n, y, err = c3_problem()
    k = g.SquaredExponentialKernel(num_dim=2, initial_params=[1.0, 0.3, 0.4], param_bounds=[(0, 10)] * 3)
    gp = g.GaussianProcess(k, X=X, y=y, err_y=err, n=n, use_hyper_deriv=True, device=local_rank)
    B = B_PER_GPU
    th = theta_batch(B, seed=1 + rank)                  # every rank owns its own batch (weak scaling)
    dev, _ = gp._sync_device()
    stream = torch.cuda.Stream(device=dev_t)          # the library launches on THIS stream; events are recorded on it
    torch.cuda.set_stream(stream)
    dev.set_stream(stream.cuda_stream)

    # ---- device-resident leg -------------------------------------------------------------------------
    th_full = np.hstack([th, np.zeros((B, 1))])         # kernel params + sigma_noise (ZeroKernel)
    d_th = torch.from_numpy(th_full).to(dev_t)
    d_ll = torch.empty(B, dtype=torch.float64, device=dev_t)
    d_grad = torch.empty((B, 3), dtype=torch.float64, device=dev_t)
    d_st = torch.empty(B, dtype=torch.int32, device=dev_t)
    d_pack = torch.empty((B, 4), dtype=torch.float64, device=dev_t)
    d_all = torch.empty((world * B, 4), dtype=torch.float64, device=dev_t) if distributed else None
    grad_idx = [0, 1, 2]

    def step_device():
        dev.ll_batched_dev(B, d_th.data_ptr(), d_ll.data_ptr(), d_st.data_ptr(), d_grad=d_grad.data_ptr(),
                           grad_idx=grad_idx)
        if distributed:
            d_pack[:, 0] = d_ll
            d_pack[:, 1:] = d_grad
            dist.all_gather_into_tensor(d_all, d_pack)

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = dev.launch_count()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    launches = dev.launch_count() - launches0
    ok = bool((d_st == 0).all().item()) and bool(torch.isfinite(d_ll).all().item())

    # kernel-only time (the dominant kernel is the only one in the step at N = 1)
    k0 = torch.cuda.Event(enable_timing=True)
    k1 = torch.cuda.Event(enable_timing=True)
    k0.record(stream)
    for _ in range(args.steps):
        dev.ll_batched_dev(B, d_th.data_ptr(), d_ll.data_ptr(), d_st.data_ptr(), d_grad=d_grad.data_ptr(),
                           grad_idx=grad_idx)
    k1.record(stream)
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / args.steps

    # ---- end-to-end leg: public API, host buffers -----------------------------------------------------
    th_pinned = torch.empty((B, 3), dtype=torch.float64).pin_memory()
    th_pinned.copy_(torch.from_numpy(th))
    th_host = th_pinned.numpy()
    for _ in range(max(1, args.warmup // 2)):
        gp.update_hyperparameters_batch(th_host, with_deriv=True)
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        neg_ll, neg_grad = gp.update_hyperparameters_batch(th_host, with_deriv=True)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clock_info = clocks.finish() if rank == 0 else None
    h2d = B * 4 * 8
    d2h = B * 8 + B * 3 * 8 + B * 4
    ok = ok and bool(np.isfinite(neg_ll).all())

    # ---- FP64 roofline denominator measured live (cuBLAS Dgemm through torch) -------------------------
    peak_tflops = None
    if rank == 0:
        a = torch.randn(8192, 8192, dtype=torch.float64, device=dev_t)
        b = torch.randn(8192, 8192, dtype=torch.float64, device=dev_t)
        torch.matmul(a, b)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            p0 = torch.cuda.Event(enable_timing=True)
            p1 = torch.cuda.Event(enable_timing=True)
            p0.record()
            torch.matmul(a, b)
            p1.record()
            torch.cuda.synchronize()
            best = min(best, p0.elapsed_time(p1))
        peak_tflops = 2.0 * 8192 ** 3 / best * 1e-9
        del a, b

    # ---- max over ranks ------------------------------------------------------------------------------------
    times = torch.tensor([ms_total, e2e_s * 1e3, kernel_ms], dtype=torch.float64, device=dev_t)
    if distributed:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, kernel_ms = (float(v) for v in times.cpu())

    if rank == 0:
        value = world * B * args.steps / (ms_total * 1e-3)
        e2e_value = world * B * args.steps / (e2e_ms * 1e-3)
        achieved = B * FLOP_PER_EVAL / (kernel_ms * 1e-3) * 1e-12
        cpu_val, cpu_dt = cpu_baseline_single_process(24)
        traffic = None
        tf = os.path.join(ROOT, "profiles", "r01_batched_kernel_traffic.json")
        if os.path.exists(tf):
            try:
                traffic = json.load(open(tf)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config 3: batched ll+grad, SE 2-D kernel, M=512 obs (256 values + 2x128 first "
                                   "derivatives), P=3 free params",
                       "thetas_per_gpu": B, "global_batch": B * world, "parallelism": "theta-sharded x%d" % world,
                       "l2": "per-CTA factor workspace 592 x 1.44 MB = 853 MB > 126 MB L2 (no flush needed)",
                       "results_ok": ok},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved / peak_tflops if peak_tflops else None, "traffic": traffic,
                         "kernel": "ll_batched4_kernel", "kernel_ms": kernel_ms,
                         "flop_per_launch": B * FLOP_PER_EVAL,
                         "peak_source": "cuBLAS Dgemm 8192^3 fp64 measured live in this run (MEASURED_PEAKS.json "
                                        "has no FP64 entry); DMMA issue peak 37.1 TFLOP/s (profiles/r01_fp64_peak_microbench.txt)"},
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": host_threads(), "kind": "port",
                             "sample": "24 thetas of the batch, one process, numpy/scipy threads on all cores (%.1f s)" % cpu_dt},
            "clocks": clock_info,
        }
        print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
