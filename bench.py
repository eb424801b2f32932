#!/usr/bin/env python
"""bench.py -- the reference's headline workload on B200: batched log-marginal-likelihood + gradient
evaluations per second (BASELINE.json metric, config 3: 4096 theta x M=512, 2-D SE kernel with first-derivative
observations, synthetic data of SURVEY.md section 8d), "sharded by theta across 1/2/4/8 B200".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-secondary]

A "step" is one pass of the hot path over the GLOBAL batch of B = 4096 hyperparameter vectors (assembly -> Cholesky
-> alpha / ll -> explicit inverse -> fused gradient): the batch is split contiguously by rank (STRONG scaling, the
named configuration), every rank runs one launch of the persistent batched kernel on its 4096 / N thetas, and one
NCCL all-gather of the ll / gradient / status words follows (the kernel writes straight into the send buffer).
  value  whole-job evals/s with the theta batch already resident in HBM, timed with CUDA events on the launching
         stream, max over ranks, the all-gather inside the timed region.
  e2e    the same metric through the product's multi-GPU API, gptools_b200.parallel.update_hyperparameters_batch_sharded
         (== GaussianProcess.update_hyperparameters_batch at N = 1) with HOST buffers: theta -> device, kernel,
         all-gather, results -> host and the host-side prior / masking arithmetic inside the timed region.
  roofline  FP64 tensor pipe: (B / N) * M^3 algorithmic flop per launch / measured launch time on this rank, against
         the cuBLAS Dgemm throughput measured in this very run (MEASURED_PEAKS.json carries no FP64 figure).
  cpu_baseline  the numpy/scipy restatement of the reference's algorithm (oracle/gp_oracle.py, "port") timed on
         the host cores on a bounded sample of the same workload.
  secondary  (same run, after the headline legs) weak scaling of the same kernel (4096 thetas PER rank); config 4:
         gpt_ll at M = 49152 (assembly + blocked Cholesky + two triangular solves) as TFLOP/s on M^3/3 and as a
         fraction of Dgemm, on rank 0's GPU (replicas only, not sharded); prediction (mean + std) on the config-4 GP
         through parallel.predict_sharded, test points split by rank, as whole-job points/s.
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
B_GLOBAL = 4096      # config 3: the theta batch, split over the ranks (strong scaling)
B_PER_GPU = 4096     # secondary weak-scaling leg: thetas per rank
C4_LOCATIONS = 16384  # config 4: value + both gradient components at every location -> M = 49152
C4_PREDICT_POINTS = 8 * 28416  # bounded sample of the 10^6 test points of config 4 (whole job, split by rank):
                               # 8 whole waves of the variance-solve GEMMs (148 SMs x 3 CTAs x 64 rows), so that
                               # every rank count 1/2/4/8 works on whole waves
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
        per_step, B_GLOBAL)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
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
def c4_problem(nloc=C4_LOCATIONS):
    """Config-4 synthetic data: value + d/dx1 + d/dx2 at every location."""
    from numpy.random import RandomState
    rs = RandomState(0)
    X0 = rs.rand(nloc, 2)
    X = np.vstack([X0, X0, X0])
    n = np.vstack([np.zeros((nloc, 2), dtype=int), np.tile([1, 0], (nloc, 1)), np.tile([0, 1], (nloc, 1))])
    y = np.concatenate([np.sin(3 * X0[:, 0]) * np.cos(2 * X0[:, 1]),
                        3 * np.cos(3 * X0[:, 0]) * np.cos(2 * X0[:, 1]),
                        -2 * np.sin(3 * X0[:, 0]) * np.sin(2 * X0[:, 1])]) + 0.05 * rs.randn(3 * nloc)
    return X, n, y, np.full(3 * nloc, 0.05)


def _event_ms(torch, stream, fn, reps):
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run_secondary(args, g, torch, dist, dev_t, stream, rank, world, peak_tflops, weak_value):
    """Config-4 legs and the weak-scaling figure; every number is max-over-ranks device/host time."""
    from gptools_b200 import parallel
    distributed = world > 1
    out = {"weak_scaling": {"value": weak_value, "unit": UNIT, "thetas_per_gpu": B_PER_GPU,
                            "note": "every rank owns its own 4096-theta batch; same kernel, same all-gather"}}
    X, n, y, err = c4_problem()
    M = len(y)
    k = g.SquaredExponentialKernel(num_dim=2, initial_params=[1.0, 0.05, 0.05], param_bounds=[(0, 10)] * 3)
    gp = g.GaussianProcess(k, X=X, y=y, err_y=err, n=n, device=dev_t.index)
    dev, _ = gp._sync_device()
    dev.set_stream(stream.cuda_stream)
    th = np.array([1.0, 0.05, 0.05])
    gp.update_hyperparameters(th)                       # allocation + first factorisation (untimed)
    ts = []
    for rep in range(2):
        gp.K_up_to_date = False
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nll = gp.update_hyperparameters(th * (1.0 + 1e-3 * (rep + 1)))   # assembly + Cholesky + solves, host call to host scalar
        ts.append(time.perf_counter() - t0)
    t_ll = min(ts)
    tt = torch.tensor([t_ll], dtype=torch.float64, device=dev_t)
    if distributed:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ll = float(tt.item())
    tfl = M ** 3 / 3.0 / t_ll * 1e-12
    out["c4_cholesky"] = {"M": M, "seconds": t_ll, "tflops": tfl, "flop": "M^3/3",
                          "frac_of_dgemm": (tfl / peak_tflops) if peak_tflops else None, "ll_finite": bool(np.isfinite(nll)),
                          "what": "GaussianProcess.update_hyperparameters = gpt_ll: covariance assembly + blocked "
                                  "Cholesky + both triangular solves + log-det, host call to host scalar; replicas "
                                  "only (every rank factors the replicated training set)"}
    from numpy.random import RandomState
    Ms = C4_PREDICT_POINTS
    Xs = RandomState(2).rand(Ms, 2)
    parallel.predict_sharded(gp, Xs[:2048 * world], n=0, return_std=True)      # warm-up: workspaces
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    t0 = time.perf_counter()
    mean, std = parallel.predict_sharded(gp, Xs, n=0, return_std=True)
    torch.cuda.synchronize()
    t_pred = time.perf_counter() - t0
    tt = torch.tensor([t_pred], dtype=torch.float64, device=dev_t)
    if distributed:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_pred = float(tt.item())
    trsm = float(M) ** 2 * Ms / t_pred * 1e-12      # the M^2 M* flop of the variance solve, whole job
    out["c4_predict"] = {"M": M, "test_points": Ms, "seconds": t_pred, "points_per_s": Ms / t_pred, "n_gpus": world,
                         "trsm_tflops_whole_job": trsm,
                         "frac_of_dgemm_per_gpu": (trsm / world / peak_tflops) if peak_tflops else None,
                         "results_finite": bool(np.isfinite(mean).all() and np.isfinite(std).all()),
                         "what": "parallel.predict_sharded mean + std, test points split by rank, host arrays in / "
                                 "out; bounded sample of the 10^6 test points of config 4"}
    del gp
    return out


def run_ours(args, rank, world, local_rank):
    import warnings
    import torch
    import torch.distributed as dist
    import gptools_b200 as g
    from gptools_b200 import parallel

    warnings.simplefilter("ignore")
    torch.cuda.set_device(local_rank)
    dev_t = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        os.environ["NCCL_DEBUG"] = os.environ.get("GPT_NCCL_DEBUG", "WARN")   # keep stdout to the ONE JSON line
        dist.init_process_group("nccl", device_id=dev_t)

    X, n, y, err = c3_problem()
    k = g.SquaredExponentialKernel(num_dim=2, initial_params=[1.0, 0.3, 0.4], param_bounds=[(0, 10)] * 3)
    gp = g.GaussianProcess(k, X=X, y=y, err_y=err, n=n, use_hyper_deriv=True, device=local_rank)
    dev, _ = gp._sync_device()
    stream = torch.cuda.Stream(device=dev_t)          # the library launches on THIS stream; events are recorded on it
    torch.cuda.set_stream(stream)
    dev.set_stream(stream.cuda_stream)
    grad_idx = [0, 1, 2]

    def make_leg(B_local, th_rows):
        """Device-resident leg for B_local thetas on this rank: kernel outputs are views of the all-gather send buffer."""
        th_full = np.hstack([th_rows, np.zeros((B_local, 1))])      # kernel params + sigma_noise (ZeroKernel)
        d_th = torch.from_numpy(th_full).to(dev_t)
        off_grad = 8 * B_local
        off_st = off_grad + 24 * B_local
        nbytes = off_st + 8 * ((4 * B_local + 7) // 8)
        send = torch.zeros(nbytes, dtype=torch.uint8, device=dev_t)
        recv = torch.empty(world * nbytes, dtype=torch.uint8, device=dev_t) if distributed else None
        base = send.data_ptr()

        def kernel_only():
            dev.ll_batched_dev(B_local, d_th.data_ptr(), base, base + off_st, d_grad=base + off_grad, grad_idx=grad_idx)

        def step():
            kernel_only()
            if distributed:
                dist.all_gather_into_tensor(recv, send)

        def check():
            ll = send[:8 * B_local].view(torch.float64)
            st = send[off_st:off_st + 4 * B_local].view(torch.int32)
            return bool((st == 0).all().item()) and bool(torch.isfinite(ll).all().item())
        return step, kernel_only, check, d_th

    # ---- headline: STRONG scaling, the global 4096-theta batch split by rank ------------------------------------
    th_global = theta_batch(B_GLOBAL, seed=1)
    lo, hi = parallel.shard_bounds(B_GLOBAL, rank, world)
    B_loc = hi - lo
    step_device, kernel_only, check, _keep = make_leg(B_loc, th_global[lo:hi])
    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = dev.launch_count()
    ms_total = _event_ms(torch, stream, step_device, args.steps) * args.steps
    if distributed:
        dist.barrier()
    launches = dev.launch_count() - launches0
    ok = check()
    kernel_ms = _event_ms(torch, stream, kernel_only, args.steps)

    # ---- end-to-end leg: the product's multi-GPU API, host buffers ----------------------------------------------
    th_pinned = torch.empty((B_GLOBAL, 3), dtype=torch.float64).pin_memory()
    th_pinned.copy_(torch.from_numpy(th_global))
    th_host = th_pinned.numpy()
    for _ in range(max(1, args.warmup // 2)):
        parallel.update_hyperparameters_batch_sharded(gp, th_host, with_deriv=True)
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        neg_ll, neg_grad = parallel.update_hyperparameters_batch_sharded(gp, th_host, with_deriv=True)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clock_info = clocks.finish() if rank == 0 else None
    h2d = B_loc * 4 * 8                                   # this rank's theta rows (kernel params + sigma_n)
    d2h = world * (B_loc * 8 + B_loc * 3 * 8 + B_loc * 4) if distributed else B_loc * 8 + B_loc * 3 * 8 + B_loc * 4
    ok = ok and bool(np.isfinite(neg_ll).all()) and neg_ll.shape == (B_GLOBAL,)

    # ---- secondary: weak scaling of the same kernel (4096 thetas per rank) --------------------------------------
    weak_steps = max(3, args.steps // 4)
    w_step, _, w_check, _keep2 = make_leg(B_PER_GPU, theta_batch(B_PER_GPU, seed=1 + rank))
    for _ in range(2):
        w_step()
    if distributed:
        dist.barrier()
    weak_ms = _event_ms(torch, stream, w_step, weak_steps)
    ok = ok and w_check()

    # ---- FP64 roofline denominator measured live (cuBLAS Dgemm through torch) -------------------------
    a = torch.randn(8192, 8192, dtype=torch.float64, device=dev_t)
    b = torch.randn(8192, 8192, dtype=torch.float64, device=dev_t)
    torch.matmul(a, b)
    best = 1e30
    for _ in range(3):
        best = min(best, _event_ms(torch, stream, lambda: torch.matmul(a, b), 1))
    peak_tflops = 2.0 * 8192 ** 3 / best * 1e-9
    del a, b

    # ---- max over ranks ------------------------------------------------------------------------------------
    times = torch.tensor([ms_total, e2e_s * 1e3, kernel_ms, weak_ms, -peak_tflops], dtype=torch.float64, device=dev_t)
    if distributed:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, kernel_ms, weak_ms, neg_peak = (float(v) for v in times.cpu())
    peak_tflops = -neg_peak                             # min over ranks of the live Dgemm figure
    weak_value = world * B_PER_GPU / (weak_ms * 1e-3)

    secondary = None
    if not args.no_secondary:
        try:
            secondary = run_secondary(args, g, torch, dist, dev_t, stream, rank, world, peak_tflops, weak_value)
        except Exception as e:   # the headline line must survive a failure of the extra legs
            secondary = {"weak_scaling": {"value": weak_value, "unit": UNIT, "thetas_per_gpu": B_PER_GPU},
                         "error": "%s: %s" % (type(e).__name__, e)}

    if rank == 0:
        value = B_GLOBAL * args.steps / (ms_total * 1e-3)
        e2e_value = B_GLOBAL * args.steps / (e2e_ms * 1e-3)
        B_max = parallel.shard_bounds(B_GLOBAL, 0, world)[1]          # the largest slice (rank 0)
        achieved = B_max * FLOP_PER_EVAL / (kernel_ms * 1e-3) * 1e-12
        cpu_val, cpu_dt = cpu_baseline_single_process(24)
        traffic = None
        tf = os.path.join(ROOT, "profiles", "r02_batched_kernel_traffic.json")
        if os.path.exists(tf):
            try:
                traffic = json.load(open(tf)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        if traffic is not None and world > 1:
            traffic = traffic * B_max / float(B_GLOBAL)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config 3: batched ll+grad, 4096 thetas, SE 2-D kernel, M=512 obs (256 values + 2x128 "
                                   "first derivatives), P=3 free params, theta batch sharded by rank",
                       "global_batch": B_GLOBAL, "thetas_per_gpu": B_max, "parallelism": "theta-sharded x%d" % world,
                       "l2": "per-CTA factor workspace (up to 592 x 2.6 MB) exceeds the 126 MB L2: no flush needed",
                       "results_ok": ok},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "gptools_b200.parallel.update_hyperparameters_batch_sharded"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved / peak_tflops if peak_tflops else None, "traffic": traffic,
                         "kernel": "ll_batched4_kernel", "kernel_ms": kernel_ms,
                         "flop_per_launch": B_max * FLOP_PER_EVAL,
                         "peak_source": "cuBLAS Dgemm 8192^3 fp64 measured live in this run (MEASURED_PEAKS.json "
                                        "has no FP64 entry); DMMA issue peak 37.1 TFLOP/s (profiles/r01_fp64_peak_microbench.txt)"},
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": host_threads(), "kind": "port",
                             "sample": "24 thetas of the batch, one process, numpy/scipy threads on all cores (%.1f s)" % cpu_dt},
            "clocks": clock_info,
            "secondary": secondary,
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
    ap.add_argument("--no-secondary", action="store_true", help="skip the config-4 legs (headline line only)")
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
