"""Batched ll+grad throughput as a function of the number of observations M (SE 2-D, value + both gradient
components at M/3 locations, like the M = 1536 variant of SURVEY 8d; any remainder rows are extra value points).
The warm-up call uses the full batch so that workspace growth is outside the timed calls.
usage: bench_m_sweep.py [B] [M,M,...]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from numpy.random import RandomState
from gptools_b200._lib import Device

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
Ms = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [129, 258, 512, 513, 576, 768, 1536, 2046]
d = Device(0)
for M in Ms:
    nloc = M // 3
    extra = M - 3 * nloc
    rs = RandomState(0)
    X0 = rs.rand(nloc, 2)
    Xe = rs.rand(extra, 2)
    f = lambda x: np.sin(3 * x[:, 0]) * np.cos(2 * x[:, 1])
    X = np.vstack([X0, X0, X0, Xe])
    n = np.vstack([np.zeros((nloc, 2), int), np.tile([1, 0], (nloc, 1)), np.tile([0, 1], (nloc, 1)),
                   np.zeros((extra, 2), int)])
    y = np.concatenate([f(X0), 3 * np.cos(3 * X0[:, 0]) * np.cos(2 * X0[:, 1]),
                        -2 * np.sin(3 * X0[:, 0]) * np.sin(2 * X0[:, 1]), f(Xe)]) + 0.05 * rs.randn(M)
    d.set_data(X, n, y, np.full(M, 0.05))
    d.set_kernel(0, 3, 1e2)
    th = np.array([1.0, 0.3, 0.4]) * np.exp(0.1 * RandomState(1).randn(B, 3))
    th = np.hstack([th, np.zeros((B, 1))])
    d.ll_batched(th, grad_idx=[0, 1, 2])
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        ll, g, st = d.ll_batched(th, grad_idx=[0, 1, 2])
        ts.append(time.perf_counter() - t0)
    t = min(ts)
    print("M=%5d B=%d: %8.2f ms  %9.0f evals/s  %6.2f TFLOP/s (M^3)  ok=%s" % (
        M, B, t * 1e3, B / t, B * float(M) ** 3 / t * 1e-12, (st == 0).all()), flush=True)
