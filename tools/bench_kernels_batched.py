import sys, time
import numpy as np
sys.path.insert(0, ".")
from gptools_b200._lib import Device
rs = np.random.RandomState(0)
M = 512
X = np.sort(rs.rand(M, 1) * 10, axis=0)
n = np.zeros((M, 1), int); n[::2] = 1
y = np.sin(X[:, 0]); err = np.full(M, 0.05)
d = Device(0)
d.set_data(X, n, y, err)
for name, kid, th0 in (("matern52", 1, [1.0, 0.8]), ("matern nu=2.5", 2, [1.0, 2.5, 0.8]), ("matern nu=2.2", 2, [1.0, 2.2, 0.8]), ("gibbs", 3, [1.5, 3.0, 1.0, 0.5, 5.0]), ("se", 0, [1.0, 0.8])):
    B = 1184
    th = np.array(th0) * np.exp(0.02 * rs.randn(B, len(th0)))
    if kid == 2: th[:, 1] = th0[1]
    th = np.hstack([th, np.zeros((B, 1))])
    d.set_kernel(kid, len(th0), 1e2)
    gi = [0, len(th0) - 1] if kid != 3 else [0, 1, 2]
    for mode, g in (("ll+grad", gi), ("ll", None)):
        d.ll_batched(th, grad_idx=g)        # full-size warm-up: workspace allocation stays out of the timing
        ts = []
        for _ in range(2):
            t0 = time.perf_counter(); ll, gr, st = d.ll_batched(th, grad_idx=g); ts.append(time.perf_counter() - t0)
        t = min(ts)
        print("%-14s %-8s B=%d %.1f ms  %.0f evals/s ok=%s" % (name, mode, B, t * 1e3, B / t, (st == 0).all()))
