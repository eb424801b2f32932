"""Quick device-side timing of the batched path on the config-3 problem (development aid)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from numpy.random import RandomState
from gptools_b200._lib import Device

rs = RandomState(0)
Xv = rs.rand(256, 2); Xd1 = rs.rand(128, 2); Xd2 = rs.rand(128, 2)
f = lambda x: np.sin(3 * x[:, 0]) * np.cos(2 * x[:, 1])
X = np.vstack([Xv, Xd1, Xd2])
n = np.vstack([np.zeros((256, 2), int), np.tile([1, 0], (128, 1)), np.tile([0, 1], (128, 1))])
y = np.concatenate([f(Xv) + 0.05 * rs.randn(256),
                    3 * np.cos(3 * Xd1[:, 0]) * np.cos(2 * Xd1[:, 1]) + 0.05 * rs.randn(128),
                    -2 * np.sin(3 * Xd2[:, 0]) * np.sin(2 * Xd2[:, 1]) + 0.05 * rs.randn(128)])
err = np.full(512, 0.05)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
th = np.array([1.0, 0.3, 0.4]) * np.exp(0.1 * RandomState(1).randn(B, 3))
th = np.hstack([th, np.zeros((B, 1))])
d = Device(0)
d.set_data(X, n, y, err)
d.set_kernel(0, 3, 1e2)
for mode, gi in (("ll+grad", [0, 1, 2]), ("ll only", None)):
    d.ll_batched(th[:296], grad_idx=gi)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        ll, g, st = d.ll_batched(th, grad_idx=gi)
        ts.append(time.perf_counter() - t0)
    t = min(ts)
    flop = B * 512.0 ** 3 * (1.0 if gi else 1 / 3.0)
    if gi and len(sys.argv) > 2:
        np.save(sys.argv[2], np.column_stack([ll, g]))
    print("%s: B=%d  %.2f ms  %.0f evals/s  %.2f TFLOP/s (M^3%s counted)  status ok=%s" % (
        mode, B, t * 1e3, B / t, flop / t * 1e-12, "" if gi else "/3", (st == 0).all()))
