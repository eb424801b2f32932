"""Latency of one single-theta call through the C-ABI and through the Python API at small sizes (the SLSQP path of
optimize_hyperparameters evaluates theta one at a time)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import bench
import gptools_b200 as g
from gptools_b200._lib import Device

X, n, y, err = bench.c3_problem()
d = Device(0)
for M in (64, 128, 256, 512):
    d.set_data(X[:M], n[:M], y[:M], err[:M])
    d.set_kernel(0, 3, 1e2)
    th = np.array([1.0, 0.3, 0.4])
    for mode, gi in (("ll", None), ("ll+grad", [0, 1, 2])):
        d.ll(th, 0.0, grad_idx=gi)
        t0 = time.perf_counter()
        for _ in range(50):
            d.ll(th, 0.0, grad_idx=gi)
        t1 = (time.perf_counter() - t0) / 50
        thb = np.array([[1.0, 0.3, 0.4, 0.0]])
        d.ll_batched(thb, grad_idx=gi)
        t0 = time.perf_counter()
        for _ in range(50):
            d.ll_batched(thb, grad_idx=gi)
        t2 = (time.perf_counter() - t0) / 50
        print("M=%4d %-8s gpt_ll %.3f ms   gpt_ll_batched(B=1) %.3f ms" % (M, mode, t1 * 1e3, t2 * 1e3))
k = g.SquaredExponentialKernel(num_dim=2, initial_params=[1.0, 0.3, 0.4], param_bounds=[(0, 10)] * 3)
gp = g.GaussianProcess(k, X=X, y=y, err_y=err, n=n, use_hyper_deriv=True)
import warnings
warnings.simplefilter("ignore")
gp.update_hyperparameters(np.array([1.0, 0.3, 0.4]))
t0 = time.perf_counter()
for i in range(50):
    gp.update_hyperparameters(np.array([1.0, 0.3, 0.4 + 1e-6 * i]))
print("GaussianProcess.update_hyperparameters (M=512, ll+grad): %.3f ms per call" % ((time.perf_counter() - t0) / 50 * 1e3))
