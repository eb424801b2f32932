"""One pass over the single-matrix kernels on a config-4-shaped problem, for per-kernel ncu captures
(assemble_kernel, potrf_diag_kernel, panel_trsm_kernel, gemm_nt_kernel, predict_mean_kernel, rowdot / row_var).
usage: prof_kernels.py [locations] [mean_only_points] [mean_var_points]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from numpy.random import RandomState
from gptools_b200._lib import Device
nloc = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
m1 = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
m2 = int(sys.argv[3]) if len(sys.argv) > 3 else 28416
rs = RandomState(0)
X0 = rs.rand(nloc, 2)
X = np.vstack([X0, X0, X0])
n = np.vstack([np.zeros((nloc, 2), int), np.tile([1, 0], (nloc, 1)), np.tile([0, 1], (nloc, 1))])
y = rs.randn(3 * nloc)
d = Device(0)
d.set_data(X, n, y, np.full(3 * nloc, 0.05))
d.set_kernel(0, 3, 1e2)
th = np.array([1.0, 0.05, 0.05]) if nloc >= 8192 else np.array([1.0, 0.1, 0.1])
print(d.ll(th, 0.0))
Xs = RandomState(2).rand(max(m1, m2), 2)
z = np.zeros((max(m1, m2), 2), dtype=int)
mean, _, _ = d.predict(Xs[:m1], z[:m1], want_var=False)
mean2, var, _ = d.predict(Xs[:m2], z[:m2], want_var=True)
print(mean[:2], mean2[:2], var[:2])
