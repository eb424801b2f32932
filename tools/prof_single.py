"""One gpt_ll (assemble + blocked Cholesky + solves) on a config-4-shaped problem (for ncu launch lists)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from numpy.random import RandomState
from gptools_b200._lib import Device
nloc = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rs = RandomState(0)
X0 = rs.rand(nloc, 2)
X = np.vstack([X0, X0, X0])
n = np.vstack([np.zeros((nloc, 2), int), np.tile([1, 0], (nloc, 1)), np.tile([0, 1], (nloc, 1))])
y = rs.randn(3 * nloc)
d = Device(0)
d.set_data(X, n, y, np.full(3 * nloc, 0.05))
d.set_kernel(0, 3, 1e2)
th = np.array([1.0, 0.05, 0.05]) if nloc >= 8192 else np.array([1.0, 0.1, 0.1])
for _ in range(2):
    print(d.ll(th, 0.0))
