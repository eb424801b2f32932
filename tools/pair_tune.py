import sys, time, os
import numpy as np
sys.path.insert(0, ".")
from numpy.random import RandomState
from gptools_b200._lib import Device
d = Device(0)
for nloc in (1366, 1707, 2048, 2731, 3414, 4096):
    rs = RandomState(0)
    X0 = rs.rand(nloc, 2)
    X = np.vstack([X0, X0, X0])
    n = np.vstack([np.zeros((nloc, 2), int), np.tile([1, 0], (nloc, 1)), np.tile([0, 1], (nloc, 1))])
    y = rs.randn(3 * nloc)
    d.set_data(X, n, y, np.full(3 * nloc, 0.05))
    d.set_kernel(0, 3, 1e2)
    th = np.array([1.0, 0.1, 0.1])
    out = []
    for pm in ("100000", "1"):
        os.environ["GPT_POTRF_PAIR_MIN"] = pm
        d.ll(th, 0.0)
        t0 = time.perf_counter()
        for _ in range(5):
            d.ll(th, 0.0)
        out.append((time.perf_counter() - t0) / 5 * 1e3)
    print("M=%5d (%3d blocks): unpaired %.2f ms   paired %.2f ms" % (3 * nloc, (3 * nloc + 127) // 128, out[0], out[1]))
