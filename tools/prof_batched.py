"""One warm-up + one measured launch of the batched kernel on the config-3 problem (for ncu).
usage: prof_batched.py [B] [grad|ll]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import bench
from gptools_b200._lib import Device

B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
mode = sys.argv[2] if len(sys.argv) > 2 else "grad"
X, n, y, err = bench.c3_problem()
th = np.hstack([bench.theta_batch(B), np.zeros((B, 1))])
d = Device(0)
d.set_data(X, n, y, err)
d.set_kernel(0, 3, 1e2)
for _ in range(2):
    ll, g, st = d.ll_batched(th, grad_idx=[0, 1, 2] if mode == "grad" else None)
print("ok", (st == 0).all(), ll[:2])
