import sys
sys.path.insert(0, ".")
import numpy as np
from gptools_b200._lib import Device, GPTLibraryError
from oracle import gp_oracle as orc
d = Device(0)
def tryit(name, f):
    try:
        r = f()
        print(name, "->", r)
    except Exception as e:
        print(name, "EXC", type(e).__name__, str(e)[:100])
for M in (1, 2, 63, 64, 65, 127, 128, 129):
    rs = np.random.RandomState(M)
    X = rs.rand(M, 1); n = np.zeros((M, 1), int); y = np.sin(3 * X[:, 0]); err = np.full(M, 0.1)
    d.set_data(X, n, y, err); d.set_kernel(0, 2, 1e2)
    th = np.array([1.0, 0.4])
    ref = orc.compute_K_L_alpha_ll(0, th, X, n, y, err, None, 0.0, 1e2, grad_idx=[0, 1])
    ll, g, st = d.ll(th, 0.0, grad_idx=[0, 1])
    llb, gb, stb = d.ll_batched(np.array([[1.0, 0.4, 0.0]]), grad_idx=[0, 1])
    Xs = rs.rand(3, 1)
    m, v, _ = d.predict(Xs, np.zeros((3, 1), int), want_var=True)
    pr = orc.predict(0, th, X, n, ref["L"], ref["alpha"], Xs, np.zeros((3, 1), int))
    pm, pc = pr[0], pr[1]
    print(M, st, abs(ll - ref["ll"]) / abs(ref["ll"]), np.abs(g - ref["ll_deriv"]).max() / np.abs(ref["ll_deriv"]).max(),
          abs(llb[0] - ref["ll"]) / abs(ref["ll"]), np.abs(gb[0] - ref["ll_deriv"]).max() / np.abs(ref["ll_deriv"]).max(),
          np.abs(m - pm).max(), np.abs(v - np.diag(pc)).max())
tryit("predict Ms=0", lambda: d.predict(np.zeros((0, 1)), np.zeros((0, 1), int), want_var=True))
tryit("batched B=0", lambda: d.ll_batched(np.zeros((0, 3)), grad_idx=[0, 1]))
tryit("cov_pairs 0", lambda: d.cov_pairs(0, np.array([1.0, 0.4]), np.zeros((0, 1)), np.zeros((0, 1)), np.zeros((0, 1), int), np.zeros((0, 1), int)))
tryit("set_data M=0", lambda: d.set_data(np.zeros((0, 1)), np.zeros((0, 1), int), np.zeros(0), np.zeros(0)))
tryit("ll after M=0", lambda: d.ll(np.array([1.0, 0.4]), 0.0))
