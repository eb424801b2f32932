"""Randomised consistency sweep (development aid): single-theta path vs batched path vs the pinned oracle over input
dimensions, derivative orders (incl. second order: the non-"low order" code paths) and sizes around the tile edges."""
import sys
import numpy as np
sys.path.insert(0, ".")
from gptools_b200._lib import Device
from oracle import gp_oracle as orc

d = Device(0)
rs = np.random.RandomState(123)
worst = 0.0
for trial in range(24):
    D = [1, 2, 3, 4][trial % 4]
    M = int(rs.choice([37, 64, 100, 129, 200, 260]))
    maxord = [1, 2][(trial // 4) % 2]
    X = rs.rand(M, D)
    n = np.zeros((M, D), dtype=int)
    for i in range(M // 3, M):
        n[i, rs.randint(D)] = rs.randint(0, maxord + 1)
    y = rs.randn(M)
    err = np.full(M, 0.2)
    th = np.concatenate([[1.0 + 0.5 * rs.rand()], 0.3 + 0.4 * rs.rand(D)])
    idx = list(range(D + 1))
    ref = orc.compute_K_L_alpha_ll(orc.KERNEL_SE, th, X, n, y, err, None, 0.0, 1e2, grad_idx=idx)
    d.set_data(X, n, y, err)
    d.set_kernel(0, D + 1, 1e2)
    ll, g, st = d.ll(th, 0.0, grad_idx=idx)
    llb, gb, stb = d.ll_batched(np.array([list(th) + [0.0]] * 2), grad_idx=idx)
    gs = np.abs(ref["ll_deriv"]).max()
    e = [abs(ll - ref["ll"]) / abs(ref["ll"]), np.abs(g - ref["ll_deriv"]).max() / gs,
         abs(llb[0] - ref["ll"]) / abs(ref["ll"]), np.abs(gb[0] - ref["ll_deriv"]).max() / gs]
    Xs = rs.rand(11, D)
    ns = np.zeros((11, D), dtype=int)
    ns[::3, 0] = 1
    d.ll(th, 0.0)
    m, v, _ = d.predict(Xs, ns, want_var=True)
    pm, ps, pc = orc.predict(orc.KERNEL_SE, th, X, n, ref["L"], ref["alpha"], Xs, ns)
    e += [np.abs(m - pm).max() / max(1.0, np.abs(pm).max()), np.abs(v - np.diag(pc)).max() / th[0] ** 2]
    worst = max(worst, max(e))
    flag = "" if max(e) < 1e-8 and st == 0 and (stb == 0).all() else "   <-- CHECK"
    print("D=%d M=%3d maxord=%d  st=%d  ll %.1e grad %.1e | batched ll %.1e grad %.1e | mean %.1e var %.1e%s" % (
        D, M, maxord, st, e[0], e[1], e[2], e[3], e[4], e[5], flag))
print("worst relative deviation: %.2e" % worst)
