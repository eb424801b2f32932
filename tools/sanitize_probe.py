"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_probe.py
Sizes are chosen to touch multi-block paths (M > 128, M > 64) while staying fast under instrumentation."""
import sys
import numpy as np
sys.path.insert(0, ".")
from gptools_b200._lib import Device

d = Device(0)
rs = np.random.RandomState(0)
for M, D, kid, th, idx in ((200, 2, 0, [1.0, 0.3, 0.4], [0, 1, 2]), (150, 1, 1, [1.0, 0.5], [0, 1]),
                           (140, 1, 3, [1.5, 0.6, 0.1, 0.05, 0.9], [0, 1, 2, 3, 4])):
    X = rs.rand(M, D)
    n = np.zeros((M, D), dtype=int)
    n[M // 2:, 0] = 1
    y = rs.randn(M)
    d.set_data(X, n, y, np.full(M, 0.1))
    d.set_kernel(kid, len(th), 1e2)
    ll, g, st = d.ll(np.array(th), 0.0, grad_idx=idx)
    llb, gb, stb = d.ll_batched(np.array([th + [0.0]] * 3), grad_idx=idx)
    Xs = rs.rand(70, D)
    ns = np.zeros((70, D), dtype=int)
    d.ll(np.array(th), 0.0)
    m, v, c = d.predict(Xs, ns, want_cov=True)
    m2, _, _ = d.predict(Xs, ns, want_var=False)
    s, sst = d.draw_sample(m, c, rs.randn(70, 5), 1e3 * 2.2e-16)
    print(M, kid, st, stb, sst, float(ll), float(llb[0]), np.abs(m - m2).max())
# T path
N, Mo = 160, 20
T = np.zeros((Mo, N))
for i in range(Mo):
    T[i, i * 7:i * 7 + 20] = 0.05
X = np.linspace(0, 1, N)[:, None]
d.set_data(X, np.zeros((N, 1), int), rs.rand(Mo), np.full(Mo, 0.05), T)
d.set_kernel(0, 2, 1e2)
print(d.ll(np.array([1.0, 0.2]), 0.0, grad_idx=[0, 1]))
print(d.predict(np.array([[0.5]]), np.zeros((1, 1), int), want_var=True))
# kernel algebra on the device: (SE + Matern52) * SE through gpt_ll, the persistent batched kernel and predict
import gptools_b200 as g
se = lambda p: g.SquaredExponentialKernel(num_dim=2, initial_params=p, param_bounds=[(0, 10)] * 3)
kc = (se([1.0, 0.4, 0.5]) + g.Matern52Kernel(num_dim=2, initial_params=[0.7, 0.8, 0.6], param_bounds=[(0, 10)] * 3)) * se([0.9, 1.5, 1.2])
Xc = rs.rand(150, 2)
gp = g.GaussianProcess(kc, use_hyper_deriv=True)
gp.add_data(Xc, np.sin(3 * Xc[:, 0]) + 0.05 * rs.randn(150), err_y=0.05)
gp.add_data(Xc[::10], 3 * np.cos(3 * Xc[::10, 0]), err_y=0.1, n=np.tile([1, 0], (15, 1)))
th = np.array(gp.free_params[:], dtype=float)
import warnings
warnings.simplefilter("ignore")
print(gp.update_hyperparameters(th)[0], gp.update_hyperparameters_batch(np.vstack([th, 1.02 * th]), with_deriv=True)[0])
print(gp.predict(rs.rand(5, 2))[0])
# batched prediction (test points as extra tile rows) for SE and Matern-5/2 (their own short-form instantiations) and
# the Matern-5/2 batched gradient
for kk in (g.SquaredExponentialKernel(num_dim=2, initial_params=[1.0, 0.4, 0.5], param_bounds=[(0, 10)] * 3),
           g.Matern52Kernel(num_dim=2, initial_params=[1.0, 0.6, 0.7], param_bounds=[(0, 10)] * 3)):
    gq = g.GaussianProcess(kk, use_hyper_deriv=True)
    gq.add_data(Xc, np.sin(3 * Xc[:, 0]) + 0.05 * rs.randn(150), err_y=0.05)
    gq.add_data(Xc[::10], 3 * np.cos(3 * Xc[::10, 0]), err_y=0.1, n=np.tile([1, 0], (15, 1)))
    tq = np.array(gq.free_params[:], dtype=float)
    tb = np.vstack([tq, 1.03 * tq, 0.97 * tq])
    fq, dq = gq.update_hyperparameters_batch(tb, with_deriv=True)
    mq, sq, okq = gq.predict_batch(tb, rs.rand(70, 2), n=np.tile([0, 1], (70, 1)))
    print(type(kk).__name__, fq[0], dq[0], float(mq[0, 0]), float(sq[0, 0]), okq.all())
