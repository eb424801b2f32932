"""Randomised cross-check of the persistent many-theta kernel against the single-theta path (which the golden / oracle
tests pin): ll, gradient and batched prediction over random sizes, dimensions, derivative orders, kernels and noise.
usage: fuzz_batched.py [trials] [seed]"""
import sys, warnings
import numpy as np
sys.path.insert(0, ".")
warnings.simplefilter("ignore")
import gptools_b200 as g

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rs = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
worst = {"ll": 0.0, "grad": 0.0, "mean": 0.0, "var": 0.0}
for trial in range(trials):
    D = int(rs.choice([1, 1, 2, 2, 3]))
    M = int(rs.choice([1, 2, 7, 31, 63, 64, 65, 100, 127, 128, 129, 200, 255, 300, 448, 513, 700]))
    kind = rs.choice(["se", "se", "m52", "matern", "comp"])
    bnd = [(0.01, 20)] * (D + 1)
    ls = list(0.3 + 0.5 * rs.rand(D))
    if kind == "se":
        k = g.SquaredExponentialKernel(num_dim=D, initial_params=[1.0 + rs.rand()] + ls, param_bounds=bnd)
    elif kind == "m52":
        k = g.Matern52Kernel(num_dim=D, initial_params=[1.0 + rs.rand()] + ls, param_bounds=bnd)
    elif kind == "matern":
        k = g.MaternKernel(num_dim=D, initial_params=[1.0 + rs.rand(), float(rs.choice([1.5, 2.5, 2.2]))] + ls,
                           param_bounds=[(0.01, 20)] * (D + 2), fixed_params=[False, True] + [False] * D)
    else:
        k = (g.SquaredExponentialKernel(num_dim=D, initial_params=[1.0] + ls, param_bounds=bnd) +
             g.SquaredExponentialKernel(num_dim=D, initial_params=[0.3] + [0.5 * v for v in ls], param_bounds=bnd))
    if k.num_params > 7 and kind == "comp":
        continue
    noise = rs.rand() < 0.5
    nk = g.DiagonalNoiseKernel(D, initial_noise=0.05 + 0.1 * rs.rand(), fixed_noise=False, noise_bound=(0, 5)) if noise else None
    gp = g.GaussianProcess(k, noise_k=nk, use_hyper_deriv=True) if noise else g.GaussianProcess(k, use_hyper_deriv=True)
    X = rs.rand(M, D)
    n = np.zeros((M, D), dtype=int)
    maxord = 1 if kind in ("m52", "matern") or rs.rand() < 0.7 else 2
    if M > 3:
        rows = rs.rand(M) < 0.4
        n[rows, rs.randint(0, D, rows.sum())] = rs.randint(1, maxord + 1, rows.sum())
    y = np.sin(3 * X[:, 0]) + 0.1 * rs.randn(M)
    gp.add_data(X, y, err_y=0.05 + 0.05 * rs.rand(M), n=n)
    th0 = np.array(gp.free_params[:], dtype=float)
    B = int(rs.choice([1, 2, 5, 9]))
    th = th0 * np.exp(0.05 * rs.randn(B, len(th0)))
    f, df = gp.update_hyperparameters_batch(th, with_deriv=True)
    Ms = int(rs.choice([1, 3, 64, 65, 150]))
    Xs = rs.rand(Ms, D)
    ns = np.zeros((Ms, D), dtype=int)
    if rs.rand() < 0.5:
        ns[::2, rs.randint(0, D)] = 1
    res = gp.predict_batch(th, Xs, n=ns)
    assert res is not None
    for b in range(B):
        fb, dfb = gp.update_hyperparameters(th[b])
        if not np.isfinite(fb):
            assert not np.isfinite(f[b]), (trial, b)
            continue
        m, s = gp.predict(Xs, n=ns)
        sc = max(1.0, abs(fb))
        worst["ll"] = max(worst["ll"], abs(f[b] - fb) / sc)
        worst["grad"] = max(worst["grad"], np.abs(df[b] - dfb).max() / max(1.0, np.abs(dfb).max()))
        worst["mean"] = max(worst["mean"], np.abs(res[0][b] - m).max() / max(1.0, np.abs(m).max()))
        pv = max(1e-12, float(np.max(s ** 2)), float(th[b][0] ** 2))
        worst["var"] = max(worst["var"], np.abs(res[1][b] ** 2 - s ** 2).max() / pv)
        assert abs(f[b] - fb) <= 1e-9 * sc, ("ll", trial, kind, M, D, b, f[b], fb)
        assert np.abs(df[b] - dfb).max() <= 1e-7 * max(1.0, np.abs(dfb).max()), ("grad", trial, kind, M, D, b, df[b], dfb)
        assert np.abs(res[0][b] - m).max() <= 1e-8 * max(1.0, np.abs(m).max()), ("mean", trial, kind, M, D, Ms, b)
        assert np.abs(res[1][b] ** 2 - s ** 2).max() <= 1e-8 * pv, ("var", trial, kind, M, D, Ms, b)
print("fuzz OK: %d trials; worst relative differences batched vs single-theta: %s" % (trials, {k_: "%.1e" % v for k_, v in worst.items()}))
