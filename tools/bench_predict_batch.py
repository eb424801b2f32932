"""compute_from_MCMC-style prediction at many hyper-parameter samples: ONE launch (gpt_predict_batched) against the
per-sample loop (update_hyperparameters + predict), config-3 problem (M = 512), development aid.
usage: bench_predict_batch.py [samples] [test points]"""
import sys, time, warnings
import numpy as np
sys.path.insert(0, ".")
warnings.simplefilter("ignore")
import bench
import gptools_b200 as g

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
Ms = int(sys.argv[2]) if len(sys.argv) > 2 else 400
X, n, y, err = bench.c3_problem()
k = g.SquaredExponentialKernel(num_dim=2, initial_params=[1.0, 0.3, 0.4], param_bounds=[(0, 10)] * 3)
gp = g.GaussianProcess(k, X=X, y=y, err_y=err, n=n)
th = bench.theta_batch(S)
Xs = np.random.RandomState(2).rand(Ms, 2)
gp.predict_batch(th[:8], Xs)
t0 = time.perf_counter()
mean, std, good = gp.predict_batch(th, Xs)
tb = time.perf_counter() - t0
nl = min(S, 64)
gp._mcmc_predict_by_loop = True
gp.compute_from_MCMC(Xs, flat_trace=th[:2])
t0 = time.perf_counter()
res = gp.compute_from_MCMC(Xs, flat_trace=th[:nl])
tl = (time.perf_counter() - t0) / nl
err_m = np.abs(mean[:nl] - np.array(res["mean"])).max()
print("M=512, %d test points: batched %d samples in %.1f ms (%.3f ms per sample); per-sample loop %.3f ms per sample "
      "(%d samples) -> %.1fx; max |mean diff| %.2e; all good %s" % (Ms, S, tb * 1e3, tb / S * 1e3, tl * 1e3, nl,
                                                                    tl / (tb / S), err_m, bool(good.all())))
