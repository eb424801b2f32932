"""Multi-start MAP optimisation on the config-3 problem (M = 512, 2-D SE, value + first-derivative observations):
lock-step batched starts against sequential starts (development / profiles aid).
usage: bench_multistart.py [nstarts_batched] [nstarts_sequential]"""
import sys, time, warnings
import numpy as np
sys.path.insert(0, ".")
import bench
import gptools_b200 as g

warnings.simplefilter("ignore")
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 8
X, n, y, err = bench.c3_problem()


def make():
    k = g.SquaredExponentialKernel(num_dim=2, initial_params=[1.0, 0.3, 0.4], param_bounds=[(0.05, 10), (0.02, 5), (0.02, 5)])
    return g.GaussianProcess(k, X=X, y=y, err_y=err, n=n, use_hyper_deriv=True)


for label, nstart, batched in (("lock-step batched", nb, True), ("sequential", nq, False), ("lock-step batched", nq, True)):
    gp = make()
    gp.update_hyperparameters(np.array([1.0, 0.3, 0.4]))          # library / workspace warm-up
    np.random.seed(3)
    t0 = time.perf_counter()
    res, nres = gp.optimize_hyperparameters(random_starts=nstart, batched_starts=batched)
    dt = time.perf_counter() - t0
    print("%-18s %3d starts: %.3f s  (%d completed)  MAP -ll %.6f at %s" % (
        label, nstart, dt, nres, res.fun, np.array2string(res.x, precision=5)), flush=True)

# where the lock-step time goes: rounds, time inside the batched evaluations, everything else (scipy + threads)
gp = make()
gp.update_hyperparameters(np.array([1.0, 0.3, 0.4]))
stats = {"rounds": 0, "t": 0.0, "rows": 0}
orig = gp._eval_batch


def timed(thetas, with_deriv):
    t0 = time.perf_counter()
    r = orig(thetas, with_deriv)
    stats["t"] += time.perf_counter() - t0
    stats["rounds"] += 1
    stats["rows"] += len(thetas)
    return r


gp._eval_batch = timed
np.random.seed(3)
t0 = time.perf_counter()
gp.optimize_hyperparameters(random_starts=nb)
dt = time.perf_counter() - t0
print("lock-step %d starts: %.3f s total, %d rounds, %d objective rows, %.3f s inside the batched evaluations "
      "(%.2f ms per round), %.3f s scipy + thread hand-over" % (nb, dt, stats["rounds"], stats["rows"], stats["t"],
                                                                 1e3 * stats["t"] / max(stats["rounds"], 1), dt - stats["t"]))
