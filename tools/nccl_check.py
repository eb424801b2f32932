"""Functional check of the multi-GPU product paths over NCCL (run under torch.distributed.run with 2+ ranks): every
sharded entry against the same call evaluated unsharded on the local GPU.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/nccl_check.py"""
import os, sys, warnings
import numpy as np
sys.path.insert(0, ".")
warnings.simplefilter("ignore")
import torch
import torch.distributed as dist
import gptools_b200 as g
from gptools_b200 import parallel

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rs = np.random.RandomState(0)
X = rs.rand(300, 2)
y = np.sin(3 * X[:, 0]) * np.cos(2 * X[:, 1]) + 0.05 * rs.randn(300)
k = g.SquaredExponentialKernel(num_dim=2, initial_params=[1.0, 0.3, 0.4], param_bounds=[(0.05, 5)] * 3)
gp = g.GaussianProcess(k, X=X, y=y, err_y=0.05, use_hyper_deriv=True, device=local)
gp.add_data(X[::10], 3 * np.cos(3 * X[::10, 0]) * np.cos(2 * X[::10, 1]), err_y=0.1, n=np.tile([1, 0], (30, 1)))
th = np.array([1.0, 0.3, 0.4]) * np.exp(0.1 * rs.randn(37, 3))       # ragged split
f_s, df_s = parallel.update_hyperparameters_batch_sharded(gp, th, with_deriv=True)
f_l, df_l = gp.update_hyperparameters_batch(th, with_deriv=True)
assert np.array_equal(f_s, f_l) and np.array_equal(df_s, df_l), "theta sharding differs from the local batch"
f1 = parallel.update_hyperparameters_batch_sharded(gp, th[:1], with_deriv=False)      # fewer rows than ranks
assert np.array_equal(f1, gp.update_hyperparameters_batch(th[:1], with_deriv=False))
thb = th.copy()
thb[5, 1] = -1.0                                                                     # outside the prior support
fb, dfb = parallel.update_hyperparameters_batch_sharded(gp, thb, with_deriv=True)
fl, dfl = gp.update_hyperparameters_batch(thb, with_deriv=True)
assert np.isinf(fb[5]) and np.array_equal(fb, fl) and np.array_equal(dfb, dfl), "prior masking through the gather"
Xs = rs.rand(1001, 2)
m_s, s_s = parallel.predict_sharded(gp, Xs)
m_l, s_l = gp.predict(Xs)
assert np.allclose(m_s, m_l, rtol=1e-12, atol=1e-13) and np.allclose(s_s, s_l, rtol=1e-9, atol=1e-12), "predict sharding"
mc = gp.compute_from_MCMC(Xs[:50], flat_trace=th[:11])
ref = gp.predict_batch(th[:11], Xs[:50])
assert len(mc["mean"]) == 11 and np.allclose(np.array(mc["mean"]), ref[0], rtol=1e-12, atol=1e-13), "MCMC prediction sharding"
np.random.seed(3)
res = gp.optimize_hyperparameters(random_starts=6, verbose=False)
par = torch.tensor(np.asarray(gp.free_params[:], dtype=float), device="cuda")
lst = [torch.empty_like(par) for _ in range(dist.get_world_size())]
dist.all_gather(lst, par)
assert all(torch.equal(lst[0], t) for t in lst), "ranks disagree on the optimum"
np.random.seed(4)
s = gp.sample_hyperparameter_posterior(nwalkers=12, nsamp=4)
ch = torch.tensor(s.chain, device="cuda")
lst = [torch.empty_like(ch) for _ in range(dist.get_world_size())]
dist.all_gather(lst, ch)
assert all(torch.equal(lst[0], t) for t in lst), "ranks disagree on the chain"
if rank == 0:
    print("nccl_check OK: world %d, MAP %s, ll %.6f" % (dist.get_world_size(), np.round(gp.free_params[:], 5), gp.ll))
dist.destroy_process_group()
