import os, sys, time, warnings
import numpy as np
sys.path.insert(0, ".")
warnings.simplefilter("ignore")
import torch, torch.distributed as dist
import bench
import gptools_b200 as g
from gptools_b200 import parallel
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
X, n, y, err = bench.c3_problem()
k = g.SquaredExponentialKernel(num_dim=2, initial_params=[1.0, 0.3, 0.4], param_bounds=[(0, 10)] * 3)
gp = g.GaussianProcess(k, X=X, y=y, err_y=err, n=n, use_hyper_deriv=True, device=local)
th = bench.theta_batch(4096)
W = dist.get_world_size()
for _ in range(3):
    parallel.update_hyperparameters_batch_sharded(gp, th, with_deriv=True)
# phase timing by monkeypatching
T = {}
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(*a, **k); torch.cuda.synchronize(); T[name] = T.get(name, 0) + time.perf_counter() - t0; return r
    return w
gp._batch_prepare = timed("prepare", gp._batch_prepare)
gp._batch_finish = timed("finish", gp._batch_finish)
gp._batch_plan_from_gathered = timed("plan_gathered", gp._batch_plan_from_gathered)
orig = parallel._theta_batch_device
parallel._theta_batch_device = timed("device+gather", orig)
N = 20
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
for _ in range(N):
    parallel.update_hyperparameters_batch_sharded(gp, th, with_deriv=True)
torch.cuda.synchronize(); tot = (time.perf_counter() - t0) / N
if rank == 0:
    print("world %d: total %.3f ms per call; %s" % (W, tot * 1e3, {k_: "%.3f ms" % (v / N * 1e3) for k_, v in T.items()}))
dist.destroy_process_group()
