"""Prediction throughput with the test points sharded over the ranks (north-star: "reported at 1, 2, 4 and 8 GPUs for the
batched-theta and prediction paths"; bench.py covers the former).  Strong scaling: the total number of test points is
fixed, every rank factors the replicated training set itself, one all-gather of mean / std.
    python tools/bench_predict_sharded.py                                        # 1 GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/bench_predict_sharded.py
Prints one JSON line per workload on rank 0; time = max over ranks of the wall time between barriers."""
import json, os, sys, time, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import gptools_b200 as g
from gptools_b200.parallel import predict_sharded

warnings.simplefilter("ignore")
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def run(name, gp, Xs, reps=2):
    gp.compute_K_L_alpha_ll()
    predict_sharded(gp, Xs[:1024 * world])       # warm-up (buffers, factor resident)
    ts = []
    for _ in range(reps):
        barrier()
        t0 = time.perf_counter()
        mean, std = predict_sharded(gp, Xs)
        barrier()
        ts.append(time.perf_counter() - t0)
    t = torch.tensor([min(ts)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"workload": name, "n_gpus": world, "test_points": int(len(Xs)), "seconds": float(t.item()),
                          "points_per_s": len(Xs) / float(t.item()), "scaling": "strong",
                          "std_finite": bool(np.isfinite(std).all())}))


# config 2: 1-D Matern-5/2, 2000 locations with value + derivative (M = 4000), mean + std at 8e5 points
rs = np.random.RandomState(0)
X = np.sort(rs.rand(2000)) * 10
k = g.Matern52Kernel(num_dim=1, initial_params=[1.0, 0.8], param_bounds=[(0, 10)] * 2)
gp = g.GaussianProcess(k, device=local)
gp.add_data(X, np.sin(X) + 0.05 * rs.randn(2000), err_y=0.05)
gp.add_data(X, np.cos(X) + 0.05 * rs.randn(2000), err_y=0.05, n=1)
run("config 2 (M=4000 Matern52), mean+std", gp, np.linspace(0, 10, 800000))

# config-4 shape at M = 12288 (4096 locations x (value, d/dx1, d/dx2)), SE 2-D, mean + std at 4e5 points
rs = np.random.RandomState(0)
X0 = rs.rand(4096, 2)
f = lambda x: np.sin(3 * x[:, 0]) * np.cos(2 * x[:, 1])
X = np.vstack([X0, X0, X0])
n = np.vstack([np.zeros((4096, 2), int), np.tile([1, 0], (4096, 1)), np.tile([0, 1], (4096, 1))])
y = np.concatenate([f(X0), 3 * np.cos(3 * X0[:, 0]) * np.cos(2 * X0[:, 1]),
                    -2 * np.sin(3 * X0[:, 0]) * np.sin(2 * X0[:, 1])]) + 0.05 * rs.randn(12288)
k2 = g.SquaredExponentialKernel(num_dim=2, initial_params=[1.0, 0.1, 0.1], param_bounds=[(0, 10)] * 3)
gp2 = g.GaussianProcess(k2, X=X, y=y, err_y=0.05, n=n, device=local)
run("config-4 shape (M=12288 SE 2-D), mean+std", gp2, np.random.RandomState(2).rand(400000, 2))
if world > 1:
    dist.destroy_process_group()
