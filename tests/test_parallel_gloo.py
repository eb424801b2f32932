"""world_size-2 gloo tests (CPU) of the N>1 path: theta rows / test points are sharded by rank with no
data-path collective, and one all-gather returns the full result on every rank."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    import warnings
    warnings.simplefilter("ignore")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import gptools_b200 as g
    from gptools_b200 import parallel
    from fake_device import FakeDevice
    rs = np.random.RandomState(0)
    X = rs.rand(10, 2)
    k = g.SquaredExponentialKernel(num_dim=2, initial_params=[1.3, 0.7, 1.1], param_bounds=[(0, 10)] * 3)
    gp = g.GaussianProcess(k, X=X, y=np.sin(X).sum(1), err_y=0.05)
    gp._dev_obj = FakeDevice()
    th = np.array([1.3, 0.7, 1.1]) * np.exp(0.1 * np.random.RandomState(1).randn(7, 3))   # 7 rows: ragged split 4 + 3
    f, df = parallel.update_hyperparameters_batch_sharded(gp, th, with_deriv=True)
    f_only = parallel.update_hyperparameters_batch_sharded(gp, th, with_deriv=False)
    nlocal = gp._dev_obj.calls.count("ll_batched")
    Xs = rs.rand(9, 2)
    gp.update_hyperparameters(np.array([1.3, 0.7, 1.1]))
    mean, std = parallel.predict_sharded(gp, Xs)
    # L4 drivers on the sharded batch call: ensemble sampler (walkers split by rank), lock-step multi-start optimiser
    # (starts split by rank), prediction over hyperparameter samples (samples split by rank)
    np.random.seed(100 + rank)                    # ranks deliberately start from DIFFERENT global RNG states
    sampler = gp.sample_hyperparameter_posterior(nwalkers=8, nsamp=3)
    ncalls0 = len(gp._dev_obj.calls)
    np.random.seed(7)                             # the user seeds the start draws identically (SPMD contract)
    gp.use_hyper_deriv = True
    res, nres = gp.optimize_hyperparameters(random_starts=4, opt_kwargs={"options": {"maxiter": 4}})
    gp.use_hyper_deriv = False
    opt_calls = gp._dev_obj.calls[ncalls0:]
    trace = sampler.chain[:, -1, :][:5]
    mc = gp.compute_from_MCMC(Xs[:4], flat_trace=trace)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), f=f, df=df, f_only=f_only, mean=mean, std=std, nlocal=nlocal,
             lo_hi=np.array(parallel.shard_bounds(7, rank, world)), chain=sampler.chain, lnp=sampler.lnprobability,
             opt_x=res.x, opt_fun=res.fun, nres=nres, opt_batched=opt_calls.count("ll_batched"),
             mc_mean=np.array(mc["mean"]), mc_std=np.array(mc["std"]))
    dist.destroy_process_group()


def test_theta_and_test_point_sharding_world2(tmp_path):
    import warnings
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "r0.npz")
    r1 = np.load(tmp_path / "r1.npz")
    # every rank holds the full, identical result
    for key in ("f", "df", "f_only", "mean", "std", "chain", "lnp", "opt_x", "opt_fun", "mc_mean", "mc_std"):
        assert np.array_equal(r0[key], r1[key]), key
    assert r0["chain"].shape == (8, 3, 3) and np.isfinite(r0["lnp"]).all()
    assert int(r0["nres"]) == 4 and int(r0["opt_batched"]) > 0 and np.isfinite(r0["opt_fun"])
    assert r0["mc_mean"].shape == (5, 4) and r0["mc_std"].shape == (5, 4)
    assert list(r0["lo_hi"]) == [0, 4] and list(r1["lo_hi"]) == [4, 7]
    # and it equals the unsharded evaluation
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    warnings.simplefilter("ignore")
    import gptools_b200 as g
    from fake_device import FakeDevice
    rs = np.random.RandomState(0)
    X = rs.rand(10, 2)
    k = g.SquaredExponentialKernel(num_dim=2, initial_params=[1.3, 0.7, 1.1], param_bounds=[(0, 10)] * 3)
    gp = g.GaussianProcess(k, X=X, y=np.sin(X).sum(1), err_y=0.05)
    gp._dev_obj = FakeDevice()
    th = np.array([1.3, 0.7, 1.1]) * np.exp(0.1 * np.random.RandomState(1).randn(7, 3))
    f, df = gp.update_hyperparameters_batch(th, with_deriv=True)
    assert np.array_equal(f, r0["f"]) and np.array_equal(df, r0["df"])
    Xs = rs.rand(9, 2)
    gp.update_hyperparameters(np.array([1.3, 0.7, 1.1]))
    mean, std = gp.predict(Xs)
    assert np.allclose(mean, r0["mean"], rtol=1e-12, atol=0) and np.allclose(std, r0["std"], rtol=1e-9, atol=0)
    mc = gp.compute_from_MCMC(Xs[:4], flat_trace=r0["chain"][:, -1, :][:5])
    assert np.allclose(np.array(mc["mean"]), r0["mc_mean"], rtol=1e-12) and np.allclose(np.array(mc["std"]), r0["mc_std"], rtol=1e-9)


def test_shard_bounds_cover_everything():
    from gptools_b200.parallel import shard_bounds
    for count in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            edges = [shard_bounds(count, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == count
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
