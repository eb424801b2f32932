"""Multi-GPU use of the hot path: one process per GPU (torchrun), units sharded by rank.

Only the two axes the path shards on naturally are split:
  * batched hyperparameter vectors theta (emcee walkers, optimizer restarts, ll grids): independent units,
    contiguous slice per rank, no data-path collective; ONE all-gather of the (1 + P) scalars per theta
    over NCCL/NVLink at the end so every rank sees the whole batch;
  * test points of predict: contiguous slice per rank, the (small) factorisation is recomputed on every
    rank from the replicated training set, one all-gather of mean / std.
Everything else (single large Cholesky, T K T^T, draw_sample) runs as independent replicas.

torch.distributed is plumbing only; on CPU (tests) the same code runs over gloo.
"""
import numpy as np


def shard_bounds(count, rank, world_size):
    """Contiguous [lo, hi) slice of ``count`` units owned by ``rank`` (first ``count % world`` ranks get one more)."""
    base, extra = divmod(int(count), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _dist():
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None
    return dist


def _all_gather_rows(local, counts, device=None):
    """All-gather row blocks of unequal length (padded to the largest block). ``local``: (n_local, C) float64."""
    import torch
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    nmax = max(counts)
    C = local.shape[1]
    use_cuda = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    buf = torch.zeros((nmax, C), dtype=torch.float64, device=dev)
    if local.shape[0] > 0:
        buf[:local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
    out = torch.empty((world * nmax, C), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, buf)
    out = out.cpu().numpy().reshape(world, nmax, C)
    return np.concatenate([out[r, :counts[r]] for r in range(world)], axis=0)


def update_hyperparameters_batch_sharded(gp, thetas, with_deriv=None):
    """``gp.update_hyperparameters_batch`` with the theta rows split contiguously over the ranks of the default
    process group and a single all-gather of the results.  Every rank passes the SAME ``thetas`` and gets the
    full (B,) / (B, P) result back."""
    thetas = np.atleast_2d(np.asarray(thetas, dtype=float))
    dist = _dist()
    if with_deriv is None:
        with_deriv = bool(gp.use_hyper_deriv)
    if dist is None or dist.get_world_size() == 1:
        return gp.update_hyperparameters_batch(thetas, with_deriv=with_deriv)
    rank, world = dist.get_rank(), dist.get_world_size()
    B = thetas.shape[0]
    counts = [shard_bounds(B, r, world)[1] - shard_bounds(B, r, world)[0] for r in range(world)]
    lo, hi = shard_bounds(B, rank, world)
    P = thetas.shape[1]
    if hi > lo:
        res = gp.update_hyperparameters_batch(thetas[lo:hi], with_deriv=with_deriv)
        if with_deriv:
            local = np.hstack([res[0][:, None], res[1]])
        else:
            local = np.asarray(res)[:, None]
    else:
        local = np.zeros((0, 1 + (P if with_deriv else 0)))
    full = _all_gather_rows(local, counts)
    if with_deriv:
        return full[:, 0], full[:, 1:]
    return full[:, 0]


def predict_sharded(gp, Xstar, n=0, return_std=True):
    """Predictive mean (and std) with the test points split over the ranks; each rank factors the replicated
    training set itself (identical bits on every rank) and one all-gather assembles the result."""
    Xstar = np.atleast_2d(np.asarray(Xstar, dtype=float))
    if gp.num_dim == 1 and Xstar.shape[0] == 1:
        Xstar = Xstar.T
    try:
        iter(n)
    except TypeError:
        n = n * np.ones(Xstar.shape, dtype=int)
    else:
        n = np.atleast_2d(np.asarray(n, dtype=int))
        if gp.num_dim == 1 and n.shape[0] == 1:
            n = n.T
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return gp.predict(Xstar, n=n, return_std=return_std)
    rank, world = dist.get_rank(), dist.get_world_size()
    Ms = Xstar.shape[0]
    counts = [shard_bounds(Ms, r, world)[1] - shard_bounds(Ms, r, world)[0] for r in range(world)]
    lo, hi = shard_bounds(Ms, rank, world)
    if hi > lo:
        res = gp.predict(Xstar[lo:hi], n=n[lo:hi], return_std=return_std)
        local = np.column_stack(res) if return_std else np.asarray(res)[:, None]
    else:
        local = np.zeros((0, 2 if return_std else 1))
    full = _all_gather_rows(local, counts)
    if return_std:
        return full[:, 0], full[:, 1]
    return full[:, 0]
