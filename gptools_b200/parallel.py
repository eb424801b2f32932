"""Multi-GPU use of the hot path: one process per GPU (torchrun), units sharded by rank.

Only the two axes the path shards on naturally are split:
  * batched hyperparameter vectors theta (emcee walkers, optimizer restarts, ll grids, MCMC-posterior prediction):
    independent units, contiguous slice per rank, no data-path collective; ONE all-gather of the (1 + P) scalars
    and the status word per theta over NCCL/NVLink at the end so every rank sees the whole batch.  On the NCCL
    backend the batched kernel writes ll / gradient / status straight into the all-gather SEND buffer (device
    memory owned by torch), the gathered buffer comes back to the host in one copy: no host round trip between
    the kernel and the collective;
  * test points of predict: contiguous slice per rank, mean / variance written by the library into device buffers
    that are gathered the same way.  The factorisation is replicated: every rank factors the (replicated) training
    set itself -- bit-identical across ranks, and all ranks finish at the time one would (broadcasting the factor
    from rank 0 instead leaves the other ranks idle for exactly that time and then moves 19 GB at config 4).
Everything else (single large Cholesky, T K T^T, draw_sample) runs as independent replicas.

torch.distributed is plumbing only (process group, NCCL communicator, device buffers); on CPU (tests) the same
code runs over gloo with host buffers.
"""
import numpy as np


def shard_bounds(count, rank, world_size):
    """Contiguous [lo, hi) slice of ``count`` units owned by ``rank`` (first ``count % world`` ranks get one more)."""
    base, extra = divmod(int(count), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_counts(count, world_size):
    return [shard_bounds(count, r, world_size)[1] - shard_bounds(count, r, world_size)[0] for r in range(world_size)]


def _dist():
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None
    return dist


def world():
    """(rank, world_size) of the default process group, (0, 1) without one."""
    dist = _dist()
    if dist is None:
        return 0, 1
    return dist.get_rank(), dist.get_world_size()


def _on_nccl(dist):
    return dist.get_backend() == "nccl"


def _all_gather_rows(local, counts):
    """All-gather row blocks of unequal length (padded to the largest block). ``local``: (n_local, C) float64 on the
    host.  Used by the gloo path and by callers whose per-rank results only exist on the host."""
    import torch
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return local
    world_size = dist.get_world_size()
    nmax = max(counts)
    C = local.shape[1]
    dev = torch.device("cuda", torch.cuda.current_device()) if _on_nccl(dist) else torch.device("cpu")
    buf = torch.zeros((nmax, C), dtype=torch.float64, device=dev)
    if local.shape[0] > 0:
        buf[:local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
    out = torch.empty((world_size * nmax, C), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, buf)
    out = out.cpu().numpy().reshape(world_size, nmax, C)
    return np.concatenate([out[r, :counts[r]] for r in range(world_size)], axis=0)


def broadcast_array(arr, src=0):
    """``arr`` of rank ``src`` on every rank (same shape and dtype float64 everywhere)."""
    import torch
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return arr
    dev = torch.device("cuda", torch.cuda.current_device()) if _on_nccl(dist) else torch.device("cpu")
    t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64)).to(dev)
    dist.broadcast(t, src=src)
    return t.cpu().numpy().reshape(np.shape(arr))


def shared_seed():
    """A random 31-bit seed drawn on rank 0 and broadcast: identical random streams on every rank."""
    seed = np.array([float(np.random.randint(0, 2 ** 31 - 1))])
    return int(broadcast_array(seed)[0])


def gather_result_lists(out, total):
    """compute_from_MCMC: every rank holds lists of equally shaped arrays for its slice of the ``total`` hyperparameter
    samples (entries of failed samples are missing); returns the concatenation over ranks in rank order."""
    dist = _dist()
    world_size = dist.get_world_size()
    merged = {}
    for key in sorted(out):
        vals = out[key]
        shape = np.array([len(vals)] + (list(np.shape(vals[0])) if len(vals) else []), dtype=float)
        # how many entries / which shape each rank holds (ranks with nothing report the shape of the others)
        meta = np.zeros(8)
        meta[:len(shape)] = shape
        meta[7] = len(shape)
        metas = _all_gather_rows(meta[None, :], [1] * world_size)
        ref = metas[np.argmax(metas[:, 0])]
        nd = int(ref[7])
        if nd == 0 or ref[0] == 0:
            merged[key] = []
            continue
        item_shape = tuple(int(v) for v in ref[1:nd])
        width = int(np.prod(item_shape)) if item_shape else 1
        counts = [int(m[0]) for m in metas]
        local = np.asarray(vals, dtype=float).reshape(len(vals), width) if len(vals) else np.zeros((0, width))
        full = _all_gather_rows(local, counts)
        merged[key] = [full[i].reshape(item_shape) for i in range(full.shape[0])]
    return merged


def _device_handles(gp):
    """(Device, torch.device, current torch stream) with the library pointed at torch's current stream, so that kernel
    launches, torch copies and NCCL collectives are ordered on ONE stream."""
    import torch
    dev = gp._dev()
    tdev = torch.device("cuda", dev.device)
    torch.cuda.set_device(tdev)
    stream = torch.cuda.current_stream(tdev)
    dev.set_stream(stream.cuda_stream)
    return dev, tdev, stream


def _theta_batch_device(gp, plan, counts, while_running=None):
    """The rows of ``plan`` (this rank's slice of a theta batch, prepared by ``gp._batch_prepare``) on this rank's GPU,
    results gathered device to device.  Returns (ll, grad, status, logp) for the WHOLE batch.

    Send-buffer layout per rank (bytes): ll  nmax x f64 | grad  nmax x P x f64 | log-prior  nmax x f64 | status
    nmax x i32 (padded to 8): the kernel's output pointers are views of it; the log-prior of the local rows (-inf where
    the row is outside the prior support or not evaluable) is copied in next to them, so that the host-side prior
    arithmetic is done once per row across the job instead of once per row on every rank."""
    import torch
    dist = _dist()
    world_size = dist.get_world_size()
    dev, tdev, stream = _device_handles(gp)
    grad_idx = plan["grad_idx"]
    P = len(grad_idx) if grad_idx else 0
    nmax = max(counts)
    off_grad = 8 * nmax
    off_logp = off_grad + 8 * nmax * P
    off_status = off_logp + 8 * nmax
    nbytes = off_status + 8 * ((4 * nmax + 7) // 8)
    send = torch.zeros(nbytes, dtype=torch.uint8, device=tdev)
    nloc = plan["B"]
    if nloc > 0:
        th = torch.from_numpy(np.ascontiguousarray(plan["full_eval"])).to(tdev, non_blocking=True)
        base = send.data_ptr()
        dev.ll_batched_dev(nloc, th.data_ptr(), base, base + off_status, d_grad=(base + off_grad) if P else 0,
                           grad_idx=grad_idx)
        logp = np.where(plan["ok"], plan["logp"], -np.inf)
        send[off_logp:off_logp + 8 * nloc].copy_(torch.from_numpy(logp.view(np.uint8)), non_blocking=True)
    recv = torch.empty(world_size * nbytes, dtype=torch.uint8, device=tdev)
    dist.all_gather_into_tensor(recv, send)
    extra = while_running() if while_running is not None else None  # host work hidden behind the kernel and the gather
    host = recv.cpu().numpy().reshape(world_size, nbytes)
    if min(counts) == nmax:  # equal slices: the gathered blocks are contiguous per field
        ll = host[:, :8 * nmax].copy().view(np.float64).reshape(-1)
        status = host[:, off_status:off_status + 4 * nmax].copy().view(np.int32).reshape(-1)
        logp_all = host[:, off_logp:off_logp + 8 * nmax].copy().view(np.float64).reshape(-1)
        grad = host[:, off_grad:off_grad + 8 * nmax * P].copy().view(np.float64).reshape(-1, P) if P else None
        return ll, grad, status, logp_all, extra
    ll = np.concatenate([host[r, :8 * counts[r]].view(np.float64) for r in range(world_size)])
    status = np.concatenate([host[r, off_status:off_status + 4 * counts[r]].view(np.int32) for r in range(world_size)])
    logp_all = np.concatenate([host[r, off_logp:off_logp + 8 * counts[r]].view(np.float64) for r in range(world_size)])
    grad = None
    if P:
        grad = np.concatenate([host[r, off_grad:off_grad + 8 * counts[r] * P].view(np.float64).reshape(counts[r], P)
                               for r in range(world_size)], axis=0)
    return ll, grad, status, logp_all, extra


def update_hyperparameters_batch_sharded(gp, thetas, with_deriv=None):
    """``gp.update_hyperparameters_batch`` with the theta rows split contiguously over the ranks of the default
    process group and a single all-gather of the results.  Every rank passes the SAME ``thetas`` and gets the
    full (B,) / (B, P) result back (bit-identical on every rank)."""
    thetas = np.atleast_2d(np.asarray(thetas, dtype=float))
    dist = _dist()
    if with_deriv is None:
        with_deriv = bool(gp.use_hyper_deriv)
    if dist is None or dist.get_world_size() == 1:
        return gp.update_hyperparameters_batch(thetas, with_deriv=with_deriv)
    rank, world_size = dist.get_rank(), dist.get_world_size()
    B = thetas.shape[0]
    counts = shard_counts(B, world_size)
    lo, hi = shard_bounds(B, rank, world_size)
    if gp._batchable(with_deriv):
        if _on_nccl(dist) and (gp.mu is None or gp.mu.num_free_params == 0):
            # every rank prepares (parameter rows, prior, validity) ITS slice only; the log-prior travels with the results
            plan = gp._batch_prepare(thetas[lo:hi] if hi > lo else thetas[:1], with_deriv)
            if hi == lo:
                plan["B"] = 0  # more ranks than rows: this rank only takes part in the gather
            # the frame of the whole-batch plan (parameter rows of every theta) is built while the GPUs work
            ll, grad, status, logp, full = _theta_batch_device(
                gp, plan, counts, while_running=lambda: gp._batch_plan_from_gathered(thetas, with_deriv, None))
            full["logp"] = logp
            full["ok"] = np.isfinite(logp)
            return gp._batch_finish(full, ll, grad, status)
        plan = gp._batch_prepare(thetas, with_deriv)  # host arithmetic over the whole batch, identical on all ranks
        # per-theta mean-function residuals / alpha (host arrays in the C-ABI), or the gloo backend
        P = len(plan["grad_idx"]) if plan["grad_idx"] else 0
        M = len(gp.y)
        na = M if plan["need_alpha"] else 0
        local = np.zeros((hi - lo, 2 + P + na))
        if hi > lo:
            res = plan["dev"].ll_batched(plan["full_eval"][lo:hi], grad_idx=plan["grad_idx"],
                                         y_batch=None if plan["y_batch"] is None else plan["y_batch"][lo:hi],
                                         return_alpha=plan["need_alpha"])
            local[:, 0] = res[0]
            local[:, 1] = res[2]
            if P:
                local[:, 2:2 + P] = res[1]
            if na:
                local[:, 2 + P:] = res[3]
        full = _all_gather_rows(local, counts)
        return gp._batch_finish(plan, full[:, 0], full[:, 2:2 + P] if P else None, full[:, 1].astype(np.int32),
                                full[:, 2 + P:] if na else None)
    # kernels the batched device entry does not take: one theta at a time on the owning rank
    nfree = thetas.shape[1]
    if hi > lo:
        res = gp.update_hyperparameters_batch(thetas[lo:hi], with_deriv=with_deriv)
        local = np.hstack([res[0][:, None], res[1]]) if with_deriv else np.asarray(res)[:, None]
    else:
        local = np.zeros((0, 1 + (nfree if with_deriv else 0)))
    full = _all_gather_rows(local, counts)
    if with_deriv:
        return full[:, 0], full[:, 1:]
    return full[:, 0]


def _canonical_test_points(gp, Xstar, n):
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
    return Xstar, n


def predict_sharded(gp, Xstar, n=0, return_std=True):
    """Predictive mean (and std) with the test points split over the ranks; each rank factors the replicated
    training set itself (identical bits on every rank) and one all-gather assembles the result."""
    Xstar, n = _canonical_test_points(gp, Xstar, n)
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return gp.predict(Xstar, n=n, return_std=return_std)
    rank, world_size = dist.get_rank(), dist.get_world_size()
    Ms = Xstar.shape[0]
    counts = shard_counts(Ms, world_size)
    lo, hi = shard_bounds(Ms, rank, world_size)
    if _on_nccl(dist) and gp._device_mode() and gp.mu is None:
        import torch
        gp.compute_K_L_alpha_ll()
        dev, tdev, stream = _device_handles(gp)
        nmax = max(counts)
        C = 2 if return_std else 1
        send = torch.zeros((C, nmax), dtype=torch.float64, device=tdev)
        if hi > lo:
            gp.k._check_orders(gp.n, n[lo:hi])
            Xs_d, ns_d = gp.k.device_points(Xstar[lo:hi], n[lo:hi])
            dev.predict_dev(Xs_d, ns_d, send[0].data_ptr(), send[1].data_ptr() if return_std else 0)
        recv = torch.empty((world_size, C, nmax), dtype=torch.float64, device=tdev)
        dist.all_gather_into_tensor(recv, send)
        host = recv.cpu().numpy()
        mean = np.concatenate([host[r, 0, :counts[r]] for r in range(world_size)])
        if not return_std:
            return mean
        with np.errstate(invalid="ignore"):
            std = np.sqrt(np.concatenate([host[r, 1, :counts[r]] for r in range(world_size)]))
        return mean, std
    if hi > lo:
        res = gp.predict(Xstar[lo:hi], n=n[lo:hi], return_std=return_std)
        local = np.column_stack(res) if return_std else np.asarray(res)[:, None]
    else:
        local = np.zeros((0, 2 if return_std else 1))
    full = _all_gather_rows(local, counts)
    if return_std:
        return full[:, 0], full[:, 1]
    return full[:, 0]
