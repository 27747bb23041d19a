"""Instance sharding across the GPUs of one box (one process per GPU).

The solve needs no collective: every batch row owns its Q, A0, p, zl, zu and state, only the LSTM
weights are shared (replicated, read-only).  These helpers split a batch into contiguous per-rank
chunks and gather the per-iteration traces for reporting; they work on any backend (NCCL on GPUs, gloo
in the CPU test-suite).
"""
import torch
import torch.distributed as dist


def shard_range(batch, rank, world):
    """Contiguous chunk [lo, hi) of `batch` instances owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(int(batch), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def balanced_shares(total, rates, cap=None):
    """Split `total` independent instances over ranks in proportion to their measured rates (instances/s): the GPUs of
    one box differ by several per cent under the power cap, and with equal shards the job runs at the pace of the slowest.
    Largest-remainder rounding; every rank gets at least one instance when there are enough of them (a rank may get 0
    when total < number of ranks), at most `cap`.  The shares ALWAYS sum to `total`; if the caps make that impossible a
    ValueError is raised -- instances are never silently dropped or double-counted."""
    n = len(rates)
    total = int(total)
    if total < 0:
        raise ValueError("balanced_shares: negative total")
    if cap is not None and int(cap) * n < total:
        raise ValueError(f"balanced_shares: {n} ranks x cap {cap} cannot hold {total} instances")
    tot = float(sum(rates))
    if tot <= 0 or any(r <= 0 for r in rates):
        base, extra = divmod(total, n)
        shares = [base + (1 if i < extra else 0) for i in range(n)]
        ideal = [float(s_) for s_ in shares]
    else:
        ideal = [total * r / tot for r in rates]
        floor = 1 if total >= n else 0
        shares = [max(floor, int(x)) for x in ideal]
    if cap is not None:
        shares = [min(s_, int(cap)) for s_ in shares]
    order = sorted(range(n), key=lambda i: ideal[i] - int(ideal[i]), reverse=True)
    floor = 1 if total >= n else 0
    while sum(shares) < total:                      # terminates: cap * n >= total was checked above
        for i in order:
            if sum(shares) < total and (cap is None or shares[i] < cap):
                shares[i] += 1
    while sum(shares) > total:                      # terminates: floor * n <= total
        for i in reversed(order):
            if sum(shares) > total and shares[i] > floor:
                shares[i] -= 1
    assert sum(shares) == total and all(s_ >= 0 for s_ in shares)
    return shares


def balance_by_rate(total, my_rate, cap=None, group=None):
    """All-gather every rank's measured rate and return (my_share, all_shares).  The only communication of a sharded
    solve, once at set-up; tiny (one float per rank)."""
    world = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.tensor([float(my_rate)], dtype=torch.float64, device=dev)
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    shares = balanced_shares(total, [float(o.item()) for o in out], cap)
    return shares[dist.get_rank(group)], shares


def shard_instances(tensors, rank, world):
    """Slice every [B, ...] tensor of a dict (or tuple) to this rank's chunk."""
    if isinstance(tensors, dict):
        B = next(iter(tensors.values())).shape[0]
        lo, hi = shard_range(B, rank, world)
        return {k: v[lo:hi].contiguous() for k, v in tensors.items()}
    B = tensors[0].shape[0]
    lo, hi = shard_range(B, rank, world)
    return tuple(v[lo:hi].contiguous() for v in tensors)


def gather_batch(local, batch, dim=0, group=None, sizes=None):
    """All-gather per-rank chunks along `dim` into the full batch, on every rank.  `sizes` = the per-rank chunk sizes
    (e.g. the shares of `balance_by_rate`); default: the equal split of `shard_range`."""
    world = dist.get_world_size(group)
    if sizes is None:
        sizes = [shard_range(batch, r, world)[1] - shard_range(batch, r, world)[0] for r in range(world)]
    sizes = [int(s_) for s_ in sizes]
    if len(sizes) != world or sum(sizes) != int(batch):
        raise ValueError(f"gather_batch: sizes {sizes} do not describe a batch of {batch} over {world} ranks")
    if local.shape[dim] != sizes[dist.get_rank(group)]:
        raise ValueError(f"gather_batch: this rank holds {local.shape[dim]} rows, expected {sizes[dist.get_rank(group)]}")
    mx = max(sizes)
    loc = local.movedim(dim, 0).contiguous()
    pad = torch.zeros((mx,) + tuple(loc.shape[1:]), dtype=loc.dtype, device=loc.device)
    pad[: loc.shape[0]] = loc
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    full = torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)
    return full.movedim(0, dim)


_COMMS = {}


def nccl_comm(group=None, device=None):
    """A raw NCCL communicator over the ranks of `group` for the library's own collective (`iadmm_allreduce_grads`): rank 0
    creates the unique id (`iadmm_nccl_unique_id`), torch.distributed only carries its 128 bytes to the other ranks (plumbing),
    every rank joins on its current CUDA device (`iadmm_nccl_comm_init`).  Returns the ncclComm_t as an integer address
    (cached per group)."""
    import ctypes
    from . import _lib
    key = id(group) if group is not None else 0
    if key not in _COMMS:
        L = _lib.lib()
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        buf = ctypes.create_string_buffer(128)
        if rank == 0:
            _lib.check(L.iadmm_nccl_unique_id(buf))
        uid = [bytes(buf.raw) if rank == 0 else None]
        dist.broadcast_object_list(uid, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        comm = ctypes.c_void_p()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        with torch.cuda.device(dev):
            _lib.check(L.iadmm_nccl_comm_init(ctypes.byref(comm), world, ctypes.create_string_buffer(uid[0], 128), rank))
        _COMMS[key] = comm.value
    return _COMMS[key]


def allreduce_gradients(module, group=None, local_batch=None, comm=None):
    """Data-parallel training: combine the LSTM weight gradients of the ranks with ONE
    all-reduce of the flat gradient buffer (2,570,601 floats at h=800, K=100) per TBPTT window, then the
    unchanged Adam step runs on every rank.  Each rank's loss is the mean over ITS instances (main.py:347), so the
    gradient of the mean over the concatenated batch is sum_r (B_r / B) * grad_r: pass `local_batch` = B_r when the
    shards are unequal (`shard_range` remainders, `balance_by_rate`); without it equal shards are assumed (plain
    average).  The flat buffer covers EVERY parameter (zeros where `.grad` is None), so all ranks issue the same
    collective whatever their local graph touched.  NCCL on GPUs, gloo in the CPU tests.
    `comm` = `nccl_comm(group)`: the all-reduce is then issued by the library itself (`iadmm_allreduce_grads` of the C ABI:
    ncclAllReduce on the current stream) instead of torch.distributed; same sum, same result."""
    world = dist.get_world_size(group)
    params = list(module.parameters())
    if not params or world == 1:
        return
    dev = params[0].device
    total = sum(p.numel() for p in params)
    # one extra slot carries the local batch size, so the weights need no second collective
    flat = torch.zeros((total + 1,), dtype=params[0].dtype, device=dev)
    w = float(local_batch) if local_batch is not None else 1.0
    off = 0
    for p in params:
        k = p.numel()
        if p.grad is not None:
            flat[off:off + k].copy_(p.grad.reshape(-1))
        off += k
    flat[:total] *= w
    flat[total] = w
    if comm is not None:
        from . import _lib
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().iadmm_allreduce_grads(_lib.ptr(flat), flat.numel(), 1.0, comm, _lib.stream_ptr()))
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat[:total] /= flat[total]
    off = 0
    for p in params:
        k = p.numel()
        g = flat[off:off + k].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += k
