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
    Largest-remainder rounding; every rank gets at least one instance, at most `cap`; the shares sum to `total`
    (if the caps allow it)."""
    n = len(rates)
    tot = float(sum(rates))
    if tot <= 0 or any(r <= 0 for r in rates):
        base, extra = divmod(int(total), n)
        return [base + (1 if i < extra else 0) for i in range(n)]
    ideal = [total * r / tot for r in rates]
    shares = [max(1, int(x)) for x in ideal]
    if cap is not None:
        shares = [min(s_, int(cap)) for s_ in shares]
    order = sorted(range(n), key=lambda i: ideal[i] - int(ideal[i]), reverse=True)
    k = 0
    while sum(shares) < total and k < 4 * n:
        i = order[k % n]
        if cap is None or shares[i] < cap:
            shares[i] += 1
        k += 1
    k = 0
    while sum(shares) > total and k < 4 * n:
        i = order[-1 - (k % n)]
        if shares[i] > 1:
            shares[i] -= 1
        k += 1
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


def gather_batch(local, batch, dim=0, group=None):
    """All-gather per-rank chunks (sizes from shard_range) along `dim` into the full batch, on every rank."""
    world = dist.get_world_size(group)
    sizes = [shard_range(batch, r, world)[1] - shard_range(batch, r, world)[0] for r in range(world)]
    mx = max(sizes)
    loc = local.movedim(dim, 0).contiguous()
    pad = torch.zeros((mx,) + tuple(loc.shape[1:]), dtype=loc.dtype, device=loc.device)
    pad[: loc.shape[0]] = loc
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    full = torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)
    return full.movedim(0, dim)


def allreduce_gradients(module, group=None):
    """Data-parallel training: average the LSTM weight gradients over the ranks with ONE
    all-reduce of the flat gradient buffer (2,570,601 floats at h=800, K=100) per TBPTT window, then the
    unchanged Adam step runs on every rank.  The loss is a batch mean (main.py:347), so with equal shards
    this equals single-process training on the concatenated batch.  NCCL on GPUs, gloo in the CPU tests."""
    world = dist.get_world_size(group)
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if not grads or world == 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= world
    off = 0
    for g in grads:
        k = g.numel()
        g.copy_(flat[off:off + k].view_as(g))
        off += k
