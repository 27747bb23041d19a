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
