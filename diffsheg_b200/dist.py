"""Multi-GPU: clips / batch rows are independent, so ranks shard the batch with no data-path
collective and exchange the generated motion with ONE all-gather at the end.
The reference has no tensor collective on this path (each rank np.saves its own files,
trainers/ddpm_show_trainer.py:924-934)."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous [lo, hi) slice of n items for `rank`; sizes differ by at most one."""
    per, extra = divmod(n, world)
    lo = rank * per + min(rank, extra)
    return lo, lo + per + (1 if rank < extra else 0)


def shard_batch(tensors, rank, world):
    """Slice every tensor (or dict of tensors) along dim 0 for this rank."""
    def cut(t):
        if isinstance(t, dict):
            return {k: cut(v) for k, v in t.items()}
        lo, hi = shard_range(t.shape[0], rank, world)
        return t[lo:hi]
    return [cut(t) for t in tensors]


def gather_motion(local, total, group=None):
    """All-gather the per-rank [b_r, T, D] results into [total, T, D] on every rank (NCCL on GPUs)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    if local.is_cuda and dist.get_backend(group) == "gloo":   # single-GPU hosts testing the sharded path: exchange through the host
        return gather_motion(local.cpu(), total, group).to(local.device)
    world = dist.get_world_size(group)
    sizes = [shard_range(total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    if all(hi - lo == mx for lo, hi in sizes):
        out = local.new_empty((world * mx,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = local.new_zeros((mx,) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)
