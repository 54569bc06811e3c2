"""
Utterance sharding for N GPUs: every stage of the separation path is per utterance,
so ranks split the batch axis and exchange NOTHING on the data path.  The only cross-rank traffic is
bookkeeping: the max-over-ranks step time and the gather of per-rank results on request.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_items, rank, world):
    """Contiguous, balanced split: the first n_items % world ranks carry one extra utterance."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError('rank %d of world %d' % (rank, world))
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value, device='cpu'):
    """max of a python float over all ranks (device-side for NCCL, host for gloo)"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_rows(local, n_total, device=None):
    """Concatenate per-rank row blocks (sized by shard_bounds) on every rank, in rank order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)
