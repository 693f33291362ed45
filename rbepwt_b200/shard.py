"""Sharding of an image batch across the GPUs of one box: by image, no collective on the data path
(SURVEY.md section 8e -- the per-level DWT couples all regions of an image, so an image is never split).
torch.distributed carries only the timing barrier and the max-over-ranks of the measured time."""


def shard_range(total, world_size, rank):
    """Contiguous slice [lo, hi) of `total` images owned by `rank`; sizes differ by at most one."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, rem = divmod(int(total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value, device="cpu"):
    """MAX all-reduce of a python float (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_throughput(images_per_rank, steps, seconds_local, device="cpu"):
    """images/s of the whole job: every rank's images over the slowest rank's time."""
    import torch
    import torch.distributed as dist

    n = float(images_per_rank) * steps
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([n], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        n = float(t.item())
    return n / max_over_ranks(seconds_local, device)
