"""Frame-parallel sharding of a detection job across ranks (one process per GPU, no data-path collective:
frames are independent, reference src/DynamicProgram.cpp:80-83)."""


def frame_range(rank, world, frames_per_rank):
    """Global frame indices processed by `rank` (weak scaling: every rank owns `frames_per_rank` frames)."""
    if not (0 <= rank < world) or frames_per_rank < 0:
        raise ValueError("bad rank/world")
    return range(rank * frames_per_rank, (rank + 1) * frames_per_rank)


def split_frames(n_frames, world):
    """Strong-scaling split of a fixed job: contiguous blocks, sizes differ by at most one."""
    base, extra = divmod(n_frames, world)
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append(range(start, start + n))
        start += n
    return out


def max_over_ranks(value, device=None):
    """Max of a per-rank scalar over the default process group (timing = slowest rank); identity without one."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
