"""Host-side logic of the N > 1 path: one process per GPU, reads sharded by rank, no collective on the data
path.  torch.distributed is used only for the barrier around the timed region and for the max-over-ranks of the
device-measured times (backend "nccl" on the GPU box, "gloo" in the CPU tests)."""
from __future__ import annotations

import os


def rank_env():
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when launched directly."""
    g = lambda k, d: int(os.environ.get(k, d) or d)
    return g("RANK", 0), g("LOCAL_RANK", 0), g("WORLD_SIZE", 1)


def shard_range(n: int, rank: int, world: int):
    """Contiguous range [begin, end) of rank `rank` when n units (read pairs of a block) are dealt to `world` GPUs;
    ranges differ by at most one unit and concatenate to [0, n) in rank order (results are merged by pair index)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    q, r = divmod(n, world)
    begin = rank * q + min(rank, r)
    return begin, begin + q + (1 if rank < r else 0)


def shard_seed(base_seed: int, rank: int) -> int:
    """Weak scaling: every rank synthesises its own batch of the same shape."""
    return base_seed + rank


def max_over_ranks(values, dist=None, device=None):
    """Element-wise maximum over all ranks of a list of floats (times measured on each rank's device)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(v) for v in values]
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def whole_job_rate(units_per_rank: int, steps: int, world: int, max_ms: float) -> float:
    """BASELINE metric: units all ranks processed / the slowest rank's time."""
    return world * units_per_rank * steps / (max_ms * 1e-3)
