"""Stream sharding for multi-GPU runs: camera streams are independent, so each rank (one process
per GPU) owns a fixed set of streams and no data-path collective exists.  torch.distributed is used only for the
barrier around the timed region and the max-over-ranks of the device time."""
from __future__ import annotations

from typing import List


def stream_ids_for_rank(streams_per_gpu: int, rank: int) -> List[int]:
    """Global ids of the streams rank `rank` owns (weak scaling: every rank holds `streams_per_gpu` streams)."""
    return list(range(rank * streams_per_gpu, (rank + 1) * streams_per_gpu))


def shard_fixed_total(total_streams: int, world: int, rank: int) -> List[int]:
    """Strong-scaling variant: `total_streams` streams split as evenly as possible (stream s -> rank s mod world)."""
    return [s for s in range(total_streams) if s % world == rank]


def reduce_max(values, dist=None, device=None):
    """max over ranks of a list of floats (device times in ms); identity without a process group"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(v) for v in values]
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]
