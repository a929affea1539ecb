"""Host-side logic for N > 1 GPUs (one process per GPU, launched by torchrun).

The path shards by independent units: every rank owns whole audio streams with a full weight copy
(BASELINE.json configs[3]); there is no data-path collective.  What the ranks do exchange is bookkeeping:
who takes which stream, and the max-over-ranks of the device time (the bench contract).  These helpers are
backend-agnostic so the world_size-2 gloo tests exercise exactly what runs under NCCL on the GPU box.
"""
from typing import List, Sequence


def assign_streams(n_streams: int, world_size: int, rank: int) -> List[int]:
    """Contiguous, balanced partition of stream ids over ranks (the first n % world ranks get one extra)."""
    base, extra = divmod(n_streams, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def reduce_max(values: Sequence[float], dist=None, device=None) -> List[float]:
    """Element-wise MAX over ranks (device time and wall time of the timed region)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(v) for v in values]
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def reduce_sum(values: Sequence[float], dist=None, device=None) -> List[float]:
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(v) for v in values]
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t.tolist()]


def aggregate_rtf(audio_seconds_this_rank: float, elapsed_ms_this_rank: float, dist=None, device=None) -> dict:
    """Whole-job real-time factor: audio seconds processed by ALL ranks / MAX elapsed time over ranks."""
    total_audio = reduce_sum([audio_seconds_this_rank], dist, device)[0]
    max_ms = reduce_max([elapsed_ms_this_rank], dist, device)[0]
    return {"audio_s": total_audio, "elapsed_ms": max_ms, "rtf": total_audio / (max_ms / 1e3) if max_ms > 0 else 0.0}
