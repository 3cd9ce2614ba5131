"""Data-parallel plumbing: frames shard by contiguous blocks across ranks.

Every frame is coded independently (the reference runs one sess.run per frame, cmrl.py:698-708), so inference
needs NO collective; the only cross-rank traffic is the max-over-ranks of the timing and -- in training -- one
all-reduce of the flat gradient buffer.  Works with any torch.distributed backend (nccl on B200, gloo in the
CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def frame_shard(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of the contiguous block of ceil(n/world) frames owned by `rank` (keeps utterance-adjacent
    frames together for the later overlap-add step)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    per = -(-n_frames // world)
    start = min(rank * per, n_frames)
    return start, min(start + per, n_frames)


def max_over_ranks(value: float, device=None) -> float:
    """MAX all-reduce of a scalar (the timing rule: a multi-GPU step takes as long as its slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
