"""Data-parallel plumbing of the image sweep: the path shards by independent units (one image / one
(identity, LoRA-variant) pair, each with its own seed, `inference_ID-Booth.py:111`), weights are replicated, there is
no collective inside the 30-step loop, and the only exchange is one all-gather of the finished uint8 images.
Works on any `torch.distributed` backend (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


def shard_units(n_units: int, rank: int, world_size: int) -> List[int]:
    """Unit indices of `rank`: round-robin (r, r + G, ...), so every rank gets ceil or floor of n / G units and the
    assignment is a pure function of (n, rank, G) -- every rank can replay it without communication."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return list(range(rank, n_units, world_size))


def unit_seed(unit: int, base_seed: int = 0) -> int:
    """Seed of a unit's generator: independent of the rank / world size that happens to process it."""
    return base_seed + unit


def gather_images(local: torch.Tensor, n_units: int, rank: int, world_size: int, group=None,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """local: uint8 [n_local, H, W, 3] images of shard_units(n_units, rank, world) in that order -> uint8
    [n_units, H, W, 3] in unit order on every rank.  One all_gather_into_tensor of equal-sized (padded) shards."""
    if local.dtype != torch.uint8:
        raise TypeError("images are gathered as uint8")
    if world_size == 1:
        return local
    per = (n_units + world_size - 1) // world_size
    shard = local
    if local.shape[0] < per:   # the last ranks may hold one unit less: pad so the collective is regular
        pad = torch.zeros((per - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        shard = torch.cat([local, pad], 0)
    shard = shard.contiguous()
    if out is None:
        out = torch.empty((world_size * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, shard, group=group)
    # rank-major [G, per, ...] -> unit order u = j * G + r
    full = out.view(world_size, per, *local.shape[1:]).transpose(0, 1).reshape(world_size * per, *local.shape[1:])
    return full[:n_units]
