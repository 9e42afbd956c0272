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


def unit_index(n_units: int, world_size: int) -> torch.Tensor:
    """Position of unit u inside the rank-major gather buffer [G * per, ...]: row r * per + j holds unit j * G + r, so
    `buffer[unit_index(n, G)]` is in unit order.  Lets a consumer read single units without reordering the buffer."""
    per = (n_units + world_size - 1) // world_size
    u = torch.arange(n_units)
    return (u % world_size) * per + u // world_size


class ImageGather:
    """The path's one exchange step, taken off the critical path: `submit(local)` issues the all-gather of this rank's
    finished uint8 images asynchronously (NCCL's own stream) into one of two rank-major buffers and returns at once;
    the previous submission is waited for first, so a rank is never more than one call ahead of the slowest rank and
    never blocks on the collective it has just issued.  `wait()` returns the last gathered buffer (rank-major; unit u
    is row `unit_index(n_units, world)[u]`): no per-call reordering copy."""

    def __init__(self, n_units: int, rank: int, world_size: int, image_shape, device, group=None):
        self.n_units, self.rank, self.world, self.group = n_units, rank, world_size, group
        self.per = (n_units + world_size - 1) // world_size
        self.bufs = [torch.empty((world_size * self.per,) + tuple(image_shape), dtype=torch.uint8, device=device)
                     for _ in range(2)]
        self.shards = [torch.zeros((self.per,) + tuple(image_shape), dtype=torch.uint8, device=device) for _ in range(2)]
        self.pending = None     # (work handle, buffer index)
        self.turn = 0

    def submit(self, local: torch.Tensor) -> None:
        if local.dtype != torch.uint8:
            raise TypeError("images are gathered as uint8")
        done = self.wait()
        del done
        i = self.turn
        self.turn ^= 1
        shard = self.shards[i]
        shard[:local.shape[0]].copy_(local)       # static source buffer: the caller may overwrite `local` at once
        work = dist.all_gather_into_tensor(self.bufs[i], shard, group=self.group, async_op=True)
        self.pending = (work, i)

    def wait(self) -> Optional[torch.Tensor]:
        if self.pending is None:
            return None
        work, i = self.pending
        work.wait()      # (NCCL: makes the current stream wait for the collective; gloo: blocks the host)
        self.pending = None
        self.last = self.bufs[i]
        return self.last


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
