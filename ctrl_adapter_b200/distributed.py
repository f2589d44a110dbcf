"""Batch-axis sharding of the denoising job over the GPUs of one node.

Every sample (image / clip with its CFG pair) is independent, so ranks never communicate inside the loop; the only
collective of a job is one all-gather of the final latents (NCCL over NVLink on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of the user batch: the first (global_batch % world) ranks take one extra sample."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    base, extra = divmod(global_batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def rank_seed(base_seed: int, rank: int) -> int:
    return base_seed + rank


def gather_latents(local: torch.Tensor, global_batch: int) -> torch.Tensor:
    """All-gather the per-rank final latents (dim 0 = local samples) into the global batch order on every rank."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(global_batch, r, world) for r in range(world)]
    max_n = max(e - s for s, e in sizes)
    pad = torch.zeros((max_n,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * max_n,) + tuple(pad.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)  # concatenated layout (accepted by both NCCL and gloo)
    out = out.reshape((world, max_n) + tuple(pad.shape[1:]))
    return torch.cat([out[r, : e - s] for r, (s, e) in enumerate(sizes)], dim=0)


def max_over_ranks(value: float, device) -> float:
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
