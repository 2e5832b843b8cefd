"""Multi-GPU: independent utterances are batch-sharded, one process per GPU,
weights replicated; there is no collective inside the sampling loop and exactly
one NCCL ``all_gather`` of the final mel.  The reference has
no distributed code (section 2.2) -- this is new, and deliberately minimal."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(B: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous near-equal split of B utterances over ``world`` ranks (first ranks get the remainder)."""
    if world <= 0:
        raise ValueError("world size must be positive")
    base, rem = divmod(B, world)
    out, lo = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((lo, lo + n))
        lo += n
    return out


def shard(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(t.shape[0], world)[rank]
    return t[lo:hi]


def gather_batch(local: torch.Tensor, B: int, group=None) -> torch.Tensor:
    """all_gather of per-rank batch shards (possibly ragged) back into [B, ...] on every rank."""
    world = dist.get_world_size(group)
    bounds = shard_bounds(B, world)
    nmax = max(hi - lo for lo, hi in bounds)
    pad = torch.zeros((nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    if all(hi - lo == nmax for lo, hi in bounds):
        return out
    return torch.cat([out[r * nmax: r * nmax + (hi - lo)] for r, (lo, hi) in enumerate(bounds)], dim=0)


def generate_mel_sharded(inference, sem_idx: torch.Tensor, num_steps: int = 4, temperature: float = 1.0,
                         x_T: Optional[torch.Tensor] = None, gather: bool = True, group=None) -> torch.Tensor:
    """Every rank holds the full ``sem_idx`` [B,S] (and optionally the full ``x_T``), computes its own
    contiguous batch shard with ``inference.generate_mel`` and, if ``gather``, all ranks receive
    the full [B, 2S, n_mels] mel.  With ``x_T`` given the result is bit-identical to the
    single-GPU call (the kernels are batch-invariant)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    B = sem_idx.shape[0]
    idx_loc = shard(sem_idx, rank, world)
    x_loc = shard(x_T, rank, world) if x_T is not None else None
    mel_loc = inference.generate_mel(idx_loc, num_steps, temperature, x_T=x_loc)
    return gather_batch(mel_loc, B, group) if gather else mel_loc
