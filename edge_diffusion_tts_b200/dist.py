"""Multi-GPU: independent utterances are batch-sharded, one process per GPU,
weights replicated; there is no collective inside the sampling loop and exactly
one NCCL collective for the final mel: ``gather_to_root`` (the
consumer -- a vocoder / writer -- sits on one rank: every other rank sends its shard
once, nobody receives what it does not need; asynchronous on NCCL's own stream so the
gather of call k runs under call k + 1) or ``gather_batch`` (all ranks receive all of
it: only when every rank really consumes the whole batch).  The reference has no
distributed code (section 2.2) -- this is new, and deliberately minimal."""
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


class RootGather:
    """Gather of per-rank batch shards onto ONE rank (``dst``), asynchronous and double-buffered.

    ``start(local)`` enqueues the NCCL gather behind the work already on the current stream and returns at once: it
    runs on the process group's own stream, so the next generate on the compute stream overlaps it.  ``finish()`` makes
    the current stream wait for the outstanding gather and returns the full ``[B, ...]`` tensor on ``dst`` (``None``
    elsewhere).  Two receive buffers alternate, so ``start`` may be called again before the previous result is read;
    it first waits for the gather that last used the same slot.  Ragged splits are padded to the largest shard on the
    wire and cut on arrival."""

    def __init__(self, B: int, dst: int = 0, group=None):
        self.B, self.dst, self.group = B, dst, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.bounds = shard_bounds(B, self.world)
        self.nmax = max(hi - lo for lo, hi in self.bounds)
        self._recv = [None, None]
        self._send = [None, None]
        self._work = [None, None]
        self._k = 0
        self._last = None

    def _buffers(self, local: torch.Tensor, slot: int):
        tail = tuple(local.shape[1:])
        if self.rank == self.dst and (self._recv[slot] is None or self._recv[slot].shape[2:] != tail
                                      or self._recv[slot].dtype != local.dtype):
            self._recv[slot] = torch.empty((self.world, self.nmax) + tail, dtype=local.dtype, device=local.device)
        if local.shape[0] != self.nmax:
            if self._send[slot] is None or self._send[slot].shape[1:] != tail:
                self._send[slot] = torch.zeros((self.nmax,) + tail, dtype=local.dtype, device=local.device)
            self._send[slot][: local.shape[0]].copy_(local)
            return self._send[slot]
        return local.contiguous()

    def start(self, local: torch.Tensor) -> None:
        slot = self._k % 2
        self._k += 1
        if self._work[slot] is not None:
            self._work[slot].wait()
        src = self._buffers(local, slot)
        lst = list(self._recv[slot].unbind(0)) if self.rank == self.dst else None
        self._work[slot] = dist.gather(src, lst, dst=self.dst, group=self.group, async_op=True)
        self._last = slot

    def finish(self) -> Optional[torch.Tensor]:
        slot = self._last
        if slot is None:
            raise RuntimeError("RootGather.finish() without start()")
        if self._work[slot] is not None:
            self._work[slot].wait()
            self._work[slot] = None
        if self.rank != self.dst:
            return None
        buf = self._recv[slot]
        if all(hi - lo == self.nmax for lo, hi in self.bounds):
            return buf.view((self.world * self.nmax,) + tuple(buf.shape[2:]))
        return torch.cat([buf[r, : hi - lo] for r, (lo, hi) in enumerate(self.bounds)], dim=0)


def gather_to_root(local: torch.Tensor, B: int, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """One synchronous gather of batch shards onto rank ``dst``: ``[B, ...]`` there, ``None`` on the other ranks."""
    g = RootGather(B, dst, group)
    g.start(local)
    return g.finish()


def generate_mel_sharded(inference, sem_idx: torch.Tensor, num_steps: int = 4, temperature: float = 1.0,
                         x_T: Optional[torch.Tensor] = None, gather=True, group=None, dst: Optional[int] = None):
    """Every rank holds the full ``sem_idx`` [B,S] (and optionally the full ``x_T``) and computes its own contiguous
    batch shard with ``inference.generate_mel``.  ``gather=True``: the full [B, 2S, n_mels] mel -- on every rank
    (``dst=None``, all_gather) or only on rank ``dst`` (gather; the other ranks get ``None``); ``gather=False``: the
    local shard.  With ``x_T`` given the result is bit-identical to the single-GPU call (the kernels are
    batch-invariant; tests/test_gpu_multi.py checks it over NCCL)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    B = sem_idx.shape[0]
    idx_loc = shard(sem_idx, rank, world)
    x_loc = shard(x_T, rank, world) if x_T is not None else None
    mel_loc = inference.generate_mel(idx_loc, num_steps, temperature, x_T=x_loc)
    if not gather:
        return mel_loc
    if dst is not None:
        return gather_to_root(mel_loc, B, dst, group)
    return gather_batch(mel_loc, B, group)
