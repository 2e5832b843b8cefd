"""CPU, world_size 2 over gloo: the batch-sharding host logic of edge_diffusion_tts_b200.dist."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from edge_diffusion_tts_b200.dist import RootGather, gather_batch, gather_to_root, generate_mel_sharded, shard_bounds


def test_shard_bounds():
    assert shard_bounds(256, 8) == [(i * 32, (i + 1) * 32) for i in range(8)]
    assert shard_bounds(5, 2) == [(0, 3), (3, 5)]
    assert shard_bounds(1, 4) == [(0, 1), (1, 1), (1, 1), (1, 1)]
    assert shard_bounds(0, 2) == [(0, 0), (0, 0)]


class _FakeInference:
    """Stands in for EdgeInference on CPU: a batch-invariant function of (sem_idx, x_T)."""

    def generate_mel(self, sem_idx, num_steps=4, temperature=1.0, x_T=None):
        return x_T * 2 + sem_idx.float().mean(dim=1)[:, None, None] + num_steps


def _worker(rank, world, port, B, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        idx = torch.randint(0, 512, (B, 6), generator=g)
        xT = torch.randn(B, 12, 80, generator=g)
        full = _FakeInference().generate_mel(idx, 4, x_T=xT)
        got = generate_mel_sharded(_FakeInference(), idx, 4, x_T=xT)
        ok = torch.equal(got, full)
        lo, hi = shard_bounds(B, world)[rank]
        ok = ok and torch.equal(gather_batch(full[lo:hi], B), full)
        # the consumer on one rank: gather to rank 0 only, synchronous and as the double-buffered asynchronous pipeline
        r = gather_to_root(full[lo:hi], B, dst=0)
        ok = ok and ((r is None) if rank else torch.equal(r, full))
        r = generate_mel_sharded(_FakeInference(), idx, 4, x_T=xT, dst=1)
        ok = ok and ((r is None) if rank != 1 else torch.equal(r, full))
        rg = RootGather(B, dst=0)
        for k in range(5):                                  # slots alternate; start() k+2 waits for gather k
            rg.start(full[lo:hi] + k)
            if k % 2:
                r = rg.finish()
                ok = ok and ((r is None) if rank else torch.equal(r, full + k))
        r = rg.finish()
        ok = ok and ((r is None) if rank else torch.equal(r, full + 4))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _run(B):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(2, port, B, out), nprocs=2, join=True)
        assert dict(out) == {0: True, 1: True}


def test_even_split_world2():
    _run(8)


def test_ragged_split_world2():
    _run(5)


def test_single_utterance_world2():
    _run(1)
