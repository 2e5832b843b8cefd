"""Multi-GPU over NCCL (pytest -m gpu on a box with >= 2 GPUs; skipped otherwise): batch shards gathered over NCCL equal
the single-GPU result BIT FOR BIT, for the all_gather and the gather-to-root paths,
even and ragged splits, fp32 and bf16 paths."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import __graft_entry__ as ge
        ge.build()
        import edge_diffusion_tts_b200 as E
        from edge_diffusion_tts_b200.dist import RootGather, generate_mel_sharded, shard
        from oracle import synth
        cfg = E.CFG(device=str(dev))
        dec = E.EdgeDiffusionDecoder(cfg).to(dev).eval()
        dec.load_state_dict(synth.synth_decoder_state(0), strict=True)
        inf = E.EdgeInference(cfg, E.DiffusionSchedule(cfg.diff_steps, device=dev), torch.nn.Identity(), dec)
        ok = True
        for prec, B, S in (("bf16", 8, 200), ("bf16", 5, 60), ("fp32", 3, 40)):
            dec.precision = prec
            idx = synth.synth_sem_idx(7, B, S).to(dev)
            xT = synth.synth_noise(7, B, 2 * S).to(dev)
            single = inf.generate_mel(idx, 4, x_T=xT)                       # every rank computes the whole batch itself
            every = generate_mel_sharded(inf, idx, 4, x_T=xT)               # all_gather: the full mel on every rank
            ok = ok and torch.equal(every, single)
            root = generate_mel_sharded(inf, idx, 4, x_T=xT, dst=0)         # gather: rank 0 only
            ok = ok and ((root is None) if rank else torch.equal(root, single))
            rg = RootGather(B, dst=0)                                       # the asynchronous, double-buffered form
            for k in range(3):
                rg.start(inf.generate_mel(shard(idx, rank, world), 4, x_T=shard(xT, rank, world)) + k)
            last = rg.finish()
            ok = ok and ((last is None) if rank else torch.equal(last, single + 2))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_generate_equals_single_gpu_bitwise_over_nccl():
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {r: True for r in range(world)}
