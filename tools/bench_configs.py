"""Throughput of the other BASELINE.json configurations on ONE B200 (the bench.py headline is config 3).
Prints one JSON line per configuration; synthetic inputs and weights (oracle/synth.py seeds), bf16 tensor-core path,
CUDA-event timing after warm-up, inputs resident in HBM.  Usage (GPU box): python tools/bench_configs.py > gpurun_out/configs.jsonl

  cfg2  1-step generation, batch 64: 768-d features -> proj -> VQ -> decoder (T = 800)
  cfg4  1000-step DDPM ancestral loop, batch 32, T = 800 (CUDA graph of 50 iterations replayed 20 times)
  cfg5  4-step DDIM on long utterances, T = 3000 / S = 1500: the per-GPU share of batch 128 over 8 GPUs (16) and the whole
        batch 128 on one GPU
  dpm   DPM-Solver++ order 2, 10 steps, sem_features conditioning, batch 256, T = 800
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__ as ge

ge.build()
import edge_diffusion_tts_b200 as E
from oracle import synth

DEV = "cuda:0"


def timed(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    cfg = E.CFG(device=DEV)
    dec = E.EdgeDiffusionDecoder(cfg).to(DEV).eval()
    dec.load_state_dict(synth.synth_decoder_state(0), strict=True)
    dec.precision = "bf16"
    sched = E.DiffusionSchedule(cfg.diff_steps, device=DEV)
    enc = E.SemanticEncoder(E.CFG(device=DEV, use_fsq=False), load_hubert=False).to(DEV).eval()
    enc.proj.load_state_dict(synth.synth_proj_state(0))
    enc.vq.load_state_dict(synth.synth_vq_state(0))
    inf = E.EdgeInference(cfg, sched, enc, dec)
    out = []

    def emit(name, frames, ms, **kw):
        line = dict(config=name, mel_frames_per_sec=frames / (ms * 1e-3), ms=ms, dtype="bf16", n_gpus=1, **kw)
        out.append(line)
        print(json.dumps(line), flush=True)

    # cfg2
    B, S = 64, 400
    h = torch.randn(B, S, 768, device=DEV, generator=torch.Generator(DEV).manual_seed(2))
    xT = torch.randn(B, 2 * S, cfg.n_mels, device=DEV)

    def cfg2():
        idx = enc.encode_features(h)
        return inf.generate_mel(idx, 1, x_T=xT)

    emit("cfg2: proj + VQ + 1-step generate, batch 64, T=800", B * 2 * S, timed(cfg2, 3, 10))
    emit("cfg2 (VQ only): proj + VQ encode, 25,600 rows", B * 2 * S, timed(lambda: enc.encode_features(h), 3, 10))

    # cfg4
    B, S = 32, 400
    idx = synth.synth_sem_idx(4, B, S).to(DEV)
    xT = torch.randn(B, 2 * S, cfg.n_mels, device=DEV)
    t0 = time.time()
    ms = timed(lambda: inf.sample_ddpm(idx, xT), 1, 2)
    emit("cfg4: 1000-step DDPM ancestral loop, batch 32, T=800", B * 2 * S, ms, frame_steps_per_sec=B * 2 * S * 1000 / (ms * 1e-3),
         ms_per_sampling_step=ms / 1000)

    # cfg5
    for B in (16, 128):
        S = 1500
        idx = synth.synth_sem_idx(5, B, S).to(DEV)
        xT = torch.randn(B, 2 * S, cfg.n_mels, device=DEV)
        ms = timed(lambda: inf.generate_mel(idx, 4, x_T=xT), 2, 5)
        emit(f"cfg5: 4-step DDIM, long utterances T=3000 / S=1500, batch {B}" + (" (one GPU's share of 128 over 8)" if B == 16 else ""),
             B * 2 * S, ms)
        del idx, xT
        inf._plans.clear()
        torch.cuda.empty_cache()

    # dpm
    B, S = 256, 400
    feats = torch.randn(B, S, cfg.semantic_dim, device=DEV)
    xT = torch.randn(B, 2 * S, cfg.n_mels, device=DEV)
    solver = E.DPMSolverPP(sched, order=2)
    ms = timed(lambda: solver.sample(dec, xT, feats, num_steps=10), 1, 3)
    emit("dpm: DPM-Solver++ order 2, 10 steps, sem_features, batch 256, T=800 (fused update, CUDA graph)", B * 2 * S, ms, ms_per_sampling_step=ms / 10)


if __name__ == "__main__":
    main()
