"""Development helper: device time of generate_mel(4 steps, bf16) for a list of B,S[,flat] shapes.  python tools/time_shapes.py 256,400 266,384 ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
import edge_diffusion_tts_b200 as E
from oracle import synth
dev = "cuda:0"
cfg = E.CFG(device=dev)
dec = E.EdgeDiffusionDecoder(cfg).to(dev).eval()
dec.load_state_dict(synth.synth_decoder_state(0), strict=True)
dec.precision = "bf16"
inf = E.EdgeInference(cfg, E.DiffusionSchedule(cfg.diff_steps, device=dev), torch.nn.Identity(), dec)
for spec in sys.argv[1:]:
    parts = spec.split(",")
    B, S = int(parts[0]), int(parts[1])
    pass
    idx = synth.synth_sem_idx(1, B, S).to(dev)
    xT = torch.randn(B, 2 * S, 80, device=dev)
    for _ in range(4):
        inf.generate_mel(idx, 4, x_T=xT)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        inf.generate_mel(idx, 4, x_T=xT)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    T = 2 * S
    tiles = (B * T + 127) // 128 if not dec.batch_invariant and T % 128 else B * ((T + 127) // 128)
    print(f"{spec:18s} B*T={B*T:8d} tiles/layer={tiles:5d} {ms:8.3f} ms  {ms*1e3/ (tiles*16):7.3f} us per tile-launch-step  {B*T/ms/1e3:7.2f} Mframes/s", flush=True)
    inf._plans.clear(); torch.cuda.empty_cache()
