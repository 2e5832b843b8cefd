import os, sys
sys.path.insert(0, "/root/repo" if os.path.isdir("/root/repo/tests") else os.getcwd())
import torch
import __graft_entry__ as ge
ge.build()
import edge_diffusion_tts_b200 as E
from oracle import synth
dev = "cuda:0"
cfg = E.CFG(device=dev)
dec = E.EdgeDiffusionDecoder(cfg).to(dev).eval(); dec.load_state_dict(synth.synth_decoder_state(0)); dec.precision = "bf16"
enc = E.SemanticEncoder(E.CFG(device=dev, use_fsq=False), load_hubert=False).to(dev).eval()
enc.proj.load_state_dict(synth.synth_proj_state(0)); enc.vq.load_state_dict(synth.synth_vq_state(0))
inf = E.EdgeInference(cfg, E.DiffusionSchedule(cfg.diff_steps, device=dev), enc, dec)
B, S = 64, 400
h = torch.randn(B, S, 768, device=dev); xT = torch.randn(B, 2 * S, 80, device=dev)
idx = enc.encode_features(h)
def t(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print("encode only      ", t(lambda: enc.encode_features(h)))
print("generate only    ", t(lambda: inf.generate_mel(idx, 1, x_T=xT)))
print("encode + generate", t(lambda: inf.generate_mel(enc.encode_features(h), 1, x_T=xT)))
for name, env in (("proj simt", "EDTTS_PROJ_SIMT"), ):
    pass
