"""Development helper: per-kernel-class device time (the library's CUDA-event profiler) of one eager generate_mel at cfg3 size.
python tools/prof_classes.py <precision> [B] [S] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
import edge_diffusion_tts_b200 as E
from edge_diffusion_tts_b200 import _lib
from oracle import synth
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
S = int(sys.argv[3]) if len(sys.argv) > 3 else 400
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = "cuda:0"
cfg = E.CFG(device=dev)
dec = E.EdgeDiffusionDecoder(cfg).to(dev).eval()
dec.load_state_dict(synth.synth_decoder_state(0), strict=True)
dec.precision = prec
inf = E.EdgeInference(cfg, E.DiffusionSchedule(cfg.diff_steps, device=dev), torch.nn.Identity(), dec, use_cuda_graph=False)
idx = synth.synth_sem_idx(1, B, S).to(dev)
xT = torch.randn(B, 2 * S, 80, device=dev)
for _ in range(2):
    inf.generate_mel(idx, 4, x_T=xT)
torch.cuda.synchronize()
_lib.prof_enable(True)
for _ in range(reps):
    inf.generate_mel(idx, 4, x_T=xT)
prof = _lib.prof_collect()
_lib.prof_enable(False)
tot = sum(ms for ms, _ in prof.values())
print(f"precision {prec} B={B} S={S}: {tot / reps:.3f} ms per generate (sum of kernel classes)")
for k, (ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:28s} {ms / reps:9.3f} ms  {n // reps:4d} launches  {ms / n * 1e3:9.1f} us each  {100 * ms / tot:5.1f} %")
