"""Numeric error report of every kernel vs the oracle (diagnostics for development; prints, never asserts).
Usage on the GPU box:  python tools/gpu_report.py > gpurun_out/report.txt"""
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__ as ge

ge.build()
import edge_diffusion_tts_b200 as E
from edge_diffusion_tts_b200 import _lib
from oracle import edtts_oracle as O, synth

DEV = "cuda:0"
lib = _lib.load()
print("device", torch.cuda.get_device_name(0), "supported", lib.edtts_device_supported())


def section(name, fn):
    t0 = time.time()
    try:
        fn()
        torch.cuda.synchronize()
        print(f"[ok ] {name} ({time.time()-t0:.1f}s)")
    except Exception:
        print(f"[ERR] {name}\n{traceback.format_exc()}")
    sys.stdout.flush()


def stats(tag, a, b):
    a, b = a.double().cpu(), b.double().cpu()
    d = (a - b).abs()
    print(f"   {tag}: max|d|={d.max().item():.3e} mean|d|={d.mean().item():.3e} "
          f"relL2={(d.norm()/b.norm().clamp_min(1e-30)).item():.3e} ref_absmax={b.abs().max().item():.3e} "
          f"nan={int(torch.isnan(a).sum())}")


def linear(prec):
    shapes = ((300, 160, 480), (129, 80, 160), (1000, 320, 160), (77, 768, 128), (5, 160, 80), (4096, 160, 640))
    if prec >= 1:
        shapes = ((128, 160, 160), (300, 160, 480), (129, 80, 160), (1000, 320, 160), (5, 160, 80), (40000, 160, 640))
    for rows, K, N in shapes:
        g = torch.Generator().manual_seed(rows)
        x, w, b = torch.randn(rows, K, generator=g), torch.randn(N, K, generator=g) * K ** -0.5, torch.randn(N, generator=g)
        y = torch.zeros(rows, N, device=DEV)
        xd, wd, bd = x.to(DEV), w.to(DEV), b.to(DEV)
        rc = lib.edtts_test_linear(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), y.data_ptr(), rows, K, N, prec,
                                   _lib.stream_ptr(DEV))
        if rc:
            print("   rc", rc, lib.edtts_last_error())
            continue
        torch.cuda.synchronize()
        stats(f"linear p{prec} {rows}x{K}x{N}", y, torch.nn.functional.linear(x.double(), w.double(), b.double()))


def attention(prec=0):
    for B, Tq, Tk, window in ((2, 200, 200, 64), (1, 333, 333, 64), (2, 150, 75, -1), (1, 800, 400, -1), (3, 800, 800, 64)):
        g = torch.Generator().manual_seed(Tq + Tk)
        q = torch.randn(B, Tq, 160, generator=g)
        kv = torch.randn(B, Tk, 320, generator=g)
        o = torch.zeros(B, Tq, 160, device=DEV)
        qd, kvd = q.to(DEV), kv.to(DEV)
        rc = lib.edtts_test_attention(qd.data_ptr(), 160, kvd.data_ptr(), kvd.data_ptr() + 640, 320, o.data_ptr(), B, Tq,
                                      Tk, window, prec, _lib.stream_ptr(DEV))
        if rc:
            print("   rc", rc, lib.edtts_last_error())
            continue
        qh = q.view(B, Tq, 4, 40).transpose(1, 2).double()
        kh = kv[..., :160].reshape(B, Tk, 4, 40).transpose(1, 2).double()
        vh = kv[..., 160:].reshape(B, Tk, 4, 40).transpose(1, 2).double()
        mask = O.band_mask(Tq, window, "cpu") if window >= 0 else None
        ref = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh, attn_mask=mask).transpose(1, 2).reshape(B, Tq, 160)
        torch.cuda.synchronize()
        stats(f"attention p{prec} B{B} Tq{Tq} Tk{Tk} w{window}", o, ref)


def make_model(prec="fp32"):
    cfg = E.CFG(device=DEV)
    sd = synth.synth_decoder_state(0)
    dec = E.EdgeDiffusionDecoder(cfg).to(DEV).eval()
    dec.load_state_dict(sd, strict=True)
    dec.precision = prec
    sched = E.DiffusionSchedule(cfg.diff_steps, device=DEV)
    return cfg, sd, dec, sched, E.EdgeInference(cfg, sched, torch.nn.Identity(), dec)


def decoder(prec="fp32"):
    cfg, sd, dec, sched, inf = make_model(prec)
    for B, S in ((2, 100), (1, 37), (1, 500)):
        idx = synth.synth_sem_idx(S, B, S)
        x = synth.synth_noise(S, B, 2 * S)
        t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(S))
        si = torch.randint(0, 16, (B,), generator=torch.Generator().manual_seed(S + 1))
        ref = O.decoder_forward(sd, x, t, idx, si)
        mod_ref = None
        eps = dec(x.to(DEV), t.to(DEV), idx.to(DEV), si.to(DEV))
        stats(f"decoder[{prec}] eps B{B} S{S}", eps, ref)
    # stage checks
    B, S = 2, 100
    idx = synth.synth_sem_idx(S, B, S)
    t = torch.tensor([999, 500]); si = torch.tensor([0, 3])
    mod = dec.prepare_cond(t.to(DEV), si.to(DEV))
    cond = O.time_condition(sd, t, si)
    ref_mod = torch.stack([torch.nn.functional.linear(cond, sd[f"layers.{l}.norm{n}.proj.weight"], sd[f"layers.{l}.norm{n}.proj.bias"])
                           for l in range(4) for n in (1, 3)], dim=1)
    stats("cond mod", mod, ref_mod)
    kv = dec.prepare_context(idx.to(DEV)).flatten()[:4 * B * S * 320].view(4, B, S, 320)   # (fp32: the operand images follow the rows)
    ctx = sd["token_emb.weight"][idx] + sd["context_pos_emb.pe"][:S]
    for l in range(4):
        k, v = O.cross_kv(ctx, sd, f"layers.{l}.cross_attn.")
        ref_kv = torch.cat([k.transpose(1, 2).reshape(B, S, 160), v.transpose(1, 2).reshape(B, S, 160)], -1)
        stats(f"context kv layer{l}", kv[l], ref_kv)


def generate(prec="fp32"):
    cfg, sd, dec, sched, inf = make_model(prec)
    tab = O.cosine_schedule(1000)
    B, S = 2, 100
    idx = synth.synth_sem_idx(5, B, S)
    xT = synth.synth_noise(5, B, 2 * S)
    trace = []
    ref = O.generate_mel(sd, tab, idx, 4, xT, trace=trace)
    for i, (x_t, eps_ref, _, _) in enumerate(trace):
        tt = torch.full((B,), [999, 749, 499, 249][i], dtype=torch.long, device=DEV)
        eps = dec(x_t.to(DEV), tt, idx.to(DEV), torch.full((B,), i, dtype=torch.long, device=DEV))
        stats(f"teacher-forced[{prec}] step{i} eps", eps, eps_ref)
    out = inf.generate_mel(idx.to(DEV), 4, x_T=xT.to(DEV))
    stats(f"free-running[{prec}] x0", out, ref)
    d = (out.cpu() - ref).abs()
    print(f"   frac |d|>1e-4: {(d > 1e-4).float().mean().item():.3e}")


def vq():
    cfg = E.CFG(device=DEV, use_fsq=False)
    enc = E.SemanticEncoder(cfg, load_hubert=False).to(DEV).eval()
    enc.proj.load_state_dict(synth.synth_proj_state(0))
    enc.vq.load_state_dict(synth.synth_vq_state(0))
    cb = synth.synth_vq_state(0)["codebook.weight"]
    for b, s in ((64, 400), (128, 1500)):
        h = synth.synth_features(21, b, s)
        z = O.encoder_proj(synth.synth_proj_state(0), h)
        stats(f"proj {b}x{s}", enc.project(h.to(DEV)), z)
        idx = enc.vq.encode(z.to(DEV)).cpu()
        r32 = O.vq_encode(cb, z)
        r64 = O.vq_encode(cb.double(), z.double())
        print(f"   vq rows={b*s}: mismatch vs fp32 oracle {(idx != r32).sum().item()}, vs fp64 {(idx != r64).sum().item()}, "
              f"oracle32 vs 64 {(r32 != r64).sum().item()}")


def hidden(shapes=((2, 100), (1, 37), (3, 400), (44, 400))):
    """Residual stream after partial / whole transformer blocks: fused kernel and separate launches vs the oracle."""
    import torch.nn.functional as F
    cfg, sd, dec, sched, inf = make_model("bf16")
    for B, S in shapes:
        T = 2 * S
        idx = synth.synth_sem_idx(S, B, S)
        x = synth.synth_noise(S, B, T)
        t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(S))
        si = torch.randint(0, 16, (B,), generator=torch.Generator().manual_seed(S + 1))
        cond = O.time_condition(sd, t, si)
        ctx = sd["token_emb.weight"][idx] + O._pe_rows(sd, "context_pos_emb.pe", S, torch.float32)
        h0 = F.linear(x, sd["in_proj.weight"], sd["in_proj.bias"]) + O._pe_rows(sd, "pos_emb.pe", T, torch.float32)
        refs = {}
        h = h0
        for l in range(4):
            pre = f"layers.{l}."
            h1 = h + O.self_attention(O.ada_rms_norm(h, cond, sd, pre + "norm1."), sd, pre + "attn.")
            h2 = h1 + O.cross_attention(O.rms_norm(h1, sd[pre + "norm2.weight"]), ctx, sd, pre + "cross_attn.")
            h3 = h2 + O.feed_forward(O.ada_rms_norm(h2, cond, sd, pre + "norm3."), sd, pre + "ffn.")
            refs[(l + 1, 1)], refs[(l + 1, 2)], refs[(l + 1, 0)] = h1, h2, h3
            h = h3
        xd = x.to(DEV)
        mod = dec.prepare_cond(t.to(DEV), si.to(DEV), T, S)
        kv = dec.prepare_context(idx.to(DEV), None, T)
        w = dec._weights(T, S)
        nbytes = lib.edtts_decoder_workspace_bytes(B, T, S, _lib.PREC_BF16)
        ws = torch.full((nbytes,), 0xFF, dtype=torch.uint8, device=DEV)   # NaN patterns: every byte read must have been written
        out = torch.zeros(B, T, 160, device=DEV)
        cases = [(1, 1), (1, 2), (1, 0), (4, 0)]
        for fused in (0, 1):
            for nl, stop in cases:
                out.zero_()
                rc = lib.edtts_test_hidden(w, xd.data_ptr(), mod.data_ptr(), kv.data_ptr(), out.data_ptr(), ws.data_ptr(),
                                           nbytes, B, T, S, nl, stop, fused, _lib.stream_ptr(DEV))
                if rc:
                    print("   rc", rc, lib.edtts_last_error())
                    continue
                torch.cuda.synchronize()
                stats(f"hidden B{B} S{S} layers={nl} stop={stop} fused={fused}", out, refs[(nl, stop)])


if "--skip-fp32" not in sys.argv:
    section("linear fp32", lambda: linear(0))
    section("attention fp32", attention)
    section("decoder fp32", decoder)
    section("generate fp32", generate)
    section("vq", vq)
if "--bf16-kernels" in sys.argv:
    section("linear bf16 (fp32 A)", lambda: linear(1))
    section("linear bf16 (chunk A)", lambda: linear(2))
    section("linear bf16 (chunk out)", lambda: linear(3))
    section("attention bf16", lambda: attention(1))
if "--bf16" in sys.argv:
    section("decoder bf16", lambda: decoder("bf16"))
    section("generate bf16", lambda: generate("bf16"))
if "--hidden" in sys.argv:
    section("hidden bf16 (fused vs separate)", hidden)
