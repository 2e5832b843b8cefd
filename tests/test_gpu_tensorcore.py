"""GPU parity tests of the bf16 tensor-core path (precision="bf16": the fused tcgen05 decoder kernel).

Everything goes through the public API / C ABI and is compared with the fp32 CPU oracle.
Tolerances (BASELINE.json north_star: "bf16 tensor-core path within 1e-2 relative L2, tolerance stated per kernel"):
  * single tcgen05 GEMM / attention kernels ............ rel-L2 <= 5e-3
  * residual stream after 1 and 4 transformer blocks ... rel-L2 <= 5e-3
  * decoder eps, teacher-forced per step ............... rel-L2 <= 1e-2   (measured 3.2e-3 .. 3.5e-3)
  * the fused DDIM / DDPM epilogue given the kernel's own eps: bit-exact vs the oracle update rule
  * free-running 4-step generate_mel: reported, bounded loosely (t=999 gain of 64,000x)
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import edtts_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def tc(lib):
    import edge_diffusion_tts_b200 as E
    assert torch.cuda.is_available() and lib.edtts_device_supported() == 1, "needs an sm_100 device"
    assert lib.edtts_packed_bf16_bytes() > 0
    cfg = E.CFG(device=DEV)
    sd = synth.synth_decoder_state(0)
    dec = E.EdgeDiffusionDecoder(cfg).to(DEV).eval()
    dec.load_state_dict(sd, strict=True)
    dec.precision = "bf16"
    sched = E.DiffusionSchedule(cfg.diff_steps, device=DEV)
    inf = E.EdgeInference(cfg, sched, torch.nn.Identity(), dec)
    return dict(E=E, cfg=cfg, sd=sd, dec=dec, sched=sched, inf=inf, tab=O.cosine_schedule(cfg.diff_steps), lib=lib)


# ------------------------------------------------------------------ single kernels
@pytest.mark.parametrize("mode", [1, 2, 3])
@pytest.mark.parametrize("rows,K,N", [(128, 160, 160), (300, 160, 480), (129, 80, 160), (5, 160, 80), (20000, 160, 640)])
def test_tcgen05_gemm(tc, rows, K, N, mode):
    """y = x W^T + b on the persistent tcgen05 GEMM: fp32 A (mode 1), bf16 chunk-major A (2), chunk-major out (3)."""
    from edge_diffusion_tts_b200 import _lib
    g = torch.Generator().manual_seed(rows + mode)
    x, w, b = torch.randn(rows, K, generator=g), torch.randn(N, K, generator=g) * K ** -0.5, torch.randn(N, generator=g)
    y = torch.full((rows, N), float("nan"), device=DEV)
    xd, wd, bd = x.to(DEV), w.to(DEV), b.to(DEV)
    _lib.check(tc["lib"].edtts_test_linear(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), y.data_ptr(), rows, K, N, mode,
                                           _lib.stream_ptr(DEV)))
    torch.cuda.synchronize()
    assert rel_l2(y, F.linear(x.double(), w.double(), b.double())) <= 5e-3


@pytest.mark.parametrize("B,Tq,Tk,window", [(2, 200, 200, 64), (1, 333, 333, 64), (2, 150, 75, -1), (1, 800, 400, -1)])
def test_tcgen05_attention(tc, B, Tq, Tk, window):
    from edge_diffusion_tts_b200 import _lib
    g = torch.Generator().manual_seed(Tq + Tk)
    q = torch.randn(B, Tq, 160, generator=g)
    kv = torch.randn(B, Tk, 320, generator=g)
    o = torch.full((B, Tq, 160), float("nan"), device=DEV)
    qd, kvd = q.to(DEV), kv.to(DEV)
    _lib.check(tc["lib"].edtts_test_attention(qd.data_ptr(), 160, kvd.data_ptr(), kvd.data_ptr() + 160 * 4, 320,
                                              o.data_ptr(), B, Tq, Tk, window, 1, _lib.stream_ptr(DEV)))
    torch.cuda.synchronize()
    qh = q.view(B, Tq, 4, 40).transpose(1, 2).double()
    kh = kv[..., :160].reshape(B, Tk, 4, 40).transpose(1, 2).double()
    vh = kv[..., 160:].reshape(B, Tk, 4, 40).transpose(1, 2).double()
    mask = O.band_mask(Tq, window, "cpu") if window >= 0 else None
    ref = F.scaled_dot_product_attention(qh, kh, vh, attn_mask=mask).transpose(1, 2).reshape(B, Tq, 160)
    assert rel_l2(o, ref) <= 5e-3


# ------------------------------------------------------------------ the fused kernel, stage by stage
def _oracle_hidden(sd, x, t, si, idx):
    """Residual stream after every stage of every block (transformer.py:146 / :151 / :158)."""
    S, T = idx.shape[1], x.shape[1]
    cond = O.time_condition(sd, t, si)
    ctx = sd["token_emb.weight"][idx] + O._pe_rows(sd, "context_pos_emb.pe", S, torch.float32)
    h = F.linear(x, sd["in_proj.weight"], sd["in_proj.bias"]) + O._pe_rows(sd, "pos_emb.pe", T, torch.float32)
    refs = {(0, 0): h}
    for l in range(4):
        pre = f"layers.{l}."
        h1 = h + O.self_attention(O.ada_rms_norm(h, cond, sd, pre + "norm1."), sd, pre + "attn.")
        h2 = h1 + O.cross_attention(O.rms_norm(h1, sd[pre + "norm2.weight"]), ctx, sd, pre + "cross_attn.")
        h = h2 + O.feed_forward(O.ada_rms_norm(h2, cond, sd, pre + "norm3."), sd, pre + "ffn.")
        refs[(l + 1, 1)], refs[(l + 1, 2)], refs[(l + 1, 0)] = h1, h2, h
    return refs


@pytest.mark.parametrize("B,S", [(2, 100), (1, 37), (3, 400), (1, 1), (44, 400)])
def test_fused_block_stages_vs_oracle(tc, B, S):
    """edtts_test_hidden: h after in_proj, after each stage of block 0 and after all 4 blocks; the fused kernel and
    the separate-launch path against the oracle.  (44, 400) = 308 tiles > 148 SMs exercises the persistent tile loop;
    the workspace is poisoned with NaN bit patterns so that any read of an unwritten byte shows."""
    from edge_diffusion_tts_b200 import _lib
    lib, dec, sd = tc["lib"], tc["dec"], tc["sd"]
    T = 2 * S
    idx = synth.synth_sem_idx(S, B, S)
    x = synth.synth_noise(S, B, T)
    t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(S))
    si = torch.randint(0, 16, (B,), generator=torch.Generator().manual_seed(S + 1))
    refs = _oracle_hidden(sd, x, t, si, idx)
    xd = x.to(DEV)
    mod = dec.prepare_cond(t.to(DEV), si.to(DEV), T, S)
    kv = dec.prepare_context(idx.to(DEV), None, T)
    w = dec._weights(T, S)
    nbytes = lib.edtts_decoder_workspace_bytes(B, T, S, _lib.PREC_BF16)
    out = torch.empty(B, T, 160, device=DEV)
    for fused in (1, 0):
        for nl, stop in ((0, 0), (1, 1), (1, 2), (1, 0), (4, 0)):
            if B > 8 and not (fused == 1 and stop == 0):
                continue
            ws = torch.full((nbytes,), 0xFF, dtype=torch.uint8, device=DEV)
            out.fill_(float("nan"))
            _lib.check(lib.edtts_test_hidden(w, xd.data_ptr(), mod.data_ptr(), kv.data_ptr(), out.data_ptr(), ws.data_ptr(),
                                             nbytes, B, T, S, nl, stop, fused, _lib.stream_ptr(DEV)), "test_hidden")
            torch.cuda.synchronize()
            assert not torch.isnan(out).any(), (fused, nl, stop)
            assert rel_l2(out, refs[(nl, stop)]) <= 5e-3, (fused, nl, stop)


# ------------------------------------------------------------------ decoder / sampler
@pytest.mark.parametrize("B,S", [(1, 1), (3, 37), (2, 300), (1, 500), (40, 400)])
def test_decoder_bf16_vs_oracle(tc, B, S):
    idx = synth.synth_sem_idx(S, B, S)
    x = synth.synth_noise(S, B, 2 * S)
    t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(S))
    si = torch.randint(0, 16, (B,), generator=torch.Generator().manual_seed(S + 1))
    ref = O.decoder_forward(tc["sd"], x, t, idx, si)
    eps = tc["dec"](x.to(DEV), t.to(DEV), idx.to(DEV), si.to(DEV))
    assert rel_l2(eps, ref) <= 1e-2


# (21, 400) / (37, 256) / (149, 64): 147 / 148 / 149 tiles per layer on 148 CTAs -- the sizes at which an item, its successor on
# the same CTA and their predecessors on the neighbouring CTAs form the tightest dependency pattern
@pytest.mark.parametrize("B,S", [(1, 30), (5, 100), (40, 400), (3, 1500), (21, 400), (37, 256), (149, 64)])
def test_merged_equals_per_layer_launches(tc, B, S, monkeypatch):
    """The decoder step as one persistent launch (head + 4 blocks, per-tile dependency flags between the layers; more
    work items than CTAs at the larger sizes, single-tile utterances at the smallest) gives the bits of one launch per
    layer: the schedule changes, the arithmetic of a tile does not."""
    idx = synth.synth_sem_idx(S + 3, B, S).to(DEV)
    x = synth.synth_noise(S + 3, B, 2 * S).to(DEV)
    t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(S)).to(DEV)
    si = torch.randint(0, 16, (B,), generator=torch.Generator().manual_seed(S + 1)).to(DEV)
    monkeypatch.setenv("EDTTS_MERGED_LAYERS", "0")
    ref = tc["dec"](x, t, idx, si).clone()
    monkeypatch.setenv("EDTTS_MERGED_LAYERS", "1")
    for _ in range(3):                                     # repeated: the flags are re-armed by every call
        out = tc["dec"](x, t, idx, si)
        assert torch.equal(out, ref)


def test_decoder_bf16_long_sequence(tc):
    """BASELINE config 5 shape class: T = 3000 frames, 1500 context tokens (24 key blocks per head)."""
    B, S = 1, 1500
    idx = synth.synth_sem_idx(7, B, S)
    x = synth.synth_noise(7, B, 2 * S)
    t = torch.tensor([499])
    ref = O.decoder_forward(tc["sd"], x, t, idx, torch.tensor([2]))
    eps = tc["dec"](x.to(DEV), t.to(DEV), idx.to(DEV), torch.tensor([2], device=DEV))
    assert rel_l2(eps, ref) <= 1e-2


def test_decoder_bf16_semantic_features(tc, golden):
    g = golden("decoder_semfeat")
    feats = synth.synth_features(12, 2, 40, 128).to(DEV)
    xs = synth.synth_noise(12, 2, 80).to(DEV)
    eps = tc["dec"](xs, torch.tensor([700, 20], device=DEV), None, None, sem_features=feats)
    assert rel_l2(eps, g["eps"]) <= 1e-2


@pytest.mark.parametrize("steps", [4, 1])
def test_generate_mel_bf16(tc, golden, steps):
    g = golden("generate_mel")
    run = g["runs"][steps]
    B = g["B"]
    idx = synth.synth_sem_idx(g["seed"], B, g["S"]).to(DEV)
    xT = synth.synth_noise(g["seed"], B, 2 * g["S"]).to(DEV)
    dec, sched = tc["dec"], tc["sched"]
    for i, tr in enumerate(run["trace"]):
        t = torch.full((B,), tr["t"], dtype=torch.long, device=DEV)
        tp = torch.full((B,), tr["t_prev"], dtype=torch.long, device=DEV)
        si = torch.full((B,), i, dtype=torch.long, device=DEV)
        x_t = tr["x_t"].to(DEV)
        eps = dec(x_t, t, idx, si)
        assert rel_l2(eps, tr["eps"]) <= 1e-2, i                          # teacher-forced
        # the update rule fused into the last kernel is bit-exact given the kernel's own eps
        from edge_diffusion_tts_b200 import _lib
        mod = dec.prepare_cond(t, si, x_t.shape[1], g["S"])
        kv = dec.prepare_context(idx, None, x_t.shape[1])
        xp, x0, e2 = torch.empty_like(x_t), torch.empty_like(x_t), torch.empty_like(x_t)
        a = _lib.StepArgs()
        a.mode, a.write_x_prev = _lib.STEP_DDIM, 1
        a.t, a.t_prev, a.alpha_bar = t.data_ptr(), tp.data_ptr(), sched.alpha_bar.data_ptr()
        a.x_prev_out, a.x0_out, a.eps_out = xp.data_ptr(), x0.data_ptr(), e2.data_ptr()
        dec.step(x_t, mod, kv, g["S"], a)
        torch.cuda.synchronize()
        assert torch.equal(e2, eps)
        xp_ref, x0_ref = O.ddim_step(tc["tab"], x_t.cpu(), t.cpu(), tp.cpu(), e2.cpu(), 0.0)
        assert torch.equal(xp.cpu(), xp_ref) and torch.equal(x0.cpu(), x0_ref)
    out = tc["inf"].generate_mel(idx, steps, x_T=xT)
    tc["inf"].use_cuda_graph = False
    try:
        out_eager = tc["inf"].generate_mel(idx, steps, x_T=xT)
    finally:
        tc["inf"].use_cuda_graph = True
    assert torch.equal(out, out_eager)                                    # graph replay == eager launches
    assert out.abs().max().item() <= 3.0 and not torch.isnan(out).any()
    # free-running vs the fp32 reference: dominated by the 64,000x gain of the t=999 step on bf16-level eps error
    # (SURVEY.md F9); bounded loosely, the per-step criterion above is the parity statement.
    assert rel_l2(out, run["x0"]) <= 0.15


def test_batch_invariance_bf16(tc):
    """A row's bits do not depend on the batch it was computed in => batch shards reproduce the 1-GPU result."""
    idx = synth.synth_sem_idx(2, 6, 90).to(DEV)
    xT = synth.synth_noise(2, 6, 180).to(DEV)
    full = tc["inf"].generate_mel(idx, 4, x_T=xT)
    parts = torch.cat([tc["inf"].generate_mel(idx[a:b], 4, x_T=xT[a:b]) for a, b in ((0, 1), (1, 4), (4, 6))])
    assert torch.equal(full, parts)


def test_sample_ddpm_bf16(tc):
    B, S, n = 2, 30, 6
    idx = synth.synth_sem_idx(4, B, S)
    xT = synth.synth_noise(4, B, 2 * S)
    noises = [synth.synth_noise(4, B, 2 * S, tag=f"n{i}") for i in range(n)]
    ref = O.ddpm_loop(tc["sd"], tc["tab"], idx, xT, noises, t_start=300, t_end=300 - n + 1)
    out = tc["inf"].sample_ddpm(idx.to(DEV), xT.to(DEV), [z.to(DEV) for z in noises], t_start=300, t_end=300 - n + 1,
                                graph_steps=3)
    assert rel_l2(out, ref) <= 2e-2
