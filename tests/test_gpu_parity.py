"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every check goes through
the public API -> ctypes -> C ABI -> CUDA kernels and is compared with the CPU oracle
(oracle/edtts_oracle.py) and with the committed golden fixtures recorded from the
unmodified reference (tests/golden, oracle/make_golden.py).

Tolerances:
  * VQ indices, DDIM / DDPM update rules ........ bit-exact
  * fp32 decoder eps, teacher-forced per step ... max-abs <= 1e-4
  * free-running generate_mel (fp32) ............ rel-L2 <= 1e-3 and every |d| > 1e-4 element traced to
                                                   the t=999 gain (F9) -- measured, see test body
"""
import pytest
import torch

from oracle import edtts_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pkg(lib):
    import edge_diffusion_tts_b200 as E
    assert torch.cuda.is_available()
    assert lib.edtts_device_supported() == 1, "not an sm_100 device"
    return E


@pytest.fixture(scope="module")
def model(pkg):
    cfg = pkg.CFG(device=DEV)
    sd = synth.synth_decoder_state(0)
    dec = pkg.EdgeDiffusionDecoder(cfg).to(DEV).eval()
    dec.load_state_dict(sd, strict=True)
    sched = pkg.DiffusionSchedule(cfg.diff_steps, device=DEV)
    inf = pkg.EdgeInference(cfg, sched, torch.nn.Identity(), dec)
    return dict(cfg=cfg, sd=sd, dec=dec, sched=sched, inf=inf, tab=O.cosine_schedule(cfg.diff_steps))


# ------------------------------------------------------------------ schedule
def test_schedule_tables_and_updates_bit_exact(pkg, golden):
    g = golden("schedule")
    s = pkg.DiffusionSchedule(1000, device=DEV)
    for k, v in g["tables"].items():
        assert torch.equal(getattr(s, k).cpu(), v), k
    x = synth.synth_noise(14, 4, 10).to(DEV)
    e = synth.synth_noise(14, 4, 10, tag="eps").to(DEV)
    n = synth.synth_noise(14, 4, 10, tag="noise").to(DEV)
    xp, x0 = s.get_ddim_step(x, g["t"].to(DEV), g["t_prev"].to(DEV), e, 0.0)
    assert torch.equal(xp.cpu(), g["ddim_x_prev"]) and torch.equal(x0.cpu(), g["ddim_x0"])
    xd = s.ddpm_step(x, g["t"].to(DEV), e, noise=n)
    assert torch.equal(xd.cpu(), g["ddpm_x_prev"])
    assert torch.equal(s.q_sample(x, g["t"].to(DEV), n)[0].cpu(), g["q_sample"])


@pytest.mark.parametrize("B,T,eta", [(3, 7, 0.0), (2, 801, 0.0), (5, 64, 0.3), (1, 1, 0.0)])
def test_ddim_ddpm_random_shapes_bit_exact(pkg, B, T, eta):
    tab = O.cosine_schedule(1000)
    s = pkg.DiffusionSchedule(1000, device=DEV)
    g = torch.Generator().manual_seed(B * 1000 + T)
    x, e, n = (torch.randn(B, T, 80, generator=g) for _ in range(3))
    t = torch.randint(0, 1000, (B,), generator=g)
    t[0] = 999
    tp = (t - 250).clamp(min=-1)
    xp_ref, x0_ref = O.ddim_step(tab, x, t, tp, e, eta, noise=n)
    xp, x0 = s.get_ddim_step(x.to(DEV), t.to(DEV), tp.to(DEV), e.to(DEV), eta, noise=n.to(DEV))
    assert torch.equal(x0.cpu(), x0_ref)
    assert torch.equal(xp.cpu(), xp_ref)
    t[-1] = 0
    xd_ref = O.ddpm_step(tab, x, t, e, n)
    xd = s.ddpm_step(x.to(DEV), t.to(DEV), e.to(DEV), noise=n.to(DEV))
    assert torch.equal(xd.cpu(), xd_ref)


# ------------------------------------------------------------------ VQ + encoder projection
def _encoder(pkg):
    cfg = pkg.CFG(device=DEV, use_fsq=False)
    enc = pkg.SemanticEncoder(cfg, load_hubert=False).to(DEV).eval()
    enc.proj.load_state_dict(synth.synth_proj_state(0))
    enc.vq.load_state_dict(synth.synth_vq_state(0))
    return enc


def test_vq_and_proj_golden(pkg, golden):
    g = golden("vq")
    enc = _encoder(pkg)
    h = synth.synth_features(15, 4, 50).to(DEV)
    z = enc.project(h)
    assert (z.cpu() - g["z"]).abs().max().item() < 2e-5
    # quantise the reference's own z so the index comparison is not confounded by the projection
    z_q, idx, loss, perp, used = enc.vq(g["z"].to(DEV))
    assert torch.equal(idx.cpu(), g["idx"])
    assert torch.equal(z_q.cpu(), g["z_q"])
    assert float(loss) == 0.0 and int(used) == int(g["used"])
    assert abs(float(perp) - float(g["perplexity"])) < 1e-3
    assert torch.equal(enc.vq.encode(g["z"].to(DEV)).cpu(), g["encode"])
    assert torch.equal(enc.decode_tokens(g["idx"][:1, :5].to(DEV)).cpu(), g["decode"])


@pytest.mark.parametrize("rows_b,rows_s", [(64, 400), (1, 1), (3, 43), (128, 1500)])
def test_vq_indices_bit_exact_vs_oracle_fp32_and_fp64(pkg, rows_b, rows_s):
    enc = _encoder(pkg)
    cb = synth.synth_vq_state(0)["codebook.weight"]
    h = synth.synth_features(21, rows_b, rows_s)
    z = O.encoder_proj(synth.synth_proj_state(0), h)               # oracle z, fed to both sides
    idx = enc.vq.encode(z.to(DEV)).cpu()
    ref32 = O.vq_encode(cb, z)
    ref64 = O.vq_encode(cb.double(), z.double())
    assert torch.equal(idx, ref64), f"{(idx != ref64).sum().item()} rows differ from the exact argmin"
    assert torch.equal(idx, ref32), f"{(idx != ref32).sum().item()} rows differ from the fp32 oracle"


def test_vq_ties_and_duplicates(pkg):
    """Collisions: duplicated codewords must resolve to the lowest index (torch.argmin)."""
    vq = pkg.VectorQuantizer(128, 512).to(DEV).eval()
    cb = synth.synth_vq_state(3)["codebook.weight"].clone()
    cb[100] = cb[7]
    cb[511] = cb[7]
    cb[300] = cb[299]
    vq.codebook.weight.data.copy_(cb)
    z = torch.cat([cb[[7, 299, 511, 100]], cb[[7]] + 1e-4, synth.synth_features(5, 1, 200, 128)[0]])[None]
    idx = vq.encode(z.to(DEV)).cpu()
    assert torch.equal(idx, O.vq_encode(cb.double(), z.double()))
    assert idx[0, :4].tolist() == [7, 299, 7, 7]
    assert vq.encode(torch.empty(0, 5, 128, device=DEV)).shape == (0, 5)


# ------------------------------------------------------------------ single kernels
@pytest.mark.parametrize("rows,K,N", [(300, 160, 480), (129, 80, 160), (1000, 320, 160), (77, 768, 128), (5, 160, 80)])
def test_linear_kernel(lib, rows, K, N):
    from edge_diffusion_tts_b200 import _lib
    g = torch.Generator().manual_seed(rows)
    x, w, b = torch.randn(rows, K, generator=g), torch.randn(N, K, generator=g) * K ** -0.5, torch.randn(N, generator=g)
    y = torch.empty(rows, N, device=DEV)
    xd, wd, bd = x.to(DEV), w.to(DEV), b.to(DEV)
    _lib.check(lib.edtts_test_linear(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), y.data_ptr(), rows, K, N, 0,
                                     _lib.stream_ptr(DEV)))
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
    assert (y.cpu().double() - ref).abs().max().item() < 2e-5


@pytest.mark.parametrize("B,Tq,Tk,window", [(2, 200, 200, 64), (1, 64, 64, 64), (1, 333, 333, 64), (2, 150, 75, -1),
                                            (1, 800, 400, -1), (1, 1, 1, 64)])
def test_attention_kernel(lib, B, Tq, Tk, window):
    from edge_diffusion_tts_b200 import _lib
    g = torch.Generator().manual_seed(Tq + Tk)
    q = torch.randn(B, Tq, 160, generator=g)
    kv = torch.randn(B, Tk, 320, generator=g)
    o = torch.empty(B, Tq, 160, device=DEV)
    qd, kvd = q.to(DEV), kv.to(DEV)
    _lib.check(lib.edtts_test_attention(qd.data_ptr(), 160, kvd.data_ptr(), kvd.data_ptr() + 160 * 4, 320,
                                        o.data_ptr(), B, Tq, Tk, window, 0, _lib.stream_ptr(DEV)))
    qh = q.view(B, Tq, 4, 40).transpose(1, 2).double()
    kh = kv[..., :160].reshape(B, Tk, 4, 40).transpose(1, 2).double()
    vh = kv[..., 160:].reshape(B, Tk, 4, 40).transpose(1, 2).double()
    mask = O.band_mask(Tq, window, "cpu") if window >= 0 else None
    ref = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh, attn_mask=mask)
    ref = ref.transpose(1, 2).reshape(B, Tq, 160)
    assert (o.cpu().double() - ref).abs().max().item() < 2e-5


# ------------------------------------------------------------------ decoder
def test_decoder_step_golden(model, golden):
    g = golden("decoder_step")
    assert g["meta"]["weights"] == synth.state_checksum(model["sd"]), "synthetic weights regenerated differently"
    idx = synth.synth_sem_idx(g["seed"], g["B"], g["S"])
    x = synth.synth_noise(g["seed"], g["B"], 2 * g["S"])
    assert int(idx.sum()) == g["idx_sum"]
    for name, c in g["cases"].items():
        si = None if c["step_idx"] is None else c["step_idx"].to(DEV)
        eps = model["dec"](x.to(DEV), c["t"].to(DEV), idx.to(DEV), si).cpu()
        err = (eps - c["eps"]).abs().max().item()
        assert err <= 1e-4, (name, err)


def test_decoder_semantic_features_golden(model, golden):
    g = golden("decoder_semfeat")
    feats = synth.synth_features(12, 2, 40, 128).to(DEV)
    xs = synth.synth_noise(12, 2, 80).to(DEV)
    eps = model["dec"](xs, torch.tensor([700, 20], device=DEV), None, None, sem_features=feats).cpu()
    assert (eps - g["eps"]).abs().max().item() <= 1e-4


@pytest.mark.parametrize("B,S", [(1, 1), (3, 37), (2, 300), (1, 500)])
def test_decoder_shapes_vs_oracle(model, B, S):
    """Ragged / tiny / maximum (T=1000, S=500: the reference's PE-table limit) sequence lengths."""
    idx = synth.synth_sem_idx(S, B, S)
    x = synth.synth_noise(S, B, 2 * S)
    t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(S))
    si = torch.randint(0, 16, (B,), generator=torch.Generator().manual_seed(S + 1))
    ref = O.decoder_forward(model["sd"], x, t, idx, si)
    eps = model["dec"](x.to(DEV), t.to(DEV), idx.to(DEV), si.to(DEV)).cpu()
    assert (eps - ref).abs().max().item() <= 1e-4


def test_decoder_long_sequence_extended_pe(model):
    """BASELINE config 5 shape class (T=3000 > 1000-row table): both sides extend the PE tables by the
    closed form (a stated deviation from the unmodified reference which raises)."""
    B, S = 1, 1500
    idx = synth.synth_sem_idx(7, B, S)
    x = synth.synth_noise(7, B, 2 * S)
    t = torch.tensor([499])
    ref = O.decoder_forward(model["sd"], x, t, idx, torch.tensor([2]))
    eps = model["dec"](x.to(DEV), t.to(DEV), idx.to(DEV), torch.tensor([2], device=DEV)).cpu()
    assert (eps - ref).abs().max().item() <= 1e-4


def test_decoder_errors(model):
    with pytest.raises(ValueError):
        model["dec"](torch.zeros(1, 4, 80, device=DEV), torch.zeros(1, dtype=torch.long, device=DEV))
    with pytest.raises(RuntimeError):
        model["dec"](torch.zeros(1, 4, 80), torch.zeros(1, dtype=torch.long), torch.zeros(1, 2, dtype=torch.long))


# ------------------------------------------------------------------ EdgeInference
@pytest.mark.parametrize("steps", [4, 1])
def test_generate_mel_golden_teacher_forced_and_free_running(model, golden, steps):
    g = golden("generate_mel")
    run = g["runs"][steps]
    idx = synth.synth_sem_idx(g["seed"], g["B"], g["S"]).to(DEV)
    xT = synth.synth_noise(g["seed"], g["B"], 2 * g["S"]).to(DEV)
    # (ii) teacher-forced per step: reference x_t in, eps within 1e-4; update rule bit-exact given reference eps
    for i, tr in enumerate(run["trace"]):
        B = g["B"]
        t = torch.full((B,), tr["t"], dtype=torch.long, device=DEV)
        tp = torch.full((B,), tr["t_prev"], dtype=torch.long, device=DEV)
        si = torch.full((B,), i, dtype=torch.long, device=DEV)
        eps = model["dec"](tr["x_t"].to(DEV), t, idx, si)
        assert (eps.cpu() - tr["eps"]).abs().max().item() <= 1e-4, i
        xp, x0 = model["sched"].get_ddim_step(tr["x_t"].to(DEV), t, tp, tr["eps"].to(DEV))
        assert torch.equal(xp.cpu(), tr["x_prev"]) and torch.equal(x0.cpu(), tr["x0"])
    # (iv) free-running: graph and eager give the same bits; vs reference judged with the F9-aware metric
    out = model["inf"].generate_mel(idx, steps, x_T=xT)
    model["inf"].use_cuda_graph = False
    try:
        out_eager = model["inf"].generate_mel(idx, steps, x_T=xT)
    finally:
        model["inf"].use_cuda_graph = True
    assert torch.equal(out, out_eager)
    ref = run["x0"]
    d = (out.cpu() - ref).abs()
    rel = ((out.cpu() - ref).norm() / ref.norm()).item()
    assert rel <= 1e-3, rel
    # Share of elements beyond 1e-4: all of them descend from clamp-edge-band elements of step 0 (F9: gain 64,000 at t = 999;
    # test_generate_mel_free_running_f9_criterion traces that element by element), so the share scales with the teacher-forced
    # eps error: 1.2e-6 on the CUDA cores (3.5e-3 of the elements), 5.8e-6 on the tf32 x 3 tensor-core path (5.1e-3).
    assert (d > 1e-4).float().mean().item() < 1e-2


def test_generate_mel_rng_and_limits(model):
    idx = synth.synth_sem_idx(1, 2, 20).to(DEV)
    torch.manual_seed(123)
    a = model["inf"].generate_mel(idx, 4)
    torch.manual_seed(123)
    x_T = torch.randn(2, 40, 80, device=DEV)
    b = model["inf"].generate_mel(idx, 4, x_T=x_T)
    assert torch.equal(a, b) and a.shape == (2, 40, 80)
    assert a.abs().max().item() <= 3.0
    model["inf"].generate_mel(idx, 16)
    with pytest.raises(IndexError):
        model["inf"].generate_mel(idx, 17)


def test_batch_invariance(model):
    """A row's bits do not depend on the rest of the batch => batch shards reproduce the 1-GPU result."""
    idx = synth.synth_sem_idx(2, 6, 90).to(DEV)
    xT = synth.synth_noise(2, 6, 180).to(DEV)
    full = model["inf"].generate_mel(idx, 4, x_T=xT)
    parts = torch.cat([model["inf"].generate_mel(idx[a:b], 4, x_T=xT[a:b]) for a, b in ((0, 1), (1, 4), (4, 6))])
    assert torch.equal(full, parts)


def test_sample_ddpm_vs_oracle(model):
    B, S, n = 2, 30, 12
    idx = synth.synth_sem_idx(4, B, S)
    xT = synth.synth_noise(4, B, 2 * S)
    noises = [synth.synth_noise(4, B, 2 * S, tag=f"n{i}") for i in range(n)]
    ref = O.ddpm_loop(model["sd"], model["tab"], idx, xT, noises, t_start=300, t_end=300 - n + 1)
    out = model["inf"].sample_ddpm(idx.to(DEV), xT.to(DEV), [z.to(DEV) for z in noises], t_start=300,
                                   t_end=300 - n + 1, graph_steps=4).cpu()
    assert ((out - ref).norm() / ref.norm()).item() < 1e-4
    last = model["inf"].sample_ddpm(idx.to(DEV), xT.to(DEV), [z.to(DEV) for z in noises[:3]], t_start=2, t_end=0).cpu()
    ref_last = O.ddpm_loop(model["sd"], model["tab"], idx, xT, noises[:3], t_start=2, t_end=0)
    assert ((last - ref_last).norm() / ref_last.norm()).item() < 1e-4


# ------------------------------------------------------------------ depthwise-separable conv
def test_dsconv_golden(pkg, golden):
    g = golden("dsconv")
    for name, c in g["cases"].items():
        cin, cout, k, stride, B, T = c["shape"]
        m = pkg.DepthwiseSeparableConv(cin, cout, k, stride).to(DEV).eval()
        m.load_state_dict(synth.synth_dsconv_state(16, cin, cout, k), strict=True)
        x = synth.synth_noise(16, B, cin, T, tag="conv_" + name).to(DEV)
        y = m(x).cpu()
        assert y.shape == c["y"].shape, name
        assert (y - c["y"]).abs().max().item() < 2e-5, name


def test_dsconv_large_vs_oracle(pkg):
    cin = cout = 160
    B, T = 4, 801
    sd = synth.synth_dsconv_state(9, cin, cout, 3)
    m = pkg.DepthwiseSeparableConv(cin, cout).to(DEV).eval()
    m.load_state_dict(sd, strict=True)
    x = synth.synth_noise(9, B, cin, T, tag="big")
    assert (m(x.to(DEV)).cpu() - O.dsconv_forward(sd, x)).abs().max().item() < 2e-5
