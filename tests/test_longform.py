"""Mel statistics and the chunk loop of the long-form pipeline (reference
edge_diffusion_tts/utils/audio.py:10-19 and inference_pipeline.py:217-393).
CPU: the oracle restatement and the host-side plan / window code against the fixture recorded from the reference's own
statements.  GPU: normalize_mel (statistics <= 1e-6 relative, fp64-accumulated; element-wise part bit-exact given the
statistics), denormalize_mel (bit-exact), MelStitcher (overlap-add and 5x3 smoothing: <= 2e-6 relative -- expf against the
CPU's vectorised exp), generate_longform against the oracle loop (fp32 rel-L2 <= 1e-3 in the linear-mel domain)."""
import pytest
import torch

from oracle import edtts_oracle as O
from oracle import synth
from oracle.make_golden import invmel_cases, longform_cases

DEV = "cuda:0"


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def test_oracle_vs_reference_fixture(golden):
    g, c = golden("longform"), longform_cases()
    mel_n, mean, std = O.normalize_mel(c["mel"])
    assert torch.equal(mean, g["mean"]) and torch.equal(std, g["std"]) and torch.equal(mel_n, g["mel_n"])
    assert std[2, 0, 7].item() == pytest.approx(1e-5)                      # the clamped constant bin
    assert torch.equal(O.denormalize_mel(mel_n, mean, std), g["back"])
    for key, plan in g["plans"].items():
        n, sr = (int(v) for v in key.split("@"))
        assert O.chunk_plan(n, sr) == plan, key
    assert torch.equal(O.crossfade_window(c["chunk_frames"], c["overlap"]), g["window"])
    final_mel, smooth = O.stitch(c["chunks"], c["stats"], c["chunk_frames"], c["overlap"], c["total"])
    assert torch.equal(final_mel, g["final_mel"]) and torch.equal(smooth, g["smooth"])


def test_host_plan_and_window(golden):
    """chunk_plan / crossfade_window of the package are host code: same numbers as the reference's statements."""
    import edge_diffusion_tts_b200 as E
    g, c = golden("longform"), longform_cases()
    for key, plan in g["plans"].items():
        n, sr = (int(v) for v in key.split("@"))
        assert [tuple(ch) for ch in E.chunk_plan(n, sr)] == plan, key
    assert torch.equal(E.crossfade_window(c["chunk_frames"], c["overlap"]), g["window"])
    with pytest.raises(RuntimeError):
        E.normalize_mel(c["mel"])                                          # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        E.MelStitcher(80, 100, 24, 6, "cpu")


@pytest.mark.gpu
def test_normalize_denormalize(golden, lib):
    import edge_diffusion_tts_b200 as E
    g, c = golden("longform"), longform_cases()
    mel = c["mel"].to(DEV)
    mel_n, mean, std = E.normalize_mel(mel)
    assert mean.shape == (3, 1, 80) and std.shape == (3, 1, 80)
    assert ((mean.cpu() - g["mean"]).abs() <= 1e-6 * g["mean"].abs().clamp_min(1.0)).all()
    assert ((std.cpu() - g["std"]).abs() <= 1e-6 * g["std"].abs()).all()
    assert std[2, 0, 7].item() == pytest.approx(1e-5)
    assert torch.equal(mel_n.cpu(), (c["mel"] - mean.cpu()) / std.cpu())    # element-wise part bit-exact given the statistics
    live = torch.ones(3, 1, 80, dtype=torch.bool)
    live[2, 0, 7] = False                                                   # (x - mean) / 1e-5 amplifies the mean's last ulp
    assert ((mel_n.cpu() - g["mel_n"]).abs() <= 2e-5)[live.expand_as(g["mel_n"])].all()
    back = E.denormalize_mel(g["mel_n"].to(DEV), g["mean"].to(DEV), g["std"].to(DEV))
    assert torch.equal(back.cpu(), g["back"])
    # ragged shapes: a single frame (std = NaN as torch), bins not a multiple of 32, shared [1,1,M] statistics
    one = torch.randn(2, 1, 80, device=DEV)
    _, m1, s1 = E.normalize_mel(one)
    assert torch.equal(m1, one) and torch.isnan(s1).all()
    odd = torch.randn(2, 301, 45, device=DEV)
    n2, m2, s2 = E.normalize_mel(odd)
    rn, rm, rs = O.normalize_mel(odd.cpu())
    assert torch.allclose(m2.cpu(), rm, rtol=0, atol=1e-6) and torch.allclose(s2.cpu(), rs, rtol=1e-6, atol=0)
    assert torch.allclose(n2.cpu(), rn, rtol=0, atol=1e-5)
    shared = E.denormalize_mel(odd, m2[:1], s2[:1])
    assert torch.equal(shared.cpu(), odd.cpu() * s2[:1].cpu() + m2[:1].cpu())


@pytest.mark.gpu
def test_stitcher_vs_fixture(golden, lib):
    import edge_diffusion_tts_b200 as E
    g, c = golden("longform"), longform_cases()
    st = E.MelStitcher(80, c["total"] + 1000, c["chunk_frames"], c["overlap"], DEV)
    assert torch.equal(st.window_mask.cpu(), g["window"])
    for i, (x, (m, s)) in enumerate(zip(c["chunks"], c["stats"])):
        st.add_chunk(i, x.to(DEV), m.to(DEV), s.to(DEV))
    final_mel, smooth = st.finalize(c["total"])
    assert final_mel.shape == g["final_mel"].shape and smooth.shape == g["smooth"].shape
    assert ((final_mel.cpu() - g["final_mel"]).abs() <= 2e-6 * g["final_mel"].abs()).all()
    assert ((smooth.cpu() - g["smooth"]).abs() <= 2e-6 * g["smooth"].abs()).all()
    # the smoothing alone is bit-exact given the stitched mel (same summation order as ATen's CPU kernel)
    ref_smooth = torch.nn.functional.avg_pool2d(final_mel.cpu()[None, None], (5, 3), 1, (2, 1)).squeeze(0)
    assert torch.equal(smooth.cpu(), ref_smooth)
    # uncovered frames have weight 0 -> 0 / 1e-5 = 0; frames past the buffer are refused as torch's += would be
    assert st.final_weights[0, c["total"] + 100:].abs().sum().item() == 0
    with pytest.raises(ValueError):
        st.add_chunk(10_000, c["chunks"][0].to(DEV), *[t.to(DEV) for t in c["stats"][0]])
    with pytest.raises(RuntimeError):
        st.add_chunk(0, c["chunks"][0][:, :5].to(DEV), *[t.to(DEV) for t in c["stats"][0]])


@pytest.mark.gpu
def test_stitcher_batched_full_size(lib):
    """BASELINE-sized property test: B utterances stitched in lock-step equal B single stitches; a constant chunk value
    under a window that sums to one reproduces the constant (partition of unity of the trapezoid)."""
    import edge_diffusion_tts_b200 as E
    B, T, L, n = 8, 172, 43, 12                                             # 2 s chunks / 0.5 s overlap at 22.05 kHz, hop 256
    hop, total = T - L, (T - L) * (n - 1) + T
    g = torch.Generator().manual_seed(3)
    xs = [torch.randn(B, T, 80, generator=g).to(DEV) for _ in range(n)]
    mean = (torch.randn(B, 1, 80, generator=g) - 5).to(DEV)
    std = (torch.rand(B, 1, 80, generator=g) + 0.5).to(DEV)
    stb = E.MelStitcher(80, total + 1000, T, L, DEV, batch=B)
    for i, x in enumerate(xs):
        stb.add_chunk(i, x, mean, std)
    mb, sb = stb.finalize(total)
    for b in (0, B - 1):
        s1 = E.MelStitcher(80, total + 1000, T, L, DEV)
        for i, x in enumerate(xs):
            s1.add_chunk(i, x[b:b + 1].contiguous(), mean[b:b + 1], std[b:b + 1])
        m1, sm1 = s1.finalize(total)
        assert torch.equal(m1, mb[b]) and torch.equal(sm1[0], sb[b])
    const = E.MelStitcher(80, total + 1000, T, L, DEV)
    zero, one = torch.zeros(1, 1, 80, device=DEV), torch.ones(1, 1, 80, device=DEV)
    for i in range(n):
        const.add_chunk(i, torch.full((1, T, 80), 0.75, device=DEV), zero, one)
    m, _ = const.finalize(total)
    inner = m[:, 1:total - 1]                                               # the very first / last frame have window weight 0
    assert torch.allclose(inner, torch.full_like(inner, float(torch.exp(torch.tensor(0.75)))), rtol=1e-6, atol=0)
    assert m[:, 0].abs().sum().item() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 1e-1)])
def test_generate_longform_vs_oracle(lib, precision, tol):
    import edge_diffusion_tts_b200 as E
    cfg = E.CFG(device=DEV)
    sd = synth.synth_decoder_state(0)
    dec = E.EdgeDiffusionDecoder(cfg).to(DEV).eval()
    dec.load_state_dict(sd, strict=True)
    dec.precision = precision
    sched = E.DiffusionSchedule(cfg.diff_steps, device=DEV)
    inf = E.EdgeInference(cfg, sched, torch.nn.Identity(), dec)
    tab = O.cosine_schedule(cfg.diff_steps)
    chunk_frames, overlap, steps = 32, 8, 3
    plan = E.chunk_plan(int(16000 * 0.9), 16000, chunk_seconds=0.4, overlap_seconds=0.1)   # 3 chunks, 20 latents each
    assert len(plan) == 3
    z = synth.synth_features(77, 1, plan[-1].end_lat, 128)
    total = (chunk_frames - overlap) * (len(plan) - 1) + chunk_frames - 3
    g = torch.Generator().manual_seed(9)
    stats = [(torch.randn(1, 1, 80, generator=g) - 5.0, 0.3 * torch.rand(1, 1, 80, generator=g) + 0.2) for _ in plan]
    noises = [(synth.synth_noise(100 + i, 1, chunk_frames, tag="xc"), synth.synth_noise(200 + i, 1, chunk_frames, tag="nz"),
               [synth.synth_noise(300 + 10 * i + k, 1, overlap, tag="kn") for k in range(steps)]) for i in range(len(plan))]
    ref_mel, ref_smooth, _ = O.longform_generate(sd, tab, z, [tuple(p) for p in plan], stats, chunk_frames, overlap, total, 0.5, steps,
                                                 1.5, noises)
    d = lambda t: t.to(DEV)
    mel, smooth = E.generate_longform(inf, d(z), plan, [(d(m), d(s)) for m, s in stats], chunk_frames, overlap, total,
                                      refine_strength=0.5, refine_steps=steps, cfg_scale=1.5,
                                      noises=[(d(a), d(b), [d(k) for k in c]) for a, b, c in noises])
    assert mel.shape == (80, total) and smooth.shape == (1, 80, total)
    assert rel_l2(mel, ref_mel) <= tol and rel_l2(smooth, ref_smooth) <= tol


def test_inverse_mel_oracle_vs_torchaudio_fixture(golden):
    g = golden("invmel")
    mel, settings = invmel_cases()
    for name in settings:
        c = g["cases"][name]
        assert rel_l2(O.inverse_mel_scale(c["fb"], mel), c["spec"]) <= 1e-6, name     # LAPACK blocking may differ with the thread count


@pytest.mark.gpu
def test_inverse_mel_scale(golden, lib):
    """relu(pinv(fb^T) mel) against torchaudio's per-frame least-squares solve (fixture recorded with torchaudio itself):
    rel-L2 <= 1e-5, same support of the clamp up to values below 1e-5 of the peak."""
    import edge_diffusion_tts_b200 as E
    g = golden("invmel")
    mel, settings = invmel_cases()
    for name, kw in settings.items():
        c = g["cases"][name]
        m = E.InverseMelScale(**kw).to(DEV)
        assert torch.equal(m.fb.cpu(), c["fb"])
        out = m(mel.to(DEV))
        assert out.shape == c["spec"].shape and (out >= 0).all()
        assert rel_l2(out, c["spec"]) <= 1e-5, name
        assert (out.cpu() - c["spec"]).abs().max().item() <= 1e-5 * c["spec"].abs().max().item(), name
    # ragged: one frame, frames not a multiple of the tile, leading batch dimensions as torchaudio accepts them
    m = E.InverseMelScale(n_stft=513, n_mels=80, sample_rate=16000).to(DEV)
    fb = g["cases"]["generate_sample"]["fb"]                # inference_pipeline.py:88: same bank (f_max defaults to 8000)
    for shape in ((80, 1), (3, 2, 80, 131)):
        x = torch.rand(*shape, generator=torch.Generator().manual_seed(5)) + 0.01
        out = m(x.to(DEV))
        assert out.shape == shape[:-2] + (513, shape[-1])
        assert rel_l2(out, O.inverse_mel_scale(fb, x)) <= 1e-5
    with pytest.raises(ValueError):
        m(torch.rand(1, 64, 10, device=DEV))
