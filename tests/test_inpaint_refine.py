"""In-painting refine loop of the long-form pipeline (reference inference_pipeline.py:145-196).
CPU: the oracle restatement against the fixture recorded from the reference's own (nested) function.
GPU: EdgeInference.inpaint_refine against the oracle -- the two streaming kernels bit-exact on random inputs, the whole
loop (5 steps from t = 500, with / without known frames and classifier-free guidance): fp32 rel-L2 <= 1e-3, bf16 <= 5e-2."""
import pytest
import torch

from oracle import edtts_oracle as O
from oracle import synth
from oracle.make_golden import inpaint_cases

DEV = "cuda:0"


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def test_oracle_vs_reference_fixture(golden):
    g = golden("inpaint")
    sd = synth.synth_decoder_state(0)
    tab = O.cosine_schedule(1000)
    feats, xc, known, noises, cases = inpaint_cases()
    for name, scale, with_known in cases:
        x = O.inpaint_refine(sd, tab, xc, feats, known if with_known else None, g["overlap"] if with_known else 0,
                             g["strength"], g["steps"], scale, noise=noises[0], known_noises=noises[1:])
        assert (x - g["x"][name]).abs().max().item() < 1e-4, name
        if with_known:
            assert torch.equal(x[:, :g["overlap"]], known)


@pytest.fixture(scope="module")
def gpu(lib):
    import edge_diffusion_tts_b200 as E
    assert torch.cuda.is_available() and lib.edtts_device_supported() == 1, "needs an sm_100 device"
    cfg = E.CFG(device=DEV)
    sd = synth.synth_decoder_state(0)
    dec = E.EdgeDiffusionDecoder(cfg).to(DEV).eval()
    dec.load_state_dict(sd, strict=True)
    sched = E.DiffusionSchedule(cfg.diff_steps, device=DEV)
    inf = E.EdgeInference(cfg, sched, torch.nn.Identity(), dec)
    return dict(E=E, cfg=cfg, sd=sd, dec=dec, sched=sched, inf=inf, lib=lib, tab=O.cosine_schedule(cfg.diff_steps))


@pytest.mark.gpu
def test_streaming_kernels_bit_exact(gpu):
    from edge_diffusion_tts_b200 import _lib
    lib, tab = gpu["lib"], gpu["tab"]
    gen = torch.Generator().manual_seed(5)
    B, T, D, L = 3, 37, 80, 9
    x, vc, vu = (torch.randn(B, T, D, generator=gen) for _ in range(3))
    known, nz = torch.randn(B, L, D, generator=gen), torch.randn(B, L, D, generator=gen)
    t = torch.tensor([500, 37, 999])
    tn = torch.tensor([400, 0, 800])
    co = torch.stack([tab["sqrt_alpha_bar"][t], tab["sqrt_one_minus_alpha_bar"][t], torch.sqrt(tab["alpha_bar"][tn]),
                      torch.sqrt(1 - tab["alpha_bar"][tn])], dim=1).contiguous()
    b = lambda v: v[:, None, None]
    d = lambda v: v.to(DEV).contiguous()
    st = _lib.stream_ptr(DEV)
    for scale, vun in ((1.0, None), (1.7, vu)):
        v = vc if vun is None else vun + scale * (vc - vun)
        x0 = torch.clamp(b(co[:, 0]) * x - b(co[:, 1]) * v, -3, 3)
        eps = b(co[:, 1]) * x + b(co[:, 0]) * v
        want = b(co[:, 2]) * x0 + b(co[:, 3]) * eps
        xd, x0d = torch.empty(B, T, D, device=DEV), torch.empty(B, T, D, device=DEV)
        xin, vcd, cod = d(x), d(vc), d(co)                    # keep the device tensors alive across the launch
        vud = d(vun) if vun is not None else None
        _lib.check(lib.edtts_vddim_step(_lib.ptr(xin), _lib.ptr(vcd), _lib.ptr(vud) if vud is not None else None, scale,
                                        _lib.ptr(cod), _lib.ptr(xd), _lib.ptr(x0d), B, T * D, st), "vddim")
        assert torch.equal(xd.cpu(), want) and torch.equal(x0d.cpu(), x0)
    xd, kd, nd, cod = d(x), d(known), d(nz), d(co)
    _lib.check(lib.edtts_inpaint_inject(_lib.ptr(xd), _lib.ptr(kd), _lib.ptr(nd), _lib.ptr(cod), B, T, L, D, st), "inject")
    want = x.clone()
    want[:, :L] = b(co[:, 0]) * known + b(co[:, 1]) * nz
    assert torch.equal(xd.cpu(), want)


@pytest.mark.gpu
@pytest.mark.parametrize("precision,bar", [("fp32", 1e-3), ("bf16", 5e-2)])
def test_inpaint_refine_vs_reference_fixture(gpu, golden, precision, bar):
    g = golden("inpaint")
    dec, inf = gpu["dec"], gpu["inf"]
    dec.precision = precision
    try:
        feats, xc, known, noises, cases = inpaint_cases()
        for name, scale, with_known in cases:
            x = inf.inpaint_refine(xc.to(DEV), feats.to(DEV), known.to(DEV) if with_known else None,
                                   g["overlap"] if with_known else 0, g["strength"], g["steps"], scale,
                                   noise=noises[0].to(DEV), known_noises=[n.to(DEV) for n in noises[1:]])
            assert rel_l2(x, g["x"][name]) <= bar, (name, rel_l2(x, g["x"][name]))
            if with_known:
                assert torch.equal(x[:, :g["overlap"]].cpu(), known)
    finally:
        dec.precision = "fp32"


@pytest.mark.gpu
def test_inpaint_refine_errors_and_rng(gpu):
    inf = gpu["inf"]
    x = torch.randn(1, 40, 80, device=DEV)
    f = torch.randn(1, 20, 128, device=DEV)
    with pytest.raises(IndexError):
        inf.inpaint_refine(x, f, strength=1.0)                # t_start = 1000: out of the tables, as in the reference
    with pytest.raises(ValueError):
        inf.inpaint_refine(x, f, known_mel=torch.zeros(1, 5, 80, device=DEV), overlap_len=8)
    torch.manual_seed(0)
    a = inf.inpaint_refine(x, f, strength=0.3, steps=3)
    torch.manual_seed(0)
    b = inf.inpaint_refine(x, f, strength=0.3, steps=3)
    assert torch.equal(a, b) and a.shape == x.shape and torch.isfinite(a).all()


@pytest.mark.gpu
def test_inpaint_refine_graph_equals_eager(gpu):
    """The refine loop replayed from a CUDA graph (static buffers per shape, re-used across the chunks of a long utterance)
    gives the bits of the eager launches, also when the same plan is replayed with new inputs and a new guidance scale."""
    inf = gpu["inf"]
    feats, xc, known, noises, _ = inpaint_cases()
    d = lambda t: t.to(DEV)
    for scale in (1.0, 1.7, 1.3):
        for shift in (0.0, 0.25):
            args = (d(xc) + shift, d(feats), d(known), 8, 0.5, 5, scale)
            kw = dict(noise=d(noises[0]), known_noises=[d(n) for n in noises[1:]])
            inf.use_cuda_graph = True
            a = inf.inpaint_refine(*args, **kw)
            inf.use_cuda_graph = False
            try:
                b = inf.inpaint_refine(*args, **kw)
            finally:
                inf.use_cuda_graph = True
            assert torch.equal(a, b), (scale, shift)
