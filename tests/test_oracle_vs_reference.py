"""CPU, build container only: the oracle against the LIVE unmodified reference at /root/reference
(skipped on the GPU box, where the reference does not exist)."""
import pytest
import torch

from oracle import edtts_oracle as O
from oracle import ref_harness as R
from oracle import synth

pytestmark = pytest.mark.skipif(not R.available(), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    return R.make_reference(0)


def test_state_dict_layout(ref):
    sd = ref["decoder"].state_dict()
    shapes = synth.decoder_shapes()
    assert list(sd.keys()) == list(shapes.keys())
    assert all(tuple(sd[k].shape) == shapes[k] for k in shapes)
    assert torch.equal(sd["pos_emb.pe"], O.positional_table(1000)) and torch.equal(sd["context_pos_emb.pe"], O.positional_table(512))


def test_decoder_and_sampler(ref):
    sd = synth.synth_decoder_state(0)
    B, S = 2, 80
    idx, x = synth.synth_sem_idx(31, B, S), synth.synth_noise(31, B, 2 * S)
    t, si = torch.tensor([999, 120]), torch.tensor([0, 9])
    with torch.no_grad():
        assert (ref["decoder"](x, t, idx, si) - O.decoder_forward(sd, x, t, idx, si)).abs().max().item() < 1e-6
        assert (ref["decoder"](x, t, idx, None) - O.decoder_forward(sd, x, t, idx, None)).abs().max().item() < 1e-6
    tab = O.cosine_schedule(1000)
    for steps in (1, 4):
        a = R.reference_generate_mel(ref, idx, steps, x)
        b = O.generate_mel(sd, tab, idx, steps, x)
        assert (a - b).abs().max().item() < 1e-5


def test_reference_limits(ref):
    """F7 / F8: the unmodified reference rejects S > 512 and num_steps > 16."""
    with pytest.raises(RuntimeError):
        R.reference_generate_mel(ref, synth.synth_sem_idx(1, 1, 513), 1, synth.synth_noise(1, 1, 1026))
    with pytest.raises(IndexError):
        R.reference_generate_mel(ref, synth.synth_sem_idx(1, 1, 8), 17, synth.synth_noise(1, 1, 16))


def test_vq_and_conv(ref):
    z = synth.synth_features(33, 3, 70, 128)
    out = ref["vq"](z)
    mine = O.vq_forward(synth.synth_vq_state(0)["codebook.weight"], z)
    assert torch.equal(out[1], mine[1]) and torch.equal(out[0], mine[0]) and torch.equal(out[4], mine[4])
    m = ref["E"].layers.DepthwiseSeparableConv(32, 48, 5, 2).eval()
    sd = synth.synth_dsconv_state(2, 32, 48, 5)
    m.load_state_dict(sd)
    x = synth.synth_noise(2, 2, 32, 61, tag="c")
    with torch.no_grad():
        assert (m(x) - O.dsconv_forward(sd, x, 2)).abs().max().item() < 1e-6


def test_dpm_solver(ref):
    """DPMSolverPP.sample of the live reference (schedule.py:440-527) == the oracle restatement, bit for bit."""
    import importlib
    S = importlib.import_module("edge_diffusion_tts.schedule")
    sd = synth.synth_decoder_state(0)
    tab = O.cosine_schedule(1000)
    feats = synth.synth_features(41, 1, 12, 128)
    xT = synth.synth_noise(41, 1, 24)
    for order, steps in ((1, 3), (2, 4), (3, 6)):
        solver = S.DPMSolverPP(ref["schedule"], order=order)
        with torch.no_grad():
            want = solver.sample(ref["decoder"], xT, feats, num_steps=steps)
        assert torch.equal(solver.get_time_steps(steps, 950), O.dpm_time_steps(tab, steps, 950))
        got = O.dpm_sample(sd, tab, xT, feats, steps, order)
        assert (got - want).abs().max().item() < 2e-5, (order, steps)


def test_inpaint_refine(ref):
    """inpaint_teacher_refine, cut out of the reference script (ref_harness.reference_inpaint_refine), == the oracle."""
    from oracle.make_golden import inpaint_cases
    fn = R.reference_inpaint_refine(ref)
    sd = synth.synth_decoder_state(0)
    tab = O.cosine_schedule(1000)
    feats, xc, known, noises, cases = inpaint_cases()
    for name, scale, with_known in cases:
        it = iter(noises)
        real = torch.randn_like
        torch.randn_like = lambda a: next(it).clone()
        try:
            with torch.no_grad():
                want = fn(xc, feats, known_mel=known if with_known else None, overlap_len=8 if with_known else 0,
                          strength=0.5, steps=5, cfg_scale=scale)
        finally:
            torch.randn_like = real
        got = O.inpaint_refine(sd, tab, xc, feats, known if with_known else None, 8 if with_known else 0, 0.5, 5, scale,
                               noise=noises[0], known_noises=noises[1:])
        assert (got - want).abs().max().item() < 2e-5, name
