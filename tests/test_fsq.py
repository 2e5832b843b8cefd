"""FSQ (reference models/fsq.py:18-132).
CPU: the oracle restatement against the fixture recorded from the unmodified reference (bit-equal).
GPU: edtts_fsq_forward / edtts_fsq_decode against the oracle.  tanh differs by ulps between the CPU reference and CUDA, so
  * indices are bit-exact on every row whose scaled values are farther than 1e-4 from a rounding boundary (k + 0.5) --
    all but a handful of rows of the seeded inputs, and the test asserts that fraction;
  * z_q (the straight-through value) is within 2e-6 of the oracle on those rows;
  * codes_to_indices / indices_to_codes (no transcendental) are bit-exact everywhere, including the reference's
    first-dimension-fastest / last-dimension-fastest mismatch."""
import pytest
import torch

from oracle import edtts_oracle as O

DEV = "cuda:0"
LEVELS = ([8, 8, 8], [8, 6, 5, 5, 5], [5, 5])


def _z(levels, rows_b=3, rows_t=41, seed=None):
    seed = 17 + len(levels) if seed is None else seed
    return torch.randn(rows_b, rows_t, len(levels), generator=torch.Generator().manual_seed(seed)) * 1.5


def test_oracle_vs_reference_fixture(golden):
    g = golden("fsq")
    for levels in LEVELS:
        c = g["cases"][tuple(levels)]
        z_q, idx, _ = O.fsq_forward(levels, _z(levels))
        assert torch.equal(z_q, c["z_q"]) and torch.equal(idx, c["idx"])
        assert torch.equal(O.fsq_indices_to_codes(levels, idx), c["codes"])


@pytest.mark.gpu
@pytest.mark.parametrize("levels", LEVELS)
@pytest.mark.parametrize("shape", [(3, 41), (1, 1), (64, 400), (0, 5)])
def test_fsq_gpu_vs_oracle(lib, levels, shape):
    import edge_diffusion_tts_b200 as E
    m = E.FSQ(levels).to(DEV)
    assert m.codebook_size == int(torch.tensor(levels).prod()) and m.num_codes == m.codebook_size
    assert torch.equal(m._basis.cpu(), torch.cumprod(torch.tensor([1] + levels[:-1]), 0))
    z = _z(levels, *shape, seed=shape[0] * 7 + len(levels))
    z_q, idx = m(z.to(DEV))
    assert z_q.shape == z.shape and idx.shape == z.shape[:-1] and idx.dtype == torch.int64
    if z.numel() == 0:
        return
    o_zq, o_idx, zs = O.fsq_forward(levels, z)
    safe = ((zs - torch.floor(zs) - 0.5).abs() > 1e-4).all(dim=-1)
    assert safe.float().mean().item() > 0.99
    assert torch.equal(idx.cpu()[safe], o_idx[safe])
    assert (z_q.cpu() - o_zq)[safe].abs().max().item() <= 2e-6
    # no transcendental: bit-exact everywhere
    assert torch.equal(m.codes_to_indices(o_zq.to(DEV)).cpu(), o_idx)
    assert torch.equal(m.indices_to_codes(o_idx.to(DEV)).cpu(), O.fsq_indices_to_codes(levels, o_idx))


@pytest.mark.gpu
def test_fsq_errors(lib):
    import edge_diffusion_tts_b200 as E
    m = E.FSQ([8, 8, 8]).to(DEV)
    with pytest.raises(ValueError):
        m(torch.zeros(2, 3, 4, device=DEV))
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 3, 3))


# ---- FSQEncoder (models/fsq.py:135-222): the quantiser SemanticEncoder builds with the reference default use_fsq=True ----
ENC_LEVELS = ([4, 4, 3, 3, 2, 2, 2, 2], [8, 6, 5, 5, 5])


def _ze(levels, b=3, t=37, seed=None):
    seed = 23 + len(levels) if seed is None else seed
    return torch.randn(b, t, 128, generator=torch.Generator().manual_seed(seed))


def test_fsq_encoder_oracle_vs_reference_fixture(golden):
    from oracle import synth
    g = golden("fsq_encoder")
    for levels in ENC_LEVELS:
        c = g["cases"][tuple(levels)]
        sd = synth.synth_fsq_encoder_state(23, levels)
        z_q, idx, loss, ppl, used, _ = O.fsq_encoder_forward(sd, levels, _ze(levels))
        assert torch.equal(idx, c["idx"]) and torch.equal(idx, c["encode"])
        assert torch.equal(z_q, c["z_q"]) and float(loss) == float(c["loss"]) == 0.0
        assert torch.equal(ppl, c["perplexity"]) and int(used) == int(c["used"])
        assert torch.equal(O.fsq_encoder_decode(sd, levels, idx), c["decode"])


def test_fsq_encoder_state_dict_keys(golden):
    """Reference encoder checkpoints (use_fsq=True, the default) load strictly: same keys under ``vq.``."""
    import edge_diffusion_tts_b200 as E
    g = golden("fsq_encoder")
    cfg = E.CFG()
    assert cfg.use_fsq and cfg.fsq_levels == [4, 4, 3, 3, 2, 2, 2, 2]                # config.py:99-100
    enc = E.SemanticEncoder(cfg, load_hubert=False)
    assert isinstance(enc.vq, E.FSQEncoder) and enc.codebook_size == 2304
    assert sorted(k[3:] for k in enc.state_dict() if k.startswith("vq.")) == g["cases"][tuple(cfg.fsq_levels)]["keys"]
    assert isinstance(E.SemanticEncoder(E.CFG(use_fsq=False), load_hubert=False).vq, E.VectorQuantizer)


@pytest.mark.gpu
@pytest.mark.parametrize("levels", ENC_LEVELS)
@pytest.mark.parametrize("shape", [(3, 37), (1, 1), (64, 400), (0, 5)])
def test_fsq_encoder_gpu_vs_oracle(lib, levels, shape):
    """The fused proj_down -> FSQ -> proj_up kernel: indices bit-exact on rows away (> 1e-4 in the scaled domain) from a
    rounding boundary (the 128-term dot product and tanh differ by ulps from ATen's), z_q within 1e-5 there; decode exact to
    2e-6 everywhere; the usage metrics equal the oracle's when no row is near a boundary."""
    import edge_diffusion_tts_b200 as E
    from oracle import synth
    sd = synth.synth_fsq_encoder_state(23, levels)
    m = E.FSQEncoder(128, levels).to(DEV).eval()
    m.load_state_dict(sd, strict=True)
    assert m.codebook_size == int(torch.tensor(levels).prod())
    z = _ze(levels, *shape, seed=shape[0] * 5 + len(levels))
    z_q, idx, loss, ppl, used = m(z.to(DEV))
    assert z_q.shape == z.shape and idx.shape == z.shape[:-1] and idx.dtype == torch.int64 and float(loss) == 0.0
    if z.numel() == 0:
        return
    o_zq, o_idx, _, o_ppl, o_used, zs = O.fsq_encoder_forward(sd, levels, z)
    safe = ((zs - torch.floor(zs) - 0.5).abs() > 1e-4).all(dim=-1)
    assert safe.float().mean().item() > 0.99
    assert torch.equal(idx.cpu()[safe], o_idx[safe])
    assert torch.equal(m.encode(z.to(DEV)).cpu(), idx.cpu())
    assert (z_q.cpu() - o_zq)[safe].abs().max().item() <= 1e-5
    assert (m.decode(o_idx.to(DEV)).cpu() - O.fsq_encoder_decode(sd, levels, o_idx)).abs().max().item() <= 2e-6
    if bool(safe.all()):
        assert int(used) == int(o_used) and abs(float(ppl) - float(o_ppl)) <= 1e-3 * float(o_ppl)


@pytest.mark.gpu
def test_semantic_encoder_use_fsq(lib):
    """SemanticEncoder with the reference default (use_fsq=True): 768-d features -> proj -> FSQEncoder 5-tuple."""
    import edge_diffusion_tts_b200 as E
    from oracle import synth
    cfg = E.CFG(device=DEV)
    enc = E.SemanticEncoder(cfg, load_hubert=False).to(DEV).eval()
    enc.proj.load_state_dict(synth.synth_proj_state(0))
    enc.vq.load_state_dict(synth.synth_fsq_encoder_state(23, cfg.fsq_levels))
    h = torch.randn(2, 50, 768, generator=torch.Generator().manual_seed(1))
    z_q, idx, loss, ppl, used = enc.quantize_features(h.to(DEV))
    z = O.encoder_proj(synth.synth_proj_state(0), h)
    o_zq, o_idx, _, _, _, zs = O.fsq_encoder_forward(synth.synth_fsq_encoder_state(23, cfg.fsq_levels), cfg.fsq_levels, z)
    safe = ((zs - torch.floor(zs) - 0.5).abs() > 1e-3).all(dim=-1)
    assert safe.float().mean().item() > 0.95 and torch.equal(idx.cpu()[safe], o_idx[safe])
    assert torch.equal(enc.encode_features(h.to(DEV)).cpu(), idx.cpu())
    # decode_tokens is NOT the inverse of the forward index for unequal levels (the reference's basis mismatch, kept)
    want = O.fsq_encoder_decode(synth.synth_fsq_encoder_state(23, cfg.fsq_levels), cfg.fsq_levels, idx.cpu())
    assert (enc.decode_tokens(idx).cpu() - want).abs().max().item() <= 2e-6
    assert 0 <= int(idx.min()) and int(idx.max()) < enc.codebook_size == 2304
