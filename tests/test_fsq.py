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
