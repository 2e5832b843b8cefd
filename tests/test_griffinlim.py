"""Griffin-Lim (reference generate_sample.py:135-141, inference_pipeline.py:89,398 via torchaudio).
CPU: the oracle restatement of torchaudio.functional.griffinlim against the fixture recorded from torchaudio's own transform
  (tests/golden/griffinlim.pt, initial phases captured from its torch.rand call): bit-equal.
GPU: edtts_griffinlim (two kernels per iteration, shared-memory FFTs) against the fixture.  The iteration feeds every
  rounding difference of the FFTs back through 32 istft / stft rounds with momentum 0.99, so the bars are: 0 and 1 iterations
  max-abs <= 1e-5 of the peak; 8 / 32 iterations relative L2 <= 1e-3 (measured values printed)."""
import pytest
import torch

from oracle import edtts_oracle as O

DEV = "cuda:0"
CASES = ("reference", "small", "one_iter", "no_iter")


def _inputs(c):
    n_fft, win, hop, n_iter, frames, B = c["cfg"]
    g = torch.Generator().manual_seed(c["seed"])
    spec = torch.rand(B, n_fft // 2 + 1, frames, generator=g) ** 2 * 3.0
    init = torch.view_as_complex(torch.rand(B, n_fft // 2 + 1, frames, 2, generator=g).contiguous())
    return spec, init


def test_oracle_vs_torchaudio_fixture(golden):
    g = golden("griffinlim")
    for name in CASES:
        c = g["cases"][name]
        n_fft, win, hop, n_iter, frames, B = c["cfg"]
        spec, init = _inputs(c)
        wave = O.griffinlim(spec, torch.hann_window(win), n_fft, hop, win, 2.0, n_iter, 0.99, init)
        assert wave.shape == c["wave"].shape == (B, hop * (frames - 1))
        assert torch.equal(wave, c["wave"]), name


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_griffinlim_gpu_vs_fixture(lib, golden, name):
    import edge_diffusion_tts_b200 as E
    c = golden("griffinlim")["cases"][name]
    n_fft, win, hop, n_iter, frames, B = c["cfg"]
    spec, init = _inputs(c)
    tr = E.GriffinLim(n_fft=n_fft, n_iter=n_iter, win_length=win, hop_length=hop, power=2.0).to(DEV)
    assert torch.equal(tr.window.cpu(), torch.hann_window(win))
    wave = tr(spec.to(DEV), angles_init=init.to(DEV)).cpu()
    ref = c["wave"]
    assert wave.shape == ref.shape
    rel = ((wave - ref).double().norm() / ref.double().norm()).item()
    mx = (wave - ref).abs().max().item() / ref.abs().max().item()
    print(f"[griffinlim {name}] n_iter={n_iter}: rel-L2 {rel:.2e}, max-abs / peak {mx:.2e}")
    if n_iter <= 1:
        assert mx <= 1e-5, mx
    else:
        assert rel <= 1e-3, rel


@pytest.mark.gpu
def test_griffinlim_api(lib):
    import edge_diffusion_tts_b200 as E
    tr = E.GriffinLim(n_fft=1024, n_iter=4, win_length=1024, hop_length=160).to(DEV)
    spec = torch.rand(2, 3, 513, 30, device=DEV)
    torch.manual_seed(5)
    a = tr(spec)
    torch.manual_seed(5)
    b = tr(spec)
    assert a.shape == (2, 3, 160 * 29) and torch.equal(a, b) and torch.isfinite(a).all()      # torch.rand initial phases: seeded
    with pytest.raises(ValueError):
        tr(torch.rand(2, 512, 30, device=DEV))
    with pytest.raises(RuntimeError):
        tr(torch.rand(513, 30))
    with pytest.raises(ValueError):
        E.GriffinLim(momentum=1.0)
