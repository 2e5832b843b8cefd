"""Mel normalisation helpers with the reference's API (edge_diffusion_tts/utils/audio.py:10-19), on libedtts kernels.

CUDA tensors only (no CPU fallback).  The statistics are reductions (fp64 accumulation, <= 1e-6 relative to torch);
the element-wise parts are bit-exact given the statistics.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib


def _bcast_stat(s: torch.Tensor, B: int, M: int) -> torch.Tensor:
    """[B,1,M] / [1,1,M] / [B,M] / [M] statistics -> contiguous [B, M]."""
    s = _lib.f32(s)
    if s.dim() == 3:
        if s.shape[1] != 1:
            raise ValueError(f"statistics must be [B,1,n_mels], got {tuple(s.shape)}")
        s = s[:, 0]
    if s.dim() == 1:
        s = s[None]
    if s.shape[-1] != M or s.shape[0] not in (1, B):
        raise ValueError(f"statistics {tuple(s.shape)} do not broadcast to [{B}, {M}]")
    return s.expand(B, M).contiguous()


def normalize_mel(mel: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """``(mel - mean) / std, mean, std`` with mean / unbiased std over dim 1, std clamped to >= 1e-5
    (utils/audio.py:10-14).  mel [B, T, n_mels] -> ([B, T, n_mels], [B, 1, n_mels], [B, 1, n_mels])."""
    mel = _lib.f32(mel)
    if mel.dim() != 3:
        raise ValueError(f"mel must be [B, T, n_mels], got {tuple(mel.shape)}")
    B, T, M = mel.shape
    out = torch.empty_like(mel)
    mean = torch.empty(B, 1, M, dtype=torch.float32, device=mel.device)
    std = torch.empty_like(mean)
    _lib.check(_lib.load().edtts_normalize_mel(_lib.ptr(mel), _lib.ptr(out), _lib.ptr(mean), _lib.ptr(std), B, T, M,
                                               _lib.stream_ptr(mel.device)), "normalize_mel")
    return out, mean, std


def denormalize_mel(mel_n: torch.Tensor, mean: torch.Tensor, std: torch.Tensor) -> torch.Tensor:
    """``mel_n * std + mean`` (utils/audio.py:17-19)."""
    mel_n = _lib.f32(mel_n)
    if mel_n.dim() != 3:
        raise ValueError(f"mel_n must be [B, T, n_mels], got {tuple(mel_n.shape)}")
    B, T, M = mel_n.shape
    out = torch.empty_like(mel_n)
    mean_b, std_b = _bcast_stat(mean, B, M), _bcast_stat(std, B, M)      # named: a temporary would be freed (and its block reused) before the launch
    _lib.check(_lib.load().edtts_denormalize_mel(_lib.ptr(mel_n), _lib.ptr(mean_b), _lib.ptr(std_b), _lib.ptr(out), B, T, M,
                                                 _lib.stream_ptr(mel_n.device)), "denormalize_mel")
    return out


class InverseMelScale(torch.nn.Module):
    """torchaudio.transforms.InverseMelScale as the reference uses it (generate_sample.py:125-141,
    inference_pipeline.py:88,395): mel [..., n_mels, time] -> relu(least-squares solution of fb^T X = mel) [..., n_stft, time].
    Same constructor arguments and ``fb`` buffer; the filter bank comes from torchaudio (library table, built once on the
    host), its fp64 pseudo-inverse is applied to every frame by one kernel (edtts_inverse_mel).  Parity with torchaudio's
    ``gels`` solve: rel-L2 <= 1e-5 (the bank has full column rank for the reference's settings; a rank-deficient bank
    raises, as the solve would be ill-defined)."""

    def __init__(self, n_stft: int, n_mels: int = 128, sample_rate: int = 16000, f_min: float = 0.0, f_max=None,
                 norm=None, mel_scale: str = "htk", driver: str = "gels"):
        super().__init__()
        import torchaudio.functional as AF
        self.n_mels, self.sample_rate, self.f_min, self.driver = n_mels, sample_rate, f_min, driver
        self.f_max = f_max or float(sample_rate // 2)
        if f_min > self.f_max:
            raise ValueError("Require f_min: {} <= f_max: {}".format(f_min, self.f_max))
        if driver not in ["gels", "gelsy", "gelsd", "gelss"]:
            raise ValueError(f'driver must be one of ["gels", "gelsy", "gelsd", "gelss"]. Found {driver}.')
        fb = AF.melscale_fbanks(n_stft, self.f_min, self.f_max, self.n_mels, self.sample_rate, norm, mel_scale)
        if int(torch.linalg.matrix_rank(fb.double())) < min(fb.shape):
            raise ValueError("the mel filter bank is rank deficient: InverseMelScale is ill-defined for these settings")
        self.register_buffer("fb", fb)                                               # [n_stft, n_mels], as torchaudio
        self.register_buffer("pinv_fb", torch.linalg.pinv(fb.double().T).float().contiguous())   # [n_stft, n_mels]

    def forward(self, melspec: torch.Tensor) -> torch.Tensor:
        shape = melspec.size()
        n_mels, time = shape[-2], shape[-1]
        if self.n_mels != n_mels:
            raise ValueError("Expected an input with {} mel bins. Found: {}".format(self.n_mels, n_mels))
        mel = _lib.f32(melspec).reshape(-1, n_mels, time).contiguous()
        freq = self.fb.shape[0]
        out = torch.empty(mel.shape[0], freq, time, dtype=torch.float32, device=mel.device)
        _lib.check(_lib.load().edtts_inverse_mel(_lib.ptr(self.pinv_fb), _lib.ptr(mel), _lib.ptr(out), mel.shape[0], freq, n_mels,
                                                 time, _lib.stream_ptr(mel.device)), "inverse_mel")
        return out.view(shape[:-2] + (freq, time))


class GriffinLim(torch.nn.Module):
    """torchaudio.transforms.GriffinLim as the reference uses it after the sampling path (generate_sample.py:135-141,
    inference_pipeline.py:89,398: n_fft 1024, n_iter 32 / 100, win_length 1024, hop_length 160, power 2): same constructor
    arguments, ``window`` buffer and ``forward(specgram [..., n_fft // 2 + 1, frames]) -> waveform [..., hop (frames - 1)]``.
    The iteration (istft -> stft -> momentum phase update, torchaudio.functional.griffinlim) runs as two CUDA kernels per
    iteration with shared-memory FFTs (edtts_griffinlim); the random initial phases are drawn with the same call
    (``torch.rand(size, complex64)``) or injected with ``angles_init`` (extension, for parity tests)."""

    def __init__(self, n_fft: int = 400, n_iter: int = 32, win_length: Optional[int] = None, hop_length: Optional[int] = None,
                 window_fn=torch.hann_window, power: float = 2.0, wkwargs: Optional[dict] = None, momentum: float = 0.99,
                 length: Optional[int] = None, rand_init: bool = True) -> None:
        super().__init__()
        if not (0 <= momentum < 1):
            raise ValueError("momentum must be in the range [0, 1). Found: {}".format(momentum))
        self.n_fft = n_fft
        self.n_iter = n_iter
        self.win_length = win_length if win_length is not None else n_fft
        self.hop_length = hop_length if hop_length is not None else self.win_length // 2
        window = window_fn(self.win_length) if wkwargs is None else window_fn(self.win_length, **wkwargs)
        self.register_buffer("window", window)
        self.length = length
        self.power = power
        self.momentum = momentum
        self.rand_init = rand_init
        self._ws = _lib.Workspace()

    @torch.no_grad()
    def forward(self, specgram: torch.Tensor, angles_init: Optional[torch.Tensor] = None) -> torch.Tensor:
        lib = _lib.load()
        if specgram.device.type != "cuda":
            raise RuntimeError("GriffinLim runs on CUDA (B200) only; there is no CPU fallback")
        n_freq = self.n_fft // 2 + 1
        if specgram.dim() < 2 or specgram.shape[-2] != n_freq:
            raise ValueError(f"specgram must be [..., {n_freq}, frames], got {tuple(specgram.shape)}")
        if self.win_length > self.n_fft:
            raise RuntimeError("win_length must not exceed n_fft")                      # as torch.stft
        shape = specgram.shape
        spec = _lib.f32(specgram).reshape(-1, n_freq, shape[-1])
        B, _, F = spec.shape
        L = self.hop_length * (F - 1)
        if self.length is not None and self.length != L:
            raise NotImplementedError(f"length={self.length}: only the natural length hop_length * (frames - 1) = {L} is supported")
        dev = spec.device
        if angles_init is None:                                                         # functional.griffinlim: rand / ones
            angles_init = (torch.rand(spec.size(), dtype=torch.complex64, device=dev) if self.rand_init
                           else torch.full(spec.size(), 1, dtype=torch.complex64, device=dev))
        ang = torch.view_as_real(angles_init.to(device=dev, dtype=torch.complex64).reshape(B, n_freq, F).transpose(1, 2).contiguous())
        pad = self.n_fft - self.win_length                                              # torch.stft pads the window to n_fft, centred
        window = torch.nn.functional.pad(_lib.f32(self.window.to(dev)), (pad // 2, pad - pad // 2)).contiguous()
        wave = torch.empty(B, L, dtype=torch.float32, device=dev)
        nbytes = lib.edtts_griffinlim_workspace_bytes(B, F, self.n_fft)
        ws = self._ws.get(nbytes, dev)
        _lib.check(lib.edtts_griffinlim(_lib.ptr(spec), _lib.ptr(ang), _lib.ptr(window), _lib.ptr(wave), _lib.ptr(ws), nbytes, B, F,
                                        self.n_fft, self.hop_length, self.n_iter, float(self.power), float(self.momentum), L,
                                        _lib.stream_ptr(dev)), "griffinlim")
        return wave.reshape(shape[:-2] + (L,))
