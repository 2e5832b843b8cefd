"""Mel normalisation helpers with the reference's API (edge_diffusion_tts/utils/audio.py:10-19), on libedtts kernels.

CUDA tensors only (no CPU fallback).  The statistics are reductions (fp64 accumulation, <= 1e-6 relative to torch);
the element-wise parts are bit-exact given the statistics.
"""
from __future__ import annotations

from typing import Tuple

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
