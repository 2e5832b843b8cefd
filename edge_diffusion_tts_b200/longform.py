"""Chunk loop of the reference's long-form script (inference_pipeline.py:217-393):
sliding 2 s chunks with a 0.5 s overlap, every chunk refined by the in-painting loop with the previous chunk's tail as
known frames, de-normalised with per-chunk statistics, exponentiated and overlap-added under a trapezoid window; then
weight normalisation, trim and a 5 x 3 smoothing.  The script does all of it inline in ``main()``; here it is a small
API on libedtts kernels: ``chunk_plan`` / ``crossfade_window`` (host integer / table code, as the script computes
them), ``MelStitcher`` (edtts_stitch_add, edtts_stitch_finalize) and ``generate_longform`` (the loop).

Audio I/O, the mel filter bank, HuBERT, InverseMelScale and Griffin-Lim stay with torchaudio / transformers (library
code either side of the path, SURVEY section 2 OUT); the caller passes the global latent sequence and the per-chunk
statistics the script derives from them.
"""
from __future__ import annotations

from typing import List, NamedTuple, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .audio import _bcast_stat


class Chunk(NamedTuple):
    start_sample: int
    end_sample: int
    start_lat: int
    end_lat: int


def chunk_plan(total_samples: int, sample_rate: int, chunk_seconds: float = 2.0, overlap_seconds: float = 0.5) -> List[Chunk]:
    """Sample and latent ranges of every chunk (inference_pipeline.py:218-225, 295-319): hop = chunk - overlap samples,
    ceil((total - overlap) / hop) chunks; the latent range is the chunk's time range at 16 kHz / 320 samples per latent."""
    chunk_samples = int(chunk_seconds * sample_rate)
    overlap_samples = int(overlap_seconds * sample_rate)
    hop_samples = chunk_samples - overlap_samples
    num_chunks = int(np.ceil((total_samples - overlap_samples) / hop_samples))
    plan = []
    for i in range(num_chunks):
        start_sample = i * hop_samples
        end_sample = start_sample + chunk_samples
        start_sec = start_sample / sample_rate
        end_sec = end_sample / sample_rate
        plan.append(Chunk(start_sample, end_sample, int(start_sec * 16000) // 320, int(end_sec * 16000) // 320))
    return plan


def crossfade_window(chunk_frames: int, overlap_frames: int, device="cpu") -> torch.Tensor:
    """[1, chunk_frames] trapezoid: linear fade-in over the first and fade-out over the last ``overlap_frames``
    (inference_pipeline.py:255-262; same tensor statements, so the table is the reference's)."""
    window_mask = torch.ones(1, chunk_frames, device=device)
    fade_len = overlap_frames
    fade_in = torch.linspace(0, 1, fade_len, device=device).unsqueeze(0)
    fade_out = torch.linspace(1, 0, fade_len, device=device).unsqueeze(0)
    window_mask[0, :fade_len] = fade_in
    window_mask[0, -fade_len:] = fade_out
    return window_mask


class MelStitcher:
    """Overlap-add accumulator of the chunk loop: ``final_mel`` [n_mels, frames] (or [B, n_mels, frames] for B
    utterances stitched in lock-step) and ``final_weights`` [1, frames] as in inference_pipeline.py:229-230."""

    def __init__(self, n_mels: int, buffer_frames: int, chunk_frames: int, overlap_frames: int, device, batch: Optional[int] = None):
        _lib.load()
        if torch.device(device).type != "cuda":
            raise RuntimeError("MelStitcher runs on CUDA (B200) only; there is no CPU fallback")
        self.n_mels, self.buffer_frames, self.chunk_frames, self.overlap_frames = n_mels, buffer_frames, chunk_frames, overlap_frames
        self.hop_frames = chunk_frames - overlap_frames                      # inference_pipeline.py:239
        self.batch = batch
        shape = (n_mels, buffer_frames) if batch is None else (batch, n_mels, buffer_frames)
        self.final_mel = torch.zeros(*shape, device=device)
        self.final_weights = torch.zeros(1, buffer_frames, device=device)
        self.window_mask = crossfade_window(chunk_frames, overlap_frames, device)

    def add_chunk(self, i: int, x_refined: torch.Tensor, mean: torch.Tensor, std: torch.Tensor) -> None:
        """``final_mel[:, s:e] += exp(denormalize_mel(x_refined, mean, std)).T * window`` and ``final_weights[:, s:e] +=
        window`` with s = i * hop_frames (inference_pipeline.py:359-375), one kernel.  x_refined [B, T, n_mels]; frames
        beyond ``chunk_frames`` are dropped as the script does."""
        x = _lib.f32(x_refined)
        B = 1 if self.batch is None else self.batch
        if x.dim() != 3 or x.shape[0] != B or x.shape[2] != self.n_mels:
            raise ValueError(f"x_refined must be [{B}, T, {self.n_mels}], got {tuple(x.shape)}")
        if x.shape[1] > self.chunk_frames:
            x = x[:, :self.chunk_frames].contiguous()
        if x.shape[1] != self.chunk_frames:
            raise RuntimeError(f"chunk has {x.shape[1]} frames, the window {self.chunk_frames}")      # the script's += would not broadcast
        start = i * self.hop_frames
        mean_b, std_b = _bcast_stat(mean, B, self.n_mels), _bcast_stat(std, B, self.n_mels)   # kept alive across the launch
        _lib.check(_lib.load().edtts_stitch_add(
            _lib.ptr(self.final_mel), _lib.ptr(self.final_weights), _lib.ptr(x), _lib.ptr(mean_b), _lib.ptr(std_b),
            _lib.ptr(self.window_mask), B, self.chunk_frames, self.n_mels, self.buffer_frames, start,
            _lib.stream_ptr(x.device)), "stitch_add")

    def finalize(self, total_frames: int, kernel_h: int = 5, kernel_w: int = 3) -> Tuple[torch.Tensor, torch.Tensor]:
        """(final_mel / clamp(final_weights, 1e-5))[..., :total_frames] and its (kernel_h x kernel_w) average-pooled copy
        (inference_pipeline.py:377-393); shapes [n_mels, total] and [1, n_mels, total] ([B, ...] with a batch)."""
        B = 1 if self.batch is None else self.batch
        dev = self.final_mel.device
        mel = torch.empty(B, self.n_mels, total_frames, device=dev)
        smooth = torch.empty_like(mel)
        _lib.check(_lib.load().edtts_stitch_finalize(
            _lib.ptr(self.final_mel), _lib.ptr(self.final_weights), _lib.ptr(mel), _lib.ptr(smooth), B, self.n_mels,
            self.buffer_frames, total_frames, kernel_h, kernel_w, _lib.stream_ptr(dev)), "stitch_finalize")
        if self.batch is None:
            return mel[0], smooth
        return mel, smooth


@torch.no_grad()
def generate_longform(inference, z_q_global: torch.Tensor, plan: Sequence[Chunk], chunk_stats: Sequence[Tuple[torch.Tensor, torch.Tensor]],
                      chunk_frames: int, overlap_frames: int, total_frames: int, refine_strength: float = 1.0,
                      refine_steps: int = 20, cfg_scale: float = 1.0, noises=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """The sliding-window loop of inference_pipeline.py:293-393 for one utterance: per chunk slice the global latents,
    draw x_T (unused by the script, drawn to keep its RNG stream) and x_coarse, refine with the previous chunk's last
    ``overlap_frames`` frames as known frames, overlap-add exp(de-normalised mel); then normalise, trim, smooth.

    ``inference``: an EdgeInference; ``z_q_global`` [1, L, semantic_dim]; ``chunk_stats[i]`` = (mean, std) of chunk i
    ([1,1,n_mels], the script takes them from the source audio); ``noises[i]`` optionally = (x_coarse, noise, known_noises)
    to inject the chunk's N(0,1) draws.  Returns (final_mel [n_mels, total_frames], smoothed [1, n_mels, total_frames])."""
    cfg = inference.cfg
    dev = z_q_global.device
    st = MelStitcher(cfg.n_mels, total_frames + 1000, chunk_frames, overlap_frames, dev)      # :228-230
    prev_mel_tail = None
    for i, ch in enumerate(plan):
        z_q_chunk = z_q_global[:, ch.start_lat:ch.end_lat, :].contiguous()
        if noises is None:
            torch.randn(1, chunk_frames, cfg.n_mels, device=dev)                              # x_T of :329
            x_coarse, nz, kn = torch.randn(1, chunk_frames, cfg.n_mels, device=dev), None, None
        else:
            x_coarse, nz, kn = noises[i]
        x_refined = inference.inpaint_refine(x_coarse, z_q_chunk, known_mel=prev_mel_tail, overlap_len=overlap_frames,
                                             strength=refine_strength, steps=refine_steps, cfg_scale=cfg_scale, noise=nz,
                                             known_noises=kn)
        prev_mel_tail = x_refined[:, -overlap_frames:, :].clone()
        mean, std = chunk_stats[i]
        st.add_chunk(i, x_refined, mean, std)
    return st.finalize(total_frames)
