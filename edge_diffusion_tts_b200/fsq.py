"""FSQ with the reference's constructor, buffers and methods (models/fsq.py:18-132); forward / codes_to_indices /
indices_to_codes run as CUDA kernels (edtts_fsq_forward / edtts_fsq_decode).  Inference only: the straight-through
value is computed (z_b + (q - z_b)) but no autograd graph is built.

Reference quirk kept as is: ``codes_to_indices`` flattens with the FIRST dimension fastest
(basis = cumprod([1] + levels[:-1])) while ``indices_to_codes`` decodes with the LAST dimension fastest, so the two are
inverse to each other only when all levels are equal.

``FSQEncoder`` (models/fsq.py:135-222) is the quantiser ``SemanticEncoder`` builds with the reference's default
``use_fsq=True``: proj_down -> FSQ -> proj_up with the VectorQuantizer 5-tuple interface, one fused kernel
(edtts_fsq_encoder)."""
from __future__ import annotations

import ctypes as C
from typing import List

import torch
import torch.nn as nn

from . import _lib


class FSQ(nn.Module):
    def __init__(self, levels: List[int]):
        super().__init__()
        self.levels = levels
        self.dim = len(levels)
        self.register_buffer("_levels", torch.tensor(levels, dtype=torch.int32))
        self.register_buffer("_basis", torch.cumprod(torch.tensor([1] + levels[:-1], dtype=torch.int64), dim=0))
        self.codebook_size = 1
        for l in levels:
            self.codebook_size *= l
        self._lv = (C.c_int32 * self.dim)(*levels)

    @property
    def num_codes(self) -> int:
        return self.codebook_size

    def _check(self, z: torch.Tensor) -> torch.Tensor:
        if z.shape[-1] != self.dim:
            raise ValueError(f"last dimension must be {self.dim}, got {tuple(z.shape)}")
        if z.device.type != "cuda":
            raise RuntimeError("FSQ runs on CUDA tensors only (no CPU fallback)")
        return _lib.f32(z)

    def bound(self, z: torch.Tensor) -> torch.Tensor:
        return torch.tanh(z)

    def quantize(self, z: torch.Tensor) -> torch.Tensor:
        """fsq.py:63-83 on an already bounded z: the quantised values are what forward() returns minus the
        straight-through rounding, i.e. indices_to_codes(codes_to_indices(.)) for equal levels; computed here by
        the forward kernel on atanh-free input is not possible, so this training-side helper stays a tensor expression."""
        half = (self._levels.float() - 1) / 2
        zq = torch.minimum(torch.clamp(torch.round((z + 1) * half), min=0), self._levels.float() - 1)
        return zq / half - 1

    @torch.no_grad()
    def forward(self, z: torch.Tensor):
        """fsq.py:84-108 -> (z_q, indices)."""
        z = self._check(z)
        lib = _lib.load()
        rows = z.numel() // self.dim
        z_q = torch.empty_like(z)
        idx = torch.empty(z.shape[:-1], dtype=torch.int64, device=z.device)
        _lib.check(lib.edtts_fsq_forward(_lib.ptr(z), C.cast(self._lv, C.c_void_p), self.dim, 0, _lib.ptr(z_q), _lib.ptr(idx),
                                         rows, _lib.stream_ptr(z.device)), "fsq_forward")
        return z_q, idx

    @torch.no_grad()
    def codes_to_indices(self, z_q: torch.Tensor) -> torch.Tensor:
        z_q = self._check(z_q)
        lib = _lib.load()
        idx = torch.empty(z_q.shape[:-1], dtype=torch.int64, device=z_q.device)
        _lib.check(lib.edtts_fsq_forward(_lib.ptr(z_q), C.cast(self._lv, C.c_void_p), self.dim, 1, None, _lib.ptr(idx),
                                         z_q.numel() // self.dim, _lib.stream_ptr(z_q.device)), "fsq_codes_to_indices")
        return idx

    @torch.no_grad()
    def indices_to_codes(self, indices: torch.Tensor) -> torch.Tensor:
        if indices.device.type != "cuda":
            raise RuntimeError("FSQ runs on CUDA tensors only (no CPU fallback)")
        indices = _lib.i64(indices)
        lib = _lib.load()
        codes = torch.empty(*indices.shape, self.dim, dtype=torch.float32, device=indices.device)
        _lib.check(lib.edtts_fsq_decode(_lib.ptr(indices), C.cast(self._lv, C.c_void_p), self.dim, _lib.ptr(codes),
                                        indices.numel(), _lib.stream_ptr(indices.device)), "fsq_decode")
        return codes


class FSQEncoder(nn.Module):
    """models/fsq.py:135-222: same constructor, sub-module names (``fsq``, ``proj_down``, ``proj_up``: reference encoder
    checkpoints load strictly), ``codebook_size``, ``forward`` 5-tuple, ``encode`` / ``decode``.  Inference only."""

    def __init__(self, input_dim: int, levels: List[int] = [8, 6, 5, 5, 5]):
        super().__init__()
        self.fsq = FSQ(list(levels))
        self.fsq_dim = len(levels)
        self.proj_down = nn.Linear(input_dim, self.fsq_dim)
        self.proj_up = nn.Linear(self.fsq_dim, input_dim)
        self.input_dim = input_dim

    @property
    def codebook_size(self) -> int:
        return self.fsq.codebook_size

    def _run(self, z, idx_in, want_zq: bool, want_idx: bool):
        lib = _lib.load()
        src = z if z is not None else idx_in
        if src.device.type != "cuda":
            raise RuntimeError("FSQEncoder runs on CUDA tensors only (no CPU fallback)")
        lead = tuple(src.shape[:-1]) if z is not None else tuple(src.shape)
        rows = 1
        for n in lead:
            rows *= n
        D = self.input_dim
        zq = torch.empty(*lead, D, dtype=torch.float32, device=src.device) if want_zq else None
        idx = torch.empty(lead, dtype=torch.int64, device=src.device) if want_idx else None
        W = [_lib.f32(t.detach()) for t in (self.proj_down.weight, self.proj_down.bias, self.proj_up.weight, self.proj_up.bias)]
        _lib.check(lib.edtts_fsq_encoder(_lib.ptr(z), _lib.ptr(idx_in), *[_lib.ptr(t) for t in W],
                                         C.cast(self.fsq._lv, C.c_void_p), self.fsq_dim, D, _lib.ptr(zq), _lib.ptr(idx), rows,
                                         _lib.stream_ptr(src.device)), "fsq_encoder")
        return zq, idx

    def _check(self, z: torch.Tensor) -> torch.Tensor:
        if z.shape[-1] != self.input_dim:
            raise ValueError(f"last dimension must be {self.input_dim}, got {tuple(z.shape)}")
        return _lib.f32(z)

    @torch.no_grad()
    def forward(self, z: torch.Tensor):
        """fsq.py:161-198 -> (z_q, idx, loss = 0, perplexity, used)."""
        z = self._check(z)
        z_q, idx = self._run(z, None, True, True)
        counts = self.fused_count_usage(idx)
        probs = counts / counts.sum().clamp_min(1.0)
        perplexity = torch.exp(-(probs * torch.log(probs.clamp_min(1e-12))).sum())
        used = (counts > 0).sum()
        return z_q, idx, torch.tensor(0.0, device=z.device), perplexity, used

    def fused_count_usage(self, indices: torch.Tensor) -> torch.Tensor:
        """fsq.py:200-209: float32 usage counts per code (edtts_vq_bincount: shared-memory histogram, no host sync)."""
        lib = _lib.load()
        flat = _lib.i64(indices).reshape(-1)
        if flat.numel() == 0:
            return torch.zeros(self.fsq.num_codes, dtype=torch.float32, device=flat.device)
        counts = torch.empty(self.fsq.num_codes, dtype=torch.int32, device=flat.device)
        _lib.check(lib.edtts_vq_bincount(_lib.ptr(flat), _lib.ptr(counts), flat.numel(), self.fsq.num_codes,
                                         _lib.stream_ptr(flat.device)), "vq_bincount")
        return counts.float()

    @torch.no_grad()
    def encode(self, z: torch.Tensor) -> torch.Tensor:
        """fsq.py:212-216."""
        return self._run(self._check(z), None, False, True)[1]

    @torch.no_grad()
    def decode(self, indices: torch.Tensor) -> torch.Tensor:
        """fsq.py:218-221: indices_to_codes (last dimension fastest, as the reference) -> proj_up."""
        return self._run(None, _lib.i64(indices), True, False)[0]
