"""VectorQuantizer with the reference's constructor, state dict and return
tuples (models/vq.py:9-163); eval-mode search/gather/histogram run as CUDA
kernels (edtts_vq_argmin / edtts_vq_gather_ste / edtts_vq_bincount)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib


class VectorQuantizer(nn.Module):
    def __init__(self, dim: int, codebook_size: int, commit: float = 0.25, decay: float = 0.99,
                 epsilon: float = 1e-5, reset_unused_every: int = 100):
        super().__init__()
        self.dim, self.codebook_size = dim, codebook_size
        self.commit, self.decay, self.epsilon = commit, decay, epsilon
        self.reset_unused_every = reset_unused_every
        self.codebook = nn.Embedding(codebook_size, dim)
        nn.init.normal_(self.codebook.weight, mean=0.0, std=1.0)            # vq.py:45
        self.register_buffer("ema_cluster_size", torch.ones(codebook_size))
        self.register_buffer("ema_w", self.codebook.weight.detach().clone())
        self.register_buffer("update_count", torch.tensor(0))
        self._ws = _lib.Workspace()

    @torch.no_grad()
    def encode(self, z: torch.Tensor) -> torch.Tensor:
        """vq.py:148-159: [B,T,D] -> int64 [B,T]."""
        if z.dim() != 3 or z.shape[-1] != self.dim:
            raise ValueError(f"z must be [B, T, {self.dim}], got {tuple(z.shape)}")
        lib = _lib.load()
        B, T, D = z.shape
        z = _lib.f32(z)
        idx = torch.empty(B, T, dtype=torch.int64, device=z.device)
        if B * T == 0:
            return idx
        cb = _lib.f32(self.codebook.weight.detach())
        # code norms + tensor-core codebook image: packed once per codebook version, then one launch per call
        key = (cb.data_ptr(), self.codebook.weight._version, str(z.device))
        c = self.__dict__.get("_packed")
        if c is None or c[0] != key:
            packed = torch.empty(max(int(lib.edtts_vq_workspace_bytes(self.codebook_size, D)), 16), dtype=torch.uint8, device=z.device)
            _lib.check(lib.edtts_vq_pack(_lib.ptr(cb), D, self.codebook_size, _lib.ptr(packed), _lib.stream_ptr(z.device)), "vq_pack")
            c = (key, packed, cb)
            self.__dict__["_packed"] = c
        _lib.check(lib.edtts_vq_argmin_packed(_lib.ptr(z), _lib.ptr(cb), _lib.ptr(c[1]), _lib.ptr(idx), B * T, D, self.codebook_size,
                                              _lib.stream_ptr(z.device)), "vq_argmin_packed")
        return idx

    def decode(self, idx: torch.Tensor) -> torch.Tensor:
        """vq.py:161-163."""
        return self.codebook(idx)

    def forward(self, z: torch.Tensor):
        """vq.py:53-107 (eval): (z_q, idx, vq_loss, perplexity, used)."""
        if self.training:
            raise NotImplementedError("VectorQuantizer training (losses, EMA update; vq.py:86-93,109-145) is outside "
                                      "the sampling path; call .eval()")
        lib = _lib.load()
        idx = self.encode(z)
        B, T, D = z.shape
        z = _lib.f32(z)
        with torch.no_grad():
            z_q = torch.empty_like(z)
            counts = torch.empty(self.codebook_size, dtype=torch.int32, device=z.device)
            cb = _lib.f32(self.codebook.weight.detach())
            st = _lib.stream_ptr(z.device)
            _lib.check(lib.edtts_vq_gather_ste(_lib.ptr(z), _lib.ptr(cb), _lib.ptr(idx), _lib.ptr(z_q), B * T, D,
                                               self.codebook_size, st), "vq_gather_ste")
            _lib.check(lib.edtts_vq_bincount(_lib.ptr(idx), _lib.ptr(counts), B * T, self.codebook_size, st),
                       "vq_bincount")
            counts = counts.float()                                          # 512 numbers: vq.py:102-105
            probs = counts / counts.sum().clamp_min(1.0)
            perplexity = torch.exp(-(probs * torch.log(probs.clamp_min(1e-12))).sum())
            used = (counts > 0).sum()
            vq_loss = torch.tensor(0.0, device=z.device)
        return z_q, idx, vq_loss, perplexity, used
