"""DiffusionSchedule: same constructor, attributes and methods as the reference
(schedule.py:11-266).  The two sampling updates run as CUDA kernels through the
C ABI (edtts_ddim_step / edtts_ddpm_step); the forward-process helpers
(q_sample, predict_*, get_v_target) are training utilities outside the sampling
path and stay one-line tensor expressions.

Tables are built with the reference's exact op sequence on the CPU and then moved
(SURVEY.md F10): cos/cumprod differ by ulps between CPU and CUDA, and parity is
defined against the CPU reference.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from . import _lib

_TABLES = ("betas", "alphas", "alpha_bar", "sqrt_alpha_bar", "sqrt_one_minus_alpha_bar", "sqrt_recip_alpha_bar",
           "sqrt_recip_alpha_bar_minus_one", "posterior_variance", "lambda_t")


class DiffusionSchedule:
    def __init__(self, T: int, beta_start: float = 1e-4, beta_end: float = 2e-2, device: str = "cpu"):
        self.T = T
        self.device = device
        s = 0.008                                             # cosine schedule; betas args ignored (F10)
        x = torch.linspace(0, T, T + 1)
        ac = torch.cos(((x / T) + s) / (1 + s) * torch.pi * 0.5) ** 2
        ac = ac / ac[0]
        self.betas = torch.clip(1 - (ac[1:] / ac[:-1]), 0.0001, 0.9999)
        self.alphas = 1.0 - self.betas
        self.alpha_bar = torch.cumprod(self.alphas, dim=0)
        self.sqrt_alpha_bar = torch.sqrt(self.alpha_bar)
        self.sqrt_one_minus_alpha_bar = torch.sqrt(1.0 - self.alpha_bar)
        self.sqrt_recip_alpha_bar = torch.sqrt(1.0 / self.alpha_bar)
        self.sqrt_recip_alpha_bar_minus_one = torch.sqrt(1.0 / self.alpha_bar - 1)
        alpha_bar_prev = F.pad(self.alpha_bar[:-1], (1, 0), value=1.0)
        self.posterior_variance = self.betas * (1.0 - alpha_bar_prev) / (1.0 - self.alpha_bar)
        self.lambda_t = torch.log(self.sqrt_alpha_bar / self.sqrt_one_minus_alpha_bar)
        if torch.device(device).type != "cpu":
            self.to(device)

    # ---- forward-process helpers (schedule.py:61-155) -------------------------
    def q_sample(self, x0, t, noise=None):
        if noise is None:
            noise = torch.randn_like(x0)
        return (self.sqrt_alpha_bar[t][:, None, None] * x0
                + self.sqrt_one_minus_alpha_bar[t][:, None, None] * noise), noise

    def predict_x0_from_eps(self, x_t, t, eps):
        return (self.sqrt_recip_alpha_bar[t][:, None, None] * x_t
                - self.sqrt_recip_alpha_bar_minus_one[t][:, None, None] * eps)

    def predict_x0_from_v(self, x_t, t, v):
        return self.sqrt_alpha_bar[t][:, None, None] * x_t - self.sqrt_one_minus_alpha_bar[t][:, None, None] * v

    def predict_eps_from_v(self, x_t, t, v):
        return self.sqrt_one_minus_alpha_bar[t][:, None, None] * x_t + self.sqrt_alpha_bar[t][:, None, None] * v

    def get_v_target(self, x0, noise, t):
        return self.sqrt_alpha_bar[t][:, None, None] * noise - self.sqrt_one_minus_alpha_bar[t][:, None, None] * x0

    # ---- sampling updates: CUDA kernels ------------------------------------------
    def _check(self, x_t: torch.Tensor, t: torch.Tensor):
        if x_t.dim() != 3:
            raise ValueError(f"x_t must be [B, T, D], got {tuple(x_t.shape)}")
        if t.shape != (x_t.shape[0],):
            raise ValueError(f"t must be [B]={x_t.shape[0]}, got {tuple(t.shape)}")
        if self.alpha_bar.device != x_t.device:
            raise RuntimeError(f"schedule tables are on {self.alpha_bar.device}, x_t on {x_t.device}; call .to()")

    def get_ddim_step(self, x_t, t, t_prev, eps_pred, eta: float = 0.0, noise: Optional[torch.Tensor] = None):
        """schedule.py:157-202 -> (x_prev, x0_pred).  ``noise`` optionally injects the
        N(0,1) draw used when eta > 0 (the reference calls randn_like)."""
        self._check(x_t, t)
        lib = _lib.load()
        x_t, eps_pred = _lib.f32(x_t), _lib.f32(eps_pred)
        t, t_prev = _lib.i64(t), _lib.i64(t_prev)
        if eta > 0 and noise is None:
            noise = torch.randn_like(x_t)
        x_prev, x0 = torch.empty_like(x_t), torch.empty_like(x_t)
        B = x_t.shape[0]
        _lib.check(lib.edtts_ddim_step(_lib.ptr(x_t), _lib.ptr(eps_pred), _lib.ptr(noise) if eta > 0 else None,
                                       _lib.ptr(self.alpha_bar), _lib.ptr(t), _lib.ptr(t_prev), float(eta),
                                       _lib.ptr(x_prev), _lib.ptr(x0), B, x_t[0].numel(),
                                       _lib.stream_ptr(x_t.device)), "ddim_step")
        return x_prev, x0

    def ddpm_step(self, x_t, t, eps_pred, noise: Optional[torch.Tensor] = None):
        """schedule.py:204-238 -> x_prev."""
        self._check(x_t, t)
        lib = _lib.load()
        x_t, eps_pred, t = _lib.f32(x_t), _lib.f32(eps_pred), _lib.i64(t)
        if noise is None:
            noise = torch.randn_like(x_t)
        x_prev = torch.empty_like(x_t)
        _lib.check(lib.edtts_ddpm_step(_lib.ptr(x_t), _lib.ptr(eps_pred), _lib.ptr(_lib.f32(noise)),
                                       _lib.ptr(self.alphas), _lib.ptr(self.alpha_bar), _lib.ptr(self.betas),
                                       _lib.ptr(self.posterior_variance), _lib.ptr(t), _lib.ptr(x_prev),
                                       x_t.shape[0], x_t[0].numel(), _lib.stream_ptr(x_t.device)), "ddpm_step")
        return x_prev

    def get_schedule_for_steps(self, num_steps: int) -> list:
        stride = self.T // num_steps
        return list(range(self.T - 1, 0, -stride))[:num_steps]

    def to(self, device) -> "DiffusionSchedule":
        self.device = device
        for k in _TABLES:
            setattr(self, k, getattr(self, k).to(device).contiguous())
        return self
