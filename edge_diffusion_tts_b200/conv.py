"""DepthwiseSeparableConv (layers/conv.py:10-64) as an operator-level drop-in.
The reference defines and exports the module but never instantiates it in the
decoder; it is provided and parity-tested standalone.  The
sub-modules only own the parameters under the reference's names
(depthwise.weight, pointwise.{weight,bias}, norm.{weight,bias})."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib


class DepthwiseSeparableConv(nn.Module):
    def __init__(self, in_ch: int, out_ch: int, kernel_size: int = 3, stride: int = 1):
        super().__init__()
        padding = kernel_size // 2
        self.depthwise = nn.Conv1d(in_ch, in_ch, kernel_size, stride=stride, padding=padding, groups=in_ch,
                                   bias=False)
        self.pointwise = nn.Conv1d(in_ch, out_ch, 1, bias=True)
        self.norm = nn.GroupNorm(min(8, out_ch), out_ch)
        self.in_ch, self.out_ch, self.kernel_size, self.stride = in_ch, out_ch, kernel_size, stride
        self._ws = _lib.Workspace()

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [B, C_in, T] -> [B, C_out, T']."""
        if x.dim() != 3 or x.shape[1] != self.in_ch:
            raise ValueError(f"expected [B, {self.in_ch}, T], got {tuple(x.shape)}")
        lib = _lib.load()
        x = _lib.f32(x)
        B, _, T = x.shape
        k, s = self.kernel_size, self.stride
        t_out = (T + 2 * (k // 2) - k) // s + 1
        y = torch.empty(B, self.out_ch, max(t_out, 0), dtype=torch.float32, device=x.device)
        if B == 0 or t_out <= 0:
            return y
        nbytes = lib.edtts_dsconv_workspace_bytes(B, self.in_ch, self.out_ch, t_out)
        ws = self._ws.get(nbytes, x.device)
        W = [_lib.f32(t.detach()) for t in (self.depthwise.weight, self.pointwise.weight, self.pointwise.bias,
                                            self.norm.weight, self.norm.bias)]
        _lib.check(lib.edtts_dsconv_forward(_lib.ptr(x), *[_lib.ptr(t) for t in W], _lib.ptr(y), _lib.ptr(ws), nbytes,
                                            B, self.in_ch, self.out_ch, T, k, s, _lib.stream_ptr(x.device)),
                   "dsconv_forward")
        return y
