"""EdgeDiffusionDecoder: drop-in for the reference module (models/decoder.py:14-109).

Same constructor (``EdgeDiffusionDecoder(cfg)``), same 92-entry ``state_dict``
(so reference checkpoints load with ``strict=True``), same
``forward`` signature and error behaviour.  The module tree below exists only to
own the parameters under the reference's names; no sub-module has a ``forward`` --
the whole evaluation is three C-ABI calls into libedtts.so:

  edtts_cond_prepare     t, step_idx -> AdaLN (scale, shift) for the 8 norms
  edtts_context_prepare  sem_idx / sem_features -> per-layer cross-attention K, V
  edtts_decoder_step     x_t -> eps  (optionally with the DDIM/DDPM update fused)

``prepare_context`` / ``prepare_cond`` / ``step`` expose the pieces so that
EdgeInference can hoist the step-invariant work out of the sampling loop
 and capture the loop in a CUDA graph.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .config import check_supported


class _Params(nn.Module):
    """Parameter holder (no forward: the math lives in libedtts.so)."""


class _RMSNormParams(_Params):
    def __init__(self, dim: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))                 # mla.py:50


class _AdaNormParams(_Params):
    def __init__(self, dim: int, cond_dim: int):
        super().__init__()
        self.norm = _RMSNormParams(dim)
        self.proj = nn.Linear(cond_dim, dim * 2)
        nn.init.zeros_(self.proj.weight)                            # transformer.py:61-62
        nn.init.zeros_(self.proj.bias)


class _PosTable(_Params):
    def __init__(self, dim: int, max_len: int):
        super().__init__()
        self.register_buffer("pe", positional_table(max_len, dim))  # persistent, as embeddings.py:130


def positional_table(max_len: int, dim: int) -> torch.Tensor:
    """embeddings.py:122-128 (also used to extend past 1000/512 rows)."""
    pe = torch.zeros(max_len, dim)
    position = torch.arange(0, max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, dim, 2) * (-math.log(10000.0) / dim))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


class _Block(_Params):
    """Parameters of DiffusionTransformerBlock (transformer.py:83-124)."""

    def __init__(self, H: int, ffn_mult: int):
        super().__init__()
        self.norm1 = _AdaNormParams(H, H)
        self.attn = _Params()
        self.attn.qkv = nn.Linear(H, 3 * H, bias=False)
        self.attn.proj = nn.Linear(H, H)
        self.norm2 = _RMSNormParams(H)
        ca = _Params()
        ca.q_proj = nn.Linear(H, H, bias=False)
        ca.kv_down_proj = nn.Linear(H, H // 2, bias=False)
        ca.kv_norm = _RMSNormParams(H // 2)
        ca.kv_up_proj = nn.Linear(H // 2, 2 * H, bias=False)
        ca.out_proj = nn.Linear(H, H, bias=False)
        self.cross_attn = ca
        self.norm3 = _AdaNormParams(H, H)
        self.ffn = _Params()
        hid = H * ffn_mult
        self.ffn.net = nn.Sequential(nn.Linear(H, 2 * hid), nn.Identity(), nn.Identity(), nn.Linear(hid, H),
                                     nn.Identity())                 # keys net.0 / net.3 (transformer.py:40-46)


_next_uid = [0]


class EdgeDiffusionDecoder(nn.Module):
    #: arithmetic of the contractions: "fp32" (max-abs 1e-4 parity bar; tf32 x 3 split products on the tcgen05 tensor cores,
    #: fp32 accumulate), "bf16" (fused tcgen05 kernel, 1e-2 rel-L2) or "fp32_simt" (the CUDA-core FFMA kernels, the checker
    #: the "fp32" path is tested against)
    precision: str = "fp32"

    def __init__(self, cfg):
        super().__init__()
        check_supported(cfg)
        H = cfg.hidden
        self.cfg = cfg
        self.token_emb = nn.Embedding(cfg.codebook_size, H)
        self.sem_proj = nn.Linear(cfg.semantic_dim, H)
        self.time_emb = nn.Sequential(nn.Identity(), nn.Linear(H, H), nn.GELU(), nn.Linear(H, H))
        self.step_emb = nn.Embedding(16, H)
        self.in_proj = nn.Linear(cfg.n_mels, H)
        self.pos_emb = _PosTable(H, 1000)
        self.context_pos_emb = _PosTable(H, 512)
        self.layers = nn.ModuleList([_Block(H, cfg.ffn_mult) for _ in range(cfg.layers)])
        self.final_norm = nn.LayerNorm(H)
        self.out_proj = nn.Linear(H, cfg.n_mels)
        nn.init.zeros_(self.out_proj.weight)                        # decoder.py:63-64
        nn.init.zeros_(self.out_proj.bias)
        half = H // 2                                               # embeddings.py:37-41
        self.register_buffer("_time_freqs", torch.exp(torch.arange(half, dtype=torch.float32)
                                                      * (-math.log(10000.0) / (half - 1))), persistent=False)
        self._wcache = None
        self.weights_epoch = 0          # bumped whenever the C view of the weights is rebuilt
        _next_uid[0] += 1
        self._uid = _next_uid[0]        # process-unique: graph / plan caches of callers are keyed on (uid, epoch)
        self._ws_ctx = _lib.Workspace()
        self._ws_step = _lib.Workspace()

    # ---- weights -> C struct ------------------------------------------------------
    def _weights(self, T: int, S: int) -> _lib.DecoderWeights:
        """Build (and cache) the edtts_decoder_weights view of the parameters; the PE
        tables are extended by their closed form when T/S exceed the stored rows (F7)."""
        ps = dict(self.named_parameters())
        dev = self.out_proj.weight.device
        if dev.type != "cuda":
            raise RuntimeError("EdgeDiffusionDecoder must be on a CUDA device (.to('cuda')); there is no CPU path")
        key = (tuple((p.data_ptr(), p._version) for p in ps.values()), self.precision)
        c = self._wcache
        if c is not None and c["key"] == key and c["pos_rows"] >= T and c["ctx_rows"] >= S:
            return c["w"]
        if c is not None:               # an extended table, once built, is kept: alternating short / long shapes
            T, S = max(T, c["pos_rows"]), max(S, c["ctx_rows"])      # must not rebuild (and re-capture) every call
        keep = []

        def P(t):
            t = t.detach()
            if t.dtype != torch.float32:
                raise TypeError("decoder parameters must be float32 (SURVEY F11)")
            t = t.contiguous()
            keep.append(t)
            return t.data_ptr()

        pos = self.pos_emb.pe if T <= self.pos_emb.pe.shape[0] else positional_table(T, self.cfg.hidden).to(dev)
        ctx = (self.context_pos_emb.pe if S <= self.context_pos_emb.pe.shape[0]
               else positional_table(S, self.cfg.hidden).to(dev))
        w = _lib.DecoderWeights()
        w.token_emb = P(self.token_emb.weight)
        w.sem_proj_w, w.sem_proj_b = P(self.sem_proj.weight), P(self.sem_proj.bias)
        w.time1_w, w.time1_b = P(self.time_emb[1].weight), P(self.time_emb[1].bias)
        w.time3_w, w.time3_b = P(self.time_emb[3].weight), P(self.time_emb[3].bias)
        w.step_emb = P(self.step_emb.weight)
        w.in_proj_w, w.in_proj_b = P(self.in_proj.weight), P(self.in_proj.bias)
        w.pos_pe, w.ctx_pe = P(pos), P(ctx)
        w.time_freqs = P(self._time_freqs)
        w.final_norm_w, w.final_norm_b = P(self.final_norm.weight), P(self.final_norm.bias)
        w.out_proj_w, w.out_proj_b = P(self.out_proj.weight), P(self.out_proj.bias)
        for i, blk in enumerate(self.layers):
            L = w.layers[i]
            L.norm1_norm_w = P(blk.norm1.norm.weight)
            L.norm1_proj_w, L.norm1_proj_b = P(blk.norm1.proj.weight), P(blk.norm1.proj.bias)
            L.attn_qkv_w = P(blk.attn.qkv.weight)
            L.attn_proj_w, L.attn_proj_b = P(blk.attn.proj.weight), P(blk.attn.proj.bias)
            L.norm2_w = P(blk.norm2.weight)
            L.q_proj_w = P(blk.cross_attn.q_proj.weight)
            L.kv_down_w = P(blk.cross_attn.kv_down_proj.weight)
            L.kv_norm_w = P(blk.cross_attn.kv_norm.weight)
            L.kv_up_w = P(blk.cross_attn.kv_up_proj.weight)
            L.cross_out_w = P(blk.cross_attn.out_proj.weight)
            L.norm3_norm_w = P(blk.norm3.norm.weight)
            L.norm3_proj_w, L.norm3_proj_b = P(blk.norm3.proj.weight), P(blk.norm3.proj.bias)
            L.ffn0_w, L.ffn0_b = P(blk.ffn.net[0].weight), P(blk.ffn.net[0].bias)
            L.ffn3_w, L.ffn3_b = P(blk.ffn.net[3].weight), P(blk.ffn.net[3].bias)
        w.codebook_size = self.token_emb.weight.shape[0]
        w.pos_rows, w.ctx_rows = pos.shape[0], ctx.shape[0]
        w.packed_bf16 = None
        w.pos_pe_cm = None
        if self.precision == "bf16":
            # the positional table once more, chunk-major [40][rows][4]: a row-per-thread walk of it is then coalesced
            w.pos_pe_cm = P(pos.detach().view(pos.shape[0], pos.shape[1] // 4, 4).permute(1, 0, 2))
            lib = _lib.load()
            packed = torch.empty(max(int(lib.edtts_packed_bf16_bytes()), 16), dtype=torch.uint8, device=dev)
            _lib.check(lib.edtts_pack_weights_bf16(w, _lib.ptr(packed), _lib.stream_ptr(dev)), "pack_weights_bf16")
            keep.append(packed)
            w.packed_bf16 = packed.data_ptr()
        self._wcache = dict(key=key, w=w, keep=keep, pos_rows=pos.shape[0], ctx_rows=ctx.shape[0])
        self.weights_epoch += 1
        return w

    def weights_token(self, T: int = 1, S: int = 1):
        """Refreshes the C view of the weights if a parameter changed (load_state_dict, in-place update, .to()) and returns
        a token that identifies decoder + weight image: callers that replay captured graphs compare it before every
        replay -- a graph holds raw pointers into the packed image of the epoch it was captured in."""
        self._weights(T, S)
        return (self._uid, self.weights_epoch)

    def _prec(self) -> int:
        try:
            return {"fp32": _lib.PREC_TF32X3, "bf16": _lib.PREC_BF16, "fp32_simt": _lib.PREC_FP32}[self.precision]
        except KeyError:
            raise ValueError(f"precision must be 'fp32', 'bf16' or 'fp32_simt', got {self.precision!r}") from None

    def alloc_kv(self, B: int, S: int, device) -> torch.Tensor:
        """The buffer prepare_context() fills and step() reads: [layers, B*S, 320] fp32 rows (k | v); for precision="fp32" a flat
        fp32 tensor that holds those rows followed by the operand images of the cross-attention (edtts_context_kv_bytes)."""
        lib = _lib.load()
        n = int(lib.edtts_context_kv_bytes(B, S, self._prec())) // 4
        if n == self.cfg.layers * B * S * 2 * self.cfg.hidden:
            return torch.empty(self.cfg.layers, B * S, 2 * self.cfg.hidden, dtype=torch.float32, device=device)
        return torch.empty(n, dtype=torch.float32, device=device)

    def workspace_bytes(self, B: int, T: int, S: int):
        """(context bytes, step bytes) a caller must provide to own the scratch itself."""
        lib = _lib.load()
        return (int(lib.edtts_context_workspace_bytes(B, S)),
                int(lib.edtts_decoder_workspace_bytes(B, T, S, self._prec())))

    # ---- the three stages -----------------------------------------------------------
    @torch.no_grad()
    def prepare_cond(self, t: torch.Tensor, step_idx: Optional[torch.Tensor], T: int = 1, S: int = 1,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """decoder.py:77-80 + the 8 AdaLN projections -> mod[B, 8, 320]."""
        lib = _lib.load()
        w = self._weights(T, S)
        t = _lib.i64(t)
        B = t.shape[0]
        if step_idx is not None:
            step_idx = _lib.i64(step_idx)
            if step_idx.shape != t.shape:
                raise ValueError("step_idx must have the shape of t")
        mod = out if out is not None else torch.empty(B, 2 * self.cfg.layers, 2 * self.cfg.hidden,
                                                      dtype=torch.float32, device=t.device)
        _lib.check(lib.edtts_cond_prepare(w, _lib.ptr(t), _lib.ptr(step_idx), None, _lib.ptr(mod), B,
                                          _lib.stream_ptr(t.device)), "cond_prepare")
        return mod

    @torch.no_grad()
    def prepare_context(self, sem_idx: Optional[torch.Tensor] = None, sem_features: Optional[torch.Tensor] = None,
                        T: int = 1, out: Optional[torch.Tensor] = None,
                        ws: Optional[torch.Tensor] = None) -> torch.Tensor:
        """decoder.py:83-93 + mla.py:144-153 for every layer -> kv[4, B*S, 320]."""
        if sem_features is None and sem_idx is None:
            raise ValueError("Either sem_idx or sem_features must be provided")       # decoder.py:90
        lib = _lib.load()
        src = sem_features if sem_features is not None else sem_idx
        B, S = src.shape[0], src.shape[1]
        w = self._weights(T, S)
        if sem_features is not None:                      # takes precedence, as decoder.py:83
            sem_features, sem_idx = _lib.f32(sem_features), None
        else:
            sem_idx = _lib.i64(sem_idx)
        kv = out if out is not None else self.alloc_kv(B, S, src.device)
        if kv.numel() * 4 < int(lib.edtts_context_kv_bytes(B, S, self._prec())) or kv.dtype != torch.float32:
            raise ValueError("context buffer too small for this precision: allocate it with alloc_kv()")
        nbytes = lib.edtts_context_workspace_bytes(B, S)
        if ws is None:
            ws = self._ws_ctx.get(nbytes, src.device)
        elif ws.numel() < nbytes:
            raise ValueError("context workspace too small")
        _lib.check(lib.edtts_context_prepare(w, _lib.ptr(sem_idx), _lib.ptr(sem_features), _lib.ptr(kv), _lib.ptr(ws),
                                             nbytes, B, S, self._prec(), _lib.stream_ptr(src.device)),
                   "context_prepare")
        return kv

    @torch.no_grad()
    def step(self, x_t: torch.Tensor, mod: torch.Tensor, kv: torch.Tensor, S: int, args: "_lib.StepArgs",
             ws: Optional[torch.Tensor] = None) -> None:
        """One evaluation + fused epilogue; outputs are the buffers named in ``args``."""
        lib = _lib.load()
        B, T, _ = x_t.shape
        w = self._weights(T, S)
        nbytes = lib.edtts_decoder_workspace_bytes(B, T, S, self._prec())
        if ws is None:
            ws = self._ws_step.get(nbytes, x_t.device)
        elif ws.numel() < nbytes:
            raise ValueError("decoder workspace too small")
        _lib.check(lib.edtts_decoder_step(w, _lib.ptr(x_t), _lib.ptr(mod), _lib.ptr(kv), args, _lib.ptr(ws), nbytes,
                                          B, T, S, self._prec(), _lib.stream_ptr(x_t.device)), "decoder_step")

    # ---- reference API ------------------------------------------------------------------
    def forward(self, x_t: torch.Tensor, t: torch.Tensor, sem_idx: Optional[torch.Tensor] = None,
                step_idx: Optional[torch.Tensor] = None, sem_features: Optional[torch.Tensor] = None) -> torch.Tensor:
        """decoder.py:66-109: x_t [B,T,n_mels], t [B] -> eps [B,T,n_mels]."""
        if sem_features is None and sem_idx is None:
            raise ValueError("Either sem_idx or sem_features must be provided")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and self.training:
            raise NotImplementedError("libedtts implements inference only; call .eval() / torch.no_grad()")
        x_t = _lib.f32(x_t)
        B, T, Mel = x_t.shape
        if Mel != self.cfg.n_mels:
            raise ValueError(f"x_t last dim must be {self.cfg.n_mels}")
        src = sem_features if sem_features is not None else sem_idx
        S = src.shape[1]
        with torch.no_grad():
            mod = self.prepare_cond(t, step_idx, T, S)
            kv = self.prepare_context(sem_idx, sem_features, T)
            eps = torch.empty_like(x_t)
            args = _lib.StepArgs()
            args.mode = _lib.STEP_EPS
            args.eps_out = eps.data_ptr()
            self.step(x_t, mod, kv, S, args)
        return eps
