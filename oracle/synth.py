"""Deterministic synthetic weights and inputs shared by the oracle, the golden
generator, the tests and bench.py.  TEST/BENCH INFRASTRUCTURE ONLY (see
oracle/edtts_oracle.py header).

Why synthetic weights rather than "construct the reference module with seed 0":
the reference is absent on the GPU box, so the weights must be reproducible
without it.  Every tensor is drawn from its own ``torch.Generator`` keyed by
(seed, state-dict key), so the values do not depend on construction order.
The zero-initialised tensors of the reference (``out_proj``, every
``AdaLayerNorm.proj``, ``final_norm.bias``) get non-zero values,
otherwise a random-init decoder returns exactly 0 and parity is vacuous.

Shapes follow the 92-entry state dict of ``EdgeDiffusionDecoder(CFG())``
(models/decoder.py:17-64).
"""
from __future__ import annotations

import hashlib
import zlib
from typing import Dict

import torch

from . import edtts_oracle as O

Tensor = torch.Tensor


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFF)
    return g


def _uniform(seed, key, shape, bound):
    return (torch.rand(shape, generator=_gen(seed, key), dtype=torch.float32) * 2 - 1) * bound


def _normal(seed, key, shape, std, mean=0.0):
    return torch.randn(shape, generator=_gen(seed, key), dtype=torch.float32) * std + mean


def decoder_shapes(H=O.HIDDEN, M=O.N_MELS, layers=O.LAYERS, rank=O.KV_RANK, ffn=O.FFN_HIDDEN,
                   sem=O.SEMANTIC_DIM, codebook=O.CODEBOOK) -> Dict[str, tuple]:
    s = {
        "token_emb.weight": (codebook, H),
        "sem_proj.weight": (H, sem), "sem_proj.bias": (H,),
        "time_emb.1.weight": (H, H), "time_emb.1.bias": (H,),
        "time_emb.3.weight": (H, H), "time_emb.3.bias": (H,),
        "step_emb.weight": (16, H),
        "in_proj.weight": (H, M), "in_proj.bias": (H,),
        "pos_emb.pe": (1000, H), "context_pos_emb.pe": (512, H),
    }
    for i in range(layers):
        p = f"layers.{i}."
        s.update({
            p + "norm1.norm.weight": (H,), p + "norm1.proj.weight": (2 * H, H), p + "norm1.proj.bias": (2 * H,),
            p + "attn.qkv.weight": (3 * H, H), p + "attn.proj.weight": (H, H), p + "attn.proj.bias": (H,),
            p + "norm2.weight": (H,),
            p + "cross_attn.q_proj.weight": (H, H), p + "cross_attn.kv_down_proj.weight": (rank, H),
            p + "cross_attn.kv_norm.weight": (rank,), p + "cross_attn.kv_up_proj.weight": (2 * H, rank),
            p + "cross_attn.out_proj.weight": (H, H),
            p + "norm3.norm.weight": (H,), p + "norm3.proj.weight": (2 * H, H), p + "norm3.proj.bias": (2 * H,),
            p + "ffn.net.0.weight": (2 * ffn, H), p + "ffn.net.0.bias": (2 * ffn,),
            p + "ffn.net.3.weight": (H, ffn), p + "ffn.net.3.bias": (H,),
        })
    s.update({"final_norm.weight": (H,), "final_norm.bias": (H,),
              "out_proj.weight": (M, H), "out_proj.bias": (M,)})
    return s


def synth_decoder_state(seed: int = 0) -> Dict[str, Tensor]:
    """92-entry decoder state dict with well-conditioned synthetic values."""
    sd = {}
    for k, shape in decoder_shapes().items():
        if k.endswith(".pe"):
            sd[k] = O.positional_table(shape[0], shape[1])
        elif k in ("token_emb.weight", "step_emb.weight"):
            sd[k] = _normal(seed, k, shape, 1.0)
        elif "norm" in k and k.endswith("weight") and len(shape) == 1:
            sd[k] = _normal(seed, k, shape, 0.1, 1.0)          # norm gains around 1
        elif k.endswith("norm1.proj.weight") or k.endswith("norm3.proj.weight"):
            sd[k] = _normal(seed, k, shape, 0.02)              # zero-init in the reference (F6)
        elif k.endswith("norm1.proj.bias") or k.endswith("norm3.proj.bias") or k == "final_norm.bias":
            sd[k] = _normal(seed, k, shape, 0.02)
        elif k == "out_proj.weight":
            sd[k] = _uniform(seed, k, shape, 1.0 / shape[1] ** 0.5)
        elif len(shape) == 2:
            sd[k] = _uniform(seed, k, shape, 1.0 / shape[1] ** 0.5)   # nn.Linear default bound
        else:
            sd[k] = _uniform(seed, k, shape, 0.05)             # biases
    assert len(sd) == 92
    return sd


def synth_vq_state(seed: int = 0, dim=O.SEMANTIC_DIM, K=O.CODEBOOK) -> Dict[str, Tensor]:
    """VectorQuantizer state (vq.py:44-50): codebook ~ N(0,1) as vq.py:45."""
    cb = _normal(seed, "vq.codebook.weight", (K, dim), 1.0)
    return {"codebook.weight": cb, "ema_cluster_size": torch.ones(K), "ema_w": cb.clone(),
            "update_count": torch.tensor(0)}


def synth_proj_state(seed: int = 0, dim=O.SEMANTIC_DIM) -> Dict[str, Tensor]:
    """SemanticEncoder.proj state (encoder.py:41-46), Sequential keys 0,2,3."""
    return {
        "0.weight": _uniform(seed, "proj.0.weight", (dim, 768), 768 ** -0.5),
        "0.bias": _uniform(seed, "proj.0.bias", (dim,), 768 ** -0.5),
        "2.weight": _normal(seed, "proj.2.weight", (dim,), 0.1, 1.0),
        "2.bias": _normal(seed, "proj.2.bias", (dim,), 0.05),
        "3.weight": _uniform(seed, "proj.3.weight", (dim, dim), dim ** -0.5),
        "3.bias": _uniform(seed, "proj.3.bias", (dim,), dim ** -0.5),
    }


def synth_fsq_encoder_state(seed: int, levels, dim=O.SEMANTIC_DIM) -> Dict[str, Tensor]:
    """FSQEncoder state (models/fsq.py:148-154): fsq._levels / fsq._basis buffers + the two projections."""
    n = len(levels)
    return {
        "fsq._levels": torch.tensor(list(levels), dtype=torch.int32),
        "fsq._basis": torch.cumprod(torch.tensor([1] + list(levels[:-1]), dtype=torch.int64), dim=0),
        "proj_down.weight": _uniform(seed, "fsq.proj_down.weight", (n, dim), 3 * dim ** -0.5),
        "proj_down.bias": _uniform(seed, "fsq.proj_down.bias", (n,), dim ** -0.5),
        "proj_up.weight": _uniform(seed, "fsq.proj_up.weight", (dim, n), n ** -0.5),
        "proj_up.bias": _uniform(seed, "fsq.proj_up.bias", (dim,), n ** -0.5),
    }


def synth_dsconv_state(seed: int, in_ch: int, out_ch: int, k: int = 3) -> Dict[str, Tensor]:
    """DepthwiseSeparableConv state (conv.py:31-49)."""
    return {
        "depthwise.weight": _uniform(seed, "dw.weight", (in_ch, 1, k), k ** -0.5),
        "pointwise.weight": _uniform(seed, "pw.weight", (out_ch, in_ch, 1), in_ch ** -0.5),
        "pointwise.bias": _uniform(seed, "pw.bias", (out_ch,), in_ch ** -0.5),
        "norm.weight": _normal(seed, "gn.weight", (out_ch,), 0.1, 1.0),
        "norm.bias": _normal(seed, "gn.bias", (out_ch,), 0.05),
    }


def synth_sem_idx(seed: int, B: int, S: int, K=O.CODEBOOK) -> Tensor:
    return torch.randint(0, K, (B, S), generator=_gen(seed, "sem_idx"), dtype=torch.long)


def synth_noise(seed: int, B: int, T: int, M=O.N_MELS, tag: str = "x_T") -> Tensor:
    return torch.randn((B, T, M), generator=_gen(seed, tag), dtype=torch.float32)


def synth_features(seed: int, B: int, S: int, D: int = 768) -> Tensor:
    return torch.randn((B, S, D), generator=_gen(seed, "hubert"), dtype=torch.float32)


def state_checksum(sd: Dict[str, Tensor]) -> str:
    """Stable digest of a state dict; stored in the golden fixtures so a test can
    tell 'weights regenerated differently' from 'kernel wrong'."""
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().contiguous().cpu().numpy().tobytes())
    return h.hexdigest()[:16]


# The five BASELINE.json configs as (B, S, steps, kind).
BASELINE_CONFIGS = {
    "cfg1": dict(B=1, S=200, steps=4, kind="ddim"),
    "cfg2": dict(B=64, S=400, steps=1, kind="vq+ddim"),
    "cfg3": dict(B=256, S=400, steps=4, kind="ddim"),
    "cfg4": dict(B=32, S=400, steps=1000, kind="ddpm"),
    "cfg5": dict(B=128, S=1500, steps=4, kind="vq+ddim"),
}
