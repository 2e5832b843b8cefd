"""CFG: the hyper-parameters the sampling path reads (reference config.py:51-153).

Field names and defaults are the reference's; an instance of the reference's
own ``CFG`` dataclass can be passed anywhere this one is accepted (only
attribute reads are performed).  One deliberate difference: constructing it
does not create ``./data`` / ``./run_edge_diffusion`` (config.py:165-166).
``use_fsq`` / ``fsq_levels`` keep the reference defaults (config.py:99-100):
``SemanticEncoder`` then builds an ``FSQEncoder``; BASELINE configs 2 and 5 name
the VectorQuantizer and pass ``use_fsq=False``.
"""
from __future__ import annotations

import random
from dataclasses import dataclass, field, fields

import torch


def get_device() -> str:
    """config.py:18-32 -- this implementation is CUDA-only."""
    return "cuda" if torch.cuda.is_available() else "cpu"


def set_seed(seed: int) -> None:
    """config.py:35-41."""
    random.seed(seed)
    try:
        import numpy as np
        np.random.seed(seed)
    except Exception:  # pragma: no cover
        pass
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


@dataclass
class CFG:
    seed: int = 42
    device: str = "cuda"
    n_mels: int = 80
    semantic_dim: int = 128
    codebook_size: int = 512
    vq_commit: float = 1.0
    use_fsq: bool = True
    fsq_levels: list = field(default_factory=lambda: [4, 4, 3, 3, 2, 2, 2, 2])
    hidden: int = 160
    layers: int = 4
    heads: int = 4
    ffn_mult: int = 2
    use_adaln: bool = True
    dropout: float = 0.2
    attn_window_size: int = 64
    diff_steps: int = 1000
    beta_start: float = 1e-4
    beta_end: float = 2e-2
    inference_steps: int = 4
    hubert_id: str = "facebook/hubert-base-ls960"
    hubert_layer: int = 9

    @classmethod
    def from_dict(cls, d: dict) -> "CFG":
        names = {f.name for f in fields(cls)}
        return cls(**{k: v for k, v in d.items() if k in names})

    def to_dict(self) -> dict:
        return {f.name: getattr(self, f.name) for f in fields(self)}


def check_supported(cfg) -> None:
    """The kernels are specialised for the reference's default architecture."""
    want = dict(n_mels=80, hidden=160, layers=4, heads=4, ffn_mult=2, attn_window_size=64, semantic_dim=128)
    bad = {k: getattr(cfg, k) for k, v in want.items() if getattr(cfg, k) != v}
    if bad or not getattr(cfg, "use_adaln", True):
        raise NotImplementedError(
            f"libedtts kernels are specialised for the default CFG {want} with use_adaln=True; got {bad}")
