"""The reference arm: the UNMODIFIED reference (baseline/_ref, placed by tools/install_reference.py) run through its own
public API -- ``EdgeInference.generate_mel`` (reference inference.py:23-53) with the reference's own ``EdgeDiffusionDecoder``
and ``DiffusionSchedule`` -- on the host cores (or, as a labelled context number, eager on a CUDA device).

None of this package's models, kernels or host code is on that path.  Only ``oracle/synth.py`` is used, to load the same
deterministic synthetic weights / tokens as the B200 arm (the reference zero-initialises out_proj and the AdaLN
projections: with construction-time weights its output is identically 0).  BENCH INFRASTRUCTURE ONLY.

Run-time shims for this image, none of which edits a reference file: ``matplotlib`` is stubbed in
sys.modules before the import (utils/visualization.py:8), the cwd is a temp dir while ``CFG()`` runs (config.py:165-166
creates ./data and ./run_edge_diffusion), ``SemanticEncoder`` is never built (encoder.py:35 downloads HuBERT; generate_mel
only calls ``encoder.eval()`` on it).
"""
from __future__ import annotations

import os
import sys
import tempfile
import time
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "edge_diffusion_tts"))


_mod = None


def import_reference():
    global _mod
    if _mod is not None:
        return _mod
    if not available():
        raise RuntimeError("baseline/_ref/edge_diffusion_tts is missing: run tools/install_reference.py where /root/reference exists")
    for m in ("matplotlib", "matplotlib.pyplot"):
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = types.ModuleType(m)
    sys.path.insert(0, REF_DIR)
    try:
        import edge_diffusion_tts as E
    finally:
        sys.path.remove(REF_DIR)
    assert os.path.realpath(os.path.dirname(E.__file__)).startswith(os.path.realpath(REF_DIR)), E.__file__
    _mod = E
    return E


class _EncoderStub(torch.nn.Module):
    """generate_mel only calls ``self.encoder.eval()`` (inference.py:27)."""


def make_inference(device: str = "cpu", seed: int = 0):
    """The reference's EdgeInference over its own decoder / schedule, synthetic weights loaded with strict=True."""
    E = import_reference()
    from oracle import synth
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix="edtts_ref_"))
    try:
        cfg = E.CFG(device=device, use_fsq=False)
    finally:
        os.chdir(cwd)
    dec = E.EdgeDiffusionDecoder(cfg)
    dec.load_state_dict(synth.synth_decoder_state(seed), strict=True)
    dec = dec.to(device).eval()
    sched = E.DiffusionSchedule(cfg.diff_steps, cfg.beta_start, cfg.beta_end, device=device)
    return E.EdgeInference(cfg, sched, _EncoderStub(), dec), cfg


def time_generate(inf, B: int, S: int, num_steps: int, steps: int, warmup: int, min_seconds: float = 0.0,
                  seed: int = 3, device: str = "cpu"):
    """(frames/s, ms per generate, repetitions) of the reference's generate_mel on [B, S] synthetic tokens."""
    from oracle import synth
    idx = synth.synth_sem_idx(seed, B, S).to(device)
    sync = (lambda: torch.cuda.synchronize(device)) if device != "cpu" else (lambda: None)
    for _ in range(warmup):
        inf.generate_mel(idx, num_steps)
    sync()
    t0 = time.perf_counter()
    n = 0
    while n < steps or (time.perf_counter() - t0) < min_seconds:
        out = inf.generate_mel(idx, num_steps)
        n += 1
    sync()
    dt = time.perf_counter() - t0
    assert tuple(out.shape) == (B, 2 * S, inf.cfg.n_mels)
    return B * 2 * S * n / dt, dt / n * 1e3, n
