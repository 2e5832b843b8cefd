"""Import and drive the UNMODIFIED reference (only where /root/reference exists,
i.e. the build container -- never on the GPU box).  TEST INFRASTRUCTURE ONLY.

Shims: stub ``matplotlib`` before import (utils/visualization.py:8
imports it), ``chdir`` to a temp dir before ``CFG()`` (config.py:165-166 creates
./data and ./run_edge_diffusion), never build ``SemanticEncoder`` (encoder.py:35
downloads HuBERT); ``use_fsq=False`` because the path uses VectorQuantizer (F5).
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import torch

from . import synth

REFERENCE_ROOT = os.environ.get("EDTTS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "edge_diffusion_tts"))


_ref = None


def import_reference():
    global _ref
    if _ref is not None:
        return _ref
    if not available():
        raise RuntimeError("reference not present at " + REFERENCE_ROOT)
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import edge_diffusion_tts as E  # noqa
    _ref = E
    return E


def make_cfg():
    E = import_reference()
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix="edtts_ref_"))
    try:
        cfg = E.CFG(device="cpu", use_fsq=False)
    finally:
        os.chdir(cwd)
    return cfg


class _EncoderStub(torch.nn.Module):
    """EdgeInference only calls ``encoder.eval()`` on the generate_mel path
    (inference.py:27); HuBERT is out of scope."""


def make_reference(seed: int = 0):
    """Reference modules carrying the synthetic weights of oracle/synth.py.
    Returns dict(cfg, decoder, vq, proj, schedule, inference, E)."""
    E = import_reference()
    cfg = make_cfg()
    dec = E.EdgeDiffusionDecoder(cfg)
    dec.load_state_dict(synth.synth_decoder_state(seed), strict=True)
    dec.eval()
    vq = E.VectorQuantizer(cfg.semantic_dim, cfg.codebook_size, commit=cfg.vq_commit)
    vq.load_state_dict(synth.synth_vq_state(seed), strict=True)
    vq.eval()
    proj = torch.nn.Sequential(torch.nn.Linear(768, cfg.semantic_dim), torch.nn.GELU(),
                               torch.nn.LayerNorm(cfg.semantic_dim),
                               torch.nn.Linear(cfg.semantic_dim, cfg.semantic_dim))  # encoder.py:41-46
    proj.load_state_dict(synth.synth_proj_state(seed), strict=True)
    proj.eval()
    sched = E.DiffusionSchedule(cfg.diff_steps, cfg.beta_start, cfg.beta_end, device="cpu")
    inf = E.EdgeInference(cfg, sched, _EncoderStub(), dec)
    return dict(cfg=cfg, decoder=dec, vq=vq, proj=proj, schedule=sched, inference=inf, E=E)


def reference_generate_mel(ref, sem_idx, num_steps, x_T):
    """Run the reference's generate_mel with injected initial noise: the
    reference draws ``torch.randn(B, 2S, n_mels, device)`` (inference.py:33);
    patch torch.randn for that one call so both sides start from the same x_T."""
    real = torch.randn
    calls = []

    def fake(*a, **k):
        calls.append(a)
        return x_T.clone()

    torch.randn = fake
    try:
        out = ref["inference"].generate_mel(sem_idx, num_steps=num_steps, temperature=1.0)
    finally:
        torch.randn = real
    assert len(calls) == 1
    return out


def reference_inpaint_refine(ref):
    """The reference's ``inpaint_teacher_refine`` is a closure nested inside ``main()`` of the top-level script
    inference_pipeline.py (lines 145-196) and cannot be imported.  Its source is cut out of the UNMODIFIED file with ``ast``
    and compiled with the closure variables it reads (cfg, schedule, device, teacher_decoder) supplied as globals -- the
    statements that run are the reference's own."""
    import ast
    import textwrap
    path = os.path.join(REFERENCE_ROOT, "inference_pipeline.py")
    src = open(path).read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "inpaint_teacher_refine")
    code = textwrap.dedent(ast.get_source_segment(src, fn))
    env = dict(torch=torch, cfg=ref["cfg"], schedule=ref["schedule"], device="cpu", teacher_decoder=ref["decoder"])
    exec(compile(code, path, "exec"), env)
    return env["inpaint_teacher_refine"]
