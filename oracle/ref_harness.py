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


def _main_statements(names, region, if_tests=()):
    """Source of the statements of ``main()`` in the UNMODIFIED inference_pipeline.py that lie ``region`` = "before" /
    "inside" / "after" its chunk loop (``for i in tqdm.tqdm(range(num_chunks))``), at any nesting depth and in source
    order, and assign to one of ``names`` (plain, subscript or augmented assignment) or are an ``if`` whose test
    mentions one of ``if_tests`` (taken whole).  The chunk loop is inline script code; this runs the reference's own
    statements on supplied inputs without the audio / HuBERT / torchaudio context around them."""
    import ast
    import textwrap
    path = os.path.join(REFERENCE_ROOT, "inference_pipeline.py")
    src = open(path).read()
    main = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "main")
    loop = next(n for n in main.body if isinstance(n, ast.For) and "num_chunks" in ast.get_source_segment(src, n.iter))
    lo, hi = {"before": (main.lineno, loop.lineno - 1), "inside": (loop.lineno + 1, loop.end_lineno),
              "after": (loop.end_lineno + 1, main.end_lineno)}[region]

    def base(t):
        while isinstance(t, ast.Subscript):
            t = t.value
        return t.id if isinstance(t, ast.Name) else None

    picked = []

    def visit(body):
        for s in body:
            if isinstance(s, ast.FunctionDef) or s.end_lineno < lo or s.lineno > hi:
                continue
            if isinstance(s, ast.Assign) and base(s.targets[0]) in names:
                picked.append(s)
            elif isinstance(s, ast.AugAssign) and base(s.target) in names:
                picked.append(s)
            elif isinstance(s, ast.If) and any(isinstance(n, ast.Name) and n.id in if_tests for n in ast.walk(s.test)):
                picked.append(s)
            elif isinstance(s, (ast.For, ast.With, ast.If)):
                visit(s.body)

    visit(main.body)
    picked.sort(key=lambda s: s.lineno)
    return "\n".join(textwrap.dedent(" " * s.col_offset + ast.get_source_segment(src, s)) for s in picked), path


def reference_chunk_plan(total_samples, sample_rate):
    """inference_pipeline.py:218-225 and :295-319 (chunk / latent ranges), the reference's own statements."""
    import numpy as np
    head, path = _main_statements({"chunk_seconds", "overlap_seconds", "chunk_samples", "overlap_samples", "hop_samples",
                                   "num_chunks"}, "before")
    body, _ = _main_statements({"start_sample", "end_sample", "start_sec", "end_sec", "start_idx_16k", "end_idx_16k", "start_lat",
                                "end_lat"}, "inside")

    class _C:
        pass
    cfg = _C()
    cfg.sample_rate = sample_rate
    env = dict(np=np, cfg=cfg, total_samples=total_samples)
    exec(compile(head, path, "exec"), env)
    plan = []
    for i in range(env["num_chunks"]):
        env["i"] = i
        exec(compile(body, path, "exec"), env)
        plan.append((env["start_sample"], env["end_sample"], env["start_lat"], env["end_lat"]))
    return plan


def reference_stitch(ref, chunks, stats, chunk_frames, overlap_frames, total_frames):
    """The overlap-add of inference_pipeline.py:228-230, 239, 255-262, 359-393 run on given refined chunks ([1,T,n_mels]
    each) and per-chunk (mean, std): returns (window_mask, final_mel [n_mels,total], lin_mel_smoothed [1,n_mels,total])."""
    import torch.nn.functional as F
    from edge_diffusion_tts.utils.audio import denormalize_mel
    init, path = _main_statements({"estimated_frames", "final_mel", "final_weights", "hop_frames", "window_mask", "fade_len",
                                   "fade_in", "fade_out"}, "before")
    loop, _ = _main_statements({"mel_denorm", "lin_mel", "output_chunk", "start_frame", "end_frame", "final_mel", "final_weights"},
                               "inside", if_tests={"output_chunk"})
    post, _ = _main_statements({"final_weights", "final_mel", "lin_mel", "lin_mel_2d", "kernel_h", "kernel_w", "lin_mel_smoothed"},
                               "after")
    env = dict(torch=torch, F=F, cfg=ref["cfg"], device="cpu", denormalize_mel=denormalize_mel, total_frames=total_frames,
               chunk_frames=chunk_frames, overlap_frames=overlap_frames)
    exec(compile(init, path, "exec"), env)
    for i, (x, (m, s)) in enumerate(zip(chunks, stats)):
        env.update(i=i, x_refined=x, real_mean=m, real_std=s)
        exec(compile(loop, path, "exec"), env)
    exec(compile(post, path, "exec"), env)
    return env["window_mask"], env["final_mel"], env["lin_mel_smoothed"]
