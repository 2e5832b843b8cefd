"""CPU: host-side logic and the C-ABI surface (no kernel is launched)."""
import ctypes
import os
import re

import pytest
import torch

from oracle import edtts_oracle as O
from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "edtts.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(edtts_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    from edge_diffusion_tts_b200 import _lib
    names = _header_functions()
    assert len(names) >= 20
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and include/edtts.h drifted apart"
    for n in names:
        assert hasattr(lib, n), n
    assert lib.edtts_version() >= 100
    assert lib.edtts_packed_bf16_bytes() >= 0
    assert lib.edtts_vq_workspace_bytes(512, 128) >= 2048
    assert lib.edtts_decoder_workspace_bytes(2, 100, 50, 0) >= 2 * 100 * 800 * 4
    assert lib.edtts_context_workspace_bytes(2, 50) >= 2 * 50 * 240 * 4


def test_struct_layout_matches_header():
    from edge_diffusion_tts_b200 import _lib
    assert ctypes.sizeof(_lib.LayerWeights) == 19 * 8
    assert ctypes.sizeof(_lib.DecoderWeights) == 17 * 8 + 4 * 19 * 8 + 16 + 8 + 8      # ... packed_bf16, pos_pe_cm
    assert ctypes.sizeof(_lib.StepArgs) == 8 + 10 * 8 + 3 * 8 + 8


def test_decoder_state_dict_is_the_reference_layout():
    import edge_diffusion_tts_b200 as E
    dec = E.EdgeDiffusionDecoder(E.CFG())
    sd = dec.state_dict()
    shapes = synth.decoder_shapes()
    assert list(sd.keys()) == list(shapes.keys()) and len(sd) == 92
    assert all(tuple(sd[k].shape) == shapes[k] for k in shapes)
    assert sum(p.numel() for p in dec.parameters()) == 1_983_440
    # reference initialisation quirks (F6): zero-init projections => eps == 0 for a fresh decoder
    assert float(sd["out_proj.weight"].abs().sum()) == 0.0 and float(sd["layers.0.norm1.proj.weight"].abs().sum()) == 0.0
    assert torch.equal(sd["pos_emb.pe"], O.positional_table(1000))
    dec.load_state_dict(synth.synth_decoder_state(0), strict=True)
    vq = E.VectorQuantizer(128, 512)
    assert sorted(vq.state_dict()) == ["codebook.weight", "ema_cluster_size", "ema_w", "update_count"]
    conv = E.DepthwiseSeparableConv(16, 24, 5, 2)
    assert sorted(conv.state_dict()) == ["depthwise.weight", "norm.bias", "norm.weight", "pointwise.bias", "pointwise.weight"]
    enc = E.SemanticEncoder(E.CFG(use_fsq=False), load_hubert=False)
    assert [k for k in enc.state_dict() if k.startswith("proj.")] == ["proj.0.weight", "proj.0.bias", "proj.2.weight",
                                                                      "proj.2.bias", "proj.3.weight", "proj.3.bias"]


def test_schedule_host_side(golden):
    import edge_diffusion_tts_b200 as E
    g = golden("schedule")
    s = E.DiffusionSchedule(1000)
    for k, v in g["tables"].items():
        assert torch.equal(getattr(s, k), v), k
    assert s.get_schedule_for_steps(4) == g["steps4"]
    x, n = synth.synth_noise(14, 4, 10), synth.synth_noise(14, 4, 10, tag="noise")
    e = synth.synth_noise(14, 4, 10, tag="eps")
    assert torch.equal(s.q_sample(x, g["t"], n)[0], g["q_sample"])
    assert torch.equal(s.get_v_target(x, n, g["t"]), g["v_target"])
    assert torch.equal(s.predict_x0_from_eps(x, g["t"], e), g["x0_from_eps"])
    assert torch.equal(s.predict_x0_from_v(x, g["t"], e), g["x0_from_v"])
    assert torch.equal(s.predict_eps_from_v(x, g["t"], e), g["eps_from_v"])


def test_no_cpu_fallback(lib):
    """CPU tensors / missing GPU must raise, never silently compute elsewhere."""
    import edge_diffusion_tts_b200 as E
    s = E.DiffusionSchedule(1000)
    x = torch.zeros(1, 2, 80)
    with pytest.raises(RuntimeError):
        s.get_ddim_step(x, torch.tensor([5]), torch.tensor([0]), x)
    with pytest.raises(RuntimeError):
        E.VectorQuantizer(128, 512).eval().encode(torch.zeros(1, 2, 128))
    dec = E.EdgeDiffusionDecoder(E.CFG())
    with pytest.raises(RuntimeError):
        dec(torch.zeros(1, 4, 80), torch.zeros(1, dtype=torch.long), torch.zeros(1, 2, dtype=torch.long))
    with pytest.raises(ValueError):
        dec(torch.zeros(1, 4, 80), torch.zeros(1, dtype=torch.long))
    with pytest.raises(NotImplementedError):
        E.EdgeDiffusionDecoder(E.CFG(hidden=256))
    with pytest.raises(NotImplementedError):
        E.VectorQuantizer(128, 512).train()(torch.zeros(1, 2, 128))


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing shipped may import, include or execute it."""
    pkg = os.path.join(ROOT, "edge_diffusion_tts_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle|#\s*include.*oracle|oracle[/.]", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f


def test_timesteps_and_cfg():
    import edge_diffusion_tts_b200 as E
    cfg = E.CFG()
    inf = E.EdgeInference(cfg, E.DiffusionSchedule(1000), torch.nn.Identity(), E.EdgeDiffusionDecoder(cfg))
    assert inf._timesteps(4) == O.ddim_timesteps(4) and inf._timesteps(1) == [(999, 0)]
    assert inf._timesteps(16) == O.ddim_timesteps(16)
    assert E.CFG.from_dict(cfg.to_dict()) == cfg


def test_precision_names_and_workspace_sizes(lib):
    """precision="fp32" (the class default) selects the tf32 x 3 tensor-core path, "fp32_simt" the CUDA-core checker, "bf16" the fused
    kernel; anything else raises.  Workspace sizes are pure host arithmetic (no GPU needed): the tensor-core fp32 path adds its weight
    images and row statistics to the CUDA-core path's buffers; the context workspace covers its scratch for every precision."""
    import pytest
    import edge_diffusion_tts_b200 as E
    from edge_diffusion_tts_b200 import _lib
    dec = E.EdgeDiffusionDecoder(E.CFG(device="cpu"))
    assert dec.precision == "fp32"
    want = {"fp32": _lib.PREC_TF32X3, "fp32_simt": _lib.PREC_FP32, "bf16": _lib.PREC_BF16}
    for name, code in want.items():
        dec.precision = name
        assert dec._prec() == code
    dec.precision = "fp16"
    with pytest.raises(ValueError):
        dec._prec()
    B, T, S = 3, 200, 100
    simt = lib.edtts_decoder_workspace_bytes(B, T, S, _lib.PREC_FP32)
    t3 = lib.edtts_decoder_workspace_bytes(B, T, S, _lib.PREC_TF32X3)
    assert simt == 2 * B * T * 160 * 4 + B * T * 480 * 4
    assert t3 >= simt + B * T * 8 + 4 * 2_400_000          # + row statistics + ~10 MB of hi | lo weight images
    assert lib.edtts_context_workspace_bytes(B, S) >= B * S * (160 + 80 + 320) * 4 + B * S * 8
    rows_bytes = 4 * B * S * 320 * 4
    assert lib.edtts_context_kv_bytes(B, S, _lib.PREC_FP32) == rows_bytes == lib.edtts_context_kv_bytes(B, S, _lib.PREC_BF16)
    # + per layer, utterance, head and 32-key block the hi | lo operand images of K (10,240 B) and V^T (12,288 B)
    assert lib.edtts_context_kv_bytes(B, S, _lib.PREC_TF32X3) >= rows_bytes + 4 * B * 4 * ((S + 31) // 32) * 22528
    # the allocator of the Python class follows it: the reference's [layers, B*S, 320] rows, or a flat buffer with the images behind
    dec.precision = "bf16"
    assert tuple(dec.alloc_kv(B, S, "cpu").shape) == (4, B * S, 320)
    dec.precision = "fp32"
    kv = dec.alloc_kv(B, S, "cpu")
    assert kv.dim() == 1 and kv.numel() * 4 == lib.edtts_context_kv_bytes(B, S, _lib.PREC_TF32X3) and kv.dtype == torch.float32
    # one N-block image of a [160 -> 160] matrix: 5 chunks x (hi | lo) x 8 slabs x 160 rows x 16 B, + the statistics of `rows` rows
    assert lib.edtts_test_gemm_workspace_bytes(1000, 160, 160, 0) == 5 * 2 * 8 * 160 * 16 + 1000 * 8
