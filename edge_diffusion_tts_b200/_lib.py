"""ctypes binding of libedtts.so (C ABI declared in include/edtts.h).

There is no CPU fallback: if the shared library is missing or a tensor is not
on a CUDA device the call raises.  Build the library with
``python -c "import __graft_entry__ as g; g.build()"`` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# EDTTS_LIB: development override to A/B a differently compiled build of the same library (tools/build_variant.sh)
LIB_PATH = os.environ.get("EDTTS_LIB") or os.path.join(_HERE, "lib", "libedtts.so")

PREC_FP32, PREC_BF16, PREC_TF32X3 = 0, 1, 2
STEP_EPS, STEP_DDIM, STEP_DDPM, STEP_DPM = 0, 1, 2, 3
N_LAYERS = 4

_f = C.c_void_p  # device pointers travel as void*


class LayerWeights(C.Structure):
    _fields_ = [(n, _f) for n in (
        "norm1_norm_w", "norm1_proj_w", "norm1_proj_b", "attn_qkv_w", "attn_proj_w", "attn_proj_b", "norm2_w",
        "q_proj_w", "kv_down_w", "kv_norm_w", "kv_up_w", "cross_out_w", "norm3_norm_w", "norm3_proj_w",
        "norm3_proj_b", "ffn0_w", "ffn0_b", "ffn3_w", "ffn3_b")]


class DecoderWeights(C.Structure):
    _fields_ = [(n, _f) for n in (
        "token_emb", "sem_proj_w", "sem_proj_b", "time1_w", "time1_b", "time3_w", "time3_b", "step_emb",
        "in_proj_w", "in_proj_b", "pos_pe", "ctx_pe", "time_freqs", "final_norm_w", "final_norm_b", "out_proj_w",
        "out_proj_b")] + [("layers", LayerWeights * N_LAYERS), ("codebook_size", C.c_int32),
                          ("pos_rows", C.c_int32), ("ctx_rows", C.c_int32), ("reserved", C.c_int32),
                          ("packed_bf16", _f), ("pos_pe_cm", _f)]


class StepArgs(C.Structure):
    _fields_ = [("mode", C.c_int32), ("write_x_prev", C.c_int32), ("t", _f), ("t_prev", _f), ("alpha_bar", _f),
                ("alphas", _f), ("betas", _f), ("posterior_var", _f), ("noise", _f), ("eps_out", _f),
                ("x_prev_out", _f), ("x0_out", _f), ("dpm_coef", _f), ("dpm_hist1", _f), ("dpm_hist2", _f),
                ("dpm_order", C.c_int32), ("dpm_predict_x0", C.c_int32)]


# name -> (restype, argtypes); must list every symbol of include/edtts.h
_i32, _i64, _p = C.c_int32, C.c_int64, C.c_void_p
SIGNATURES = {
    "edtts_version": (C.c_int, []),
    "edtts_last_error": (C.c_char_p, []),
    "edtts_device_supported": (C.c_int, []),
    "edtts_kernel_classes": (C.c_int, []),
    "edtts_kernel_class_name": (C.c_char_p, [C.c_int]),
    "edtts_launch_counts": (C.c_int, [_p, C.c_int]),
    "edtts_prof_enable": (C.c_int, [C.c_int]),
    "edtts_prof_collect": (C.c_int, [_p, _p, C.c_int]),
    "edtts_vq_argmin": (C.c_int, [_p, _p, _p, _i64, _i32, _i32, _p, _p]),
    "edtts_vq_workspace_bytes": (_i64, [_i32, _i32]),
    "edtts_encoder_proj_workspace_bytes": (_i64, [_i64, _i32]),
    "edtts_encoder_proj_image_bytes": (_i64, [_i32]),
    "edtts_encoder_proj_pack": (C.c_int, [_p, _p, _i32, _p, _p]),
    "edtts_encoder_proj_packed": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i32, _p]),
    "edtts_vq_pack": (C.c_int, [_p, _i32, _i32, _p, _p]),
    "edtts_vq_argmin_packed": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _i32, _p]),
    "edtts_vq_gather_ste": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _i32, _p]),
    "edtts_vq_bincount": (C.c_int, [_p, _p, _i64, _i32, _p]),
    "edtts_encoder_proj": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i32, _p]),
    "edtts_cond_prepare": (C.c_int, [C.POINTER(DecoderWeights), _p, _p, _p, _p, _i32, _p]),
    "edtts_context_prepare": (C.c_int, [C.POINTER(DecoderWeights), _p, _p, _p, _p, _i64, _i32, _i32, _i32, _p]),
    "edtts_context_workspace_bytes": (_i64, [_i32, _i32]),
    "edtts_context_kv_bytes": (_i64, [_i32, _i32, _i32]),
    "edtts_decoder_step": (C.c_int, [C.POINTER(DecoderWeights), _p, _p, _p, C.POINTER(StepArgs), _p, _i64, _i32,
                                      _i32, _i32, _i32, _p]),
    "edtts_decoder_workspace_bytes": (_i64, [_i32, _i32, _i32, _i32]),
    "edtts_packed_bf16_bytes": (_i64, []),
    "edtts_pack_weights_bf16": (C.c_int, [C.POINTER(DecoderWeights), _p, _p]),
    "edtts_ddim_step": (C.c_int, [_p, _p, _p, _p, _p, _p, C.c_float, _p, _p, _i32, _i64, _p]),
    "edtts_ddpm_step": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i32, _i64, _p]),
    "edtts_dpm_step": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _p, _p, _i32, _i64, _p]),
    "edtts_vddim_step": (C.c_int, [_p, _p, _p, C.c_float, _p, _p, _p, _i32, _i64, _p]),
    "edtts_inpaint_inject": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _i32, _i32, _p]),
    "edtts_normalize_mel": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _i32, _p]),
    "edtts_denormalize_mel": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _i32, _p]),
    "edtts_stitch_add": (C.c_int, [_p, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i64, _i64, _p]),
    "edtts_stitch_finalize": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _i64, _i64, _i32, _i32, _p]),
    "edtts_inverse_mel": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _i64, _p]),
    "edtts_fsq_forward": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _i64, _p]),
    "edtts_fsq_decode": (C.c_int, [_p, _p, _i32, _p, _i64, _p]),
    "edtts_fsq_encoder": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _i32, _i32, _p, _p, _i64, _p]),
    "edtts_griffinlim_workspace_bytes": (_i64, [_i32, _i32, _i32]),
    "edtts_griffinlim": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _i32, _i32, C.c_float, C.c_float, _i64, _p]),
    "edtts_dsconv_forward": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "edtts_dsconv_workspace_bytes": (_i64, [_i32, _i32, _i32, _i32]),
    "edtts_test_linear": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _i32, _i32, _p]),
    "edtts_test_attention": (C.c_int, [_p, _i32, _p, _p, _i32, _p, _i32, _i32, _i32, _i32, _i32, _p]),
    "edtts_test_gemm": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _i32, _i32, _i32, _p, _p, C.c_float, _p, _i32, _p, _p, _i32, _i32,
                                  _p, _i64, _p]),
    "edtts_test_gemm_workspace_bytes": (_i64, [_i64, _i32, _i32, _i32]),
    "edtts_test_hidden": (C.c_int, [C.POINTER(DecoderWeights), _p, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _i32, _i32,
                                    _i32, _p]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load libedtts.so and bind every symbol; raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library drift apart
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().edtts_last_error().decode(errors="replace")
        kind = {-1: ValueError, -4: NotImplementedError}.get(rc, RuntimeError)
        raise kind(f"libedtts {what} failed ({rc}): {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("edge_diffusion_tts_b200 runs on CUDA (B200) tensors only; got a CPU tensor "
                           "(there is no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("tensor must be contiguous")
    return t.data_ptr()


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"expected float32, got {t.dtype} (the path is fp32 at the API, SURVEY F11)")
    return t.contiguous()


def i64(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.int64:
        t = t.long()
    return t.contiguous()


def launch_counts() -> dict:
    """Kernel launches made through libedtts so far, by kernel class (includes launches recorded
    into CUDA graphs at capture time, not their replays)."""
    lib = load()
    n = lib.edtts_kernel_classes()
    arr = (C.c_uint64 * n)()
    check(lib.edtts_launch_counts(arr, n), "launch_counts")
    return {lib.edtts_kernel_class_name(i).decode(): int(arr[i]) for i in range(n)}


def prof_enable(on: bool) -> None:
    load().edtts_prof_enable(1 if on else 0)


def prof_collect() -> dict:
    """{class: (total_ms, launches)} of the eager launches since the last collect (synchronises)."""
    lib = load()
    n = lib.edtts_kernel_classes()
    ms, cnt = (C.c_double * n)(), (C.c_uint64 * n)()
    check(lib.edtts_prof_collect(ms, cnt, n), "prof_collect")
    return {lib.edtts_kernel_class_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n) if cnt[i]}


class Workspace:
    """Grow-only device scratch owned by the caller side of the ABI."""

    def __init__(self):
        self._buf = None

    def get(self, nbytes: int, device) -> torch.Tensor:
        if self._buf is None or self._buf.numel() < nbytes or self._buf.device != torch.device(device):
            self._buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        return self._buf
