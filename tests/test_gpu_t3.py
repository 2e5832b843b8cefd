"""GPU parity tests of the fp32-grade tensor-core path (precision="fp32" -> EDTTS_PREC_TF32X3: tf32 x 3 GEMMs and attention,
csrc/t3_decoder.cu).  Single kernels against fp64 torch and against the CUDA-core kernels of the same library
(precision="fp32_simt", themselves pinned to the oracle in test_gpu_parity.py); the decoder against the oracle with the
north_star bar for fp32: max-abs <= 1e-4 on teacher-forced eps.  Tolerance of the kernel-level checks: attention 2e-5 max-abs
(the bar the CUDA-core kernel is held to); GEMMs 4e-6 of the largest output magnitude (a tf32 x 3 product carries ~2^-22
relative per operand, an FFMA 2^-24: measured 1e-6 of the output scale, the CUDA cores 3e-7)."""
import pytest
import torch

from oracle import edtts_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PRO_NONE, PRO_RMS, PRO_ADARMS, PRO_LN = 0, 1, 2, 3
EPI_STORE, EPI_GELU, EPI_RESID, EPI_PE, EPI_SWIGLU = 0, 1, 2, 3, 4


def _gemm(lib, x, w, b, N, pro, epi, norm_w, norm_b, eps, mod, rpb, resid, pe, period, use_tc):
    from edge_diffusion_tts_b200 import _lib
    rows, K = x.shape
    y = torch.full((rows, N), float("nan"), device=DEV)
    nbytes = int(lib.edtts_test_gemm_workspace_bytes(rows, K, N, epi))
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=DEV)
    p = lambda t: None if t is None else t.data_ptr()
    _lib.check(lib.edtts_test_gemm(p(x), p(w), p(b), p(y), rows, K, N, pro, epi, p(norm_w), p(norm_b), eps, p(mod), rpb, p(resid),
                                   p(pe), period, use_tc, p(ws), nbytes, _lib.stream_ptr(DEV)), "test_gemm")
    torch.cuda.synchronize()
    return y


def _ref_gemm(x, w, b, N, pro, epi, norm_w, norm_b, eps, mod, rpb, resid, pe, period):
    x, w = x.double().cpu(), w.double().cpu()
    rows, K = x.shape
    if pro in (PRO_RMS, PRO_ADARMS):
        x = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * norm_w.double().cpu()
        if pro == PRO_ADARMS:
            m = mod.double().cpu()[torch.arange(rows) // rpb]
            x = x * (1 + m[:, :K]) + m[:, K:]
    elif pro == PRO_LN:
        x = torch.nn.functional.layer_norm(x, (K,), norm_w.double().cpu(), norm_b.double().cpu(), eps)
    y = x @ w.t()
    if b is not None:
        y = y + b.double().cpu()
    if epi == EPI_GELU:
        y = torch.nn.functional.gelu(y)
    elif epi == EPI_RESID:
        y = resid.double().cpu() + y
    elif epi == EPI_PE:
        y = y + pe.double().cpu()[torch.arange(rows) % period]
    elif epi == EPI_SWIGLU:
        y = y[:, :N] * torch.nn.functional.silu(y[:, N:])
    return y


# every (K, N, prologue, epilogue) the decoder step launches (t3_decoder.cu), on row counts that are not tile multiples and
# utterance lengths that make a 128-row tile span several utterances
CASES = [
    ("in_proj", 80, 160, PRO_NONE, EPI_PE),
    ("qkv", 160, 480, PRO_ADARMS, EPI_STORE),
    ("proj", 160, 160, PRO_NONE, EPI_RESID),
    ("q_proj", 160, 160, PRO_RMS, EPI_STORE),
    ("ffn0", 160, 320, PRO_ADARMS, EPI_SWIGLU),
    ("ffn3", 320, 160, PRO_NONE, EPI_RESID),
    ("out_proj_ln", 160, 80, PRO_LN, EPI_STORE),
    ("gelu", 160, 160, PRO_NONE, EPI_GELU),
]


@pytest.mark.parametrize("name,K,N,pro,epi", CASES)
@pytest.mark.parametrize("rows,rpb", [(1000, 50), (128, 128), (77, 77), (2600, 650)])
def test_t3_gemm(lib, name, K, N, pro, epi, rows, rpb):
    g = torch.Generator().manual_seed(rows * 7 + K + N)
    R = lambda *s: torch.randn(*s, generator=g)
    wrows = 2 * N if epi == EPI_SWIGLU else N
    x = (R(rows, K) * 1.5 + 0.2).to(DEV)
    w = (R(wrows, K) * K ** -0.5).to(DEV)
    b = R(wrows).to(DEV) if name not in ("qkv", "q_proj") else None
    norm_w = (1 + 0.3 * R(K)).to(DEV)
    norm_b = (0.2 * R(K)).to(DEV)
    nb = (rows + rpb - 1) // rpb
    mod = (0.5 * R(nb, 2 * K)).to(DEV)
    resid = R(rows, N).to(DEV)
    pe = R(rpb, N).to(DEV)
    eps = 1e-5 if pro == PRO_LN else 1e-6
    args = (x, w, b, N, pro, epi, norm_w, norm_b, eps, mod, rpb, resid, pe, rpb)
    y_tc = _gemm(lib, *args, use_tc=1)
    y_cc = _gemm(lib, *args, use_tc=0)
    ref = _ref_gemm(*args)
    assert not torch.isnan(y_tc).any()
    e_tc = (y_tc.cpu().double() - ref).abs().max().item()
    e_cc = (y_cc.cpu().double() - ref).abs().max().item()
    print(f"{name} rows={rows}: tf32x3 max|d| {e_tc:.2e}, CUDA cores {e_cc:.2e}")
    assert e_tc < 4e-6 * max(1.0, ref.abs().max().item()), (name, e_tc, e_cc, ref.abs().max().item())


def test_t3_gemm_batch_invariance(lib):
    """A row's bits do not depend on where it sits in a launch (other rows, tile position)."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(700, 160, generator=g).to(DEV)
    w = (torch.randn(160, 160, generator=g) * 160 ** -0.5).to(DEV)
    nw = (1 + 0.3 * torch.randn(160, generator=g)).to(DEV)
    full = _gemm(lib, x, w, None, 160, PRO_RMS, EPI_STORE, nw, None, 1e-6, None, 1, None, None, 1, use_tc=1)
    part = _gemm(lib, x[333:500].contiguous(), w, None, 160, PRO_RMS, EPI_STORE, nw, None, 1e-6, None, 1, None, None, 1, use_tc=1)
    assert torch.equal(full[333:500], part)


@pytest.mark.parametrize("B,Tq,Tk,window", [(2, 200, 200, 64), (1, 64, 64, 64), (1, 333, 333, 64), (2, 150, 75, -1),
                                            (1, 800, 400, -1), (1, 1, 1, 64), (3, 129, 33, -1), (1, 1000, 1000, 64)])
def test_t3_attention(lib, B, Tq, Tk, window):
    from edge_diffusion_tts_b200 import _lib
    g = torch.Generator().manual_seed(Tq + Tk)
    q = torch.randn(B, Tq, 160, generator=g)
    kv = torch.randn(B, Tk, 320, generator=g)
    qd, kvd = q.to(DEV), kv.to(DEV)
    outs = []
    for prec in (2, 0):
        o = torch.full((B, Tq, 160), float("nan"), device=DEV)
        _lib.check(lib.edtts_test_attention(qd.data_ptr(), 160, kvd.data_ptr(), kvd.data_ptr() + 160 * 4, 320,
                                            o.data_ptr(), B, Tq, Tk, window, prec, _lib.stream_ptr(DEV)))
        torch.cuda.synchronize()
        outs.append(o.cpu().double())
    qh = q.view(B, Tq, 4, 40).transpose(1, 2).double()
    kh = kv[..., :160].reshape(B, Tk, 4, 40).transpose(1, 2).double()
    vh = kv[..., 160:].reshape(B, Tk, 4, 40).transpose(1, 2).double()
    mask = O.band_mask(Tq, window, "cpu") if window >= 0 else None
    ref = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh, attn_mask=mask)
    ref = ref.transpose(1, 2).reshape(B, Tq, 160)
    e_tc, e_cc = (outs[0] - ref).abs().max().item(), (outs[1] - ref).abs().max().item()
    print(f"attention B={B} Tq={Tq} Tk={Tk} w={window}: tf32x3 max|d| {e_tc:.2e}, CUDA cores {e_cc:.2e}")
    assert not torch.isnan(outs[0]).any()
    assert e_tc < 2e-5, (e_tc, e_cc)


@pytest.fixture(scope="module")
def model(lib):
    import edge_diffusion_tts_b200 as E
    cfg = E.CFG(device=DEV)
    sd = synth.synth_decoder_state(0)
    dec = E.EdgeDiffusionDecoder(cfg).to(DEV).eval()
    dec.load_state_dict(sd, strict=True)
    return dict(cfg=cfg, sd=sd, dec=dec)


@pytest.mark.parametrize("B,S", [(3, 40), (2, 203), (1, 500)])
def test_t3_decoder_vs_oracle_and_cuda_cores(model, B, S):
    """Teacher-forced eps of one decoder evaluation: tf32 x 3 path vs the oracle (north_star fp32 bar, 1e-4) and vs the
    CUDA-core path of the same library."""
    dec = model["dec"]
    T = 2 * S
    idx = synth.synth_sem_idx(11, B, S)
    x = synth.synth_noise(11, B, T)
    t = torch.tensor([999, 500, 3][:B])
    si = torch.tensor([0, 1, 3][:B])
    ref = O.decoder_forward(model["sd"], x, t, idx, si)
    keep = dec.precision
    try:
        out = {}
        for prec in ("fp32", "fp32_simt"):
            dec.precision = prec
            out[prec] = dec(x.to(DEV), t.to(DEV), idx.to(DEV), si.to(DEV)).cpu()
    finally:
        dec.precision = keep
    d = (out["fp32"] - out["fp32_simt"]).abs().max().item()
    print(f"decoder B={B} S={S}: tf32x3 vs CUDA cores max|d| {d:.2e}")
    assert d <= 2e-5
    e = (out["fp32"] - ref).abs().max().item()
    print(f"decoder B={B} S={S}: tf32x3 vs oracle max|d| {e:.2e}")
    assert e <= 1e-4
