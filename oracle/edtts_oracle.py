"""CPU oracle for the edge-diffusion-tts few-step sampling path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``edge_diffusion_tts_b200/`` may import
this file; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker /
reported baseline -- never as the thing shipped.

It is a *functional* restatement (state-dict in, tensors out) of the reference's
PyTorch modules, written from SURVEY.md appendix A; every function cites the
reference ``file:line`` it follows (paths relative to
``/root/reference/edge_diffusion_tts``).  It uses the same ATen ops the
reference uses (``F.linear``, ``F.scaled_dot_product_attention`` with the dense
band mask, ``F.layer_norm`` ...) so that timing it on host cores is
representative of the reference's own CPU path.

Parity pinning: the reference ships no tests or golden vectors for this path
("parity unpinned" by the reference itself).  The oracle is
therefore pinned against the *reference run in the build container*:
``oracle/make_golden.py`` imports the unmodified reference, loads the synthetic
weights of ``oracle/synth.py`` into its modules and records outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement
against those fixtures everywhere, and ``tests/test_oracle_vs_reference.py``
checks it against the live reference wherever ``/root/reference`` exists.

All functions run in the dtype of the tensors they are given, so passing an
fp64 state dict yields the fp64 restatement used to measure the fp32 noise
floor.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
State = Dict[str, Tensor]

# Hyper-parameters of the default CFG (config.py:97-111) that the path reads.
N_MELS = 80
HIDDEN = 160
LAYERS = 4
HEADS = 4
HEAD_DIM = 40
KV_RANK = 80
FFN_HIDDEN = 320
WINDOW = 64
SEMANTIC_DIM = 128
CODEBOOK = 512
DIFF_STEPS = 1000


# ----------------------------------------------------------------------------
# schedule.py
# ----------------------------------------------------------------------------
def cosine_schedule(T: int = DIFF_STEPS, device: str = "cpu") -> State:
    """Cosine schedule tables, fp32 (schedule.py:36-59).  beta_start/beta_end
    are ignored by the reference, so they are not parameters."""
    s = 0.008
    x = torch.linspace(0, T, T + 1, device=device)
    ac = torch.cos(((x / T) + s) / (1 + s) * torch.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = torch.clip(1 - (ac[1:] / ac[:-1]), 0.0001, 0.9999)
    alphas = 1.0 - betas
    alpha_bar = torch.cumprod(alphas, dim=0)
    alpha_bar_prev = F.pad(alpha_bar[:-1], (1, 0), value=1.0)
    sqrt_ab = torch.sqrt(alpha_bar)
    sqrt_1mab = torch.sqrt(1.0 - alpha_bar)
    return {
        "betas": betas,
        "alphas": alphas,
        "alpha_bar": alpha_bar,
        "sqrt_alpha_bar": sqrt_ab,
        "sqrt_one_minus_alpha_bar": sqrt_1mab,
        "sqrt_recip_alpha_bar": torch.sqrt(1.0 / alpha_bar),
        "sqrt_recip_alpha_bar_minus_one": torch.sqrt(1.0 / alpha_bar - 1),
        "posterior_variance": betas * (1.0 - alpha_bar_prev) / (1.0 - alpha_bar),
        "lambda_t": torch.log(sqrt_ab / sqrt_1mab),
    }


def ddim_step(tab: State, x_t: Tensor, t: Tensor, t_prev: Tensor, eps: Tensor,
              eta: float = 0.0, noise: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """DDIM update (schedule.py:179-202).  Returns (x_prev, x0_pred)."""
    ab = tab["alpha_bar"].to(x_t.dtype)
    ab_t = ab[t][:, None, None]
    ab_p = torch.where(t_prev[:, None, None] >= 0, ab[t_prev.clamp(min=0)][:, None, None],
                       torch.ones_like(ab_t))
    x0 = (x_t - torch.sqrt(1 - ab_t) * eps) / torch.sqrt(ab_t)
    x0 = torch.clamp(x0, -3, 3)
    sigma = eta * torch.sqrt((1 - ab_p) / (1 - ab_t) * (1 - ab_t / ab_p))
    dir_xt = torch.sqrt(1 - ab_p - sigma ** 2) * eps
    if eta > 0:
        if noise is None:
            noise = torch.randn_like(x_t)
    else:
        noise = 0
    x_prev = torch.sqrt(ab_p) * x0 + dir_xt + sigma * noise
    return x_prev, x0


def ddpm_step(tab: State, x_t: Tensor, t: Tensor, eps: Tensor,
              noise: Optional[Tensor] = None) -> Tensor:
    """Ancestral DDPM update (schedule.py:221-238).  ``noise`` may be injected
    (the reference draws ``torch.randn_like(x_t)`` at :232)."""
    dt = x_t.dtype
    alpha = tab["alphas"].to(dt)[t][:, None, None]
    alpha_bar = tab["alpha_bar"].to(dt)[t][:, None, None]
    beta = tab["betas"].to(dt)[t][:, None, None]
    coef1 = 1.0 / torch.sqrt(alpha)
    coef2 = beta / torch.sqrt(1.0 - alpha_bar)
    mean = coef1 * (x_t - coef2 * eps)
    var = tab["posterior_variance"].to(dt)[t][:, None, None]
    if noise is None:
        noise = torch.randn_like(x_t)
    nonzero = (t > 0).to(dt)[:, None, None]
    return mean + nonzero * torch.sqrt(var) * noise


# ----------------------------------------------------------------------------
# layers/embeddings.py
# ----------------------------------------------------------------------------
def time_freqs(dim: int = HIDDEN) -> Tensor:
    """fp32 frequency vector of SinusoidalTimeEmb (embeddings.py:37-41)."""
    half = dim // 2
    return torch.exp(torch.arange(half, dtype=torch.float32) * (-math.log(10000.0) / (half - 1)))


def time_embedding(t: Tensor, dim: int = HIDDEN, dtype=torch.float32) -> Tensor:
    """SinusoidalTimeEmb.forward (embeddings.py:37-43): cat[sin, cos]."""
    args = t.to(dtype).unsqueeze(1) * time_freqs(dim).to(dtype).unsqueeze(0)
    return torch.cat([torch.sin(args), torch.cos(args)], dim=1)


def positional_table(max_len: int, dim: int = HIDDEN) -> Tensor:
    """SinusoidalPositionalEmb.__init__ (embeddings.py:122-128), fp32.
    Used both to check the persistent ``pe`` buffers and to extend them past
    1000 / 512 rows for BASELINE config 5 (a stated deviation)."""
    pe = torch.zeros(max_len, dim)
    position = torch.arange(0, max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, dim, 2) * (-math.log(10000.0) / dim))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def _pe_rows(sd: State, key: str, n: int, dtype) -> Tensor:
    pe = sd[key]
    if n > pe.shape[0]:  # F7 deviation: same closed form, more rows
        pe = positional_table(n, pe.shape[1]).to(pe.device)
    return pe[:n].to(dtype)


# ----------------------------------------------------------------------------
# layers/mla.py, layers/transformer.py, layers/attention.py
# ----------------------------------------------------------------------------
def rms_norm(x: Tensor, weight: Tensor, eps: float = 1e-6) -> Tensor:
    """RMSNorm.forward (mla.py:53-58); normalises in >= fp32."""
    xf = x.float() if x.dtype in (torch.float16, torch.bfloat16) else x
    out = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).type_as(x)
    return out * weight


def ada_rms_norm(x: Tensor, cond: Tensor, sd: State, prefix: str) -> Tensor:
    """AdaLayerNorm.forward (transformer.py:64-68)."""
    mod = F.linear(cond, sd[prefix + "proj.weight"], sd[prefix + "proj.bias"])
    scale, shift = mod.chunk(2, dim=-1)
    x = rms_norm(x, sd[prefix + "norm.weight"])
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def band_mask(T: int, window: int, device) -> Tensor:
    """create_local_attention_mask (attention.py:27-30): |i-j| <= window."""
    r = torch.arange(T, device=device)
    return ((r[None, :] - r[:, None]).abs() <= window)[None, None]


def self_attention(x: Tensor, sd: State, prefix: str, heads: int = HEADS,
                   window: Optional[int] = WINDOW) -> Tensor:
    """EfficientAttention.forward (attention.py:87-123)."""
    B, T, C = x.shape
    d = C // heads
    qkv = F.linear(x, sd[prefix + "qkv.weight"]).reshape(B, T, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    mask = band_mask(T, window, x.device) if window is not None else None
    out = F.scaled_dot_product_attention(q, k, v, attn_mask=mask)
    out = out.transpose(1, 2).reshape(B, T, C)
    return F.linear(out, sd[prefix + "proj.weight"], sd[prefix + "proj.bias"])


def cross_kv(context: Tensor, sd: State, prefix: str, heads: int = HEADS) -> Tuple[Tensor, Tensor]:
    """The step-invariant half of MultiHeadLatentAttention.forward (mla.py:144-153)."""
    B, S, C = context.shape
    d = C // heads
    c_kv = rms_norm(F.linear(context, sd[prefix + "kv_down_proj.weight"]), sd[prefix + "kv_norm.weight"])
    kv = F.linear(c_kv, sd[prefix + "kv_up_proj.weight"]).reshape(B, S, 2, heads, d).permute(2, 0, 3, 1, 4)
    return kv[0], kv[1]


def cross_attention(x: Tensor, context: Tensor, sd: State, prefix: str, heads: int = HEADS) -> Tensor:
    """MultiHeadLatentAttention.forward with context given, cond=None
    (mla.py:133-194; RoPE and mask branches are dead on this path, F3)."""
    B, T, C = x.shape
    d = C // heads
    q = F.linear(x, sd[prefix + "q_proj.weight"]).reshape(B, T, heads, d).transpose(1, 2)
    k, v = cross_kv(context, sd, prefix, heads)
    out = F.scaled_dot_product_attention(q, k, v)
    out = out.transpose(1, 2).reshape(B, T, C)
    return F.linear(out, sd[prefix + "out_proj.weight"])


def feed_forward(x: Tensor, sd: State, prefix: str) -> Tensor:
    """FeedForward / SwiGLU (transformer.py:13-49): first half * silu(second half)."""
    u = F.linear(x, sd[prefix + "net.0.weight"], sd[prefix + "net.0.bias"])
    a, gate = u.chunk(2, dim=-1)
    return F.linear(a * F.silu(gate), sd[prefix + "net.3.weight"], sd[prefix + "net.3.bias"])


def transformer_block(x: Tensor, context: Tensor, cond: Tensor, sd: State, prefix: str,
                      heads: int = HEADS, window: Optional[int] = WINDOW) -> Tensor:
    """DiffusionTransformerBlock.forward with use_adaln=True (transformer.py:141-160)."""
    x = x + self_attention(ada_rms_norm(x, cond, sd, prefix + "norm1."), sd, prefix + "attn.", heads, window)
    x = x + cross_attention(rms_norm(x, sd[prefix + "norm2.weight"]), context, sd, prefix + "cross_attn.", heads)
    x = x + feed_forward(ada_rms_norm(x, cond, sd, prefix + "norm3."), sd, prefix + "ffn.")
    return x


# ----------------------------------------------------------------------------
# models/decoder.py
# ----------------------------------------------------------------------------
def time_condition(sd: State, t: Tensor, step_idx: Optional[Tensor]) -> Tensor:
    """t_cond of EdgeDiffusionDecoder.forward (decoder.py:77-80)."""
    w = sd["time_emb.1.weight"]
    e = time_embedding(t, w.shape[1], w.dtype)
    c = F.linear(F.gelu(F.linear(e, w, sd["time_emb.1.bias"])), sd["time_emb.3.weight"], sd["time_emb.3.bias"])
    if step_idx is not None:
        c = c + sd["step_emb.weight"][step_idx]
    return c


def decoder_forward(sd: State, x_t: Tensor, t: Tensor, sem_idx: Optional[Tensor] = None,
                    step_idx: Optional[Tensor] = None, sem_features: Optional[Tensor] = None,
                    heads: int = HEADS, window: Optional[int] = WINDOW,
                    return_hidden: bool = False):
    """EdgeDiffusionDecoder.forward (decoder.py:66-109)."""
    cond = time_condition(sd, t, step_idx)
    if sem_features is not None:
        context = F.linear(sem_features, sd["sem_proj.weight"], sd["sem_proj.bias"])
    elif sem_idx is not None:
        context = sd["token_emb.weight"][sem_idx]
    else:
        raise ValueError("Either sem_idx or sem_features must be provided")
    dt = context.dtype
    context = context + _pe_rows(sd, "context_pos_emb.pe", context.shape[1], dt)
    h = F.linear(x_t, sd["in_proj.weight"], sd["in_proj.bias"])
    h = h + _pe_rows(sd, "pos_emb.pe", h.shape[1], dt)
    hidden = [h]
    n_layers = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("layers."))
    for i in range(n_layers):
        h = transformer_block(h, context, cond, sd, f"layers.{i}.", heads, window)
        hidden.append(h)
    h = F.layer_norm(h, (h.shape[-1],), sd["final_norm.weight"], sd["final_norm.bias"], 1e-5)
    eps = F.linear(h, sd["out_proj.weight"], sd["out_proj.bias"])
    return (eps, hidden) if return_hidden else eps


# ----------------------------------------------------------------------------
# models/vq.py, models/encoder.py
# ----------------------------------------------------------------------------
def vq_distances(codebook: Tensor, flat: Tensor) -> Tensor:
    """vq.py:75-79; note Python precedence: ``2 * flat @ W.t()`` == ``(2*flat) @ W.t()``."""
    return (flat.pow(2).sum(1, keepdim=True) - 2 * flat @ codebook.t()
            + codebook.pow(2).sum(1, keepdim=True).t())


def vq_encode(codebook: Tensor, z: Tensor) -> Tensor:
    """VectorQuantizer.encode (vq.py:148-159): first-minimum argmin, int64."""
    B, T, D = z.shape
    return vq_distances(codebook, z.reshape(-1, D)).argmin(dim=1).view(B, T)


def vq_forward(codebook: Tensor, z: Tensor):
    """VectorQuantizer.forward in eval mode (vq.py:53-107) -> 5-tuple."""
    B, T, D = z.shape
    idx = vq_distances(codebook, z.reshape(-1, D)).argmin(dim=1)
    z_q = codebook[idx].view(B, T, D)
    vq_loss = torch.tensor(0.0, device=z.device)
    z_q = z + (z_q - z)  # straight-through estimator rounding (vq.py:98)
    counts = torch.bincount(idx, minlength=codebook.shape[0]).float()
    probs = counts / counts.sum().clamp_min(1.0)
    perplexity = torch.exp(-(probs * torch.log(probs.clamp_min(1e-12))).sum())
    used = (counts > 0).sum()
    return z_q, idx.view(B, T), vq_loss, perplexity, used


def encoder_proj(sd: State, h: Tensor) -> Tensor:
    """SemanticEncoder.proj (encoder.py:41-46): Linear, GELU, LayerNorm, Linear."""
    z = F.gelu(F.linear(h, sd["0.weight"], sd["0.bias"]))
    z = F.layer_norm(z, (z.shape[-1],), sd["2.weight"], sd["2.bias"], 1e-5)
    return F.linear(z, sd["3.weight"], sd["3.bias"])


# ----------------------------------------------------------------------------
# layers/conv.py (operator-level only)
# ----------------------------------------------------------------------------
def dsconv_forward(sd: State, x: Tensor, stride: int = 1) -> Tensor:
    """DepthwiseSeparableConv.forward (conv.py:61-64), x: [B, C_in, T]."""
    wd = sd["depthwise.weight"]
    k = wd.shape[-1]
    y = F.conv1d(x, wd, None, stride=stride, padding=k // 2, groups=wd.shape[0])
    y = F.conv1d(y, sd["pointwise.weight"], sd["pointwise.bias"])
    groups = min(8, y.shape[1])
    y = F.group_norm(y, groups, sd["norm.weight"], sd["norm.bias"], 1e-5)
    return F.gelu(y)


# ----------------------------------------------------------------------------
# inference.py
# ----------------------------------------------------------------------------
def ddim_timesteps(num_steps: int, diff_steps: int = DIFF_STEPS):
    """inference.py:35-36,41: (t, t_prev) pairs of generate_mel."""
    stride = diff_steps // num_steps
    ts = list(range(diff_steps - 1, 0, -stride))[:num_steps]
    return [(t, max(t - stride, 0)) for t in ts]


def generate_mel(sd: State, tab: State, sem_idx: Tensor, num_steps: int = 4,
                 x_T: Optional[Tensor] = None, temperature: float = 1.0,
                 trace: Optional[list] = None) -> Tensor:
    """EdgeInference.generate_mel (inference.py:23-53).  ``x_T`` injects the
    initial noise (the reference draws randn(B, 2S, n_mels) * temperature at :33).
    ``trace`` (optional list) receives (x_t, eps, x_prev, x0) per step for
    teacher-forced parity."""
    B, S = sem_idx.shape
    if x_T is None:
        x_T = torch.randn(B, 2 * S, N_MELS) * temperature
    x = x_T
    x0 = None
    with torch.no_grad():
        for i, (t, t_prev) in enumerate(ddim_timesteps(num_steps, tab["alpha_bar"].shape[0])):
            tt = torch.full((B,), t, dtype=torch.long)
            si = torch.full((B,), i, dtype=torch.long)
            tp = torch.full((B,), t_prev, dtype=torch.long)
            eps = decoder_forward(sd, x, tt, sem_idx, si)
            x_in = x
            x, x0 = ddim_step(tab, x, tt, tp, eps, 0.0)
            if trace is not None:
                trace.append((x_in, eps, x, x0))
    return x0


def ddpm_loop(sd: State, tab: State, sem_idx: Tensor, x_T: Tensor, noises, t_start: Optional[int] = None,
              t_end: int = 0) -> Tensor:
    """Harness-driven ancestral loop for BASELINE config 4:
    ``eps = decoder(x, t, sem_idx, step_idx=None); x = ddpm_step(x, t, eps)`` for
    t = t_start .. t_end.  ``noises[i]`` is the pre-drawn N(0,1) tensor of step i."""
    B = sem_idx.shape[0]
    T = tab["alpha_bar"].shape[0]
    t_start = T - 1 if t_start is None else t_start
    x = x_T
    with torch.no_grad():
        for i, t in enumerate(range(t_start, t_end - 1, -1)):
            tt = torch.full((B,), t, dtype=torch.long)
            eps = decoder_forward(sd, x, tt, sem_idx, None)
            x = ddpm_step(tab, x, tt, eps, noises[i])
    return x


# ----------------------------------------------------------------------------
# schedule.py:269-531  DPMSolverPP (the sampler train_v2.py / inference_pipeline.py use)
# ----------------------------------------------------------------------------
def dpm_time_steps(tab: State, num_steps: int, max_t: Optional[int] = None) -> Tensor:
    """DPMSolverPP.get_time_steps (schedule.py:299-324): timesteps uniformly spaced in lambda = log(alpha/sigma)."""
    lam = tab["lambda_t"]
    max_t = max_t or (lam.shape[0] - 1)
    lambda_max = lam[1].item()
    lambda_min = lam[max_t].item()
    lambdas = torch.linspace(lambda_min, lambda_max, num_steps + 1)
    ts = []
    for l in lambdas[:-1]:
        t = (lam - l).abs().argmin().item()
        ts.append(max(1, min(t, max_t)))
    return torch.tensor(ts, dtype=torch.long)


def dpm_coefficients(tab: State, t: Tensor, t_prev: Tensor, t_prev2: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """The [B] scalars of the three update rules (schedule.py:339-438), each built with the reference's op order:
    c0 = sigma_prev / sigma_t, c1 = alpha_prev (1 - e^-h), c2 = alpha_prev ((1 - e^-h) / h + 1),
    c3 = alpha_prev ((1 - e^-h) / h^2 + 0.5 / h + 0.5), inv_r = 1 / (h_prev / h) (second order only)."""
    alpha_prev = tab["sqrt_alpha_bar"][t_prev]
    sigma_t = tab["sqrt_one_minus_alpha_bar"][t]
    sigma_prev = tab["sqrt_one_minus_alpha_bar"][t_prev]
    lambda_t = tab["lambda_t"][t]
    lambda_prev = tab["lambda_t"][t_prev]
    h = lambda_prev - lambda_t
    out = {
        "sa": tab["sqrt_alpha_bar"][t], "sb": tab["sqrt_one_minus_alpha_bar"][t],
        "c0": sigma_prev / sigma_t,
        "c1": alpha_prev * (1 - torch.exp(-h)),
        "c2": alpha_prev * ((1 - torch.exp(-h)) / h + 1),
        "c3": alpha_prev * ((1 - torch.exp(-h)) / (h ** 2) + 0.5 / h + 0.5),
    }
    if t_prev2 is not None:
        h_prev = tab["lambda_t"][t_prev2] - lambda_prev
        r = h_prev / h
        out["inv_r"] = 1 / r
    return out


def dpm_update(x: Tensor, x0_pred: Tensor, hist: list, co: Dict[str, Tensor], order_used: int) -> Tensor:
    """first / second / third_order_update (schedule.py:339-438) given the per-row coefficients.
    ``hist`` is x0_history as the reference keeps it (oldest first)."""
    b = lambda v: v[:, None, None]
    if order_used == 1:
        return b(co["c0"]) * x + b(co["c1"]) * x0_pred
    if order_used == 2:
        D1 = b(co["inv_r"]) * (x0_pred - hist[-1])
        return b(co["c0"]) * x + b(co["c1"]) * x0_pred + b(co["c2"]) * D1 * 0.5
    p = [x0_pred] + hist[-2:]                         # the reference's list order: [now, older, newer] (schedule.py:510)
    D1 = p[0] - p[1]
    D2 = p[0] - 2 * p[1] + p[2]
    return (b(co["c0"]) * x + b(co["c1"]) * p[0] + b(co["c2"]) * D1 * 0.5 + b(co["c3"]) * D2 / 6)


def dpm_sample(sd: State, tab: State, x_T: Tensor, sem_features: Tensor, num_steps: int = 10, order: int = 2,
               predict_x0: bool = False, max_t: Optional[int] = None, trace: Optional[list] = None) -> Tensor:
    """DPMSolverPP.sample (schedule.py:440-527) with the decoder's sem_features conditioning (decoder.py:83-85).
    ``trace`` receives (x_t, model_output, x0_pred, x_next, order_used) per step for teacher-forced parity."""
    max_t = max_t or 950
    ts = dpm_time_steps(tab, num_steps, max_t)
    x = x_T
    B = x.shape[0]
    x0_hist, t_hist = [], []
    with torch.no_grad():
        for i, t in enumerate(ts):
            tt = torch.full((B,), t.item(), dtype=torch.long)
            si = torch.full((B,), i, dtype=torch.long)
            out = decoder_forward(sd, x, tt, None, si, sem_features=sem_features)
            n = min(out.shape[1], x.shape[1])
            out, x = out[:, :n], x[:, :n]
            if predict_x0:
                x0 = out
            else:
                x0 = tab["sqrt_alpha_bar"][tt][:, None, None] * x - tab["sqrt_one_minus_alpha_bar"][tt][:, None, None] * out
            x0 = torch.clamp(x0, -3, 3)
            tp = torch.full((B,), ts[i + 1].item() if i < len(ts) - 1 else 0, dtype=torch.long)
            if order == 1 or len(x0_hist) == 0:
                used, co = 1, dpm_coefficients(tab, tt, tp)
            elif order == 2 or len(x0_hist) == 1:
                used, co = 2, dpm_coefficients(tab, tt, tp, t_hist[-1])
            else:
                used, co = 3, dpm_coefficients(tab, tt, tp)
            x_in = x
            x = dpm_update(x, x0, x0_hist, co, used)
            if trace is not None:
                trace.append((x_in, out, x0, x, used))
            x0_hist.append(x0)
            t_hist.append(tp)
            if len(x0_hist) > 2:
                x0_hist.pop(0)
                t_hist.pop(0)
    return x


# ----------------------------------------------------------------------------
# models/fsq.py:18-132  FSQ
# ----------------------------------------------------------------------------
def fsq_forward(levels, z: Tensor):
    """FSQ.forward (fsq.py:84-108) -> (z_q, indices, z_scaled): z_scaled = (tanh(z) + 1) * half is returned so that tests
    can exclude inputs that sit on a rounding boundary (tanh differs by ulps between CPU and CUDA)."""
    lv = torch.tensor(levels, dtype=torch.int32)
    basis = torch.cumprod(torch.tensor([1] + list(levels[:-1]), dtype=torch.int64), dim=0)
    half = (lv.float() - 1) / 2
    zb = torch.tanh(z)
    zs = (zb + 1) * half
    q = torch.minimum(torch.clamp(torch.round(zs), min=0), lv.float() - 1) / half - 1
    z_q = zb + (q - zb)
    idx = (((z_q + 1) * half).round().long() * basis).sum(dim=-1)
    return z_q, idx, zs


def fsq_indices_to_codes(levels, indices: Tensor) -> Tensor:
    """FSQ.indices_to_codes (fsq.py:121-132): last dimension fastest (not the inverse of the forward basis)."""
    lv = torch.tensor(levels, dtype=torch.int32)
    codes = []
    for i in range(len(levels) - 1, -1, -1):
        codes.append(indices % lv[i])
        indices = indices // lv[i]
    codes = torch.stack(codes[::-1], dim=-1)
    return codes.float() / ((lv.float() - 1) / 2) - 1


def fsq_encoder_forward(sd: State, levels, z: Tensor):
    """FSQEncoder.forward (fsq.py:161-198) -> (z_q, indices, loss, perplexity, used, z_scaled); sd holds proj_down.* /
    proj_up.* (nn.Linear).  z_scaled as in fsq_forward, for boundary exclusion in the GPU tests."""
    z_low = F.linear(z, sd["proj_down.weight"], sd["proj_down.bias"])
    z_q_low, idx, zs = fsq_forward(levels, z_low)
    z_q = F.linear(z_q_low, sd["proj_up.weight"], sd["proj_up.bias"])
    n = 1
    for l in levels:
        n *= l
    counts = torch.zeros(n, dtype=torch.float32).scatter_add_(0, idx.flatten(), torch.ones(idx.numel()))
    probs = counts / counts.sum().clamp_min(1.0)
    perplexity = torch.exp(-(probs * torch.log(probs.clamp_min(1e-12))).sum())
    return z_q, idx, torch.tensor(0.0), perplexity, (counts > 0).sum(), zs


def fsq_encoder_decode(sd: State, levels, indices: Tensor) -> Tensor:
    """FSQEncoder.decode (fsq.py:218-221)."""
    return F.linear(fsq_indices_to_codes(levels, indices), sd["proj_up.weight"], sd["proj_up.bias"])


# ----------------------------------------------------------------------------
# inference_pipeline.py:145-196  in-painting refine loop of the long-form pipeline
# ----------------------------------------------------------------------------
def inpaint_refine(sd: State, tab: State, x_coarse: Tensor, sem_features: Tensor, known_mel: Optional[Tensor] = None,
                   overlap_len: int = 0, strength: float = 0.2, steps: int = 10, cfg_scale: float = 1.0,
                   noise: Optional[Tensor] = None, known_noises=None, trace: Optional[list] = None) -> Tensor:
    """inpaint_teacher_refine (inference_pipeline.py:145-196): diffuse x_coarse to t_start = int(T * strength), then
    ``steps`` v-prediction DDIM steps on a linear time grid; before every step the first ``overlap_len`` frames are
    replaced by a freshly noised copy of ``known_mel`` (the previous chunk's tail); optional classifier-free guidance
    against zero conditioning.  ``noise`` / ``known_noises[i]`` inject the reference's randn_like draws."""
    B = x_coarse.shape[0]
    T = tab["alpha_bar"].shape[0]
    t_start = int(T * strength)
    z_null = torch.zeros_like(sem_features)
    if noise is None:
        noise = torch.randn_like(x_coarse)
    ts = torch.full((B,), t_start, dtype=torch.long)
    x = tab["sqrt_alpha_bar"][ts][:, None, None] * x_coarse + tab["sqrt_one_minus_alpha_bar"][ts][:, None, None] * noise
    times = torch.linspace(t_start, 0, steps + 1).long()[:-1]
    s_idx = torch.full((B,), 0, dtype=torch.long)
    with torch.no_grad():
        for i in range(len(times)):
            t_curr = times[i]
            t_next = times[i + 1] if i < len(times) - 1 else torch.tensor(0)
            tt = torch.full((B,), int(t_curr), dtype=torch.long)
            sa = tab["sqrt_alpha_bar"][tt][:, None, None]
            sb = tab["sqrt_one_minus_alpha_bar"][tt][:, None, None]
            if known_mel is not None:
                nk = known_noises[i] if known_noises is not None else torch.randn_like(known_mel)
                x[:, :overlap_len, :] = sa * known_mel + sb * nk
            v_cond = decoder_forward(sd, x, tt, None, s_idx, sem_features=sem_features)
            if cfg_scale != 1.0:
                v_unc = decoder_forward(sd, x, tt, None, s_idx, sem_features=z_null)
                v = v_unc + cfg_scale * (v_cond - v_unc)
            else:
                v = v_cond
            x0 = torch.clamp(sa * x - sb * v, -3, 3)
            eps = sb * x + sa * v
            a_next = tab["alpha_bar"][t_next]
            x_in = x
            x = torch.sqrt(a_next) * x0 + torch.sqrt(1 - a_next) * eps
            if trace is not None:
                trace.append((x_in.clone(), v_cond, x0, x.clone()))
    if known_mel is not None:
        x[:, :overlap_len, :] = known_mel
    return x


# ----------------------------------------------------------------------------
# utils/audio.py:10-19 and inference_pipeline.py:217-393  mel statistics, chunk plan, cross-fade stitch
# ----------------------------------------------------------------------------
def normalize_mel(mel: Tensor):
    """utils/audio.py:10-14: statistics over dim 1, unbiased std clamped to >= 1e-5."""
    mean = mel.mean(dim=1, keepdim=True)
    std = mel.std(dim=1, keepdim=True).clamp_min(1e-5)
    return (mel - mean) / std, mean, std


def denormalize_mel(mel_n: Tensor, mean: Tensor, std: Tensor) -> Tensor:
    """utils/audio.py:17-19."""
    return mel_n * std + mean


def chunk_plan(total_samples: int, sample_rate: int, chunk_seconds: float = 2.0, overlap_seconds: float = 0.5):
    """inference_pipeline.py:218-225, 295-319: (start_sample, end_sample, start_lat, end_lat) per chunk."""
    chunk_samples = int(chunk_seconds * sample_rate)
    overlap_samples = int(overlap_seconds * sample_rate)
    hop = chunk_samples - overlap_samples
    n = int(math.ceil((total_samples - overlap_samples) / hop))
    out = []
    for i in range(n):
        a = i * hop
        b = a + chunk_samples
        out.append((a, b, int(a / sample_rate * 16000) // 320, int(b / sample_rate * 16000) // 320))
    return out


def crossfade_window(chunk_frames: int, overlap_frames: int) -> Tensor:
    """inference_pipeline.py:255-262: ones with a linear fade-in / fade-out of overlap_frames at the two ends."""
    w = torch.ones(1, chunk_frames)
    w[0, :overlap_frames] = torch.linspace(0, 1, overlap_frames)
    w[0, -overlap_frames:] = torch.linspace(1, 0, overlap_frames)
    return w


def stitch(chunks, stats, chunk_frames: int, overlap_frames: int, total_frames: int, kernel=(5, 3)):
    """inference_pipeline.py:228-230, 359-393: overlap-add exp(denormalised chunk) under the window at i * hop frames,
    divide by the clamped accumulated weights, trim, average-pool.  chunks[i] [1,T,n_mels]; stats[i] = (mean, std).
    Returns (final_mel [n_mels,total], smoothed [1,n_mels,total])."""
    n_mels = chunks[0].shape[2]
    final_mel = torch.zeros(n_mels, total_frames + 1000)
    final_w = torch.zeros(1, total_frames + 1000)
    win = crossfade_window(chunk_frames, overlap_frames)
    hop = chunk_frames - overlap_frames
    for i, (x, (mean, std)) in enumerate(zip(chunks, stats)):
        out = torch.exp(denormalize_mel(x, mean, std)).transpose(1, 2).squeeze(0)[:, :chunk_frames]
        final_mel[:, i * hop:i * hop + chunk_frames] += out * win
        final_w[:, i * hop:i * hop + chunk_frames] += win
    final_mel = (final_mel / torch.clamp(final_w, min=1e-5))[:, :total_frames]
    kh, kw = kernel
    smooth = torch.nn.functional.avg_pool2d(final_mel[None, None], kernel_size=(kh, kw), stride=1,
                                            padding=(kh // 2, kw // 2)).squeeze(0)
    return final_mel, smooth


def longform_generate(sd: State, tab: State, z_q_global: Tensor, plan, chunk_stats, chunk_frames: int, overlap_frames: int,
                      total_frames: int, refine_strength: float, refine_steps: int, cfg_scale: float, noises):
    """The chunk loop of inference_pipeline.py:293-393 with injected draws: noises[i] = (x_coarse, noise, known_noises)."""
    prev_tail, chunks = None, []
    for i, (_, _, a, b) in enumerate(plan):
        xc, nz, kn = noises[i]
        x = inpaint_refine(sd, tab, xc, z_q_global[:, a:b, :], prev_tail, overlap_frames, refine_strength, refine_steps,
                           cfg_scale, noise=nz, known_noises=kn)
        prev_tail = x[:, -overlap_frames:, :].clone()
        chunks.append(x)
    return stitch(chunks, chunk_stats, chunk_frames, overlap_frames, total_frames) + (chunks,)


def inverse_mel_scale(fb: Tensor, melspec: Tensor, driver: str = "gels") -> Tensor:
    """torchaudio.transforms.InverseMelScale.forward (torchaudio 2.x, the reference's dependency; call sites
    generate_sample.py:125-141, inference_pipeline.py:88,395): per frame the least-squares solution of fb^T X = mel, clamped at 0.
    fb [n_stft, n_mels] (torchaudio.functional.melscale_fbanks), melspec [..., n_mels, time] -> [..., n_stft, time]."""
    shape = melspec.size()
    mel = melspec.reshape(-1, shape[-2], shape[-1])
    spec = torch.relu(torch.linalg.lstsq(fb.transpose(-1, -2)[None], mel, driver=driver).solution)
    return spec.view(shape[:-2] + (fb.shape[0], shape[-1]))


def to_dtype(sd: State, dtype) -> State:
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


# ----------------------------------------------------------------------------
# torchaudio.functional.griffinlim (third-party: torchaudio, in-image 2.11; the reference calls it through
# T.GriffinLim at generate_sample.py:135-141 and inference_pipeline.py:89,398)  -- SURVEY.md section 8f-3
# ----------------------------------------------------------------------------
def griffinlim(spec: Tensor, window: Tensor, n_fft: int, hop_length: int, win_length: int, power: float, n_iter: int,
               momentum: float, angles_init: Tensor) -> Tensor:
    """The published algorithm with the initial phases injected (the library draws torch.rand(size, complex64)):
    spec [..., n_fft // 2 + 1, frames] -> waveform [..., hop_length * (frames - 1)]."""
    mom = momentum / (1 + momentum)
    shape = spec.shape
    mag = spec.reshape(-1, shape[-2], shape[-1]).pow(1 / power)
    angles = angles_init.reshape(mag.shape).to(torch.complex64)
    tprev = torch.tensor(0.0, dtype=mag.dtype)
    for _ in range(n_iter):
        inverse = torch.istft(mag * angles, n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window)
        rebuilt = torch.stft(inverse, n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window, center=True,
                             pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
        angles = rebuilt
        if mom:
            angles = angles - tprev * mom
        angles = angles / (angles.abs() + 1e-16)
        tprev = rebuilt
    wave = torch.istft(mag * angles, n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window)
    return wave.reshape(shape[:-2] + wave.shape[-1:])
