"""DiffusionSchedule: same constructor, attributes and methods as the reference
(schedule.py:11-266).  The two sampling updates run as CUDA kernels through the
C ABI (edtts_ddim_step / edtts_ddpm_step); the forward-process helpers
(q_sample, predict_*, get_v_target) are training utilities outside the sampling
path and stay one-line tensor expressions.

Tables are built with the reference's exact op sequence on the CPU and then moved
(SURVEY.md F10): cos/cumprod differ by ulps between CPU and CUDA, and parity is
defined against the CPU reference.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from . import _lib

_TABLES = ("betas", "alphas", "alpha_bar", "sqrt_alpha_bar", "sqrt_one_minus_alpha_bar", "sqrt_recip_alpha_bar",
           "sqrt_recip_alpha_bar_minus_one", "posterior_variance", "lambda_t")


class DiffusionSchedule:
    def __init__(self, T: int, beta_start: float = 1e-4, beta_end: float = 2e-2, device: str = "cpu"):
        self.T = T
        self.device = device
        s = 0.008                                             # cosine schedule; betas args ignored (F10)
        x = torch.linspace(0, T, T + 1)
        ac = torch.cos(((x / T) + s) / (1 + s) * torch.pi * 0.5) ** 2
        ac = ac / ac[0]
        self.betas = torch.clip(1 - (ac[1:] / ac[:-1]), 0.0001, 0.9999)
        self.alphas = 1.0 - self.betas
        self.alpha_bar = torch.cumprod(self.alphas, dim=0)
        self.sqrt_alpha_bar = torch.sqrt(self.alpha_bar)
        self.sqrt_one_minus_alpha_bar = torch.sqrt(1.0 - self.alpha_bar)
        self.sqrt_recip_alpha_bar = torch.sqrt(1.0 / self.alpha_bar)
        self.sqrt_recip_alpha_bar_minus_one = torch.sqrt(1.0 / self.alpha_bar - 1)
        alpha_bar_prev = F.pad(self.alpha_bar[:-1], (1, 0), value=1.0)
        self.posterior_variance = self.betas * (1.0 - alpha_bar_prev) / (1.0 - self.alpha_bar)
        self.lambda_t = torch.log(self.sqrt_alpha_bar / self.sqrt_one_minus_alpha_bar)
        if torch.device(device).type != "cpu":
            self.to(device)

    # ---- forward-process helpers (schedule.py:61-155) -------------------------
    def q_sample(self, x0, t, noise=None):
        if noise is None:
            noise = torch.randn_like(x0)
        return (self.sqrt_alpha_bar[t][:, None, None] * x0
                + self.sqrt_one_minus_alpha_bar[t][:, None, None] * noise), noise

    def predict_x0_from_eps(self, x_t, t, eps):
        return (self.sqrt_recip_alpha_bar[t][:, None, None] * x_t
                - self.sqrt_recip_alpha_bar_minus_one[t][:, None, None] * eps)

    def predict_x0_from_v(self, x_t, t, v):
        return self.sqrt_alpha_bar[t][:, None, None] * x_t - self.sqrt_one_minus_alpha_bar[t][:, None, None] * v

    def predict_eps_from_v(self, x_t, t, v):
        return self.sqrt_one_minus_alpha_bar[t][:, None, None] * x_t + self.sqrt_alpha_bar[t][:, None, None] * v

    def get_v_target(self, x0, noise, t):
        return self.sqrt_alpha_bar[t][:, None, None] * noise - self.sqrt_one_minus_alpha_bar[t][:, None, None] * x0

    # ---- sampling updates: CUDA kernels ------------------------------------------
    def _check(self, x_t: torch.Tensor, t: torch.Tensor):
        if x_t.dim() != 3:
            raise ValueError(f"x_t must be [B, T, D], got {tuple(x_t.shape)}")
        if t.shape != (x_t.shape[0],):
            raise ValueError(f"t must be [B]={x_t.shape[0]}, got {tuple(t.shape)}")
        if self.alpha_bar.device != x_t.device:
            raise RuntimeError(f"schedule tables are on {self.alpha_bar.device}, x_t on {x_t.device}; call .to()")

    def get_ddim_step(self, x_t, t, t_prev, eps_pred, eta: float = 0.0, noise: Optional[torch.Tensor] = None):
        """schedule.py:157-202 -> (x_prev, x0_pred).  ``noise`` optionally injects the
        N(0,1) draw used when eta > 0 (the reference calls randn_like)."""
        self._check(x_t, t)
        lib = _lib.load()
        x_t, eps_pred = _lib.f32(x_t), _lib.f32(eps_pred)
        t, t_prev = _lib.i64(t), _lib.i64(t_prev)
        if eta > 0 and noise is None:
            noise = torch.randn_like(x_t)
        x_prev, x0 = torch.empty_like(x_t), torch.empty_like(x_t)
        B = x_t.shape[0]
        _lib.check(lib.edtts_ddim_step(_lib.ptr(x_t), _lib.ptr(eps_pred), _lib.ptr(noise) if eta > 0 else None,
                                       _lib.ptr(self.alpha_bar), _lib.ptr(t), _lib.ptr(t_prev), float(eta),
                                       _lib.ptr(x_prev), _lib.ptr(x0), B, x_t[0].numel(),
                                       _lib.stream_ptr(x_t.device)), "ddim_step")
        return x_prev, x0

    def ddpm_step(self, x_t, t, eps_pred, noise: Optional[torch.Tensor] = None):
        """schedule.py:204-238 -> x_prev."""
        self._check(x_t, t)
        lib = _lib.load()
        x_t, eps_pred, t = _lib.f32(x_t), _lib.f32(eps_pred), _lib.i64(t)
        if noise is None:
            noise = torch.randn_like(x_t)
        x_prev = torch.empty_like(x_t)
        _lib.check(lib.edtts_ddpm_step(_lib.ptr(x_t), _lib.ptr(eps_pred), _lib.ptr(_lib.f32(noise)),
                                       _lib.ptr(self.alphas), _lib.ptr(self.alpha_bar), _lib.ptr(self.betas),
                                       _lib.ptr(self.posterior_variance), _lib.ptr(t), _lib.ptr(x_prev),
                                       x_t.shape[0], x_t[0].numel(), _lib.stream_ptr(x_t.device)), "ddpm_step")
        return x_prev

    def get_schedule_for_steps(self, num_steps: int) -> list:
        stride = self.T // num_steps
        return list(range(self.T - 1, 0, -stride))[:num_steps]

    def to(self, device) -> "DiffusionSchedule":
        self.device = device
        for k in _TABLES:
            setattr(self, k, getattr(self, k).to(device).contiguous())
        return self


class DPMSolverPP:
    """DPM-Solver++ sampler with the reference's constructor, methods and arithmetic (schedule.py:269-531): lambda-spaced
    timesteps, v-prediction -> x0, first / second / third order updates with the reference's history handling.

    Each update (model_to_x0 + clamp + the order's update rule) is ONE streaming CUDA kernel (``edtts_dpm_step``).  The
    per-step scalars (sigma ratio, alpha_prev (1 - e^-h), ...) are evaluated with the reference's own tensor expressions
    on the CPU copy of the schedule tables -- like the tables themselves, because exp/log differ by ulps
    between CPU and CUDA and parity is defined against the CPU reference -- and uploaded as an [B, 8] coefficient block.
    """

    def __init__(self, schedule: DiffusionSchedule, order: int = 2, predict_x0: bool = False):
        self.schedule = schedule
        self.order = order
        self.predict_x0 = predict_x0
        self.device = schedule.device
        self._host = None

    # CPU copies of the three tables the solver reads (bit-identical: the tables were built on the CPU and moved)
    def _tables(self):
        if self._host is None:
            self._host = {k: getattr(self.schedule, k).detach().cpu() for k in
                          ("sqrt_alpha_bar", "sqrt_one_minus_alpha_bar", "lambda_t")}
        return self._host

    def get_time_steps(self, num_steps: int, max_t: Optional[int] = None) -> torch.Tensor:
        """schedule.py:299-324."""
        lam = self._tables()["lambda_t"]
        max_t = max_t or (self.schedule.T - 1)
        lambda_max = lam[1].item()
        lambda_min = lam[max_t].item()
        lambdas = torch.linspace(lambda_min, lambda_max, num_steps + 1)
        timesteps = []
        for l in lambdas[:-1]:
            t = (lam - l).abs().argmin().item()
            timesteps.append(max(1, min(t, max_t)))
        return torch.tensor(timesteps, device=self.device, dtype=torch.long)

    def model_to_x0(self, model_output, x_t, t):
        """schedule.py:326-337."""
        if self.predict_x0:
            return model_output
        return self.schedule.predict_x0_from_v(x_t, t, model_output)

    def _coef(self, t: torch.Tensor, t_prev: torch.Tensor, t_prev2: Optional[torch.Tensor]) -> torch.Tensor:
        """[B, 8] fp32 coefficient block of edtts_dpm_step (include/edtts.h), reference op order (schedule.py:350-431)."""
        tb = self._tables()
        t, t_prev = t.detach().cpu(), t_prev.detach().cpu()
        alpha_prev = tb["sqrt_alpha_bar"][t_prev]
        sigma_t = tb["sqrt_one_minus_alpha_bar"][t]
        sigma_prev = tb["sqrt_one_minus_alpha_bar"][t_prev]
        lambda_t = tb["lambda_t"][t]
        lambda_prev = tb["lambda_t"][t_prev]
        h = lambda_prev - lambda_t
        co = torch.zeros(t.shape[0], 8, dtype=torch.float32)
        co[:, 0] = tb["sqrt_alpha_bar"][t]
        co[:, 1] = sigma_t
        co[:, 2] = sigma_prev / sigma_t
        co[:, 3] = alpha_prev * (1 - torch.exp(-h))
        co[:, 4] = alpha_prev * ((1 - torch.exp(-h)) / h + 1)
        if t_prev2 is not None:
            h_prev = tb["lambda_t"][t_prev2.detach().cpu()] - lambda_prev
            r = h_prev / h
            co[:, 5] = 1 / r
        co[:, 6] = alpha_prev * ((1 - torch.exp(-h)) / (h ** 2) + 0.5 / h + 0.5)
        return co

    def _step(self, x, model_output, hist, co, order_used, mode, want_x0=True):
        """mode 0: model_output is v; 1: x0 (to be clamped); 2: x0 used as given."""
        lib = _lib.load()
        x, model_output = _lib.f32(x), _lib.f32(model_output)
        if x.device.type != "cuda":
            raise RuntimeError("DPMSolverPP runs on CUDA tensors only (no CPU fallback)")
        co = co.to(x.device, non_blocking=True).contiguous()
        x_prev = torch.empty_like(x)
        x0 = torch.empty_like(x) if want_x0 else None
        h1 = _lib.f32(hist[0]) if order_used >= 2 else None
        h2 = _lib.f32(hist[1]) if order_used >= 3 else None
        _lib.check(lib.edtts_dpm_step(_lib.ptr(x), _lib.ptr(model_output), _lib.ptr(h1) if h1 is not None else None,
                                      _lib.ptr(h2) if h2 is not None else None, _lib.ptr(co), order_used,
                                      mode, _lib.ptr(x_prev), _lib.ptr(x0) if want_x0 else None,
                                      x.shape[0], x[0].numel(), _lib.stream_ptr(x.device)), "dpm_step")
        return x_prev, x0

    # the three update rules on a given x0 prediction (schedule.py:339-438)
    def first_order_update(self, x, x0_pred, t, t_prev):
        return self._step(x, x0_pred, [], self._coef(t, t_prev, None), 1, 2, want_x0=False)[0]

    def second_order_update(self, x, x0_pred, x0_prev, t, t_prev, t_prev2):
        return self._step(x, x0_pred, [x0_prev], self._coef(t, t_prev, t_prev2), 2, 2, want_x0=False)[0]

    def third_order_update(self, x, x0_preds, t, t_prev, ts_history):
        return self._step(x, x0_preds[0], [x0_preds[1], x0_preds[2]], self._coef(t, t_prev, None), 3, 2, want_x0=False)[0]

    @torch.no_grad()
    def sample(self, model, x_T, sem_features, num_steps: int = 10, max_t: Optional[int] = None,
               return_intermediates: bool = False):
        """schedule.py:440-527.  With this package's decoder the whole loop runs fused and graph-captured
        (``_sample_fused``); any other callable model takes the reference's step-by-step route with one update kernel
        per step."""
        max_t = max_t or 950
        ts = self.get_time_steps(num_steps, max_t).tolist()
        if (hasattr(model, "prepare_context") and hasattr(model, "step") and x_T.dim() == 3 and x_T.is_cuda
                and sem_features.shape[0] == x_T.shape[0]):
            return self._sample_fused(model, x_T, sem_features, ts, return_intermediates)
        x = x_T
        x0_history, t_history, intermediates = [], [], []
        B = x.shape[0]
        for i, t in enumerate(ts):
            t_tensor = torch.full((B,), t, device=x.device, dtype=torch.long)
            step_idx = torch.full((B,), i, device=x.device, dtype=torch.long)
            model_output = model(x, t_tensor, sem_features=sem_features, step_idx=step_idx)
            min_len = min(model_output.shape[1], x.shape[1])
            model_output = model_output[:, :min_len, :]
            x = x[:, :min_len, :]
            t_prev = ts[i + 1] if i < len(ts) - 1 else 0
            t_prev_tensor = torch.full((B,), t_prev, device=x.device, dtype=torch.long)
            used, nh, tp2 = self._order_used(len(x0_history), t_history)
            x, x0_pred = self._step(x, model_output, x0_history[-nh:] if nh else [],
                                    self._coef(t_tensor, t_prev_tensor, tp2), used, 1 if self.predict_x0 else 0)
            if return_intermediates:
                intermediates.append(x0_pred.clone())
            x0_history.append(x0_pred)
            t_history.append(t_prev_tensor)
            if len(x0_history) > 2:
                x0_history.pop(0)
                t_history.pop(0)
        if return_intermediates:
            return x, intermediates
        return x

    def _order_used(self, n_hist: int, t_history):
        """The reference's choice of update rule (schedule.py:495-508) -> (order used, history tensors, t_prev2)."""
        if self.order == 1 or n_hist == 0:
            return 1, 0, None
        if self.order == 2 or n_hist == 1:
            return 2, 1, t_history[-1]
        return 3, 2, None

    use_cuda_graph = True

    def _sample_fused(self, dec, x_T, sem_features, ts, return_intermediates):
        """The same loop with (a) the context K/V prepared once -- the reference re-projects sem_features inside every
        model call (decoder.py:83-93) although they do not depend on the step, (b) model_to_x0 + clamp + update rule
        fused into the last decoder kernel of each step (EDTTS_STEP_DPM), x updated in place, x0 history in a ring of
        buffers, and (c) the whole loop replayed from a CUDA graph (per-shape plan, like EdgeInference.generate_mel)."""
        dev = x_T.device
        B, T, _ = x_T.shape
        S = sem_features.shape[1]
        n = len(ts)
        # the plan holds a graph with the decoder's weight pointers baked in: keyed on the decoder as well (a teacher and a
        # student decoder share one solver in the reference pipeline)
        key = (dec._uid, B, T, S, tuple(ts), self.order, self.predict_x0, bool(return_intermediates), dec.precision, str(dev))
        plans = self.__dict__.setdefault("_plans", {})
        p = plans.get(key)
        if p is None:
            p = {"graph": None, "epoch": None}
            p["x"] = torch.empty(B, T, x_T.shape[2], dtype=torch.float32, device=dev)
            p["feats"] = torch.empty(B, S, sem_features.shape[2], dtype=torch.float32, device=dev)
            p["x0"] = [torch.empty_like(p["x"]) for _ in range(n if return_intermediates else min(n, 3))]
            cfg = dec.cfg
            p["kv"] = dec.alloc_kv(B, S, dev)
            p["mod_all"] = torch.empty(n * B, 2 * cfg.layers, 2 * cfg.hidden, dtype=torch.float32, device=dev)
            p["mods"] = [p["mod_all"][i * B:(i + 1) * B] for i in range(n)]
            nb_ctx, nb_step = dec.workspace_bytes(B, T, S)
            p["ws_ctx"] = torch.empty(max(nb_ctx, 256), dtype=torch.uint8, device=dev)
            p["ws_step"] = torch.empty(max(nb_step, 256), dtype=torch.uint8, device=dev)
            p["t"], p["si"], p["coef"], p["used"] = [], [], [], []
            t_hist, n_hist = [], 0
            for i, t in enumerate(ts):
                tt = torch.full((B,), t, dtype=torch.long)
                tp = torch.full((B,), ts[i + 1] if i < n - 1 else 0, dtype=torch.long)
                used, nh, tp2 = self._order_used(n_hist, t_hist)
                p["t"].append(tt.to(dev))
                p["si"].append(torch.full((B,), i, dtype=torch.long, device=dev))
                p["coef"].append(self._coef(tt, tp, tp2).to(dev).contiguous())
                p["used"].append(used)
                t_hist = (t_hist + [tp])[-2:]
                n_hist = min(n_hist + 1, 2)
            # one (t, step index) per step: they are torch.full over the batch (reference schedule.py:476-477), and a
            # conditioning row depends on nothing else -> computed once per step, broadcast over the batch (bit-identical)
            p["t_uni"] = torch.stack([t[0] for t in p["t"]])
            p["si_uni"] = torch.stack([s_[0] for s_ in p["si"]])
            p["mod_uni"] = torch.empty(n, 2 * cfg.layers, 2 * cfg.hidden, dtype=torch.float32, device=dev)
            plans[key] = p

        def run():
            dec.prepare_context(None, p["feats"], T, out=p["kv"], ws=p["ws_ctx"])
            dec.prepare_cond(p["t_uni"], p["si_uni"], T, S, out=p["mod_uni"])   # all steps in one launch, one row per step
            p["mod_all"].view(n, B, *p["mod_uni"].shape[1:]).copy_(p["mod_uni"][:, None].expand(n, B, *p["mod_uni"].shape[1:]))
            nb = len(p["x0"])
            for i in range(n):
                a = _lib.StepArgs()
                a.mode = _lib.STEP_DPM
                a.dpm_coef = p["coef"][i].data_ptr()
                a.dpm_order = p["used"][i]
                a.dpm_predict_x0 = 1 if self.predict_x0 else 0
                if p["used"][i] == 2:
                    a.dpm_hist1 = p["x0"][(i - 1) % nb].data_ptr()
                elif p["used"][i] == 3:                               # the reference's order: [older, newer]
                    a.dpm_hist1 = p["x0"][(i - 2) % nb].data_ptr()
                    a.dpm_hist2 = p["x0"][(i - 1) % nb].data_ptr()
                a.x_prev_out = p["x"].data_ptr()                      # in place: element-wise read-then-write
                a.x0_out = p["x0"][i % nb].data_ptr()
                dec.step(p["x"], p["mods"][i], p["kv"], S, a, ws=p["ws_step"])

        p["feats"].copy_(sem_features)
        p["x"].copy_(x_T)
        if not self.use_cuda_graph:
            run()
        else:
            token = dec.weights_token(T, S)                           # repacks + bumps the epoch if a parameter changed
            if p["graph"] is None or p["epoch"] != token:
                run()                                                 # warm-up: builds weight views / packed images
                p["x"].copy_(x_T)
                torch.cuda.synchronize(dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    run()
                p["graph"], p["epoch"] = g, token
                p["x"].copy_(x_T)
            p["graph"].replay()
        x = p["x"].clone()
        if return_intermediates:
            return x, [b.clone() for b in p["x0"]]
        return x

    def to(self, device) -> "DPMSolverPP":
        self.device = device
        self.schedule.to(device)
        return self
