#!/usr/bin/env python
"""bench.py -- mel-frames/sec of the few-step sampling path (BASELINE.json metric).

A "step" is one full ``EdgeInference.generate_mel`` (context prep + 4 DDIM steps) over BASELINE config 3: default CFG,
4-step DDIM, GLOBAL batch 256, 800 mel frames per utterance.  At N GPUs the 256 utterances are batch-sharded (256 / N per
rank: **strong scaling**, the configuration north_star / SURVEY 8(e) name), no collective inside the sampling loop, one NCCL
gather of the final mel onto rank 0 (asynchronous: the gather of step k runs under step k + 1; the timed region ends after
the last one has landed).  Before timing, at N > 1, the gathered shards are asserted bit-equal to rank 0's own single-GPU
generate of all 256.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp32|bf16] [--impl reference]

One JSON line on stdout (rank 0).  Keys beyond the base contract:
  roofline      dominant kernel, timed live with CUDA events by the library's per-class profiler in an eager pass of the same
                workload; ``traffic`` is NOT measured in this run -- it is the ncu ``dram__bytes`` of the committed capture
                named in ``traffic_source`` (profiles/traffic.json)
  e2e           pinned host sem_idx -> public API -> pinned host mel per rank, copies inside the timed region (no collective:
                at N > 1 every rank reads its own shard back, so ``value`` and ``e2e`` time different last hops -- stated
                in ``e2e.path`` / ``config.parallelism``)
  weak          (N > 1) the same step with 256 utterances PER GPU
  cfg5          (N > 1) BASELINE config 5, batch 128 x 3000 frames sharded over the N GPUs
  configs       (N = 1) the other BASELINE configurations and cfg3 on both fp32-grade paths (cfg3_fp32_path: precision="fp32", tf32 x 3
                on the tensor cores, the class default; cfg3_fp32_simt_path: the CUDA-core checker), each with ms / frames per
                second / fraction of the BASELINE.md section 3 ideal
  cpu_baseline  (N = 1) the UNMODIFIED reference (baseline/_ref) on the host cores: all threads and one thread, cfg1 in full,
                the same modules eager on cuda:0 with TF32 off as a labelled context number, and the free-running parity of
                this package's fp32 / bf16 paths against it (oracle/parity.py criterion)

``--impl reference`` times the unmodified reference's own ``EdgeInference.generate_mel`` on the host cores with the same
``config`` (each step a bounded sample of the workload, stated in ``cpu_baseline.sample``).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "mel_frames_per_sec"
UNIT = "mel frames/s"
B_GLOBAL, S_TOK, N_STEPS = 256, 400, 4            # BASELINE config 3
T_MEL = 2 * S_TOK
IDEAL_MS = {"cfg2": 0.148, "cfg3": 2.37, "cfg4_per_step": 0.0741, "cfg5": 7.58}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def workload_config(world: int) -> dict:
    """The ``config`` object of the line -- identical for the B200 arm and the reference arm."""
    return {"workload": "cfg3: default CFG (hidden 160, 4 layers, 4 heads), 4-step DDIM generate_mel, global batch 256, "
                        "800 mel frames (400 semantic tokens) per utterance",
            "global_batch": B_GLOBAL, "batch_per_gpu": B_GLOBAL // world, "T_mel": T_MEL, "S_tokens": S_TOK,
            "ddim_steps": N_STEPS, "n_gpus": world,
            "parallelism": (f"batch-sharded x{world} (strong scaling: {B_GLOBAL // world} utterances per GPU), one "
                            f"asynchronous NCCL gather of the final mel onto rank 0 per step inside the timed region"
                            if world > 1 else "single GPU"),
            "l2_policy": "no flush: the per-step working set (activations 0.66-0.9 GB at batch 256, >= 0.1 GB at batch 32) "
                         "exceeds or matches the 126 MB L2 and every step rewrites it"}


# ----------------------------------------------------------------------------- algorithmic work
def flops_per_generate(B: int, T: int, S: int, steps: int) -> dict:
    """Algorithmic FLOPs (2*M*N*K per GEMM; band attention counts 129T-4160 keys/head), SURVEY 8(d)."""
    R = B * T
    H, M, F = 160, 80, 320
    gemm_step = 2 * R * (M * H + 4 * (H * 3 * H + H * H + H * H + H * H + H * 2 * F + F * H) + H * M)
    win = B * 4 * (129 * T - 4160) * 40 * 2 * 2 * 4           # 4 layers
    cross = B * 4 * T * S * 40 * 2 * 2 * 4
    ctx = 2 * B * S * 4 * (H * 80 + 80 * 2 * H)
    return dict(gemm=gemm_step * steps + ctx, attn_window=win * steps, attn_cross=cross * steps, ctx=ctx,
                total=(gemm_step + win + cross) * steps + ctx)


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index: int, period: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv = None
            log("[bench] NVML unavailable:", e)

    NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
             0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
             0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def run(self):
        if self.nv is None:
            return
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.NAMES.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(self.period)

    def stop(self) -> dict:
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------- reference arm (unmodified reference)
def _cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def reference_cpu_rate(batch: int, steps: int, warmup: int, threads: int, min_seconds: float = 0.0, S: int = S_TOK):
    """The unmodified reference's generate_mel on the host cores (baseline/_ref); falls back to the oracle port only when
    baseline/_ref did not travel.  Returns (frames/s, ms per generate, repetitions, kind)."""
    from baseline import reference_arm as RA
    torch.set_num_threads(threads)
    if RA.available():
        inf, _ = RA.make_inference("cpu")
        rate, ms, n = RA.time_generate(inf, batch, S, N_STEPS, steps, warmup, min_seconds)
        return rate, ms, n, "reference"
    from oracle import edtts_oracle as O, synth
    sd, tab = synth.synth_decoder_state(0), O.cosine_schedule(1000)
    idx, xT = synth.synth_sem_idx(3, batch, S), synth.synth_noise(5, batch, 2 * S)
    for _ in range(warmup):
        O.generate_mel(sd, tab, idx, N_STEPS, xT)
    t0, n = time.perf_counter(), 0
    while n < steps or (time.perf_counter() - t0) < min_seconds:
        O.generate_mel(sd, tab, idx, N_STEPS, xT)
        n += 1
    dt = time.perf_counter() - t0
    return batch * 2 * S * n / dt, dt / n * 1e3, n, "port"


def run_reference(args):
    """``--impl reference``: rank 0 alone; each step = the reference's generate_mel over a bounded sample of cfg3 (as many
    utterances of the cfg3 shape as keep (steps + warmup) generates within ~2.5 minutes, at most the arm's batch)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    probe_rate, _, _, kind = reference_cpu_rate(4, 1, 1, threads)                 # frames/s is flat in B
    budget_s = float(os.environ.get("EDTTS_REF_BUDGET_S", "150"))
    total = max(args.steps + args.warmup, 1)
    b_fit = int(probe_rate * budget_s / total / T_MEL)
    batch = max(1, min(B_GLOBAL, 1 << max(b_fit, 1).bit_length() - 1))
    rate, ms, n, kind = reference_cpu_rate(batch, args.steps, args.warmup, threads)
    what = ("the UNMODIFIED reference (baseline/_ref): EdgeInference.generate_mel, inference.py:23-53" if kind == "reference"
            else "oracle port (baseline/_ref absent)")
    sample = (f"{what}, CPU fp32, torch {torch.__version__}, {threads} threads on {_cpu_model()}; each step = {batch} of the "
              f"256 utterances x {T_MEL} frames x {N_STEPS} DDIM steps ({ms:.0f} ms), {n} steps; frames/s is flat in the "
              f"batch size on CPU (SURVEY 8d), so the sample's rate stands for the full batch")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": n,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
                         "sample_batch": batch, "host_cpus": os.cpu_count()},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def cpu_baseline_leg(inf_gpu, dec_gpu, dev) -> dict:
    """N = 1, rank 0: the unmodified reference on the host cores next to the B200 numbers, plus the context figures
    BASELINE.md section 4 lists and the free-running parity of this package against it."""
    from baseline import reference_arm as RA
    threads = os.cpu_count() or 1
    rate, ms, n, kind = reference_cpu_rate(8, 2, 1, threads, min_seconds=8.0)
    out = {"value": rate, "unit": UNIT, "cores": threads, "kind": kind, "host_cpus": os.cpu_count(), "cpu_model": _cpu_model(),
           "sample": f"{'unmodified reference (baseline/_ref)' if kind == 'reference' else 'oracle port'} generate_mel on CPU "
                     f"fp32, 8 utterances x {T_MEL} frames x {N_STEPS} steps, {n} repetitions ({ms:.0f} ms each), "
                     f"torch {torch.__version__}"}
    try:
        r1, ms1, n1, _ = reference_cpu_rate(2, 1, 1, 1)
        out["one_thread"] = {"value": r1, "cores": 1, "sample": f"2 utterances x {T_MEL} frames, {n1} repetition ({ms1:.0f} ms)"}
        c1, cms, cn, _ = reference_cpu_rate(1, 3, 1, threads, S=200)
        c1s, cms1, _, _ = reference_cpu_rate(1, 2, 1, 1, S=200)
        out["cfg1"] = {"workload": "cfg1: batch 1, 400 mel frames, 4-step DDIM, CPU, in full",
                       "all_threads": {"value": c1, "ms": cms, "cores": threads}, "one_thread": {"value": c1s, "ms": cms1}}
    except Exception as e:  # pragma: no cover
        out["context_error"] = repr(e)
    if kind != "reference":
        return out
    try:
        # context, clearly not the baseline: the same reference modules eager on the B200, TF32 off
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        ginf, _ = RA.make_inference(str(dev))
        g, gms, gn = RA.time_generate(ginf, 64, S_TOK, N_STEPS, 3, 2, device=str(dev))
        out["reference_eager_on_b200"] = {"value": g, "ms": gms, "sample": f"context only: the unmodified reference modules, "
                                          f"eager PyTorch on {dev}, fp32 with TF32 off, 64 utterances x {T_MEL} frames, {gn} repetitions"}
        del ginf
        torch.cuda.empty_cache()
    except Exception as e:  # pragma: no cover
        out["reference_eager_on_b200"] = {"error": repr(e)}
    try:
        out["parity_vs_reference"] = parity_vs_reference(inf_gpu, dec_gpu, dev)
    except Exception as e:  # pragma: no cover
        out["parity_vs_reference"] = {"error": repr(e)}
    return out


def parity_vs_reference(inf_gpu, dec_gpu, dev, B: int = 16, S: int = 40) -> dict:
    """Free-running 4-step generate_mel of this package (fp32 and bf16 paths) against the unmodified reference run on the
    CPU here, same weights / tokens / x_T; the F9-aware criterion of oracle/parity.py (checker only)."""
    from baseline import reference_arm as RA
    from oracle import parity, synth
    torch.set_num_threads(os.cpu_count() or 1)
    rinf, _ = RA.make_inference("cpu")
    idx, xT = synth.synth_sem_idx(11, B, S), synth.synth_noise(12, B, 2 * S)
    rec = []
    hook = rinf.decoder.register_forward_hook(lambda m, a, o: rec.append((a[0].clone(), o.clone())))
    real = torch.randn
    torch.randn = lambda *a, **k: xT.clone()                   # inference.py:33 draws x_T; inject the same one
    try:
        ref = rinf.generate_mel(idx, N_STEPS)
    finally:
        torch.randn = real
        hook.remove()
    ab = float(rinf.schedule.alpha_bar[999])
    eps_ref0 = rec[0][1]
    x0_ref0 = torch.clamp((xT - (1 - ab) ** 0.5 * eps_ref0) / ab ** 0.5, -3, 3)
    res = {}
    keep = dec_gpu.precision
    for prec, tol in (("fp32", 1e-4), ("bf16", 1e-4)):
        dec_gpu.precision = prec
        t = torch.full((B,), 999, dtype=torch.long, device=dev)
        eps0 = dec_gpu(xT.to(dev), t, idx.to(dev), torch.zeros_like(t)).cpu()
        one = inf_gpu.generate_mel(idx.to(dev), 1, x_T=xT.to(dev)).cpu()          # the clamped x0 of step 0
        got = inf_gpu.generate_mel(idx.to(dev), N_STEPS, x_T=xT.to(dev)).cpu()
        r = parity.f9_report(ab, xT, eps_ref0, eps0, x0_ref0, one, ref, got, tol)
        r["step0_rel_l2_eps"] = ((eps0 - eps_ref0).norm() / eps_ref0.norm()).item()
        r["step0_max_abs_eps"] = (eps0 - eps_ref0).abs().max().item()
        res[prec] = r
    dec_gpu.precision = keep
    res["sample"] = f"{B} utterances x {2 * S} frames, 4-step, same weights / tokens / x_T as the unmodified reference on CPU"
    return res


# ----------------------------------------------------------------------------- GPU arm
def cuda_time(fn, warm: int, reps: int) -> float:
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def other_configs(E, synth, cfg, dec, sched, inf, dev) -> dict:
    """The other BASELINE configurations on this one GPU (median of a few CUDA-event timings each, inputs resident)."""
    out = {}

    def put(name, frames, ms, ideal_ms=None, **kw):
        d = {"ms": ms, "mel_frames_per_sec": frames / (ms * 1e-3), **kw}
        if ideal_ms:
            d["ideal_ms_at_sustained_bf16_peak"] = ideal_ms
            d["frac_of_ideal"] = ideal_ms / ms
        out[name] = d

    enc = E.SemanticEncoder(E.CFG(device=str(dev), use_fsq=False), load_hubert=False).to(dev).eval()
    enc.proj.load_state_dict(synth.synth_proj_state(0))
    enc.vq.load_state_dict(synth.synth_vq_state(0))
    inf2 = E.EdgeInference(cfg, sched, enc, dec)
    # cfg2: 768-d features -> proj -> VQ -> 1-step generate, batch 64
    B, S = 64, 400
    h = torch.randn(B, S, 768, device=dev, generator=torch.Generator(dev).manual_seed(2))
    xT = torch.randn(B, 2 * S, cfg.n_mels, device=dev)
    put("cfg2", B * 2 * S, cuda_time(lambda: inf2.generate_mel(enc.encode_features(h), 1, x_T=xT), 3, 10), IDEAL_MS["cfg2"],
        workload="768-d features -> proj -> VQ -> 1-step generate_mel, batch 64, 800 frames")
    put("cfg2_proj_vq_only", B * 2 * S, cuda_time(lambda: enc.encode_features(h), 3, 10), workload="proj + VQ encode, 25,600 rows")
    del h
    # cfg4: 1000-step DDPM, batch 32
    B = 32
    idx = synth.synth_sem_idx(4, B, S).to(dev)
    xT = torch.randn(B, 2 * S, cfg.n_mels, device=dev)
    ms = cuda_time(lambda: inf2.sample_ddpm(idx, xT), 1, 2)
    put("cfg4", B * 2 * S, ms, IDEAL_MS["cfg4_per_step"] * 1000, ms_per_sampling_step=ms / 1000,
        frame_steps_per_sec=B * 2 * S * 1000 / (ms * 1e-3), workload="1000-step DDPM ancestral loop, batch 32, 800 frames")
    # cfg5: long utterances; one GPU's share of 128 over 8, and all 128 on this GPU
    S = 1500
    for B in (16, 128):
        idx = synth.synth_sem_idx(5, B, S).to(dev)
        xT = torch.randn(B, 2 * S, cfg.n_mels, device=dev)
        ms = cuda_time(lambda: inf2.generate_mel(idx, 4, x_T=xT), 2, 4)
        put("cfg5" if B == 128 else "cfg5_share_of_8", B * 2 * S, ms, IDEAL_MS["cfg5"] * B / 128,
            workload=f"4-step DDIM, 3000 frames / 1500 tokens, batch {B}")
        del idx, xT
        inf2._plans.clear()
        torch.cuda.empty_cache()
    # cfg3 on the fp32-grade (max-abs 1e-4) paths: precision="fp32" = tf32 x 3 split products on the tensor cores (the class default),
    # precision="fp32_simt" = the CUDA-core kernels it is tested against
    keep = dec.precision
    idx = synth.synth_sem_idx(100, B_GLOBAL, S_TOK).to(dev)
    xT = torch.randn(B_GLOBAL, T_MEL, cfg.n_mels, device=dev)
    for prec, key, what in (("fp32", "cfg3_fp32_path", "tf32 x 3 tensor-core GEMMs and attention (the class default; max-abs 1e-4 parity path)"),
                            ("fp32_simt", "cfg3_fp32_simt_path", "CUDA-core FFMA kernels (the checker of the fp32 path)")):
        dec.precision = prec
        ms = cuda_time(lambda: inf2.generate_mel(idx, 4, x_T=xT), 1, 3)
        put(key, B_GLOBAL * T_MEL, ms, workload=f"cfg3 with precision='{prec}': {what}")
        inf2._plans.clear()
    dec.precision = keep
    inf2._plans.clear()
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--precision", default=os.environ.get("EDTTS_BENCH_PRECISION", "auto"))
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=0, help="utterances per GPU (default: 256 / N, BASELINE cfg3 sharded)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configurations (N = 1)")
    ap.add_argument("--no-graph", action="store_true", help="eager launches (for ncu launch lists)")
    ap.add_argument("--timed-only", action="store_true", help="skip the profiler / e2e / cpu passes (ncu runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    stdout_fd = None
    if world > 1:
        # NCCL prints its version banner (and, with NCCL_DEBUG, its log) on file descriptor 1 when the communicator is
        # created: point fd 1 at stderr until the ONE JSON line of the contract is printed
        sys.stdout.flush()
        stdout_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__ as ge
    ge.build()
    import edge_diffusion_tts_b200 as E
    from edge_diffusion_tts_b200 import _lib
    from edge_diffusion_tts_b200.dist import RootGather, shard_bounds
    from oracle import synth

    lib = _lib.load()
    assert lib.edtts_device_supported() == 1, "bench needs an sm_100 (B200) device"
    prec = args.precision
    if prec == "auto":
        prec = "bf16" if lib.edtts_packed_bf16_bytes() > 0 else "fp32"
    assert B_GLOBAL % world == 0, "cfg3's 256 utterances split evenly over 1/2/4/8 GPUs"
    B = args.batch or B_GLOBAL // world
    Bg = B * world
    cfg = E.CFG(device=str(dev))
    dec = E.EdgeDiffusionDecoder(cfg).to(dev).eval()
    dec.load_state_dict(synth.synth_decoder_state(0), strict=True)
    dec.precision = prec
    sched = E.DiffusionSchedule(cfg.diff_steps, device=dev)
    inf = E.EdgeInference(cfg, sched, torch.nn.Identity(), dec, use_cuda_graph=not args.no_graph)

    # synthetic inputs of the BASELINE shape: the same global batch on every rank, each rank works on its contiguous shard
    idx_glob = synth.synth_sem_idx(100, Bg, S_TOK)
    xT_glob = synth.synth_noise(200, Bg, T_MEL)
    lo, hi = shard_bounds(Bg, world)[rank]
    idx_host = idx_glob[lo:hi].clone().pin_memory()
    idx = idx_host.to(dev)
    x_T = xT_glob[lo:hi].to(dev)
    frames_per_step = Bg * T_MEL

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sharded_equals_single = None
    if world > 1:
        # the gathered shards are the single-GPU result, bit for bit (batch invariance)
        rg0 = RootGather(Bg, dst=0)
        rg0.start(inf.generate_mel(idx, N_STEPS, x_T=x_T))
        full = rg0.finish()
        if rank == 0:
            single = inf.generate_mel(idx_glob.to(dev), N_STEPS, x_T=xT_glob.to(dev))
            sharded_equals_single = bool(torch.equal(full, single))
            assert sharded_equals_single, "gathered shards differ from the single-GPU generate"
            del single
            inf._plans.pop(("ddim", Bg, S_TOK, N_STEPS, prec, str(dev)), None)
            torch.cuda.empty_cache()
        del full, rg0
    del idx_glob, xT_glob

    def timed_steps(n_warm: int, n_steps: int, idx_d, xT_d, batch_global: int):
        """(ms per step, max over ranks) of generate_mel [+ asynchronous gather to rank 0] on resident inputs."""
        rg = RootGather(batch_global, dst=0) if world > 1 else None

        def one():
            mel = inf.generate_mel(idx_d, N_STEPS, x_T=xT_d)
            if rg is not None:
                rg.start(mel)

        for _ in range(n_warm):
            one()
        if rg is not None and n_warm:
            rg.finish()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(n_steps):
            one()
        if rg is not None:
            rg.finish()                                       # the last gather has landed on rank 0
        ev1.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() / n_steps

    # ---- device-resident throughput ("value") ---------------------------------
    timed_steps(args.warmup, 1, idx, x_T, Bg)                  # warm-up (graph capture, NCCL channels)
    sampler = ClockSampler(local)
    sampler.start()
    ms_step = timed_steps(0, args.steps, idx, x_T, Bg)
    clocks = sampler.stop()
    value = frames_per_step / (ms_step * 1e-3)

    if args.timed_only:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "ms_per_step": ms_step, "dtype": prec,
                              "note": "timed-only run (profiling helper), not a bench line"}))
        if world > 1:
            dist.destroy_process_group()
        return

    # launches per generate: the graph was captured during warm-up; count one eager generate
    inf.use_cuda_graph = False
    c1 = _lib.launch_counts()
    inf.generate_mel(idx, N_STEPS, x_T=x_T)
    c2 = _lib.launch_counts()
    per_generate = {k: c2[k] - c1[k] for k in c2 if c2[k] - c1[k]}
    launches = sum(per_generate.values()) * args.steps

    # ---- per-kernel-class timing, live, eager pass of the same workload -----------
    _lib.prof_enable(True)
    n_prof = min(args.steps, 5)
    for _ in range(n_prof):
        inf.generate_mel(idx, N_STEPS, x_T=x_T)
    prof = _lib.prof_collect()
    _lib.prof_enable(False)
    inf.use_cuda_graph = not args.no_graph
    tot_ms = sum(ms for ms, _ in prof.values()) or 1.0
    kernels = {k: {"ms_per_step": ms / n_prof, "launches_per_step": n // n_prof, "share": ms / tot_ms}
               for k, (ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0])}
    fl = flops_per_generate(B, T_MEL, S_TOK, N_STEPS)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tensor_peak = peaks.get("bf16_tflops_sustained", 1590.0 * 0.88)
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback"
    # the fused kernel (tc_layer_bf16) executes every decoder FLOP of the step: GEMMs, window and cross attention
    class_flops = {"gemm_simt_fp32": fl["gemm"], "tc_gemm_bf16": fl["ctx"] if "tc_layer_bf16" in kernels else fl["gemm"],
                   "tc_layer_bf16": fl["total"] - fl["ctx"],
                   "attn_window_simt_fp32": fl["attn_window"], "tc_attn_window_bf16": fl["attn_window"],
                   "attn_cross_simt_fp32": fl["attn_cross"], "tc_attn_cross_bf16": fl["attn_cross"]}
    dom = next((k for k in kernels if k in class_flops), None)
    roofline = None
    if dom:
        k = kernels[dom]
        achieved = class_flops[dom] / (k["ms_per_step"] * 1e-3) / 1e12
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic, traffic_src = tj.get(dom), tj.get("source")
        except Exception:
            pass
        # the kernel's nearest hardware floor is the MUFU pipe (softmax ex2: 16 ops/clk/SM, measured in tools/ubench/mufu.cu)
        n_exp = B * T_MEL * 4 * 4 * ((129 * T_MEL - 4160) / T_MEL + S_TOK) * N_STEPS + B * T_MEL * 320 * 4 * N_STEPS
        sm_mhz = (clocks.get("sm_mhz") or 1965) * 1e6
        mufu_floor_ms = n_exp / (16 * 148 * sm_mhz) * 1e3
        roofline = {"kernel": dom, "bound": "tensor", "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
                    "mufu": {"ops_per_step": n_exp, "floor_ms_per_step": mufu_floor_ms,
                             "frac_of_mufu_floor": mufu_floor_ms / k["ms_per_step"],
                             "note": "exp2 (softmax) + tanh (SiLU) operations / (16 per clk per SM x 148 SMs x SM clock)"},
                    "frac": achieved / tensor_peak, "traffic": traffic,
                    "traffic_source": (traffic_src or "profiles/traffic.json") + " (static: ncu dram__bytes of the committed "
                                      "capture at batch 256, NOT measured in this run)",
                    "peak_source": peak_src,
                    "avg_launch_ms": k["ms_per_step"] / max(k["launches_per_step"], 1),
                    "share_of_step": k["share"], "flops_per_step": class_flops[dom]}

    # ---- end to end through the public API with HOST buffers ----------------------
    mel_host = torch.empty(B, T_MEL, cfg.n_mels, dtype=torch.float32).pin_memory()

    def e2e_step():
        d_idx = idx_host.to(dev, non_blocking=True)           # H2D of this step's inputs (pinned)
        mel = inf.generate_mel(d_idx, N_STEPS)                # public API; x_T drawn on device as the reference does
        mel_host.copy_(mel, non_blocking=True)                # D2H of the result
        torch.cuda.current_stream(dev).synchronize()
        return mel_host

    for _ in range(args.warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize(dev)
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e = {"value": frames_per_step * args.steps / dt.item(), "unit": UNIT,
           "h2d_bytes_per_step": idx_host.numel() * 8, "d2h_bytes_per_step": mel_host.numel() * 4,
           "path": "per rank: pinned host sem_idx shard -> generate_mel(sem_idx, 4) -> pinned host mel shard; every call "
                   "synchronised before the next one starts; no collective (each rank reads its own shard back), bytes are per rank"}

    # the same traffic as a serving loop would issue it: the D2H of call k runs on a copy stream under call k + 1
    # (two pinned result buffers); reported next to the synchronous number, not instead of it
    copy_stream = torch.cuda.Stream(dev)
    mel_hosts = [mel_host, torch.empty_like(mel_host).pin_memory()]
    done = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_pipelined(k):
        d_idx = idx_host.to(dev, non_blocking=True)
        mel = inf.generate_mel(d_idx, N_STEPS)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))
        done[k % 2].synchronize()                             # the buffer's previous copy has landed
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            mel_hosts[k % 2].copy_(mel, non_blocking=True)
            mel.record_stream(copy_stream)
            done[k % 2].record(copy_stream)

    for k in range(args.warmup):
        e2e_pipelined(k)
    torch.cuda.synchronize(dev)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        e2e_pipelined(k)
    torch.cuda.synchronize(dev)
    dtp = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dtp, op=dist.ReduceOp.MAX)
    e2e["pipelined_value"] = frames_per_step * args.steps / dtp.item()

    # ---- extra multi-GPU lines: weak scaling (256 per GPU) and BASELINE cfg5 sharded over the N GPUs -------------------
    weak = cfg5 = None
    if world > 1 and not args.batch:
        inf._plans.clear()
        torch.cuda.empty_cache()
        idx_w = synth.synth_sem_idx(100 + rank, B_GLOBAL, S_TOK).to(dev)
        xT_w = synth.synth_noise(200 + rank, B_GLOBAL, T_MEL).to(dev)
        ms_w = timed_steps(3, max(args.steps // 2, 3), idx_w, xT_w, B_GLOBAL * world)
        weak = {"value": B_GLOBAL * world * T_MEL / (ms_w * 1e-3), "unit": UNIT, "ms_per_step": ms_w, "batch_per_gpu": B_GLOBAL,
                "global_batch": B_GLOBAL * world, "scaling": "weak"}
        del idx_w, xT_w
        inf._plans.clear()
        torch.cuda.empty_cache()
        B5, S5 = 128 // world, 1500
        idx_5 = synth.synth_sem_idx(500 + rank, B5, S5).to(dev)
        xT_5 = synth.synth_noise(600 + rank, B5, 2 * S5).to(dev)
        ms_5 = timed_steps(3, max(args.steps // 4, 3), idx_5, xT_5, 128)
        cfg5 = {"workload": f"cfg5: 4-step DDIM, 3000 frames / 1500 tokens, batch 128 sharded over {world} GPUs "
                            f"({B5} per GPU), gather to rank 0", "value": 128 * 2 * S5 / (ms_5 * 1e-3), "unit": UNIT,
                "ms_per_step": ms_5, "ideal_ms_at_sustained_bf16_peak": IDEAL_MS["cfg5"] / world,
                "frac_of_ideal": IDEAL_MS["cfg5"] / world / ms_5}
        del idx_5, xT_5
        inf._plans.clear()
        torch.cuda.empty_cache()

    configs = None
    if rank == 0 and world == 1 and not args.no_configs and not args.batch and prec == "bf16":
        try:
            configs = other_configs(E, synth, cfg, dec, sched, inf, dev)
            configs["cfg3"] = {"ms": ms_step, "mel_frames_per_sec": value, "ideal_ms_at_sustained_bf16_peak": IDEAL_MS["cfg3"],
                               "frac_of_ideal": IDEAL_MS["cfg3"] / ms_step, "workload": "this line's headline"}
        except Exception as e:  # pragma: no cover
            configs = {"error": repr(e)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_leg(inf, dec, dev)

    if stdout_fd is not None:
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        os.close(stdout_fd)
    if rank == 0:
        conf = workload_config(world)
        if args.batch:
            conf.update({"batch_per_gpu": B, "global_batch": Bg})
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": prec, "data": "synthetic", "config": conf, "cuda_graph": not args.no_graph,
            "e2e": e2e, "gpu_launches": launches, "launches_per_step": per_generate, "clocks": clocks,
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu,
            "sharded_equals_single_gpu_bitwise": sharded_equals_single, "weak": weak, "cfg5": cfg5, "configs": configs,
            "algorithmic_tflop_per_step": fl["total"] * world / 1e12,
            "achieved_tflops_whole_step": fl["total"] * world / (ms_step * 1e-3) / 1e12,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
