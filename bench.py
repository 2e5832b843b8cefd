#!/usr/bin/env python
"""bench.py -- mel-frames/sec of the few-step sampling path (BASELINE.json metric).

A "step" is one full ``EdgeInference.generate_mel`` (context prep + 4 DDIM steps) over one batch
of synthetic semantic tokens: BASELINE config 3 (default CFG, 4-step DDIM, batch 256, 800 mel
frames) per GPU.  Independent utterances are batch-sharded: every rank runs its own 256-utterance
batch (weak scaling), no collective inside the sampling loop, one NCCL all_gather of the final mel
inside the timed region when N > 1.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp32|bf16] [--impl reference]

One JSON line on stdout (rank 0).  Keys beyond the base contract: ``roofline`` (dominant kernel,
timed live with CUDA events by the library's per-class profiler in an eager pass of the same
workload), ``cpu_baseline`` (the oracle port of the reference on the host cores, bounded sample),
``e2e`` (host buffers -> public API -> host buffers), ``clocks``, ``gpu_launches``, ``kernels``.
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
B_PER_GPU, S_TOK, N_STEPS = 256, 400, 4          # BASELINE config 3
T_MEL = 2 * S_TOK
CPU_SAMPLE_B = 8                                  # bounded CPU sample: 8 utterances of the same shape


def log(*a):
    print(*a, file=sys.stderr, flush=True)


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


# ----------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_generate_rate(steps: int, warmup: int, min_seconds: float = 0.0):
    """Times the oracle port of EdgeInference.generate_mel on the host cores, all threads, on a
    bounded sample (CPU_SAMPLE_B utterances of the cfg3 shape).  Returns (frames/s, ms/step, threads)."""
    from oracle import edtts_oracle as O, synth
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth.synth_decoder_state(0)
    tab = O.cosine_schedule(1000)
    idx = synth.synth_sem_idx(3, CPU_SAMPLE_B, S_TOK)
    xT = synth.synth_noise(5, CPU_SAMPLE_B, T_MEL)
    for _ in range(max(warmup, 1)):
        O.generate_mel(sd, tab, idx, N_STEPS, xT)
    t0 = time.perf_counter()
    n = 0
    while n < steps or (time.perf_counter() - t0) < min_seconds:
        O.generate_mel(sd, tab, idx, N_STEPS, xT)
        n += 1
    dt = time.perf_counter() - t0
    return CPU_SAMPLE_B * T_MEL * n / dt, dt / n * 1e3, torch.get_num_threads(), n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, ms, threads, n = cpu_generate_rate(args.steps, args.warmup)
    sample = (f"oracle port of EdgeInference.generate_mel (CPU fp32, torch {torch.__version__}), "
              f"{CPU_SAMPLE_B} utterances x {T_MEL} frames x {N_STEPS} DDIM steps per step, {n} steps")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": n,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "cfg3: default CFG, 4-step DDIM generate_mel, 800 mel frames/utterance",
                   "sample_batch": CPU_SAMPLE_B, "T_mel": T_MEL, "S_tokens": S_TOK, "ddim_steps": N_STEPS},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--precision", default=os.environ.get("EDTTS_BENCH_PRECISION", "auto"))
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="utterances per GPU (default: BASELINE cfg3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    from edge_diffusion_tts_b200.dist import gather_batch
    from oracle import synth

    lib = _lib.load()
    assert lib.edtts_device_supported() == 1, "bench needs an sm_100 (B200) device"
    prec = args.precision
    if prec == "auto":
        prec = "bf16" if lib.edtts_packed_bf16_bytes() > 0 else "fp32"
    B = args.batch
    cfg = E.CFG(device=str(dev))
    dec = E.EdgeDiffusionDecoder(cfg).to(dev).eval()
    dec.load_state_dict(synth.synth_decoder_state(0), strict=True)
    dec.precision = prec
    sched = E.DiffusionSchedule(cfg.diff_steps, device=dev)
    inf = E.EdgeInference(cfg, sched, torch.nn.Identity(), dec, use_cuda_graph=not args.no_graph)

    # synthetic inputs of the BASELINE shape; a different batch per rank
    idx_host = synth.synth_sem_idx(100 + rank, B, S_TOK).pin_memory()
    idx = idx_host.to(dev)
    x_T = synth.synth_noise(200 + rank, B, T_MEL).to(dev)
    frames_per_step = B * T_MEL * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one_step():
        mel = inf.generate_mel(idx, N_STEPS, x_T=x_T)
        if world > 1:
            mel = gather_batch(mel, B * world)
        return mel

    # ---- device-resident throughput ("value") ---------------------------------
    for _ in range(args.warmup):
        one_step()
    c0 = _lib.launch_counts()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        one_step()
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
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
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
        except Exception:
            pass
        # the kernel's real limiter is the MUFU pipe (softmax ex2: 16 ops/clk/SM, measured in tools/ubench/mufu.cu)
        n_exp = B * T_MEL * 4 * 4 * ((129 * T_MEL - 4160) / T_MEL + S_TOK) * N_STEPS + B * T_MEL * 320 * 4 * N_STEPS
        sm_mhz = (clocks.get("sm_mhz") or 1965) * 1e6
        mufu_floor_ms = n_exp / (16 * 148 * sm_mhz) * 1e3
        roofline = {"kernel": dom, "bound": "tensor", "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
                    "mufu": {"ops_per_step": n_exp, "floor_ms_per_step": mufu_floor_ms,
                             "frac_of_mufu_floor": mufu_floor_ms / k["ms_per_step"],
                             "note": "exp2 (softmax) + tanh (SiLU) operations / (16 per clk per SM x 148 SMs x SM clock)"},
                    "frac": achieved / tensor_peak, "traffic": traffic, "peak_source": peak_src,
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
           "note": "pinned host sem_idx -> generate_mel(sem_idx, 4) -> pinned host mel, per rank; every call "
                   "synchronised before the next one starts"}

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

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, ms_cpu, threads, n = cpu_generate_rate(2, 1, min_seconds=10.0)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"oracle port of generate_mel on CPU fp32, {CPU_SAMPLE_B} utterances x {T_MEL} frames x "
                         f"{N_STEPS} steps, {n} repetitions ({ms_cpu:.0f} ms each), torch {torch.__version__}",
               "host_cpus": os.cpu_count()}

    if stdout_fd is not None:
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        os.close(stdout_fd)
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": prec, "data": "synthetic",
            "config": {"workload": "cfg3: default CFG (hidden 160, 4 layers, 4 heads), 4-step DDIM generate_mel, "
                                   "batch 256 per GPU, 800 mel frames (400 semantic tokens)",
                       "batch_per_gpu": B, "global_batch": B * world, "T_mel": T_MEL, "S_tokens": S_TOK,
                       "ddim_steps": N_STEPS, "parallelism": f"batch-sharded x{world}, all_gather of mel" if world > 1
                       else "single GPU", "cuda_graph": not args.no_graph,
                       "l2_policy": "no flush: per-step working set (activations 0.66-0.9 GB) exceeds the 126 MB L2"},
            "e2e": e2e, "gpu_launches": launches, "launches_per_step": per_generate, "clocks": clocks,
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu,
            "algorithmic_tflop_per_step": fl["total"] / 1e12,
            "achieved_tflops_whole_step": fl["total"] / (ms_step * 1e-3) / 1e12 * world,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
