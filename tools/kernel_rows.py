"""Every NON-fused kernel class of the path at its BASELINE size, one call each (for `ncu`) or timed with CUDA events.

    python tools/kernel_rows.py              # CUDA-event timing, prints a table with achieved GB/s (algorithmic bytes)
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,\
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/rows.csv \
        python tools/kernel_rows.py --once
    python tools/kernel_rows.py --summarise gpurun_out/rows.csv   # ncu csv -> per-kernel rows (achieved vs 6547.8 GB/s)

Algorithmic bytes per call (SURVEY 8d "per memory-bound kernel"): read every input once, write every output once.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
HBM_PEAK = 6547.8
try:
    HBM_PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def cases():
    """[(name, callable, algorithmic bytes, algorithmic flops)]"""
    import torch

    import __graft_entry__ as ge
    ge.build()
    import edge_diffusion_tts_b200 as E
    from oracle import synth
    dev = "cuda:0"
    f = dict(device=dev, dtype=torch.float32)
    g = torch.Generator(dev).manual_seed(0)
    out = []
    # ---- VQ + encoder projection (cfg2: 25,600 rows; cfg5: 192,000 rows)
    enc = E.SemanticEncoder(E.CFG(device=dev, use_fsq=False), load_hubert=False).to(dev).eval()
    enc.proj.load_state_dict(synth.synth_proj_state(0))
    enc.vq.load_state_dict(synth.synth_vq_state(0))
    for rows_b, rows_s, tag in ((64, 400, "cfg2"), (128, 1500, "cfg5")):
        h = torch.randn(rows_b, rows_s, 768, generator=g, **f)
        z = enc.project(h)
        R = rows_b * rows_s
        out.append((f"encoder_proj {tag} ({R} rows)", lambda h=h: enc.project(h), R * (768 + 128) * 4, R * (768 * 128 + 128 * 128) * 2))
        out.append((f"vq_argmin {tag} ({R} rows)", lambda z=z: enc.vq.encode(z), R * 128 * 4 + 512 * 128 * 4 + R * 8, R * 512 * 128 * 2))
    fq = E.SemanticEncoder(E.CFG(device=dev), load_hubert=False).to(dev).eval()
    z = torch.randn(64, 400, 128, generator=g, **f)
    out.append(("fsq_encoder cfg2 (25600 rows)", lambda: fq.vq(z), 25600 * (128 * 2 * 4 + 8), 25600 * 128 * 8 * 4))
    # ---- stand-alone update rules at the cfg3 shape
    B, T, M = 256, 800, 80
    n = B * T * M * 4
    sch = E.DiffusionSchedule(1000, device=dev)
    x, e, nz = (torch.randn(B, T, M, generator=g, **f) for _ in range(3))
    t = torch.full((B,), 749, dtype=torch.long, device=dev)
    tp = torch.full((B,), 499, dtype=torch.long, device=dev)
    out.append(("ddim_step [256,800,80]", lambda: sch.get_ddim_step(x, t, tp, e), 4 * n, 0))
    out.append(("ddpm_step [256,800,80]", lambda: sch.ddpm_step(x, t, e, noise=nz), 4 * n, 0))
    sol = E.DPMSolverPP(sch, order=2)
    out.append(("dpm second_order_update [256,800,80]", lambda: sol.second_order_update(x, e, nz, t, tp, t + 100), 4 * n, 0))
    from edge_diffusion_tts_b200 import _lib
    lib = _lib.load()
    co = torch.rand(B, 4, **f)
    xo = torch.empty_like(x)
    out.append(("vddim_step [256,800,80]", lambda: _lib.check(lib.edtts_vddim_step(_lib.ptr(x), _lib.ptr(e), None, 1.0, _lib.ptr(co), _lib.ptr(xo),
                                                                                  None, B, T * M, _lib.stream_ptr(dev))), 3 * n, 0))
    out.append(("inpaint_inject [256,800,80]", lambda: _lib.check(lib.edtts_inpaint_inject(_lib.ptr(xo), _lib.ptr(x), _lib.ptr(nz), _lib.ptr(co), B, T, T,
                                                                                         M, _lib.stream_ptr(dev))), 3 * n, 0))
    # ---- mel statistics / stitch / inverse mel
    mel = torch.randn(B, T, M, generator=g, **f)
    out.append(("normalize_mel [256,800,80]", lambda: E.normalize_mel(mel), 3 * n, 0))
    mn, sd = torch.zeros(B, 1, M, **f), torch.ones(B, 1, M, **f)
    out.append(("denormalize_mel [256,800,80]", lambda: E.denormalize_mel(mel, mn, sd), 2 * n, 0))
    st = E.MelStitcher(M, 4000, 800, 200, dev, batch=16)
    xs = torch.randn(16, 800, M, generator=g, **f)
    m16, s16 = torch.zeros(16, 1, M, **f), torch.ones(16, 1, M, **f)
    out.append(("stitch_add 16 x [800,80]", lambda: st.add_chunk(0, xs, m16, s16), 16 * 800 * M * 4 * 3, 0))
    out.append(("stitch_finalize 16 x [80,4000]", lambda: st.finalize(3800), 16 * M * 4000 * 4 * 3, 0))
    inv = E.InverseMelScale(n_stft=513, n_mels=M).to(dev)
    m8 = torch.rand(8, M, 800, generator=g, **f)
    out.append(("inverse_mel 8 x [80,800] -> [513,800]", lambda: inv(m8), 8 * 800 * (M + 513) * 4, 8 * 800 * 513 * M * 2))
    # ---- depthwise-separable convolution, operator level
    conv = E.DepthwiseSeparableConv(160, 160, 3, 1).to(dev).eval()
    conv.load_state_dict(synth.synth_dsconv_state(16, 160, 160, 3))
    xc = torch.randn(256, 160, 800, generator=g, **f)
    out.append(("dsconv [256,160,800]", lambda: conv(xc), 2 * 256 * 160 * 800 * 4, 256 * 800 * (160 * 160 + 160 * 3) * 2))
    # ---- conditioning / context of cfg3
    dec = E.EdgeDiffusionDecoder(E.CFG(device=dev)).to(dev).eval()
    dec.load_state_dict(synth.synth_decoder_state(0))
    dec.precision = "bf16"
    idx = synth.synth_sem_idx(1, 256, 400).to(dev)
    out.append(("context_prepare cfg3 (102,400 tokens, 4 layers)", lambda: dec.prepare_context(idx, None, 800), 256 * 400 * (8 + 4 * 320 * 2), 0))
    tt = torch.full((4,), 999, dtype=torch.long, device=dev)
    out.append(("cond_prepare (4 rows)", lambda: dec.prepare_cond(tt, torch.zeros_like(tt), 800, 400), 1.6e6, 0))
    return out


def main():
    import torch
    once = "--once" in sys.argv
    cs = cases()
    if "--only" in sys.argv:
        pat = sys.argv[sys.argv.index("--only") + 1]
        cs = [c for c in cs if pat in c[0]]
    torch.cuda.synchronize()
    rows = []
    for name, fn, nbytes, flops in cs:
        fn()
        torch.cuda.synchronize()
        if once:
            torch.cuda.nvtx.range_push(name)
            fn()
            torch.cuda.synchronize()
            torch.cuda.nvtx.range_pop()
            print("ROW", name)
            continue
        ts = []
        for _ in range(7):
            flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0").fill_(1)   # evict L2 between repetitions
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
            del flush
        ms = sorted(ts)[len(ts) // 2]
        gbs = nbytes / (ms * 1e-3) / 1e9
        rows.append(dict(kernel=name, ms=ms, algorithmic_bytes=nbytes, achieved_gbs=gbs, frac_of_hbm_peak=gbs / HBM_PEAK,
                         tflops=flops / (ms * 1e-3) / 1e12 if flops else None))
        print(f"{name:52s} {ms * 1e3:9.1f} us  {gbs:8.1f} GB/s ({100 * gbs / HBM_PEAK:5.1f} % of {HBM_PEAK:.0f})"
              + (f"  {flops / (ms * 1e-3) / 1e12:6.2f} TFLOP/s" if flops else ""))
    if not once:
        print("JSON", json.dumps(rows))


def summarise(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iid = hdr.index("ID")
    by = {}
    for r in rows[1:]:
        by.setdefault((int(r[iid]), r[ik]), {})[r[im]] = float(r[iv].replace(",", ""))
    print(f"{'kernel':60s} {'us':>9s} {'dram MB':>9s} {'GB/s':>8s} {'% HBM':>6s} {'fma %':>6s} {'xu %':>6s}")
    for (i, k), m in sorted(by.items()):
        ns = m.get("gpu__time_duration.sum", 0.0)
        by_ = m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        gbs = by_ / ns if ns else 0.0
        print(f"{k[:60]:60s} {ns / 1e3:9.1f} {by_ / 1e6:9.2f} {gbs:8.1f} {100 * gbs / HBM_PEAK:6.1f} "
              f"{m.get('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 0):6.1f} "
              f"{m.get('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 0):6.1f}")


if __name__ == "__main__":
    if "--summarise" in sys.argv:
        summarise(sys.argv[sys.argv.index("--summarise") + 1])
    else:
        main()
