"""Generate tests/golden/*.pt by running the UNMODIFIED reference (build
container only; needs /root/reference).  Run:  python -m oracle.make_golden

Inputs and weights are regenerated from seeds (oracle/synth.py) at test time;
the fixtures hold the reference's OUTPUTS plus checksums of the regenerated
inputs, so they stay small.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os

import torch

from . import ref_harness as R
from . import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
ROWS = [0, 1, 63, 64, 65, 127, 128, 129, 198, 199]   # band-edge rows kept from hidden states


def make_dpm(ref, meta):
    """DPMSolverPP.sample (schedule.py:440-527) through the reference decoder's sem_features path: final sample and the
    per-step (x_t, model_output) trace for orders 1-3, plus the lambda-spaced timesteps."""
    import importlib
    S = importlib.import_module("edge_diffusion_tts.schedule")
    dec, sched = ref["decoder"], ref["schedule"]
    feats = synth.synth_features(21, 2, 20, 128)
    xT = synth.synth_noise(21, 2, 40)
    runs = {}
    for order, steps in ((1, 5), (2, 5), (3, 5), (2, 3), (3, 10)):
        solver = S.DPMSolverPP(sched, order=order)
        trace = []

        def spy(x, t, sem_idx=None, step_idx=None, sem_features=None):
            out = dec(x, t, sem_idx, step_idx, sem_features=sem_features)
            trace.append(dict(x_t=x.clone(), t=int(t[0]), step=int(step_idx[0]), out=out.clone()))
            return out

        with torch.no_grad():
            x, inter = solver.sample(spy, xT, feats, num_steps=steps, return_intermediates=True)
        if steps == 10:
            trace, inter = [trace[0], trace[-1]], [inter[0], inter[-1]]
        elif (order, steps) not in ((2, 5), (3, 5)):      # full traces only where the tests teacher-force
            trace, inter = [], []
        runs[(order, steps)] = dict(x=x, trace=trace, x0=inter, timesteps=solver.get_time_steps(steps, 950))
    ts = {n: S.DPMSolverPP(sched).get_time_steps(n) for n in (1, 4, 10, 20)}
    torch.save(dict(meta=meta, seed=21, B=2, S=20, runs=runs, timesteps_full=ts), os.path.join(OUT, "dpm.pt"))


def make_fsq(ref, meta):
    """FSQ.forward / indices_to_codes of the reference (models/fsq.py) on seeded inputs."""
    import importlib
    F = importlib.import_module("edge_diffusion_tts.models.fsq")
    cases = {}
    for levels in ([8, 8, 8], [8, 6, 5, 5, 5], [5, 5]):
        m = F.FSQ(levels)
        z = torch.randn(3, 41, len(levels), generator=torch.Generator().manual_seed(17 + len(levels))) * 1.5
        z_q, idx = m(z)
        cases[tuple(levels)] = dict(z_q=z_q, idx=idx, codes=m.indices_to_codes(idx), seed=17 + len(levels))
    torch.save(dict(meta=meta, cases=cases), os.path.join(OUT, "fsq.pt"))


def make_griffinlim(ref, meta):
    """torchaudio.transforms.GriffinLim as the reference configures it (generate_sample.py:135-141: n_fft 1024, n_iter 32,
    win_length 1024, hop_length 160, power 2), the library's torch.rand initial phases captured by patching torch.rand for the
    call; plus a short-window / odd-size case."""
    import torchaudio.transforms as T
    cases = {}
    for name, (n_fft, win, hop, n_iter, frames, B) in dict(reference=(1024, 1024, 160, 32, 60, 2), small=(256, 200, 64, 8, 33, 3),
                                                           one_iter=(1024, 1024, 160, 1, 40, 1), no_iter=(512, 512, 128, 0, 20, 1)).items():
        g = torch.Generator().manual_seed(31 + frames)
        spec = torch.rand(B, n_fft // 2 + 1, frames, generator=g) ** 2 * 3.0
        init = torch.rand(B, n_fft // 2 + 1, frames, 2, generator=g)
        init = torch.view_as_complex(init.contiguous())
        tr = T.GriffinLim(n_fft=n_fft, n_iter=n_iter, win_length=win, hop_length=hop, power=2.0)
        real = torch.rand
        torch.rand = lambda *a, **k: init.clone()
        try:
            wave = tr(spec)
        finally:
            torch.rand = real
        cases[name] = dict(cfg=(n_fft, win, hop, n_iter, frames, B), seed=31 + frames, wave=wave)
    torch.save(dict(meta=meta, cases=cases), os.path.join(OUT, "griffinlim.pt"))


def make_fsq_encoder(ref, meta):
    """FSQEncoder.forward / encode / decode of the reference (models/fsq.py:135-222) with synthetic projections, and the
    reference SemanticEncoder's quantiser choice (encoder.py:49-57) recorded as state-dict keys."""
    import importlib
    F = importlib.import_module("edge_diffusion_tts.models.fsq")
    cases = {}
    for levels in ([4, 4, 3, 3, 2, 2, 2, 2], [8, 6, 5, 5, 5]):
        m = F.FSQEncoder(128, levels).eval()
        m.load_state_dict(synth.synth_fsq_encoder_state(23, levels), strict=True)
        z = torch.randn(3, 37, 128, generator=torch.Generator().manual_seed(23 + len(levels)))
        with torch.no_grad():
            z_q, idx, loss, ppl, used = m(z)
            cases[tuple(levels)] = dict(z_q=z_q, idx=idx, loss=loss, perplexity=ppl, used=used, encode=m.encode(z),
                                        decode=m.decode(idx), keys=sorted(m.state_dict()), codebook_size=m.codebook_size)
    torch.save(dict(meta=meta, seed=23, cases=cases), os.path.join(OUT, "fsq_encoder.pt"))


def inpaint_cases():
    """(name, cfg_scale, with known frames) and the seeded inputs shared by the fixture generator and the tests."""
    feats = synth.synth_features(61, 2, 16, 128)
    xc = synth.synth_noise(61, 2, 32)
    known = synth.synth_noise(62, 2, 8)
    noises = [synth.synth_noise(70 + i, 2, 32 if i == 0 else 8, tag=f"n{i}") for i in range(6)]
    return feats, xc, known, noises, (("plain", 1.0, False), ("inpaint", 1.0, True), ("inpaint_cfg", 1.7, True))


def make_inpaint(ref, meta):
    """inpaint_teacher_refine (inference_pipeline.py:145-196), cut out of the unmodified script (ref_harness)."""
    fn = R.reference_inpaint_refine(ref)
    feats, xc, known, noises, cases = inpaint_cases()
    out = {}
    for name, scale, with_known in cases:
        it = iter(noises)
        real = torch.randn_like
        torch.randn_like = lambda a: next(it).clone()
        try:
            with torch.no_grad():
                out[name] = fn(xc, feats, known_mel=known if with_known else None, overlap_len=8 if with_known else 0,
                               strength=0.5, steps=5, cfg_scale=scale)
        finally:
            torch.randn_like = real
    torch.save(dict(meta=meta, strength=0.5, steps=5, overlap=8, x=out), os.path.join(OUT, "inpaint.pt"))


def longform_cases():
    """Seeded inputs shared by the fixture generator and the tests: 4 chunks of 24 frames with an overlap of 6, a mel with
    realistic (log-domain) statistics, and the per-chunk noise draws of a short refine loop."""
    chunk_frames, overlap, n_chunks = 24, 6, 4
    hop = chunk_frames - overlap
    total = hop * (n_chunks - 1) + chunk_frames - 5                       # trimmed inside the last chunk
    g = torch.Generator().manual_seed(81)
    mel = torch.randn(3, 57, 80, generator=g) * torch.linspace(0.5, 2.5, 80) - 4.0 + torch.randn(3, 1, 80, generator=g)
    mel[2, :, 7] = mel[2, 0, 7]                                           # a constant bin: std clamps to 1e-5
    chunks = [synth.synth_noise(90 + i, 1, chunk_frames + (2 if i == 1 else 0), tag=f"c{i}") for i in range(n_chunks)]   # one over-long chunk
    stats = [(torch.randn(1, 1, 80, generator=g) - 5.0, torch.rand(1, 1, 80, generator=g) + 0.5) for _ in range(n_chunks)]
    return dict(chunk_frames=chunk_frames, overlap=overlap, total=total, mel=mel, chunks=chunks, stats=stats)


def make_longform(ref, meta):
    """normalize_mel / denormalize_mel (utils/audio.py, imported) and the chunk plan / window / overlap-add statements of
    inference_pipeline.py:217-393 (cut out of the unmodified script, ref_harness._main_statements)."""
    from edge_diffusion_tts.utils.audio import normalize_mel, denormalize_mel
    c = longform_cases()
    mel_n, mean, std = normalize_mel(c["mel"])
    back = denormalize_mel(mel_n, mean, std)
    plans = {f"{n}@{sr}": R.reference_chunk_plan(n, sr) for n, sr in ((22050 * 7, 22050), (16000 * 5 + 123, 16000), (24000 * 3, 24000))}
    win, final_mel, smooth = R.reference_stitch(ref, c["chunks"], c["stats"], c["chunk_frames"], c["overlap"], c["total"])
    torch.save(dict(meta=meta, mel_n=mel_n, mean=mean, std=std, back=back, plans=plans, window=win, final_mel=final_mel,
                    smooth=smooth), os.path.join(OUT, "longform.pt"))


def invmel_cases():
    """A linear-mel input (exp of a normalised-mel-like signal) and the filter-bank settings of generate_sample.py:125-132
    (inference_pipeline.py:88 leaves f_min / f_max at their defaults, 0 and sample_rate // 2 = 8000: the same bank)."""
    g = torch.Generator().manual_seed(91)
    mel = torch.exp(torch.randn(1, 80, 21, generator=g) * 1.5 - 4.0)
    return mel, dict(generate_sample=dict(n_stft=513, n_mels=80, sample_rate=16000, f_min=0.0, f_max=8000.0, norm=None))


def make_invmel(ref, meta):
    """torchaudio.transforms.InverseMelScale exactly as generate_sample.py:125-132 / inference_pipeline.py:88 construct it."""
    import torchaudio
    import torchaudio.transforms as T
    mel, settings = invmel_cases()
    out = {}
    for name, kw in settings.items():
        m = T.InverseMelScale(**kw)
        out[name] = dict(fb=m.fb.clone(), spec=m(mel))
    torch.save(dict(meta=dict(meta, torchaudio=torchaudio.__version__), cases=out), os.path.join(OUT, "invmel.pt"))


def main(only=None):
    torch.manual_seed(0)
    torch.set_num_threads(1)          # fixed summation order for the recorded outputs
    os.makedirs(OUT, exist_ok=True)
    ref = R.make_reference(0)
    E, dec, sched = ref["E"], ref["decoder"], ref["schedule"]
    sd = synth.synth_decoder_state(0)
    meta = dict(weights=synth.state_checksum(sd), torch=torch.__version__)
    if only == "dpm":                 # fixtures added later are generated alone; the older files stay byte-identical
        make_dpm(ref, meta)
        return
    if only == "fsq":
        make_fsq(ref, meta)
        return
    if only == "griffinlim":
        make_griffinlim(ref, meta)
        return
    if only == "fsq_encoder":
        make_fsq_encoder(ref, meta)
        return
    if only == "inpaint":
        make_inpaint(ref, meta)
        return
    if only == "longform":
        make_longform(ref, meta)
        return
    if only == "invmel":
        make_invmel(ref, meta)
        return

    # --- decoder.forward, one step, mixed t / step_idx, with per-layer hidden rows
    B, S = 2, 100
    idx = synth.synth_sem_idx(11, B, S)
    x = synth.synth_noise(11, B, 2 * S)
    cases = {}
    for name, t, si in (("mixed", [999, 500], [0, 3]), ("nostep", [3, 749], None), ("late", [249, 0], [15, 1])):
        tt = torch.tensor(t)
        ss = None if si is None else torch.tensor(si)
        hidden = []
        hooks = [l.register_forward_hook(lambda m, i, o: hidden.append(o[:, ROWS].clone())) for l in dec.layers]
        with torch.no_grad():
            eps = dec(x, tt, idx, ss)
            cond = dec.time_emb(tt)
            if ss is not None:
                cond = cond + dec.step_emb(ss)
        for h in hooks:
            h.remove()
        cases[name] = dict(t=tt, step_idx=ss, eps=eps, hidden_rows=torch.stack(hidden), cond=cond)
    torch.save(dict(meta=meta, B=B, S=S, seed=11, rows=ROWS, cases=cases,
                    idx_sum=int(idx.sum()), x_sum=float(x.double().sum())),
               os.path.join(OUT, "decoder_step.pt"))

    # --- decoder.forward with sem_features (v-path conditioning, decoder.py:83-85)
    feats = synth.synth_features(12, 2, 40, 128)
    xs = synth.synth_noise(12, 2, 80)
    with torch.no_grad():
        eps_f = dec(xs, torch.tensor([700, 20]), None, None, sem_features=feats)
    torch.save(dict(meta=meta, seed=12, eps=eps_f), os.path.join(OUT, "decoder_semfeat.pt"))

    # --- EdgeInference.generate_mel, 4-step and 1-step, with the per-step trace
    B, S = 1, 72
    idx = synth.synth_sem_idx(13, B, S)
    xT = synth.synth_noise(13, B, 2 * S)
    gm = {}
    for steps in (4, 1, 16):
        trace = []
        real_ddim = sched.get_ddim_step

        def spy(x_t, t, t_prev, eps, eta=0.0):
            xp, x0 = real_ddim(x_t, t, t_prev, eps, eta)
            trace.append(dict(x_t=x_t.clone(), t=int(t[0]), t_prev=int(t_prev[0]), eps=eps.clone(),
                              x_prev=xp.clone(), x0=x0.clone()))
            return xp, x0

        sched.get_ddim_step = spy
        try:
            out = R.reference_generate_mel(ref, idx, steps, xT)
        finally:
            sched.get_ddim_step = real_ddim
        if steps == 16:   # keep only the ends of the long trace
            trace = [trace[0], trace[-1]]
        gm[steps] = dict(x0=out, trace=trace)
    torch.save(dict(meta=meta, B=B, S=S, seed=13, runs=gm), os.path.join(OUT, "generate_mel.pt"))

    # --- DiffusionSchedule tables and the two update rules
    tab = {k: getattr(sched, k).clone() for k in
           ("betas", "alphas", "alpha_bar", "sqrt_alpha_bar", "sqrt_one_minus_alpha_bar", "sqrt_recip_alpha_bar",
            "sqrt_recip_alpha_bar_minus_one", "posterior_variance", "lambda_t")}
    xs = synth.synth_noise(14, 4, 10)
    es = synth.synth_noise(14, 4, 10, tag="eps")
    ns = synth.synth_noise(14, 4, 10, tag="noise")
    t = torch.tensor([999, 749, 1, 0])
    tp = torch.tensor([749, 499, 0, 0])
    xp, x0 = sched.get_ddim_step(xs, t, tp, es, 0.0)
    real = torch.randn_like
    torch.randn_like = lambda a: ns.clone()
    try:
        xd = sched.ddpm_step(xs, t, es)
    finally:
        torch.randn_like = real
    torch.save(dict(meta=meta, tables=tab, seed=14, t=t, t_prev=tp, ddim_x_prev=xp, ddim_x0=x0, ddpm_x_prev=xd,
                    q_sample=sched.q_sample(xs, t, ns)[0], v_target=sched.get_v_target(xs, ns, t),
                    x0_from_eps=sched.predict_x0_from_eps(xs, t, es), x0_from_v=sched.predict_x0_from_v(xs, t, es),
                    eps_from_v=sched.predict_eps_from_v(xs, t, es), steps4=sched.get_schedule_for_steps(4)),
               os.path.join(OUT, "schedule.pt"))

    # --- SemanticEncoder.proj + VectorQuantizer (forward 5-tuple, encode, decode)
    h = synth.synth_features(15, 4, 50)
    with torch.no_grad():
        z = ref["proj"](h)
        z_q, vidx, vq_loss, perp, used = ref["vq"](z)
        enc = ref["vq"].encode(z)
        decd = ref["vq"].decode(vidx[:1, :5])
    torch.save(dict(meta=dict(meta, vq=synth.state_checksum(synth.synth_vq_state(0)),
                              proj=synth.state_checksum(synth.synth_proj_state(0))),
                    seed=15, z=z, z_q=z_q, idx=vidx, vq_loss=vq_loss, perplexity=perp, used=used,
                    encode=enc, decode=decd), os.path.join(OUT, "vq.pt"))

    # --- DepthwiseSeparableConv (operator level, F2)
    conv = {}
    for name, (cin, cout, k, stride, B, T) in dict(default=(160, 160, 3, 1, 2, 50), odd=(80, 96, 5, 2, 3, 37),
                                                   tiny=(4, 4, 3, 1, 1, 9)).items():
        m = E.layers.DepthwiseSeparableConv(cin, cout, k, stride).eval()
        m.load_state_dict(synth.synth_dsconv_state(16, cin, cout, k), strict=True)
        xin = synth.synth_noise(16, B, cin, T, tag="conv_" + name)
        with torch.no_grad():
            conv[name] = dict(shape=(cin, cout, k, stride, B, T), y=m(xin))
    torch.save(dict(meta=meta, seed=16, cases=conv), os.path.join(OUT, "dsconv.pt"))

    make_dpm(ref, meta)
    make_fsq(ref, meta)
    make_fsq_encoder(ref, meta)
    make_griffinlim(ref, meta)
    make_inpaint(ref, meta)
    make_longform(ref, meta)
    make_invmel(ref, meta)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    import sys
    main(sys.argv[1] if len(sys.argv) > 1 else None)
