"""GPU parity at the BASELINE.json configurations themselves (run on the B200 box: pytest -m gpu).

  * cfg2 / cfg5 chain from RAW 768-d features: proj -> VQ indices against the fp32 oracle; a row may differ only where the
    GPU projection's rounding (<= 2e-5 vs ATen's) flips a proven near-tie, and every such row is checked against that bound
  * teacher-forced eps at FULL cfg3 (B = 256, T = 800) and cfg5-share (B = 16, T = 3000) size, fp32 (max-abs <= 1e-4) and
    bf16 (rel-L2 <= 1e-2), the oracle evaluating 4 sampled utterances (batch invariance: an utterance's bits do not depend on
    the rest of the batch, test_batch_invariance)
  * free-running generate_mel with the F9-aware criterion of SURVEY 8(c)(iv) (oracle/parity.py)
  * generate_from_audio through a stub encoder; CUDA-graph replay after load_state_dict; two decoders on one solver
  * the merged launch on a grid smaller than the SM count and with a second stream holding SMs
"""
import os
import subprocess
import sys

import pytest
import torch

from oracle import edtts_oracle as O
from oracle import parity, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def env(lib):
    import edge_diffusion_tts_b200 as E
    assert torch.cuda.is_available() and lib.edtts_device_supported() == 1, "needs an sm_100 device"
    cfg = E.CFG(device=DEV)
    sd = synth.synth_decoder_state(0)
    dec = E.EdgeDiffusionDecoder(cfg).to(DEV).eval()
    dec.load_state_dict(sd, strict=True)
    sched = E.DiffusionSchedule(cfg.diff_steps, device=DEV)
    inf = E.EdgeInference(cfg, sched, torch.nn.Identity(), dec)
    return dict(E=E, cfg=cfg, sd=sd, dec=dec, sched=sched, inf=inf, tab=O.cosine_schedule(cfg.diff_steps), lib=lib)


# ------------------------------------------------------------------ (a) raw features -> proj -> VQ
@pytest.mark.parametrize("B,S", [(64, 400), (128, 1500)])
def test_raw_features_to_vq_indices(env, B, S):
    """north_star cfg2 / cfg5: 768-d features -> SemanticEncoder.proj -> VectorQuantizer indices, end to end on the GPU,
    against the fp32 oracle chain on the CPU.  d_a(z') - d_b(z') = d_a(z) - d_b(z) - 2 (z' - z).(e_a - e_b): a row whose
    winner differs must have an exact (fp64) top-2 gap at the oracle's z no larger than 2 |z' - z| |e_a - e_b| plus the fp32
    rounding of the distance itself; anything else is a kernel error."""
    E = env["E"]
    enc = E.SemanticEncoder(E.CFG(device=DEV, use_fsq=False), load_hubert=False).to(DEV).eval()
    enc.proj.load_state_dict(synth.synth_proj_state(0))
    enc.vq.load_state_dict(synth.synth_vq_state(0))
    cb = synth.synth_vq_state(0)["codebook.weight"]
    h = synth.synth_features(31, B, S)
    z_ref = O.encoder_proj(synth.synth_proj_state(0), h)
    z_gpu = enc.project(h.to(DEV)).cpu()
    idx = enc.encode_features(h.to(DEV)).cpu()
    assert idx.shape == (B, S) and idx.dtype == torch.int64
    dz = (z_gpu - z_ref).abs().max().item()
    assert dz <= 2e-5, dz
    ref = torch.cat([O.vq_encode(cb, z_ref[i:i + 8]) for i in range(0, B, 8)])
    bad = (idx != ref).nonzero()
    print(f"[cfg rows {B * S}] proj max|dz| = {dz:.2e}; rows whose index differs from the fp32 oracle chain: {len(bad)}")
    # the GPU's answer is the exact argmin for the z the GPU computed (fp64 distances on its own z), everywhere
    ex = torch.cat([O.vq_encode(cb.double(), z_gpu[i:i + 8].double()) for i in range(0, B, 8)])
    assert torch.equal(idx, ex), f"{(idx != ex).sum().item()} rows are not the exact argmin of the GPU's own z"
    for b, s in bad.tolist():
        zr, zg = z_ref[b, s].double(), z_gpu[b, s].double()
        ea, eb = cb[idx[b, s]].double(), cb[ref[b, s]].double()
        gap = abs(((zr - ea) ** 2).sum() - ((zr - eb) ** 2).sum()).item()
        bound = 2 * (zg - zr).norm().item() * (ea - eb).norm().item() + 1e-4      # + fp32 rounding of a ~130-magnitude distance
        assert gap <= bound, (b, s, gap, bound)
    assert len(bad) <= max(2, B * S // 20000), len(bad)


# ------------------------------------------------------------------ (b) full BASELINE sizes, teacher-forced
@pytest.mark.parametrize("B,S,prec", [(256, 400, "bf16"), (256, 400, "fp32"), (16, 1500, "bf16"), (16, 1500, "fp32")])
def test_eps_at_baseline_size(env, B, S, prec):
    """cfg3 (256 x 800: 1792 tiles per layer, 12 rounds per CTA of the merged launch) and one GPU's share of cfg5
    (16 x 3000): the whole batch through the GPU, 4 sampled utterances through the oracle."""
    dec = env["dec"]
    T = 2 * S
    idx = synth.synth_sem_idx(41, B, S)
    x = synth.synth_noise(41, B, T)
    g = torch.Generator().manual_seed(B + S)
    t = torch.randint(0, 1000, (B,), generator=g)
    si = torch.randint(0, 4, (B,), generator=g)
    rows = sorted({0, B - 1, B // 2 + 1, int(torch.randint(0, B, (1,), generator=g))})
    keep = dec.precision
    dec.precision = prec
    try:
        eps = dec(x.to(DEV), t.to(DEV), idx.to(DEV), si.to(DEV)).cpu()
    finally:
        dec.precision = keep
    assert not torch.isnan(eps).any()
    ref = O.decoder_forward(env["sd"], x[rows], t[rows], idx[rows], si[rows])
    if prec == "fp32":
        d = (eps[rows] - ref).abs().max().item()
        print(f"[B={B} T={T} fp32] max|d eps| on utterances {rows}: {d:.2e}")
        assert d <= 1e-4, d
    else:
        r = rel_l2(eps[rows], ref)
        print(f"[B={B} T={T} bf16] rel-L2(eps) on utterances {rows}: {r:.2e}")
        assert r <= 1e-2, r


# ------------------------------------------------------------------ (c) free-running, F9-aware
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_generate_mel_free_running_f9_criterion(env, prec):
    """at step 0 every element whose clamped x0 differs by more than tol lies in the clamp-edge band; the
    utterances without such an element agree with the oracle to tol at the end of the 4 steps (fp32: max-abs 1e-4).  64 short
    utterances (80 frames) so that clean utterances exist: a band element is expected once per ~27,000 elements in fp32."""
    dec, inf, tab = env["dec"], env["inf"], env["tab"]
    B, S = 64, 40
    idx, xT = synth.synth_sem_idx(51, B, S), synth.synth_noise(52, B, 2 * S)
    trace = []
    ref = O.generate_mel(env["sd"], tab, idx, 4, xT, trace=trace)
    keep = dec.precision
    dec.precision = prec
    try:
        t = torch.full((B,), 999, dtype=torch.long, device=DEV)
        eps0 = dec(xT.to(DEV), t, idx.to(DEV), torch.zeros_like(t)).cpu()
        x0_0 = inf.generate_mel(idx.to(DEV), 1, x_T=xT.to(DEV)).cpu()
        out = inf.generate_mel(idx.to(DEV), 4, x_T=xT.to(DEV)).cpu()
    finally:
        dec.precision = keep
    rep = parity.f9_report(float(tab["alpha_bar"][999]), xT, trace[0][1], eps0, trace[0][3], x0_0, ref, out, 1e-4)
    print(f"[free-running {prec}] " + ", ".join(f"{k}={v:.3e}" if isinstance(v, float) else f"{k}={v}" for k, v in rep.items()))
    assert rep["step0_over_tol_outside_band"] == 0, rep
    if prec == "fp32":
        assert rep["clean_utterances"] >= B // 2, rep
        assert rep["clean_max_abs"] <= 1e-4, rep
        assert rep["rel_l2"] <= 1e-3 or rep["step0_over_tol"] > 0, rep
    else:
        assert rep["rel_l2"] <= 0.15, rep                                  # every utterance descends from band elements


# ------------------------------------------------------------------ (d) generate_from_audio
class _StubEncoder(torch.nn.Module):
    """Returns the 5-tuple of SemanticEncoder.forward (encoder.py:74-100): (z_q, idx, loss, perplexity, used)."""

    def __init__(self, idx):
        super().__init__()
        self.idx = idx
        self.calls = []

    def forward(self, wav):
        self.calls.append(tuple(wav.shape))
        B = wav.shape[0]
        z = torch.zeros(B, self.idx.shape[1], 128, device=wav.device)
        return z, self.idx[:B].to(wav.device), torch.tensor(0.0), torch.tensor(1.0), torch.tensor(1)


def test_generate_from_audio_stub_encoder(env):
    """inference.py:55-62: a 1-D wav gains a batch dimension, moves to the device, the encoder's 2nd output feeds generate_mel."""
    E = env["E"]
    idx = synth.synth_sem_idx(61, 2, 25)
    enc = _StubEncoder(idx)
    inf = E.EdgeInference(env["cfg"], env["sched"], enc, env["dec"])
    torch.manual_seed(7)
    a = inf.generate_from_audio(torch.zeros(16000), num_steps=4)
    assert enc.calls == [(1, 16000)] and a.shape == (1, 50, 80) and a.device.type == "cuda"
    torch.manual_seed(7)
    b = inf.generate_mel(idx[:1].to(DEV), 4)
    assert torch.equal(a, b)
    c = inf.generate_from_audio(torch.zeros(2, 8000), num_steps=1)
    assert c.shape == (2, 50, 80) and enc.calls[-1] == (2, 8000) and not enc.training


# ------------------------------------------------------------------ ADVICE: stale graphs
def test_graph_replay_sees_new_weights(env):
    """A captured graph holds pointers into the packed bf16 weight image of its epoch: after load_state_dict the next
    generate_mel must repack and re-capture (EdgeInference, inpaint_refine and DPMSolverPP all check weights_token)."""
    E = env["E"]
    sd1, sd2 = synth.synth_decoder_state(0), synth.synth_decoder_state(5)
    idx = synth.synth_sem_idx(71, 2, 30).to(DEV)
    xT = synth.synth_noise(71, 2, 60).to(DEV)
    feats = synth.synth_features(71, 2, 30, 128).to(DEV)

    def fresh(sd):
        d = E.EdgeDiffusionDecoder(env["cfg"]).to(DEV).eval()
        d.load_state_dict(sd, strict=True)
        d.precision = "bf16"
        return d

    dec = fresh(sd1)
    inf = E.EdgeInference(env["cfg"], env["sched"], torch.nn.Identity(), dec)
    solver = E.DPMSolverPP(env["sched"], order=2)
    a1 = inf.generate_mel(idx, 4, x_T=xT)
    r1 = inf.inpaint_refine(xT, feats, steps=3, noise=xT * 0.5)
    s1 = solver.sample(dec, xT, feats, num_steps=3)
    assert torch.equal(a1, inf.generate_mel(idx, 4, x_T=xT))                # replay
    dec.load_state_dict(sd2, strict=True)
    a2 = inf.generate_mel(idx, 4, x_T=xT)
    r2 = inf.inpaint_refine(xT, feats, steps=3, noise=xT * 0.5)
    s2 = solver.sample(dec, xT, feats, num_steps=3)
    d2 = fresh(sd2)
    inf2 = E.EdgeInference(env["cfg"], env["sched"], torch.nn.Identity(), d2)
    assert torch.equal(a2, inf2.generate_mel(idx, 4, x_T=xT)) and not torch.equal(a1, a2)
    assert torch.equal(r2, inf2.inpaint_refine(xT, feats, steps=3, noise=xT * 0.5)) and not torch.equal(r1, r2)
    assert torch.equal(s2, E.DPMSolverPP(env["sched"], order=2).sample(d2, xT, feats, num_steps=3)) and not torch.equal(s1, s2)
    with torch.no_grad():                                                   # an in-place parameter update
        dec.out_proj.bias.add_(0.25)
        d2.out_proj.bias.add_(0.25)
    assert torch.equal(inf.generate_mel(idx, 4, x_T=xT), inf2.generate_mel(idx, 4, x_T=xT))


def test_two_decoders_share_one_solver(env):
    """The reference pipeline keeps a teacher and a student decoder on one schedule / solver: same shapes, same epoch
    number, different weights -- the second must not replay the first one's graph."""
    E = env["E"]
    decs = []
    for seed in (0, 9):
        d = E.EdgeDiffusionDecoder(env["cfg"]).to(DEV).eval()
        d.load_state_dict(synth.synth_decoder_state(seed), strict=True)
        d.precision = "bf16"
        decs.append(d)
    xT = synth.synth_noise(81, 2, 60).to(DEV)
    feats = synth.synth_features(81, 2, 30, 128).to(DEV)
    shared = E.DPMSolverPP(env["sched"], order=2)
    got = [shared.sample(d, xT, feats, num_steps=4) for d in decs]
    got2 = [shared.sample(d, xT, feats, num_steps=4) for d in decs]
    want = [E.DPMSolverPP(env["sched"], order=2).sample(d, xT, feats, num_steps=4) for d in decs]
    for g, g2, w in zip(got, got2, want):
        assert torch.equal(g, w) and torch.equal(g2, w)
    assert not torch.equal(got[0], got[1])
    assert decs[0].weights_token() != decs[1].weights_token()


# ------------------------------------------------------------------ merged launch: grid size / co-residency
_CHILD = r"""
import sys, torch
sys.path.insert(0, {root!r})
import __graft_entry__ as ge
ge.build()
import edge_diffusion_tts_b200 as E
from oracle import synth
dev = "cuda:0"
cfg = E.CFG(device=dev)
dec = E.EdgeDiffusionDecoder(cfg).to(dev).eval()
dec.load_state_dict(synth.synth_decoder_state(0), strict=True)
dec.precision = "bf16"
inf = E.EdgeInference(cfg, E.DiffusionSchedule(cfg.diff_steps, device=dev), torch.nn.Identity(), dec)
B, S = 12, 200
out = inf.generate_mel(synth.synth_sem_idx(91, B, S).to(dev), 4, x_T=synth.synth_noise(91, B, 2 * S).to(dev))
torch.save(out.cpu(), sys.argv[1])
"""


def test_merged_launch_on_a_smaller_grid(env, tmp_path):
    """EDTTS_MAX_CTAS caps the persistent grid (as an MPS / green-context SM limit would): 37 and 5 CTAs for 5 x 48 items
    give the bits of the full-size grid -- the per-item flags order any static round-robin deal without deadlock."""
    dec, inf = env["dec"], env["inf"]
    B, S = 12, 200
    keep = dec.precision
    dec.precision = "bf16"
    try:
        want = inf.generate_mel(synth.synth_sem_idx(91, B, S).to(DEV), 4, x_T=synth.synth_noise(91, B, 2 * S).to(DEV)).cpu()
    finally:
        dec.precision = keep
    for cap in ("37", "5"):
        f = str(tmp_path / f"out_{cap}.pt")
        e = dict(os.environ, EDTTS_MAX_CTAS=cap)
        r = subprocess.run([sys.executable, "-c", _CHILD.format(root=ROOT), f], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        assert torch.equal(torch.load(f), want), cap


def test_merged_launch_with_a_second_stream_holding_sms(env):
    """Two decoders of one process on two streams (the advisor's scenario): each merged launch needs every SM for itself;
    launched cooperatively, the driver runs each only when its whole grid is resident, so neither can starve the other's
    flag waits.  Results equal the serial ones."""
    E = env["E"]
    decs, infs = [], []
    for seed in (0, 9):
        d = E.EdgeDiffusionDecoder(env["cfg"]).to(DEV).eval()
        d.load_state_dict(synth.synth_decoder_state(seed), strict=True)
        d.precision = "bf16"
        decs.append(d)
        infs.append(E.EdgeInference(env["cfg"], env["sched"], torch.nn.Identity(), d, use_cuda_graph=False))
    B, S = 40, 400
    idx = synth.synth_sem_idx(95, B, S).to(DEV)
    xT = synth.synth_noise(95, B, 2 * S).to(DEV)
    serial = [i.generate_mel(idx, 4, x_T=xT) for i in infs]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(DEV), torch.cuda.Stream(DEV)]
    outs = [[], []]
    for rep in range(3):
        for k in (0, 1):
            with torch.cuda.stream(streams[k]):
                outs[k].append(infs[k].generate_mel(idx, 4, x_T=xT))
    torch.cuda.synchronize()
    for k in (0, 1):
        for o in outs[k]:
            assert torch.equal(o, serial[k])
