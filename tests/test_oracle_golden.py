"""CPU: the oracle restatement against the golden fixtures recorded from the UNMODIFIED reference
(oracle/make_golden.py).  This is what pins the oracle (the reference ships no tests of its own)."""
import torch

from oracle import edtts_oracle as O
from oracle import synth

TOL = 2e-5   # fixtures were recorded single-threaded; thread count changes fp32 summation order


def test_weights_regenerate_identically(golden):
    g = golden("decoder_step")
    assert g["meta"]["weights"] == synth.state_checksum(synth.synth_decoder_state(0))
    v = golden("vq")
    assert v["meta"]["vq"] == synth.state_checksum(synth.synth_vq_state(0))
    assert v["meta"]["proj"] == synth.state_checksum(synth.synth_proj_state(0))


def test_schedule_tables_bit_equal(golden):
    g = golden("schedule")
    tab = O.cosine_schedule(1000)
    for k, v in g["tables"].items():
        assert torch.equal(tab[k], v), k
    assert O.ddim_timesteps(4) == [(999, 749), (749, 499), (499, 249), (249, 0)]
    assert [t for t, _ in O.ddim_timesteps(4)] == g["steps4"]
    assert O.ddim_timesteps(1) == [(999, 0)]


def test_update_rules_bit_equal(golden):
    g = golden("schedule")
    tab = O.cosine_schedule(1000)
    x = synth.synth_noise(14, 4, 10)
    e = synth.synth_noise(14, 4, 10, tag="eps")
    n = synth.synth_noise(14, 4, 10, tag="noise")
    xp, x0 = O.ddim_step(tab, x, g["t"], g["t_prev"], e, 0.0)
    assert torch.equal(xp, g["ddim_x_prev"]) and torch.equal(x0, g["ddim_x0"])
    assert torch.equal(O.ddpm_step(tab, x, g["t"], e, n), g["ddpm_x_prev"])


def test_decoder_step(golden):
    g = golden("decoder_step")
    sd = synth.synth_decoder_state(0)
    idx = synth.synth_sem_idx(g["seed"], g["B"], g["S"])
    x = synth.synth_noise(g["seed"], g["B"], 2 * g["S"])
    assert int(idx.sum()) == g["idx_sum"] and abs(float(x.double().sum()) - g["x_sum"]) < 1e-6
    for name, c in g["cases"].items():
        eps, hidden = O.decoder_forward(sd, x, c["t"], idx, c["step_idx"], return_hidden=True)
        assert (eps - c["eps"]).abs().max().item() < TOL, name
        got = torch.stack([h[:, g["rows"]] for h in hidden[1:]])
        assert (got - c["hidden_rows"]).abs().max().item() < 1e-4, name
        assert (O.time_condition(sd, c["t"], c["step_idx"]) - c["cond"]).abs().max().item() < TOL


def test_decoder_semantic_features(golden):
    g = golden("decoder_semfeat")
    sd = synth.synth_decoder_state(0)
    eps = O.decoder_forward(sd, synth.synth_noise(12, 2, 80), torch.tensor([700, 20]), None, None,
                            sem_features=synth.synth_features(12, 2, 40, 128))
    assert (eps - g["eps"]).abs().max().item() < TOL


def test_generate_mel_teacher_forced(golden):
    g = golden("generate_mel")
    sd = synth.synth_decoder_state(0)
    tab = O.cosine_schedule(1000)
    idx = synth.synth_sem_idx(g["seed"], g["B"], g["S"])
    for steps in (4, 1):
        for i, tr in enumerate(g["runs"][steps]["trace"]):
            B = g["B"]
            t = torch.full((B,), tr["t"])
            eps = O.decoder_forward(sd, tr["x_t"], t, idx, torch.full((B,), i))
            assert (eps - tr["eps"]).abs().max().item() < TOL
            xp, x0 = O.ddim_step(tab, tr["x_t"], t, torch.full((B,), tr["t_prev"]), tr["eps"])
            assert torch.equal(xp, tr["x_prev"]) and torch.equal(x0, tr["x0"])
    tr16 = g["runs"][16]["trace"]
    assert (tr16[0]["t"], tr16[0]["t_prev"]) == (999, 937) and tr16[1]["t_prev"] == 7


def test_vq_and_proj(golden):
    g = golden("vq")
    z = O.encoder_proj(synth.synth_proj_state(0), synth.synth_features(15, 4, 50))
    assert (z - g["z"]).abs().max().item() < TOL
    cb = synth.synth_vq_state(0)["codebook.weight"]
    z_q, idx, loss, perp, used = O.vq_forward(cb, g["z"])
    assert torch.equal(idx, g["idx"]) and torch.equal(z_q, g["z_q"])
    assert float(loss) == 0.0 and int(used) == int(g["used"]) and abs(float(perp - g["perplexity"])) < 1e-4
    assert torch.equal(O.vq_encode(cb, g["z"]), g["encode"])
    assert torch.equal(O.vq_encode(cb.double(), g["z"].double()), g["encode"])   # fp32 argmin == exact argmin


def test_dsconv(golden):
    g = golden("dsconv")
    for name, c in g["cases"].items():
        cin, cout, k, stride, B, T = c["shape"]
        y = O.dsconv_forward(synth.synth_dsconv_state(16, cin, cout, k), synth.synth_noise(16, B, cin, T, tag="conv_" + name),
                             stride)
        assert y.shape == c["y"].shape and (y - c["y"]).abs().max().item() < TOL, name
