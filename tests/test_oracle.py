"""Pins the CPU oracle against golden vectors produced by the real reference
(tests/golden/make_golden.py).  No access to /root/reference here."""
import os

import numpy as np
import pytest
import torch

import hvae_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = {"tiny_ukbb": 3, "tiny_morphomnist": 3, "tiny_cmnist": 3, "morphomnist": 2, "cmnist": 2,
         "ukbb192": 1, "mimic192": 1, "mimic224": 1}


def sub(t, n=4096):
    flat = t.detach().reshape(-1)
    if flat.numel() <= n:
        return flat.numpy()
    return flat[np.linspace(0, flat.numel() - 1, n).astype(np.int64)].numpy()


def setup(name):
    cfg = O.make_cfg(name)
    sd = O.seeded_state_dict(cfg, seed=7)
    x8, pa, cf = O.synthetic_batch(cfg, CASES[name], seed=11)
    x = O.normalise_x(x8)
    return cfg, sd, x, O.expand_parents(pa, cfg.input_res), O.expand_parents(cf, cfg.input_res)


@pytest.mark.parametrize("name", list(CASES))
def test_elbo_and_grads_match_reference(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cfg, sd, x, pa, _ = setup(name)
    for p in sd.values():
        p.requires_grad_(True)
    out = O.hvae_forward(sd, cfg, x, pa, O.NoiseTape(seed=101), beta=cfg.beta, detail=True)
    np.testing.assert_allclose(out["elbo"].item(), g["elbo"], rtol=2e-5)
    np.testing.assert_allclose(out["nll"].item(), g["nll"], rtol=2e-5)
    np.testing.assert_allclose(out["kl"].item(), g["kl"], rtol=2e-5)
    np.testing.assert_allclose(out["block_kl"].detach().numpy(), g["block_kl"], rtol=2e-4, atol=1e-4)
    np.testing.assert_allclose(sub(out["h"]), g["h_sub"], rtol=1e-3, atol=2e-4)
    out["elbo"].backward()
    names = list(g["grad_names"])
    assert names == list(sd.keys())
    mine = np.array([float(sd[n].grad.norm()) if sd[n].grad is not None else -1.0 for n in names])
    np.testing.assert_allclose(mine, g["grad_norm"], rtol=2e-3, atol=1e-6)
    for k in g.files:
        if k.startswith("grad::"):
            np.testing.assert_allclose(sd[k[6:]].grad.reshape(-1)[:64].numpy(), g[k], rtol=2e-3, atol=1e-5)


@pytest.mark.parametrize("name", ["tiny_ukbb", "tiny_morphomnist"])
def test_free_bits_elbo_and_grads_match_reference(name):
    """kl_free_bits > 0 (src/vae.py:443-449): floor = midpoint of the per-channel batch-mean KLs, so half of the channels are
    gated (tests/golden/make_golden_freebits.py ran the real reference)."""
    g = np.load(os.path.join(GOLD, f"freebits_{name}.npz"))
    cfg, sd, x, pa, _ = setup(name)
    cfg.kl_free_bits = float(g["free_bits"])
    for p in sd.values():
        p.requires_grad_(True)
    out = O.hvae_forward(sd, cfg, x, pa, O.NoiseTape(seed=101), beta=cfg.beta)
    for k in ("elbo", "nll", "kl"):
        np.testing.assert_allclose(out[k].item(), g[k], rtol=2e-5)
    out["elbo"].backward()
    assert list(g["grad_names"]) == list(sd.keys())
    mine = np.array([float(sd[n].grad.norm()) if sd[n].grad is not None else -1.0 for n in sd])
    np.testing.assert_allclose(mine, g["grad_norm"], rtol=2e-3, atol=1e-6)
    for k in g.files:
        if k.startswith("grad::"):
            np.testing.assert_allclose(sd[k[6:]].grad.reshape(-1)[:64].numpy(), g[k], rtol=2e-3, atol=1e-5)


@pytest.mark.parametrize("name", list(CASES))
def test_counterfactual_path_matches_reference(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cfg, sd, x, pa, cf = setup(name)
    with torch.no_grad():
        zs = O.hvae_abduct(sd, cfg, x, pa, O.NoiseTape(seed=202), t=0.9)
        zs = [z["z"] for z in zs] if cfg.cond_prior else zs
        st = np.array([[float(z.mean()), float(z.std())] for z in zs])
        np.testing.assert_allclose(st, g["z_stats"], rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(sub(zs[-1], 1024), g["z_last_sub"], rtol=1e-3, atol=1e-3)
        cf_loc, cf_scale = O.hvae_forward_latents(sd, cfg, zs, cf)
        rec_loc, rec_scale = O.hvae_forward_latents(sd, cfg, zs, pa)
        u = (x - rec_loc) / rec_scale.clamp(min=1e-12)
        cf_x = torch.clamp(cf_loc + cf_scale * u, -1, 1)
        np.testing.assert_allclose(sub(rec_loc), g["rec_loc_sub"], atol=5e-4)
        np.testing.assert_allclose(sub(rec_scale), g["rec_scale_sub"], rtol=2e-3, atol=1e-6)
        np.testing.assert_allclose(sub(cf_x), g["cf_x_sub"], atol=2e-3)
        half = zs[: len(zs) // 2]
        pl, _ = O.hvae_forward_latents(sd, cfg, half, pa, O.NoiseTape(seed=303), t=0.7)
        np.testing.assert_allclose(sub(pl), g["partial_loc_sub"], atol=5e-4)
        if cfg.cond_prior:
            cz = O.hvae_abduct(sd, cfg, x, pa, O.NoiseTape(seed=404), cf_parents=cf, alpha=0.65, t=0.8)
            st = np.array([[float(z.mean()), float(z.std())] for z in cz])
            np.testing.assert_allclose(st, g["cfz_stats"], rtol=1e-3, atol=1e-4)
            np.testing.assert_allclose(sub(cz[-1], 1024), g["cfz_last_sub"], rtol=1e-3, atol=1e-3)
        sx, ss = O.hvae_sample(sd, cfg, pa, O.NoiseTape(seed=505), t=0.5)
        np.testing.assert_allclose(sub(sx), g["sample_sub"], atol=5e-4)
        np.testing.assert_allclose(sub(ss), g["sample_scale_sub"], rtol=2e-3, atol=1e-6)


def test_counterfactual_helper_equals_inline_lines():
    cfg, sd, x, pa, cf = setup("tiny_ukbb")
    g = np.load(os.path.join(GOLD, "tiny_ukbb.npz"))
    with torch.no_grad():
        cf_x, var = O.counterfactual(sd, cfg, x, pa, cf, O.NoiseTape(seed=202), t_abduct=0.9)
    assert var is None
    np.testing.assert_allclose(sub(cf_x), g["cf_x_sub"], atol=2e-3)


def test_conditioning_dropout_morphomnist():
    for name in ["tiny_morphomnist", "morphomnist"]:
        g = np.load(os.path.join(GOLD, name + ".npz"))
        cfg, sd, x, pa, _ = setup(name)
        with torch.no_grad():
            for i, drop in enumerate([(0, 1), (1, 0)]):
                out = O.hvae_forward(sd, cfg, x, pa, O.NoiseTape(seed=101), beta=cfg.beta, drop=drop)
                np.testing.assert_allclose(out["elbo"].item(), g[f"elbo_drop{i}"], rtol=2e-5)


@pytest.mark.parametrize("name", ["tiny_cmnist", "cmnist"])
def test_dmol_head_composition(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = O.make_cfg(name, x_like="diag_dmol")
    sd = O.seeded_state_dict(O.make_cfg(name), seed=7)
    dsd = O.seeded_state_dict(cfg, seed=7)
    sd = {k: v for k, v in sd.items() if not k.startswith("likelihood.")}
    sd["likelihood.conv.weight"] = dsd["likelihood.conv.weight"]
    sd["likelihood.conv.bias"] = dsd["likelihood.conv.bias"]
    x8, pa, _ = O.synthetic_batch(cfg, CASES[name], seed=11)
    x = O.normalise_x(x8)
    pa = O.expand_parents(pa, cfg.input_res)
    with torch.no_grad():
        out = O.hvae_forward(sd, cfg, x, pa, O.NoiseTape(seed=101), beta=cfg.beta)
        np.testing.assert_allclose(out["elbo"].item(), g["dmol_elbo"], rtol=2e-5)
        np.testing.assert_allclose(out["nll"].item(), g["dmol_nll"], rtol=2e-5)
        zs = O.hvae_abduct(sd, cfg, x, pa, O.NoiseTape(seed=202))
        loc, scale = O.hvae_forward_latents(sd, cfg, zs, pa)
        np.testing.assert_allclose(sub(loc), g["dmol_rec_loc_sub"], atol=5e-4)
        np.testing.assert_allclose(sub(scale), g["dmol_rec_scale_sub"], rtol=2e-3, atol=1e-6)


def test_dmol_functions():
    g = np.load(os.path.join(GOLD, "dmol_unit.npz"))
    l = torch.from_numpy(g["l"]).requires_grad_(True)
    x = torch.from_numpy(g["x"])
    loss = O.dmol_loss(x, l)
    np.testing.assert_allclose(loss.detach().numpy(), g["loss"], rtol=1e-5)
    loss.sum().backward()
    np.testing.assert_allclose(l.grad.numpy(), g["dl"], rtol=1e-4, atol=1e-7)
    with torch.no_grad():
        for mask in ["soft", "hard", "top3"]:
            m, s = O.dmol_mean(l.detach(), mask=mask)
            np.testing.assert_allclose(m.numpy(), g[f"mean_{mask}"], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(s.numpy(), g[f"scale_{mask}"], rtol=1e-5, atol=1e-6)
        tape = O.NoiseTape([torch.from_numpy(g["gumbel_u"]), torch.from_numpy(g["logistic_u"])])
        sx, ss = O.dmol_sample(l.detach(), tape, t=0.7)
        np.testing.assert_allclose(sx.numpy(), g["sample"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(ss.numpy(), g["sample_scale"], rtol=1e-5, atol=1e-6)


def test_kl_forms_agree():
    r = np.random.default_rng(0)
    a, b, c, d = (torch.from_numpy(r.standard_normal((4, 16, 8, 8)).astype(np.float64)) for _ in range(4))
    np.testing.assert_allclose(O.gaussian_kl(a, b, c, d).numpy(), O.gaussian_kl_ref_form(a, b, c, d).numpy(), rtol=1e-10)


def test_param_counts_match_survey():
    counts = {"morphomnist": 2047074, "cmnist": 2062649, "ukbb192": 17371122, "mimic192": 7976658,
              "mimic224": 8096850}
    for n, c in counts.items():
        shapes = O.param_shapes(O.make_cfg(n))
        assert sum(int(np.prod(s)) for s in shapes.values()) == c


def test_ema_decay_schedule_matches_reference_trajectory():
    """src/utils.py EMA warm-up (copy phase, inverse-decay ramp, cap at beta) replayed on the reference's own
    trajectory (tests/golden/make_golden_ema.py); the optimiser kernel implements the same rule"""
    g = np.load(os.path.join(GOLD, "ema_schedule.npz"))
    p, want = g["p"].astype(np.float64), g["ema"]
    ema = 0.0
    got = np.zeros_like(want)
    for s in range(len(p)):
        d = O.ema_decay(s, float(g["beta"]), int(g["update_after"]))
        ema = p[s] if d == 0.0 else ema - (1.0 - d) * (ema - p[s])
        got[s] = ema
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-5)
    assert O.ema_decay(101) == 0.0 and abs(O.ema_decay(102) - 2.0 / 3.0) < 1e-12 and O.ema_decay(5000) == 0.999
    # the arithmetic of csrc/optim.cu (optim_advance_kernel), restated in float32
    f = np.float32
    for s in list(range(0, 140)) + [1098, 1099, 1100, 5000]:
        dev = f(0.0)
        if s > 100 + 1:
            dev = min(max(f(1.0) - f(1.0) / (f(1.0) + f(s - 100)), f(0.0)), f(0.999))
        assert abs(float(dev) - O.ema_decay(s)) <= 1e-6, s
