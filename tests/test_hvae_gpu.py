"""Model-level parity on a real B200: the CUDA path (through the HVAE surface -> C ABI) against the CPU
oracle on the same seeded weights, inputs and eps.

The CUDA path stores activations in bf16 (fp32 accumulation, fp32 latent statistics / KL / likelihood);
the oracle -- like the reference -- is fp32.  With the all-paths-active seeded weights used here the
40-block residual stream amplifies any 2^-9 storage perturbation to ~1e-2 on pixels, so each case also
runs the oracle with the same bf16 storage points emulated (``hvae_oracle.EMULATE_BF16``) and states the
tolerance against that measured, inherent drift:

    elbo / nll / kl        rel <= 5e-3 vs the fp32 oracle                      (measured <= 2.4e-3)
    per-block KL sums      |d| <= 1e-2*|ref| + 1e-3*sum|ref| per (sample, block), rel-L2 over blocks <= 5e-3
                           (rel 1e-2; measured max rel 4.6e-3 on blocks >= 1% of the largest)
    gradients              global rel-L2 <= 3.5e-2 (fixed; measured 0.5e-2 ... 2.5e-2) AND <= max(1.5e-2, 2 * drift);
                           per tensor (norm > 1% of the largest) rel-L2 <= max(5e-2, 2 * drift_tensor + 2e-2); the
                           single-sample 192x192 cases allow outliers <= 0.3 holding <= 2 % of the gradient energy (ReLU
                           gates of the 1x1 ... 6x6 blocks flipping under bf16 storage, see the comment in the test)
                           (activation gradients are also stored in bf16, which the forward-only emulation omits)
    abducted z             rel-L2 <= max(5e-3, 2 * drift)
    rec / cf / sampled px  fixed caps: mean |d| <= 1.5/255 and p99 <= 6/255 at 16x16 / 32x32, mean <= 3/255 and
                           p99 <= 13/255 at 192x192 / 224x224 (40 blocks; measured worst 2.6/255 and 11.3/255), AND
                           mean <= max(1/255, 1.5 * drift_mean), p99 <= max(2/255, 1.5 * drift_p99)

where drift = the same statistic of (bf16-storage oracle - fp32 oracle): it tells "implementation differs" from "bf16
storage differs" and is reported next to every measured deviation in the parity report (tests/conftest.py); the fixed
numbers are what fails the test when the emulation itself would be wrong.
"""
import numpy as np
import pytest
import torch

import hvae_oracle as O
from conftest import parity_report

pytestmark = pytest.mark.gpu
DEV = "cuda"

CASES = {"tiny_ukbb": 3, "tiny_morphomnist": 3, "tiny_cmnist": 3, "morphomnist": 2, "cmnist": 2, "ukbb192": 1,
         "mimic192": 1}


def build(name, **over):
    from causalgen_b200 import HVAE
    cfg = O.make_cfg(name, **over)
    sd = O.seeded_state_dict(cfg, seed=7)
    model = HVAE(cfg)
    model.load_state_dict(sd, strict=True)
    model.to(DEV).eval()
    x8, pa, cf = O.synthetic_batch(cfg, CASES[name], seed=11)
    x = O.normalise_x(x8)
    return cfg, sd, model, x, pa, cf


def draw_eps(cfg, sd, x, pa_full, seed):
    tape = O.NoiseTape(seed=seed)
    with torch.no_grad():
        O.hvae_forward(sd, cfg, x, pa_full, tape)
    return tape.drawn


def rel_l2(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


def oracle_elbo(cfg, sd, x, pa_full, emulate):
    O.EMULATE_BF16 = emulate
    try:
        sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        tape = O.NoiseTape(seed=101)
        out = O.hvae_forward(sdr, cfg, x, pa_full, tape, beta=cfg.beta, detail=True)
        out["elbo"].backward()
    finally:
        O.EMULATE_BF16 = False
    return out, sdr, tape


@pytest.mark.parametrize("name", list(CASES) + ["tiny_ukbb+q_correction"])
def test_elbo_kl_and_gradients(name):
    over = {}
    if name.endswith("+q_correction"):  # src/vae.py:255-263,297: prior reads h, no z stream / z_feat_proj
        name, over = name.split("+")[0], dict(q_correction=True)
    cfg, sd, model, x, pa, _ = build(name, **over)
    name = name + ("+qc" if over else "")
    pa_full = O.expand_parents(pa, cfg.input_res)
    ref, sd32, tape = oracle_elbo(cfg, sd, x, pa_full, False)
    emu, sd16, _ = oracle_elbo(cfg, sd, x, pa_full, True)
    eps = [e.to(DEV) for e in tape.drawn]
    model.zero_grad()
    out = model(x.to(DEV), pa_full.to(DEV), beta=cfg.beta, eps=eps)
    out["elbo"].backward()
    torch.cuda.synchronize()
    T = f"elbo[{name}]"
    for k in ("elbo", "nll", "kl"):
        parity_report(T, f"{k} rel", abs(out[k].item() - ref[k].item()) / abs(ref[k].item()), 5e-3,
                      f"bf16-emulated oracle: {abs(emu[k].item() - ref[k].item()) / abs(ref[k].item()):.2e}")
        np.testing.assert_allclose(out[k].item(), ref[k].item(), rtol=5e-3, err_msg=f"{name} {k}")
    # per-block KL sums (rel <= 1e-2): every (sample, block) within 1e-2 of its reference value, plus an
    # absolute floor of 1e-3 of the sample's total KL for blocks whose KL is ~0
    bk, rk = model.block_kl().cpu(), ref["block_kl"].detach()
    tol = 1e-2 * rk.abs() + 1e-3 * rk.abs().sum(1, keepdim=True)
    worst = float(((bk - rk).abs() / tol).max())
    big = rk.abs() >= 1e-2 * rk.abs().max()
    parity_report(T, "block KL max rel (blocks>=1% max)", float(((bk - rk).abs() / rk.abs())[big].max()), 1e-2)
    parity_report(T, "block KL worst |d|/tol", worst, 1.0, "tol = 1e-2*|ref| + 1e-3*sum|ref|")
    parity_report(T, "block KL rel-L2", rel_l2(bk, rk), 5e-3)
    assert worst <= 1.0, f"{name} block KL: {(bk - rk).abs().max()} vs {rk.abs().max()}"
    assert rel_l2(bk, rk) <= 5e-3, f"{name} block KL rel-L2 {rel_l2(bk, rk):.4g}"
    # gradients vs fp32, bounded by the measured bf16-storage drift of the algorithm itself
    named = dict(model.named_parameters())
    num = den = dnum = 0.0
    gmax = max(float(p.grad.norm()) for p in sd32.values() if p.grad is not None)
    bad, bad_energy = [], 0.0
    for k, p in sd32.items():
        if p.grad is None:
            continue
        g, g16 = named[k].grad.cpu(), sd16[k].grad
        num += float((g - p.grad).pow(2).sum())
        dnum += float((g16 - p.grad).pow(2).sum())
        den += float(p.grad.pow(2).sum())
        if float(p.grad.norm()) > 1e-2 * gmax:
            r, drift = rel_l2(g, p.grad), rel_l2(g16, p.grad)
            if r > max(5e-2, 2 * drift + 2e-2):
                bad.append((k, round(r, 4), round(drift, 4)))
                bad_energy += float(p.grad.pow(2).sum())
    glob, gdrift = (num / den) ** 0.5, (dnum / den) ** 0.5
    parity_report(T, "grad global rel-L2", glob, min(3.5e-2, max(1.5e-2, 2 * gdrift)), f"bf16-emulated oracle drift {gdrift:.2e}")
    assert glob <= 3.5e-2, f"{name} global grad rel-L2 {glob:.4f}"
    assert glob <= max(1.5e-2, 2 * gdrift), f"{name} global grad rel-L2 {glob:.4f} (drift {gdrift:.4f})"
    if CASES[name.split("+")[0]] == 1:
        # ONE sample through the 192x192 hierarchy: its top blocks see a 1x1 ... 6x6 image, i.e. a few hundred ReLU units
        # per block, and bf16 storage moves a pre-activation that sits within rounding of zero across it.  The unit's
        # gradient gate then differs from the fp32 oracle's (measured with tools/grad_diag.py, profiles/r4c_grad_diag.txt:
        # the same library lands on 0.8e-2 or 1.35e-2 global deviation depending on which rounding-equivalent kernel
        # variant ran; decoder.blocks.1.posterior.conv.1 and everything upstream of it in the backward pass move by 11..22 %
        # together).  Such tensors carry < 1 % of the gradient energy.  A wiring error shows as a deviation of order 1 or
        # as an outlier among the large tensors, so: no outlier above 0.3, and outliers hold <= 2 % of the energy.
        worst = max((r for _, r, _ in bad), default=0.0)
        parity_report(T, "grad per-tensor outliers: worst rel-L2", worst, 0.3, f"{len(bad)} tensors over max(5e-2, 2*drift+2e-2)")
        parity_report(T, "grad per-tensor outliers: energy share", bad_energy / den, 2e-2)
        assert worst <= 0.3 and bad_energy / den <= 2e-2, f"{name} per-tensor grad outliers (name, rel, drift): {bad[:8]}"
    else:
        assert not bad, f"{name} per-tensor grad outliers (name, rel, drift): {bad[:8]}"


@pytest.mark.parametrize("name", ["tiny_ukbb", "tiny_morphomnist"])
@pytest.mark.parametrize("explicit", [True, False])
def test_free_bits_elbo_and_gradients(name, explicit):
    """kl_free_bits > 0 (src/vae.py:443-449) against the oracle (pinned to the real reference by freebits_<name>.npz): the
    per-channel batch-mean statistics (cg_latent_fwd kl_ch), the floor / gate (cg_free_bits) and the gated KL gradient
    (cg_latent_bwd kl_gate).  explicit=False runs the Philox stream kernels (statistics only: the noise differs)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"freebits_{name}.npz"))
    fb = float(g["free_bits"])
    cfg, sd, model, x, pa, _ = build(name, kl_free_bits=fb)
    pa_full = O.expand_parents(pa, cfg.input_res)
    ref, sd32, tape = oracle_elbo(cfg, sd, x, pa_full, False)
    np.testing.assert_allclose(ref["elbo"].item(), g["elbo"], rtol=2e-5)   # the oracle IS the reference here
    T = f"freebits[{name},{'eps' if explicit else 'philox'}]"
    model.zero_grad()
    if explicit:
        out = model(x.to(DEV), pa_full.to(DEV), beta=cfg.beta, eps=[e.to(DEV) for e in tape.drawn])
    else:
        out = model(x.to(DEV), pa_full.to(DEV), beta=cfg.beta)
    out["elbo"].backward()
    torch.cuda.synchronize()
    prog = [p for k, p in model.engine().programs.items() if k[0] == "elbo"][-1]
    ch = prog.kl_ch.flatten().cpu().numpy() / x.shape[0]
    gate = prog.kl_gate.flatten().cpu().numpy()
    if not explicit:
        # different noise: the statistics must still be self-consistent (kl = sum max(fb, mean) / npix, gate = mean > fb)
        npix = float(np.prod(x.shape[1:]))
        np.testing.assert_allclose(out["kl"].item(), np.maximum(fb, ch).sum() / npix, rtol=1e-4)
        assert ((ch > fb) == (gate > 0.5)).all()
        return
    parity_report(T, "per-channel mean KL max rel", float(np.max(np.abs(ch - g["kl_ch"]) / np.abs(g["kl_ch"]))), 1e-2)
    np.testing.assert_allclose(ch, g["kl_ch"], rtol=1e-2)
    assert ((g["kl_ch"] > fb) == (gate > 0.5)).all(), "free-bits gate differs from the reference"
    for k in ("elbo", "nll", "kl"):
        parity_report(T, f"{k} rel", abs(out[k].item() - ref[k].item()) / abs(ref[k].item()), 5e-3)
        np.testing.assert_allclose(out[k].item(), ref[k].item(), rtol=5e-3, err_msg=f"{name} {k}")
    named = dict(model.named_parameters())
    num = den = 0.0
    for k, p in sd32.items():
        if p.grad is not None:
            num += float((named[k].grad.cpu() - p.grad).pow(2).sum())
            den += float(p.grad.pow(2).sum())
    parity_report(T, "grad global rel-L2", (num / den) ** 0.5, 3.5e-2)
    assert (num / den) ** 0.5 <= 3.5e-2


def px_stats(a, b):
    d = (a - b).abs().flatten()
    return float(d.mean()), float(torch.quantile(d[:: max(1, d.numel() // 200000)], 0.99))


def assert_pixels(ours, ref, emu, what):
    m, p99 = px_stats(ours, ref)
    dm, dp99 = px_stats(emu, ref)
    big = ours.shape[-1] >= 192
    cap_m, cap_p = (3.0, 13.0) if big else (1.5, 6.0)
    parity_report("pixels", f"{what} mean|d| *255", m * 255, min(cap_m, max(1.0, 1.5 * dm * 255)), f"emulated drift {dm * 255:.3f}")
    parity_report("pixels", f"{what} p99|d| *255", p99 * 255, min(cap_p, max(2.0, 1.5 * dp99 * 255)), f"emulated drift {dp99 * 255:.3f}")
    assert m * 255 <= cap_m and p99 * 255 <= cap_p, f"{what}: mean {m * 255:.2f}/255 p99 {p99 * 255:.2f}/255 over the fixed caps"
    assert m <= max(1 / 255, 1.5 * dm), f"{what}: mean |d| {m:.5f} vs bf16 drift {dm:.5f}"
    assert p99 <= max(2 / 255, 1.5 * dp99), f"{what}: p99 |d| {p99:.5f} vs bf16 drift {dp99:.5f}"


def oracle_cf(cfg, sd, x, pa_full, cf_full, emulate):
    O.EMULATE_BF16 = emulate
    try:
        with torch.no_grad():
            r = {}
            r["tape"] = O.NoiseTape(seed=202)
            zs = O.hvae_abduct(sd, cfg, x, pa_full, r["tape"], t=0.9)
            r["zs"] = [z["z"] for z in zs] if cfg.cond_prior else zs
            cf_loc, cf_scale = O.hvae_forward_latents(sd, cfg, r["zs"], cf_full)
            r["rec_loc"], r["rec_scale"] = O.hvae_forward_latents(sd, cfg, r["zs"], pa_full)
            u = (x - r["rec_loc"]) / r["rec_scale"].clamp(min=1e-12)
            r["cf"] = torch.clamp(cf_loc + cf_scale * u, -1, 1)
            r["tape2"] = O.NoiseTape(seed=303)
            half = r["zs"][: len(r["zs"]) // 2]
            r["partial"], _ = O.hvae_forward_latents(sd, cfg, half, pa_full, r["tape2"], t=0.7)
            r["tape3"] = O.NoiseTape(seed=505)
            r["sample"], r["sample_scale"] = O.hvae_sample(sd, cfg, pa_full, r["tape3"], t=0.5)
    finally:
        O.EMULATE_BF16 = False
    return r


@pytest.mark.parametrize("name", list(CASES))
def test_abduct_forward_latents_counterfactual(name):
    from causalgen_b200 import counterfactual
    cfg, sd, model, x, pa, cf = build(name)
    R = cfg.input_res
    pa_full, cf_full = O.expand_parents(pa, R), O.expand_parents(cf, R)
    ref = oracle_cf(cfg, sd, x, pa_full, cf_full, False)
    emu = oracle_cf(cfg, sd, x, pa_full, cf_full, True)
    eps = [e.to(DEV) for e in ref["tape"].drawn]
    xd, pad, cfd = x.to(DEV), pa.to(DEV), cf_full.to(DEV)  # (B,ctx) and (B,ctx,R,R) forms both accepted
    zs = model.abduct(xd, pad, t=0.9, eps=eps)
    zs = [z["z"] for z in zs] if cfg.cond_prior else zs
    assert len(zs) == len(ref["zs"])
    for i, (a, b, e) in enumerate(zip(zs, ref["zs"], emu["zs"])):
        assert a.shape == b.shape
        assert rel_l2(a.cpu(), b) <= max(5e-3, 2 * rel_l2(e, b)), f"{name} z[{i}] rel-L2 {rel_l2(a.cpu(), b):.4f}"
    loc, scale = model.forward_latents(zs, pad)
    assert_pixels(loc.cpu(), ref["rec_loc"], emu["rec_loc"], f"{name} rec loc")
    assert rel_l2(scale.cpu(), ref["rec_scale"]) <= max(5e-3, 2 * rel_l2(emu["rec_scale"], ref["rec_scale"]))
    cf_x, var = counterfactual(model, xd, pad, cfd, t_abduct=0.9, eps=[eps])
    assert var is None
    assert_pixels(cf_x.cpu(), ref["cf"], emu["cf"], f"{name} cf_x")
    # partially given latents + temperature: the rest is sampled from the prior with the given eps
    half = ref["zs"][: len(ref["zs"]) // 2]
    pl, _ = model.forward_latents([z.to(DEV) for z in half], pad, t=0.7, eps=[e.to(DEV) for e in ref["tape2"].drawn])
    assert_pixels(pl.cpu(), ref["partial"], emu["partial"], f"{name} partial latents")
    # unconditional sample
    sx, ss = model.sample(pad, t=0.5, eps=[e.to(DEV) for e in ref["tape3"].drawn])
    assert_pixels(sx.cpu(), ref["sample"], emu["sample"], f"{name} sample")
    assert rel_l2(ss.cpu(), ref["sample_scale"]) <= max(5e-3, 2 * rel_l2(emu["sample_scale"], ref["sample_scale"]))


def test_counterfactual_mimic224_custom_arch():
    """config 5 of BASELINE.json: the 224x224 DSCM abduct -> predict pass.  The shipped mimic192 arch cannot run at 224
    (SURVEY section 0); the custom arch has odd resolutions (7 -> zero-padded 8, src/vae.py:130-132) and the
    4-conv GELU block."""
    from causalgen_b200 import counterfactual
    CASES["mimic224"] = 1
    cfg, sd, model, x, pa, cf = build("mimic224")
    R = cfg.input_res
    pa_full, cf_full = O.expand_parents(pa, R), O.expand_parents(cf, R)
    ref = oracle_cf(cfg, sd, x, pa_full, cf_full, False)
    emu = oracle_cf(cfg, sd, x, pa_full, cf_full, True)
    eps = [e.to(DEV) for e in ref["tape"].drawn]
    cf_x, _ = counterfactual(model, x.to(DEV), pa.to(DEV), cf.to(DEV), t_abduct=0.9, eps=[eps])
    assert cf_x.shape == (1, 1, 224, 224)
    assert_pixels(cf_x.cpu(), ref["cf"], emu["cf"], "mimic224 cf_x")


def test_counterfactual_particles():
    from causalgen_b200 import counterfactual
    cfg, sd, model, x, pa, cf = build("tiny_ukbb")
    R = cfg.input_res
    pa_full, cf_full = O.expand_parents(pa, R), O.expand_parents(cf, R)
    with torch.no_grad():
        tape = O.NoiseTape(seed=77)
        mref, vref = O.counterfactual(sd, cfg, x, pa_full, cf_full, tape, t_abduct=1.0, particles=3)
    n = len(tape.drawn) // 3
    eps = [[e.to(DEV) for e in tape.drawn[i * n:(i + 1) * n]] for i in range(3)]
    m, v = counterfactual(model, x.to(DEV), pa.to(DEV), cf.to(DEV), t_abduct=1.0, particles=3, eps=eps)
    assert float((m.cpu() - mref).abs().mean()) <= 2 / 255
    assert float((v.cpu() - vref).abs().mean()) <= 1e-3 + 0.1 * float(vref.abs().mean())


@pytest.mark.parametrize("name", ["tiny_ukbb", "tiny_morphomnist", "tiny_cmnist"])
def test_counterfactual_fused_program_matches_reference_interface_path(name):
    """serving path (one program, bf16 latents shared in place, in-kernel noise) against the path that goes through
    the reference's fp32 latent interface (abduct -> forward_latents x2 -> combine), at a temperature where the
    noise is negligible"""
    from causalgen_b200 import counterfactual
    cfg, sd, model, x, pa, cf = build(name)
    xd, pad, cfd = x.to(DEV), pa.to(DEV), cf.to(DEV)
    zero_eps = [torch.zeros_like(e).to(DEV) for e in draw_eps(cfg, sd, x, O.expand_parents(pa, cfg.input_res), 5)]
    want, _ = counterfactual(model, xd, pad, cfd, t_abduct=1e-4, eps=[zero_eps])
    got, var = counterfactual(model, xd, pad, cfd, t_abduct=1e-4)
    torch.cuda.synchronize()
    assert var is None and got.shape == x.shape
    # the two paths round the abducted latents to bf16 at different points (1e-4-scaled noise flips last bits)
    d = (got - want).abs()
    assert float(d.mean()) <= 0.5 / 255 and float(d.max()) <= 6.0 / 255, (float(d.mean()), float(d.max()))
    mean, var = counterfactual(model, xd, pad, cfd, t_abduct=1.0, particles=3)
    assert bool(torch.isfinite(mean).all()) and bool((var >= -1e-6).all()) and float(var.mean()) > 0


@pytest.mark.parametrize("name,t_abduct", [("tiny_ukbb", 1.0), ("tiny_morphomnist", 0.8), ("tiny_cmnist", 1.0)])
def test_counterfactual_gradients_reach_the_hvae(name, t_abduct):
    """counterfactual fine-tuning (src/pgm/train_cf.py:159-180): aux_loss(cf_x).backward() must reach every HVAE weight
    the reference's autograd graph reaches through abduct -> forward_latents x 2 -> combine (src/pgm/dscm.py:52-56).
    Checked against the oracle's autograd on the eps the abduction kernels drew.

    The counterfactual is a DIFFERENCE of two decodes of the same latents divided by a predicted scale, and its clamps /
    ReLU masks are discontinuous, so its gradient is ill-conditioned under any storage rounding (the oracle with bf16
    storage emulated deviates from the fp32 oracle by 6 ... 30 % on the benchmark inputs).  The test therefore uses a smooth
    regime -- mid-range pixels, means scaled into (-1, 1), a full parent swap as intervention -- where that inherent
    drift is 2 ... 11 % (reported next to the measurement).  Tolerances: gradient global rel-L2 <= max(0.12, 1.3 x drift), cosine >= 0.985,
    no gradient where the reference has none; the new kernels are checked tightly on their own
    (test_kernels_gpu.py::test_cf_combine_and_sample_backward)."""
    from causalgen_b200 import HVAE, counterfactual
    torch.manual_seed(1234)  # the abduction draws Philox noise keyed by torch's seed: fixed, so the measured drift is too
    cfg = O.make_cfg(name)
    sd = O.seeded_state_dict(cfg, seed=7)
    sd["likelihood.x_loc.weight"] = sd["likelihood.x_loc.weight"] * 0.25
    sd["likelihood.x_loc.bias"] = sd["likelihood.x_loc.bias"] * 0.25
    model = HVAE(cfg)
    model.load_state_dict(sd, strict=True)
    model.to(DEV).train()
    model.__dict__["export_eps"] = True
    if cfg.cond_prior:
        model.drop_cond = lambda: (1, 1)
    x8, pa, _ = O.synthetic_batch(cfg, CASES[name], seed=11)
    x = O.normalise_x((x8.float() * 0.5 + 64).to(torch.uint8))
    cf = O.synthetic_batch(cfg, CASES[name], seed=99)[1]
    R = cfg.input_res
    pa_full, cf_full = O.expand_parents(pa, R), O.expand_parents(cf, R)
    w = torch.from_numpy(np.random.default_rng(5).standard_normal(tuple(x.shape)).astype(np.float32))
    model.zero_grad()
    cf_x, var = counterfactual(model, x.to(DEV), pa.to(DEV), cf.to(DEV), t_abduct=t_abduct)
    assert var is None and cf_x.requires_grad
    (cf_x * w.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    prog = model.engine().programs[("cf_train", x.shape[0])]
    eps = [e.cpu().clone() for e in prog.eps_out]

    def oracle(emulate):
        O.EMULATE_BF16 = emulate
        try:
            sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
            zs = O.hvae_abduct(sdr, cfg, x, pa_full, O.NoiseTape(tensors=eps), t=t_abduct)
            zs = [z["z"] for z in zs] if cfg.cond_prior else zs
            cf_loc, cf_scale = O.hvae_forward_latents(sdr, cfg, zs, cf_full)
            rec_loc, rec_scale = O.hvae_forward_latents(sdr, cfg, zs, pa_full)
            u = (x - rec_loc) / rec_scale.clamp(min=1e-12)
            out = torch.clamp(cf_loc + cf_scale * u, -1, 1)
            (out * w).sum().backward()
        finally:
            O.EMULATE_BF16 = False
        return out.detach(), sdr
    ref, sd32 = oracle(False)
    emu, sd16 = oracle(True)
    assert_pixels(cf_x.detach().cpu(), ref, emu, f"{name} cf_x (grad path)")
    named = dict(model.named_parameters())
    num = den = dnum = 0.0
    gmax = max(float(p.grad.norm()) for p in sd32.values() if p.grad is not None)
    bad, reached = [], 0
    for k, p in sd32.items():
        if p.grad is None or float(p.grad.norm()) <= 1e-6 * gmax:  # untouched, or cancelling analytically (scale bias)
            assert named[k].grad is None or float(named[k].grad.norm()) <= 1e-3 * gmax, f"{k}: gradient where the reference has none"
            continue
        reached += 1
        g = named[k].grad.cpu()
        num += float((g - p.grad).pow(2).sum())
        dnum += float((sd16[k].grad - p.grad).pow(2).sum())
        den += float(p.grad.pow(2).sum())
        if float(p.grad.norm()) > 5e-2 * gmax and rel_l2(g, p.grad) > 0.3:
            bad.append((k, round(rel_l2(g, p.grad), 4)))
    glob, drift = (num / den) ** 0.5, (dnum / den) ** 0.5
    ga = torch.cat([named[k].grad.cpu().flatten() for k, p in sd32.items() if p.grad is not None and float(p.grad.norm()) > 0])
    gb = torch.cat([p.grad.flatten() for k, p in sd32.items() if p.grad is not None and float(p.grad.norm()) > 0])
    cos = float((ga * gb).sum() / (ga.norm() * gb.norm()))
    parity_report(f"cf-grad[{name}]", "grad global rel-L2", glob, 0.12, f"bf16-emulated oracle drift {drift:.2e}; {reached} tensors")
    parity_report(f"cf-grad[{name}]", "grad 1 - cosine", 1 - cos, 0.015)
    # the fixed caps hold in the smooth regime; a noise draw that pushes the ORACLE's own bf16-storage drift above them is
    # judged against that drift (the implementation must not be worse than 1.3 x the emulation)
    assert glob <= max(0.12, 1.3 * drift) and cos >= 0.985, \
        f"{name} counterfactual grad rel-L2 {glob:.4f} cos {cos:.4f} (drift {drift:.4f})"
    assert not bad, f"{name} per-tensor outliers: {bad[:8]}"
    assert reached > 0.8 * len(sd32)
    # the ELBO node and the counterfactual node may be evaluated before one backward (src/pgm/dscm.py:41-88)
    model.zero_grad()
    out = model(x.to(DEV), pa.to(DEV), beta=cfg.beta)
    cf2, _ = counterfactual(model, x.to(DEV), pa.to(DEV), cf.to(DEV), t_abduct=t_abduct)
    (out["elbo"] + 0.1 * (cf2 * w.to(DEV)).sum()).backward()
    assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)


def test_counterfactual_graph_replay():
    """CUDA-graph replay of the abduct -> predict pass: null intervention returns the observation (both decodes
    share latents), a real intervention matches the eager helper up to the noise redrawn per call"""
    from causalgen_b200 import CounterfactualGraph, counterfactual
    cfg, sd, model, x, pa, cf = build("tiny_ukbb")
    xd, pad, cfd = x.to(DEV), pa.to(DEV), cf.to(DEV)
    run = CounterfactualGraph(model, x.shape[0], t_abduct=0.1)
    same, _ = run(xd, pad, pad)
    torch.cuda.synchronize()
    assert float((same - xd).abs().max()) <= 1e-4
    got = run(xd, pad, cfd)[0].clone()
    got2 = run(xd, pad, cfd)[0].clone()
    want, _ = counterfactual(model, xd, pad, cfd, t_abduct=0.1)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(got).all()) and got.shape == x.shape
    assert not torch.equal(got, got2), "noise must be redrawn on every replay"
    # low abduction temperature: the counterfactual is dominated by the posterior means
    assert float((got - want).abs().mean()) < 0.05


def test_mediator_mixture_abduction():
    cfg, sd, model, x, pa, cf = build("morphomnist")
    R = cfg.input_res
    pa_full, cf_full = O.expand_parents(pa, R), O.expand_parents(cf, R)
    with torch.no_grad():
        tape = O.NoiseTape(seed=404)
        ref = O.hvae_abduct(sd, cfg, x, pa_full, tape, cf_parents=cf_full, alpha=0.65, t=0.8)
    out = model.abduct(x.to(DEV), pa.to(DEV), cf_parents=cf.to(DEV), alpha=0.65, t=0.8,
                       eps=[e.to(DEV) for e in tape.drawn])
    for i, (a, b) in enumerate(zip(out, ref)):
        assert rel_l2(a.cpu(), b) <= 1e-2, f"cf z[{i}] rel-L2 {rel_l2(a.cpu(), b):.4f}"


def test_dmol_likelihood_swap():
    from causalgen_b200 import HVAE
    cfg = O.make_cfg("cmnist", x_like="diag_dmol")
    sd = O.seeded_state_dict(cfg, seed=7)
    model = HVAE(cfg)
    model.load_state_dict(sd, strict=True)
    model.to(DEV).eval()
    x8, pa, _ = O.synthetic_batch(cfg, 2, seed=11)
    x = O.normalise_x(x8)
    pa_full = O.expand_parents(pa, cfg.input_res)
    for p in sd.values():
        p.requires_grad_(True)
    tape = O.NoiseTape(seed=101)
    ref = O.hvae_forward(sd, cfg, x, pa_full, tape, beta=1.0)
    ref["elbo"].backward()
    out = model(x.to(DEV), pa.to(DEV), beta=1.0, eps=[e.to(DEV) for e in tape.drawn])
    out["elbo"].backward()
    for k in ("elbo", "nll", "kl"):
        np.testing.assert_allclose(out[k].item(), ref[k].item(), rtol=5e-3, err_msg=k)
    g = dict(model.named_parameters())["likelihood.conv.weight"].grad.cpu()
    assert rel_l2(g, sd["likelihood.conv.weight"].grad) <= 5e-2
    with torch.no_grad():
        tape = O.NoiseTape(seed=202)
        zs = O.hvae_abduct(sd, cfg, x, pa_full, tape)
        loc_ref, scale_ref = O.hvae_forward_latents(sd, cfg, zs, pa_full)
        O.EMULATE_BF16 = True
        try:
            loc_emu, _ = O.hvae_forward_latents(sd, cfg, zs, pa_full)
        finally:
            O.EMULATE_BF16 = False
    loc, scale = model.forward_latents([z.to(DEV) for z in zs], pa.to(DEV))
    assert_pixels(loc.cpu(), loc_ref, loc_emu, "dmol rec loc")
    # DmolNet.mask (src/dmol.py:226,164-190) is read at call time: hard / top-k means through the same cached program
    for mask in ("hard", "top3", "soft"):
        model.likelihood.mask = mask
        cfg.dmol_mask = mask
        with torch.no_grad():
            loc_ref, _ = O.hvae_forward_latents(sd, cfg, zs, pa_full)
            O.EMULATE_BF16 = True
            try:
                loc_emu, _ = O.hvae_forward_latents(sd, cfg, zs, pa_full)
            finally:
                O.EMULATE_BF16 = False
        loc, _ = model.forward_latents([z.to(DEV) for z in zs], pa.to(DEV))
        if mask == "soft":
            assert_pixels(loc.cpu(), loc_ref, loc_emu, f"dmol rec loc mask={mask}")
            continue
        # discrete component selection: a bf16-level logit difference flips the argmax / the top-k set at near-ties (the
        # bf16-emulated ORACLE flips as often), so the bulk of the pixels must agree and outliers stay a small fraction;
        # exact agreement on identical logits is covered by test_kernels_gpu.py::test_dmol_against_oracle_functions
        d, de = (loc.cpu() - loc_ref).abs().flatten(), (loc_emu - loc_ref).abs().flatten()
        med, frac = float(d.median()) * 255, float((d > 8 / 255).float().mean())
        frac_emu = float((de > 8 / 255).float().mean())
        parity_report("pixels", f"dmol rec loc mask={mask} median|d| *255", med, 1.0)
        parity_report("pixels", f"dmol rec loc mask={mask} frac(|d| > 8/255)", frac, max(0.05, 2 * frac_emu),
                      f"bf16-emulated oracle: {frac_emu:.4f}")
        assert med <= 1.0 and frac <= max(0.05, 2 * frac_emu), (mask, med, frac, frac_emu)


def test_conditioning_dropout_and_philox_noise():
    cfg, sd, model, x, pa, _ = build("tiny_morphomnist")
    pa_full = O.expand_parents(pa, cfg.input_res)
    model.train()
    tape = O.NoiseTape(seed=101)
    with torch.no_grad():
        refs = {d: O.hvae_forward(sd, cfg, x, pa_full, O.NoiseTape(seed=101), beta=1.0, drop=d)["elbo"].item()
                for d in [(0, 1), (1, 0), (1, 1)]}
        O.hvae_forward(sd, cfg, x, pa_full, tape)
    eps = [e.to(DEV) for e in tape.drawn]
    for d, want in refs.items():
        model.drop_cond = lambda d=d: d
        with torch.no_grad():
            got = model(x.to(DEV), pa.to(DEV), beta=1.0, eps=eps)["elbo"].item()
        np.testing.assert_allclose(got, want, rtol=5e-3, err_msg=str(d))
    # in-kernel Philox noise: finite, differs call to call, statistically close to the explicit-eps value
    model.eval()
    with torch.no_grad():
        a = model(x.to(DEV), pa.to(DEV))["elbo"].item()
        b = model(x.to(DEV), pa.to(DEV))["elbo"].item()
    assert np.isfinite(a) and np.isfinite(b) and a != b
    assert abs(a - refs[(1, 1)]) < 0.25 * abs(refs[(1, 1)])


def test_state_dict_roundtrip_and_deepcopy():
    import copy
    cfg, sd, model, x, pa, _ = build("tiny_ukbb")
    got = model.state_dict()
    assert list(got.keys()) == list(sd.keys())
    for k in sd:
        assert got[k].shape == sd[k].shape
    with torch.no_grad():
        e1 = model(x.to(DEV), pa.to(DEV), eps=None)
    ema = copy.deepcopy(model)  # src/utils.py:125
    ema.requires_grad_(False)
    tape_eps = [torch.zeros(1)]  # placeholder so both see deterministic noise below
    eps = [e.to(DEV) for e in draw_eps(cfg, sd, x, O.expand_parents(pa, cfg.input_res), 1)]
    with torch.no_grad():
        a = model(x.to(DEV), pa.to(DEV), eps=eps)["elbo"].item()
        b = ema(x.to(DEV), pa.to(DEV), eps=eps)["elbo"].item()
    assert a == b


def test_cpu_model_fails_loudly():
    from causalgen_b200 import HVAE
    cfg = O.make_cfg("tiny_ukbb")
    model = HVAE(cfg)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.zeros(1, 1, 16, 16), torch.zeros(1, 4))


@pytest.mark.parametrize("name,B", [("tiny_ukbb", 4), ("tiny_morphomnist", 12)])
def test_write_images_grid(name, B, tmp_path):
    """visual-regression grid (src/utils.py:231-419): row layout, blank rows, originals, determinism, PNG on disk"""
    from causalgen_b200 import HVAE
    from causalgen_b200.viz import TEMPS, intervened_parents, write_images
    cfg = O.make_cfg(name)
    model = HVAE(cfg)
    model.load_state_dict(O.seeded_state_dict(cfg, seed=7))
    model.to(DEV).eval()
    x8, pa, _ = O.synthetic_batch(cfg, B, seed=11)
    x = O.normalise_x(x8).to(DEV)
    batch = {"x": x, "pa": O.expand_parents(pa, cfg.input_res).to(DEV)}   # args.expand_pa layout of the reference
    cfg.save_dir, cfg.iter = str(tmp_path), 3
    torch.manual_seed(5)
    grid = write_images(cfg, model, batch)
    torch.manual_seed(5)
    again = write_images(cfg, model, batch, save=False)
    assert (grid == again).all(), "same torch seed, same grid"
    h = w = cfg.input_res
    per_image = 1 + (6 if cfg.cond_prior else 2) + 0     # effects (+ diffs) per image, then one blank row
    n_rows = 3 + len(TEMPS) + 1 + B * (per_image)
    assert grid.shape == (n_rows * h, B * w, cfg.input_channels) and grid.dtype == np.uint8
    orig = ((x.permute(0, 2, 3, 1) + 1.0) * 127.5).cpu().numpy().astype(np.uint8)
    rows = grid.reshape(n_rows, h, B, w, -1).transpose(0, 2, 1, 3, 4)
    assert (rows[0] == orig).all()
    assert (rows[2] == 0).all() and (rows[3 + len(TEMPS)] == 0).all() and (rows[-1] == 0).all()
    assert rows[1].std() > 0 and rows[3].std() > 0
    # one column per intervened attribute, the rest of the row is padding
    assert (rows[3 + len(TEMPS) + 1][cfg.context_dim:] == 0).all()
    assert (tmp_path / "viz-3.png").exists()
    # interventions touch exactly one attribute (group) per row
    cf = intervened_parents(cfg, pa[0], pa, 1)
    assert cf.shape == (cfg.context_dim, cfg.context_dim)
    if name == "tiny_ukbb":
        assert float(cf[0, 0]) == pytest.approx(1 - float(pa[0, 0]), rel=1e-6) and float(cf[1, 1]) == float(pa[1, 1])
        assert float(cf[3, 3]) == pytest.approx(1 - float(pa[0, 3]), rel=1e-6)
        assert (cf[0, 1:] == pa[0, 1:]).all()


@pytest.mark.parametrize("name", ["tiny_ukbb", "tiny_morphomnist", "tiny_cmnist"])
def test_submodule_calls_compose_like_the_reference(name):
    """model.encoder(x), model.decoder(parents, x=acts, ...), model.likelihood.nll / .sample as stand-alone calls
    (SURVEY 8b attrs row; src/vae.py:440-442 composes HVAE.forward from exactly these) against the oracle's functions on
    the same eps; also the binding of a deep copy (EMA, src/utils.py:125)."""
    import copy
    cfg, sd, model, x, pa, _ = build(name)
    pa_full = O.expand_parents(pa, cfg.input_res)
    arch = O.build_arch(cfg)
    T = f"submodules[{name}]"
    with torch.no_grad():
        acts_ref = O.encoder(sd, cfg, arch, x)
        tape = O.NoiseTape(seed=303)
        h_ref, stats_ref = O.decoder(sd, cfg, arch, pa_full, tape, acts=acts_ref, abduct=True, t=0.8)
        nll_ref = O.likelihood_nll(sd, cfg, h_ref, x)
        loc_ref, scale_ref = O.likelihood_sample(sd, cfg, h_ref)
    # encoder
    acts = model.encoder(x.to(DEV))
    assert sorted(acts) == sorted(acts_ref)
    worst = max(float((acts[r].cpu() - acts_ref[r]).abs().max() / (acts_ref[r].abs().max() + 1e-6)) for r in acts_ref)
    parity_report(T, "encoder acts max err / scale (worst res)", worst, 1.5e-2)
    assert worst <= 1.5e-2
    # decoder on the ORACLE's activations (isolates the decoder), abduction mode with temperature
    eps = [e.to(DEV) for e in tape.drawn]
    h, stats = model.decoder(pa_full.to(DEV), x={r: a.to(DEV) for r, a in acts_ref.items()}, abduct=True, t=0.8, eps=eps)
    eh = float((h.cpu() - h_ref).abs().max() / (h_ref.abs().max() + 1e-6))
    parity_report(T, "decoder h max err / scale", eh, 1.5e-2)
    assert eh <= 1.5e-2 and len(stats) == len(stats_ref)
    kl_rel, z_err = 0.0, 0.0
    for s, sr in zip(stats, stats_ref):
        assert s["kl"].shape == sr["kl"].shape
        kl_rel = max(kl_rel, abs(float(s["kl"].sum()) - float(sr["kl"].sum())) / (abs(float(sr["kl"].sum())) + 1e-3))
        z, zr = (s["z"]["z"], sr["z"]["z"]) if cfg.cond_prior else (s["z"], sr["z"])
        if cfg.cond_prior:
            assert set(s["z"]) == {"z", "q_loc", "q_logscale"}
            z_err = max(z_err, float((s["z"]["q_logscale"].cpu() - sr["z"]["q_logscale"]).abs().max()))
        z_err = max(z_err, float((z.cpu() - zr).abs().max() / (zr.abs().max() + 1e-6)))
    parity_report(T, "decoder stats: block KL sum rel (worst)", kl_rel, 1e-2)
    parity_report(T, "decoder stats: z / q_logscale max err", z_err, 1.5e-2)
    assert kl_rel <= 1e-2 and z_err <= 1.5e-2
    # prior-only decode of given latents == forward_latents' features
    zs = [(s["z"]["z"] if cfg.cond_prior else s["z"]) for s in stats_ref]
    with torch.no_grad():
        h2_ref, st2 = O.decoder(sd, cfg, arch, pa_full, O.NoiseTape(seed=1), latents=zs)
    h2, st2o = model.decoder(pa_full.to(DEV), latents=[z.to(DEV) for z in zs])
    e2 = float((h2.cpu() - h2_ref).abs().max() / (h2_ref.abs().max() + 1e-6))
    parity_report(T, "decoder(latents) h max err / scale", e2, 1.5e-2)
    assert e2 <= 1.5e-2 and st2o == [] and st2 == []
    # likelihood on the ORACLE's features
    nll = model.likelihood.nll(h_ref.to(DEV), x.to(DEV))
    en = float((nll.cpu() - nll_ref).abs().max() / nll_ref.abs().max())
    loc, scale = model.likelihood.sample(h_ref.to(DEV))
    el = float((loc.cpu() - loc_ref).abs().max()) * 255
    es = float(((scale.cpu() - scale_ref).abs() / scale_ref).max())
    parity_report(T, "likelihood.nll rel", en, 5e-3)
    parity_report(T, "likelihood.sample loc |d| *255", el, 2.0)
    parity_report(T, "likelihood.sample scale rel", es, 3e-2)
    assert en <= 5e-3 and el <= 2.0 and es <= 3e-2
    # a deep copy binds its OWN sub-modules (and weights)
    m2 = copy.deepcopy(model)
    with torch.no_grad():
        m2.encoder.stem.weight.mul_(0.5)
    a2 = m2.encoder(x.to(DEV))
    r0 = max(acts)
    assert float((a2[r0] - acts[r0]).abs().max()) > 1e-3, "the copy must use its own parameters"
    assert float((model.encoder(x.to(DEV))[r0] - acts[r0]).abs().max()) == 0.0
