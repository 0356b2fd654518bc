"""Generate golden vectors by running the REAL reference (imported from /root/reference/src).

Run only in the build container (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

For every case it builds the reference ``HVAE`` from the reference's own ``hps`` presets and
flag sets, loads the deterministic weights of ``oracle.hvae_oracle.seeded_state_dict`` into it
(``load_state_dict(strict=True)`` -- also pins key names/shapes), patches
``vae.sample_gaussian`` to draw from a pre-generated eps tape, runs forward / backward /
abduct / forward_latents / sample / the DSCM combine lines, and stores the (small) outputs in
``tests/golden/<case>.npz``.  The oracle is then checked against these files by
``tests/test_oracle.py`` with no access to the reference.
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, "/root/reference/src")

import hvae_oracle as O  # noqa: E402
import hps as ref_hps  # noqa: E402
import vae as ref_vae  # noqa: E402
import dmol as ref_dmol  # noqa: E402

CASES = {
    # name: (oracle cfg name, reference --hps, extra reference flags, batch)
    "tiny_ukbb": ("tiny_ukbb", "ukbb64", ["--input_res", "16", "--enc_arch", "16b1d2,8b2d2,4b1d4,1b1",
                                          "--dec_arch", "1b1,4b2,8b2,16b1", "--widths", "16", "32", "48", "64",
                                          "--context_dim", "4", "--z_max_res", "8", "--beta", "5"], 3),
    "tiny_morphomnist": ("tiny_morphomnist", "morphomnist",
                         ["--input_res", "16", "--enc_arch", "16b1d2,8b1d2,4b1d4,1b1",
                          "--dec_arch", "1b1,4b1,8b1,16b1", "--widths", "16", "32", "48", "64",
                          "--context_dim", "12", "--cond_prior"], 3),
    "tiny_cmnist": ("tiny_cmnist", "cmnist",
                    ["--input_res", "16", "--enc_arch", "16b1d2,8b1d2,4b1d4,1b1",
                     "--dec_arch", "1b1,4b1,8b1,16b1", "--widths", "16", "32", "48", "64",
                     "--context_dim", "20"], 3),
    "morphomnist": ("morphomnist", "morphomnist", ["--context_dim", "12", "--cond_prior"], 2),
    "cmnist": ("cmnist", "cmnist", ["--context_dim", "20"], 2),
    "ukbb192": ("ukbb192", "ukbb192", ["--context_dim", "4", "--z_max_res", "96", "--beta", "5"], 1),
    "mimic192": ("mimic192", "mimic192", ["--context_dim", "6", "--z_max_res", "96", "--beta", "9"], 1),
    # BASELINE.json configs[4]: 224x224.  The shipped mimic192 arch cannot run at 224 (KeyError 6);
    # this arch has odd resolutions (7 zero-padded to 8, src/vae.py:130-132, then pooled by 7 with floor semantics)
    "mimic224": ("mimic224", "mimic192", ["--input_res", "224", "--enc_arch", "224b1d2,112b3d2,56b7d2,28b11d2,14b7d2,7b3d7,1b2",
                                          "--dec_arch", "1b2,8b4,14b8,28b12,56b8,112b4,224b2", "--context_dim", "6",
                                          "--z_max_res", "112", "--beta", "9"], 1),
}


def ref_args(hps_name, extra):
    p = argparse.ArgumentParser()
    ref_hps.add_arguments(p)
    p.set_defaults(**ref_hps.HPARAMS_REGISTRY[hps_name].__dict__)
    a = ref_hps.Hparams()
    a.update(p.parse_args(["--hps", hps_name] + extra).__dict__)
    return a


class Tape:
    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        self.drawn = []

    def __call__(self, loc, logscale):
        e = torch.from_numpy(self.rng.standard_normal(tuple(loc.shape)).astype(np.float32))
        self.drawn.append(e)
        return loc + logscale.exp() * e


def subsample(t, n=4096):
    flat = t.detach().reshape(-1)
    if flat.numel() <= n:
        return flat.numpy().copy()
    idx = np.linspace(0, flat.numel() - 1, n).astype(np.int64)
    return flat[idx].numpy().copy()


def run_case(name):
    cfg_name, hps_name, extra, B = CASES[name]
    cfg = O.make_cfg(cfg_name)
    args = ref_args(hps_name, extra)
    torch.manual_seed(0)
    model = ref_vae.HVAE(args)
    sd = O.seeded_state_dict(cfg, seed=7)
    model.load_state_dict(sd, strict=True)
    x8, pa, cf = O.synthetic_batch(cfg, B, seed=11)
    x = O.normalise_x(x8)
    pa_f = O.expand_parents(pa, cfg.input_res)
    cf_f = O.expand_parents(cf, cfg.input_res)
    out = {}

    # --- ELBO forward + backward (eval mode: no conditioning dropout) -------------------
    model.eval()
    tape = Tape(101)
    ref_vae.sample_gaussian = tape
    captured = {}
    orig_dec_fwd = model.decoder.forward

    def dec_hook(*a, **k):
        h, stats = orig_dec_fwd(*a, **k)
        captured["stats"] = stats
        captured["h"] = h
        return h, stats

    model.decoder.forward = dec_hook
    res = model(x, pa_f, beta=args.beta)
    res["elbo"].backward()
    out["elbo"] = res["elbo"].detach().numpy()
    out["nll"] = res["nll"].detach().numpy()
    out["kl"] = res["kl"].detach().numpy()
    out["block_kl"] = torch.stack([s["kl"].sum(dim=(1, 2, 3)) for s in captured["stats"]], 1).detach().numpy()
    out["h_sub"] = subsample(captured["h"])
    names = [n for n, _ in model.named_parameters()]
    out["grad_names"] = np.array(names)
    out["grad_norm"] = np.array([float(p.grad.norm()) if p.grad is not None else -1.0
                                 for _, p in model.named_parameters()], dtype=np.float64)
    gsub = {}
    for n, p in model.named_parameters():
        if p.grad is not None and (p.numel() <= 256 or n.endswith("encoder.stem.weight")):
            gsub[n] = p.grad.detach().reshape(-1)[:64].numpy().copy()
    for n, v in list(gsub.items())[:24]:
        out["grad::" + n] = v
    model.zero_grad()
    model.decoder.forward = orig_dec_fwd

    # --- train mode with conditioning dropout (morphomnist only) ----------------------
    if "morphomnist" in hps_name:
        for opt, drop in enumerate([(0, 1), (1, 0)]):
            model.train()
            model.decoder.drop_cond = lambda d=drop: d
            ref_vae.sample_gaussian = Tape(101)
            with torch.no_grad():
                r = model(x, pa_f, beta=args.beta)
            out[f"elbo_drop{opt}"] = r["elbo"].numpy()
        model.eval()

    with torch.no_grad():
        # --- abduct (t) + forward_latents + DSCM combine -----------------------------
        ref_vae.sample_gaussian = Tape(202)
        zs = model.abduct(x, parents=pa_f, t=0.9)
        zs_plain = [z["z"] for z in zs] if model.cond_prior else zs
        out["z_stats"] = np.array([[float(z.mean()), float(z.std())] for z in zs_plain])
        out["z_last_sub"] = subsample(zs_plain[-1], 1024)
        cf_loc, cf_scale = model.forward_latents(zs_plain, parents=cf_f)
        rec_loc, rec_scale = model.forward_latents(zs_plain, parents=pa_f)
        u = (x - rec_loc) / rec_scale.clamp(min=1e-12)
        cf_x = torch.clamp(cf_loc + cf_scale * u, min=-1, max=1)
        out["rec_loc_sub"] = subsample(rec_loc)
        out["rec_scale_sub"] = subsample(rec_scale)
        out["cf_x_sub"] = subsample(cf_x)
        # partial latents: only the first half given, rest sampled from the prior
        ref_vae.sample_gaussian = Tape(303)
        half = zs_plain[: len(zs_plain) // 2]
        pl_loc, _ = model.forward_latents(half, parents=pa_f, t=0.7)
        out["partial_loc_sub"] = subsample(pl_loc)
        # --- mediator mixture abduction (conditional prior only) -----------------------
        if model.cond_prior:
            ref_vae.sample_gaussian = Tape(404)
            cf_zs = model.abduct(x, parents=pa_f, cf_parents=cf_f, alpha=0.65, t=0.8)
            out["cfz_stats"] = np.array([[float(z.mean()), float(z.std())] for z in cf_zs])
            out["cfz_last_sub"] = subsample(cf_zs[-1], 1024)
        # --- unconditional sample ------------------------------------------------------
        ref_vae.sample_gaussian = Tape(505)
        sx, sscale = model.sample(pa_f, return_loc=True, t=0.5)
        out["sample_sub"] = subsample(sx)
        out["sample_scale_sub"] = subsample(sscale)

        # --- DmolNet head swapped in as the likelihood (3-channel cases) ---------------
        if cfg.input_channels == 3:
            dcfg = O.make_cfg(cfg_name, x_like="diag_dmol")
            dsd = O.seeded_state_dict(dcfg, seed=7)
            head = ref_dmol.DmolNet(args)
            head.load_state_dict({"conv.weight": dsd["likelihood.conv.weight"],
                                  "conv.bias": dsd["likelihood.conv.bias"]})
            model.likelihood = head
            ref_vae.sample_gaussian = Tape(101)
            r = model(x, pa_f, beta=args.beta)
            out["dmol_elbo"] = r["elbo"].numpy()
            out["dmol_nll"] = r["nll"].numpy()
            ref_vae.sample_gaussian = Tape(202)
            zs = model.abduct(x, parents=pa_f)
            loc, scale = model.forward_latents(zs, parents=pa_f)
            out["dmol_rec_loc_sub"] = subsample(loc)
            out["dmol_rec_scale_sub"] = subsample(scale)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in list(out.items())[:6]},
          float(out["elbo"]), float(out["nll"]), float(out["kl"]))


def dmol_unit():
    """Stand-alone DMoL functions on random parameters (src/dmol.py)."""
    rng = np.random.default_rng(5)
    B, H, W = 2, 8, 8
    l = torch.from_numpy(rng.standard_normal((B, H, W, 100)).astype(np.float32)) * 1.5
    x8 = rng.integers(0, 256, (B, H, W, 3))
    x8[rng.random((B, H, W, 3)) < 0.3] = 0
    x8[rng.random((B, H, W, 3)) < 0.1] = 255
    x = (torch.from_numpy(x8).float() - 127.5) / 127.5
    out = {"l": l.numpy(), "x": x.numpy()}
    lg = l.clone().requires_grad_(True)
    loss = ref_dmol.discretized_mix_logistic_loss(x, lg)
    loss.sum().backward()
    out["loss"] = loss.detach().numpy()
    out["dl"] = lg.grad.numpy()
    for mask in ["soft", "hard", "top3"]:
        m, s = ref_dmol.mean_discretized_mix_logistic(l.clone(), 10, mask=mask, return_scale=True)
        out[f"mean_{mask}"] = m.numpy()
        out[f"scale_{mask}"] = s.numpy()
    # sampling: reproduce the reference's two uniform_ draws with the global torch RNG
    torch.manual_seed(99)
    g = torch.empty(B, H, W, 10).uniform_(1e-5, 1.0 - 1e-5)
    u = torch.empty(B, H, W, 3).uniform_(1e-5, 1.0 - 1e-5)
    torch.manual_seed(99)
    sx, ss = ref_dmol.sample_from_discretized_mix_logistic(l, 10, return_scale=True, t=0.7)
    out.update(gumbel_u=g.numpy(), logistic_u=u.numpy(), sample=sx.numpy(), sample_scale=ss.numpy())
    np.savez_compressed(os.path.join(HERE, "dmol_unit.npz"), **out)
    print("dmol_unit", out["loss"])


if __name__ == "__main__":
    torch.set_num_threads(8)
    which = sys.argv[1:] or (list(CASES) + ["dmol_unit"])
    for n in which:
        dmol_unit() if n == "dmol_unit" else run_case(n)
