"""BASELINE.json configs[0]: the reference's single-latent ``simple_vae.VAE`` (Morpho-MNIST 32x32, CPU).
The restatement in oracle/simple_vae_oracle.py is pinned against fixtures generated from the real reference
(tests/golden/make_golden_simple.py); there is no GPU kernel for this plumbing config."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import simple_vae_oracle as S  # noqa: E402


def load(case):
    g = np.load(os.path.join(ROOT, "tests", "golden", f"simple_vae_{case}.npz"))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    t = {k: torch.from_numpy(g[k]) for k in g.files if not k.startswith("sd.")}
    return sd, t, bool(int(g["cond_prior"]))


@pytest.mark.parametrize("case", ["morphomnist", "morphomnist_cond"])
def test_elbo_and_gradients(case):
    sd, t, cond = load(case)
    sdr = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    out = S.forward(sdr, cond, t["x"], t["pa"], t["eps"], beta=2.0)
    out["elbo"].backward()
    for k in ("elbo", "nll", "kl"):
        np.testing.assert_allclose(out[k].item(), t[k].item(), rtol=2e-5, atol=1e-8, err_msg=k)
    grads = [k for k in t if k.startswith("grad.")]
    assert grads
    for k in grads:
        np.testing.assert_allclose(sdr[k[5:]].grad.numpy(), t[k].numpy(), rtol=2e-4, atol=1e-7, err_msg=k)


@pytest.mark.parametrize("case", ["morphomnist", "morphomnist_cond"])
def test_abduct_decode_sample(case):
    sd, t, cond = load(case)
    pa_full = t["pa"][:, :, None, None].repeat(1, 1, 32, 32)  # both parent forms are accepted (src/simple_vae.py:65)
    with torch.no_grad():
        z = S.abduct(sd, cond, t["x"], pa_full, t["eps"], t=0.7)
        zs = z[0]["z"] if isinstance(z[0], dict) else z[0]
        np.testing.assert_allclose(zs.numpy(), t["abduct_z"].numpy(), rtol=1e-5, atol=1e-6)
        loc, scale = S.forward_latents(sd, cond, [zs], t["pa"])
        np.testing.assert_allclose(loc.numpy(), t["rec_loc"].numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(scale.numpy(), t["rec_scale"].numpy(), rtol=1e-4, atol=1e-6)
        if cond:  # mediator mixture with alpha (not alpha^2) on the variances
            zc = S.abduct(sd, cond, t["x"], t["pa"], t["eps"], cf_parents=t["cf"], alpha=0.3, t=0.7)[0]
            np.testing.assert_allclose(zc.numpy(), t["abduct_cf"].numpy(), rtol=1e-5, atol=1e-6)
        sx, ss = S.sample(sd, cond, t["cf"], t["eps"], t=0.5)
        np.testing.assert_allclose(sx.numpy(), t["sample_loc"].numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(ss.numpy(), t["sample_scale"].numpy(), rtol=1e-4, atol=1e-6)
