"""Golden vectors for BASELINE.json configs[0] (reference ``simple_vae.VAE``, Morpho-MNIST 32x32, CPU).
Run only in the build container:  python tests/golden/make_golden_simple.py
Imports the real reference, draws a seeded model + batch, patches ``sample_gaussian`` to consume a fixed eps and
stores weights, inputs, eps and outputs in tests/golden/simple_vae_<case>.npz."""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/src")
import hps as ref_hps  # noqa: E402
import simple_vae as ref  # noqa: E402


def ref_args(extra):
    p = argparse.ArgumentParser()
    ref_hps.add_arguments(p)
    p.set_defaults(**ref_hps.HPARAMS_REGISTRY["morphomnist"].__dict__)
    a = ref_hps.Hparams()
    a.update(p.parse_args(["--hps", "morphomnist", "--vae", "simple", "--context_dim", "12"] + extra).__dict__)
    return a


def run(case, extra):
    torch.manual_seed(7)
    args = ref_args(extra)
    model = ref.VAE(args)
    if model.cond_prior:  # zero-initialised prior heads would hide prior-path bugs (src/simple_vae.py:85-88)
        for n in ("z_loc", "z_logscale"):
            torch.nn.init.normal_(getattr(model.decoder.prior, n).weight, std=0.05)
    model.eval()
    rng = np.random.default_rng(11)
    B = 4
    x8 = rng.integers(0, 256, (B, 1, 32, 32), dtype=np.uint8)
    x8[rng.random(x8.shape) < 0.8] = 0
    x = (torch.from_numpy(x8).float() - 127.5) / 127.5
    pa = torch.from_numpy(rng.uniform(-1, 1, (B, 12)).astype(np.float32))
    cf = torch.from_numpy(rng.uniform(-1, 1, (B, 12)).astype(np.float32))
    pa_full = pa[:, :, None, None].repeat(1, 1, 32, 32)
    eps = torch.from_numpy(rng.standard_normal((B, args.z_dim)).astype(np.float32))
    ref.sample_gaussian = lambda loc, logscale: loc + logscale.exp() * eps
    out = {"x": x.numpy(), "pa": pa.numpy(), "cf": cf.numpy(), "eps": eps.numpy(), "cond_prior": np.array(int(model.cond_prior))}
    for k, v in model.state_dict().items():
        out["sd." + k] = v.numpy()
    res = model(x, pa_full, beta=2.0)
    res["elbo"].backward()
    for k in ("elbo", "nll", "kl"):
        out[k] = res[k].detach().numpy()
    for k, p in model.named_parameters():
        if p.grad is not None and k in ("encoder.conv.0.weight", "decoder.fc.0.weight", "likelihood.x_loc.weight",
                                        "encoder.z_logscale.bias"):
            out["grad." + k] = p.grad.numpy().copy()
    with torch.no_grad():
        z = model.abduct(x, pa_full, t=0.7)
        zs = z[0]["z"] if isinstance(z[0], dict) else z[0]
        out["abduct_z"] = zs.numpy()
        loc, scale = model.forward_latents([zs], pa_full)
        out["rec_loc"], out["rec_scale"] = loc.numpy(), scale.numpy()
        if model.cond_prior:
            out["abduct_cf"] = model.abduct(x, pa_full, cf_parents=cf, alpha=0.3, t=0.7)[0].numpy()
        sx, ss = model.sample(cf, t=0.5)
        out["sample_loc"], out["sample_scale"] = sx.numpy(), ss.numpy()
    path = os.path.join(HERE, f"simple_vae_{case}.npz")
    np.savez_compressed(path, **out)
    print(case, {k: float(out[k]) for k in ("elbo", "nll", "kl")}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    run("morphomnist", [])
    run("morphomnist_cond", ["--cond_prior"])
