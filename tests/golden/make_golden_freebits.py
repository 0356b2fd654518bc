"""Golden vectors for kl_free_bits > 0 (src/vae.py:443-449) from the REAL reference (build container only):

    python tests/golden/make_golden_freebits.py

Same recipe as make_golden.py (seeded weights, eps tape patched into vae.sample_gaussian).  The floor is chosen per case as
the MEDIAN (midpoint of the two middle values) of the per-(block, channel) batch-mean KL sums of the plain run, so that about half of the channels sit below
it (their KL gradient is gated off) and half above."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (sets sys.path for the oracle and /root/reference/src)

O, ref_vae = G.O, G.ref_vae


def run(name):
    cfg_name, hps_name, extra, B = G.CASES[name]
    cfg = O.make_cfg(cfg_name)
    args = G.ref_args(hps_name, extra)
    torch.manual_seed(0)
    model = ref_vae.HVAE(args)
    model.load_state_dict(O.seeded_state_dict(cfg, seed=7), strict=True)
    model.eval()
    x8, pa, _ = O.synthetic_batch(cfg, B, seed=11)
    x, pa_f = O.normalise_x(x8), O.expand_parents(pa, cfg.input_res)
    # per-channel batch means of the plain run -> floor
    captured = {}
    orig = model.decoder.forward

    def hook(*a, **k):
        h, stats = orig(*a, **k)
        captured["stats"] = stats
        return h, stats

    model.decoder.forward = hook
    ref_vae.sample_gaussian = G.Tape(101)
    with torch.no_grad():
        model(x, pa_f, beta=args.beta)
    ch = torch.cat([s["kl"].sum(dim=(2, 3)).mean(dim=0) for s in captured["stats"]])
    srt = ch.sort().values
    fb = float(0.5 * (srt[ch.numel() // 2 - 1] + srt[ch.numel() // 2]))  # midpoint of the two middle values: no channel on the floor
    model.decoder.forward = orig
    model.free_bits = fb
    ref_vae.sample_gaussian = G.Tape(101)
    res = model(x, pa_f, beta=args.beta)
    res["elbo"].backward()
    out = {"free_bits": np.float64(fb), "kl_ch": ch.numpy(), "elbo": res["elbo"].detach().numpy(),
           "nll": res["nll"].detach().numpy(), "kl": res["kl"].detach().numpy()}
    names = [n for n, _ in model.named_parameters()]
    out["grad_names"] = np.array(names)
    out["grad_norm"] = np.array([float(p.grad.norm()) if p.grad is not None else -1.0 for _, p in model.named_parameters()])
    k = 0
    for n, p in model.named_parameters():
        if p.grad is not None and (p.numel() <= 256 or n.endswith("encoder.stem.weight")) and k < 24:
            out["grad::" + n] = p.grad.detach().reshape(-1)[:64].numpy().copy()
            k += 1
    np.savez_compressed(os.path.join(HERE, f"freebits_{name}.npz"), **out)
    print(name, "free_bits", fb, "channels below floor", int((ch <= fb).sum()), "of", ch.numel(), float(res["elbo"]),
          float(res["kl"]))


if __name__ == "__main__":
    for n in ("tiny_ukbb", "tiny_morphomnist"):
        run(n)
