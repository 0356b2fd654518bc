"""Visual-regression grid of reconstructions, prior samples and per-attribute pseudo counterfactuals
(reference ``write_images``, src/utils.py:231-419).  Host-side composition only: every image comes out
of ``HVAE.abduct / forward_latents / sample`` (the CUDA path), rows are stacked in the reference's order:

    originals | reconstruction from all latents (t = 0.1) | blank | prior samples at t = 0.1 .. 1.0 |
    then, per image i of the batch:  blank | [direct effect x*, x* - rec] (+ [indirect, diff], [total, diff] for
    conditional priors)  with one column per intervened attribute | blank

and written to ``<save_dir>/viz-<iter>.png`` (PIL; the reference uses imageio).  Returns the uint8 grid."""
from __future__ import annotations

import os
from typing import Dict, List

import numpy as np
import torch

TEMPS = (0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0)


def _u8(x: torch.Tensor) -> np.ndarray:
    """[-1, 1] NCHW -> channels-last [0, 255] float array (src/utils.py:238-240)"""
    return ((x.permute(0, 2, 3, 1) + 1.0) * 127.5).detach().cpu().numpy()


def intervened_parents(args, pa_i: torch.Tensor, pa_all: torch.Tensor, donor: int) -> torch.Tensor:
    """One row per attribute, each with that attribute alone changed (src/utils.py:336-369).  pa_i: (context_dim,) parents
    of the image, pa_all: (B, context_dim), donor: index of the image the continuous attributes are borrowed from."""
    ctx = args.context_dim
    cf = pa_i[None].repeat(ctx, 1).clone()
    if "ukbb" in args.hps:
        if ctx == 4:    # mri_seq, brain_volume, ventricle_volume, sex
            cf[0, 0] = 1 - cf[0, 0]
            cf[1, 1] = pa_all[donor, 1]
            cf[2, 2] = pa_all[donor, 2]
            cf[3, 3] = 1 - cf[3, 3]
        elif ctx == 3:
            cf[0, 0] = 1 - cf[0, 0]
            cf[1, 1] = pa_all[donor, 1]
            cf[2, 2] = pa_all[donor, 2]
    elif "morphomnist" in args.hps:
        assert ctx == 12
        cf[0, 0] = pa_all[donor, 0]
        cf[1, 1] = pa_all[donor, 1]
        cf[2:, 2:] = torch.eye(10, device=cf.device, dtype=cf.dtype)
    elif "cmnist" in args.hps:
        assert ctx == 20
        cf[:10, :10] = torch.eye(10, device=cf.device, dtype=cf.dtype)
        cf[10:, 10:] = torch.eye(10, device=cf.device, dtype=cf.dtype)
    return cf  # other datasets: the reference leaves the parents unchanged


@torch.no_grad()
def write_images(args, model, batch: Dict[str, torch.Tensor], save: bool = True) -> np.ndarray:
    x, pa_in = batch["x"], batch["pa"]
    bs, c, h, w = x.shape
    pa = pa_in[:, :, 0, 0] if pa_in.dim() == 4 else pa_in       # the engine broadcasts parents itself
    rows: List[np.ndarray] = []
    orig = _u8(x).astype(np.uint8)
    blank = np.zeros_like(orig)
    rows.append(orig)
    zs = model.abduct(x=x, parents=pa)
    z_all = [z["z"] for z in zs] if model.cond_prior else list(zs)
    rec, _ = model.forward_latents(latents=z_all, parents=pa, t=0.1)
    rows.append(_u8(rec).astype(np.uint8))
    rows.append(blank)
    for temp in TEMPS:
        smp, _ = model.sample(parents=pa, return_loc=True, t=temp)
        rows.append(_u8(smp).astype(np.uint8))
    order = np.arange(bs)
    np.random.RandomState(1).shuffle(order)                     # donors of the continuous interventions
    alpha, t = 0.6, 0.5
    ctx = args.context_dim
    rows.append(blank)
    for i in range(bs):
        pa_rep = pa[i][None].repeat(ctx, 1)
        cf_pa = intervened_parents(args, pa[i], pa, int(order[i]))
        z_i = [z[i][None].repeat(ctx, 1, 1, 1) for z in z_all]
        x_rec = _u8(model.forward_latents(latents=z_i, parents=pa_rep, t=t)[0])

        def effect(latents, parents):
            img = _u8(model.forward_latents(latents=latents, parents=parents, t=t)[0])
            rows.append(img.astype(np.uint8))
            rows.append((img - x_rec).astype(np.uint8))

        effect(z_i, cf_pa)                                      # direct effect  x* = g(pa*, z)
        if model.cond_prior:
            x_rep = x[i][None].repeat(ctx, 1, 1, 1)
            cf_z = model.abduct(x=x_rep, parents=pa_rep, cf_parents=cf_pa, alpha=alpha, t=t)
            effect(cf_z, pa_rep)                                # indirect effect x* = g(pa, z*)
            effect(cf_z, cf_pa)                                 # total effect    x* = g(pa*, z*)
        rows.append(blank)
    for j, r in enumerate(rows):                                # rows with fewer than bs images are zero padded
        if r.shape[0] < bs:
            rows[j] = np.concatenate([r, np.zeros((bs - r.shape[0],) + r.shape[1:], np.uint8)], 0)
        elif r.shape[0] > bs:
            raise ValueError(f"context_dim {r.shape[0]} exceeds the batch size {bs}: the grid has one column per image")
    n = len(rows)
    grid = np.concatenate(rows, 0).reshape(n, bs, h, w, c).transpose(0, 2, 1, 3, 4).reshape(n * h, bs * w, c)
    if save:
        from PIL import Image
        os.makedirs(args.save_dir, exist_ok=True)
        Image.fromarray(grid[..., 0] if c == 1 else grid).save(os.path.join(args.save_dir, f"viz-{args.iter}.png"))
    return grid
