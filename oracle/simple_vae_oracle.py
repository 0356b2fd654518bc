"""CPU oracle for BASELINE.json configs[0]: the reference's single-latent ``simple_vae.VAE`` (the reference's own
CPU-runnable plumbing case; it has no GPU counterpart in this repo -- SURVEY.md section 8a marks it oracle-only).

TEST INFRASTRUCTURE ONLY (same rule as hvae_oracle.py: only tests/ and bench.py's CPU legs may import it).

Functional restatement driven by a ``state_dict`` with the reference's key names; each function cites the
reference lines it follows (paths relative to /root/reference).  Parity status: PINNED against
``tests/golden/simple_vae_*.npz`` produced by ``tests/golden/make_golden_simple.py`` from the real reference
(weights included in the fixture).  Noise is explicit (``eps`` arguments), like in hvae_oracle.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
EPS = -9.0  # src/simple_vae.py:11-12 (logscale clamps)


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _conv(sd, name, x, stride=1, pad=0):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=stride, padding=pad)


def _vec(y: Tensor) -> Tensor:  # (B,ctx,R,R) spatially constant parents -> (B,ctx), src/simple_vae.py:65-66
    return y[:, :, 0, 0] if y.dim() > 2 else y


def encoder(sd, x: Tensor, y: Tensor, t: Optional[float] = None):  # src/simple_vae.py:34-70
    lrelu = F.leaky_relu
    h = lrelu(_conv(sd, "encoder.conv.0", x, 2, 1))
    h = lrelu(_conv(sd, "encoder.conv.2", h, 2, 1))
    h = lrelu(_conv(sd, "encoder.conv.4", h, 2, 1))
    h = lrelu(_lin(sd, "encoder.fc.0", h.reshape(h.shape[0], -1)))
    h = lrelu(_lin(sd, "encoder.embed.0", torch.cat((h, _vec(y)), dim=-1)))
    loc, ls = _lin(sd, "encoder.z_loc", h), _lin(sd, "encoder.z_logscale", h).clamp(min=EPS)
    if t is not None:
        ls = ls + math.log(t)
    return loc, ls


def cond_prior(sd, y: Tensor, t: Optional[float] = None):  # src/simple_vae.py:73-100
    h = F.leaky_relu(_lin(sd, "decoder.prior.fc.0", _vec(y)))
    h = F.leaky_relu(_lin(sd, "decoder.prior.fc.2", h))
    loc, ls = _lin(sd, "decoder.prior.z_loc", h), _lin(sd, "decoder.prior.z_logscale", h).clamp(min=EPS)
    if t is not None:
        ls = ls + math.log(t)
    return loc, ls, _lin(sd, "decoder.prior.p_feat", h)


def decoder(sd, cond: bool, y: Tensor, z: Optional[Tensor] = None, t: Optional[float] = None, eps: Optional[Tensor] = None,
            drop=(1, 1)):  # src/simple_vae.py:282-311 (drop = (p1, p2) of drop_cond, (1, 1) outside training)
    y = _vec(y)
    y1, y2 = y.clone(), y.clone()
    y1[:, 2:] = y1[:, 2:] * drop[0]
    y2[:, 2:] = y2[:, 2:] * drop[1]
    if cond:
        p_loc, p_ls, p_feat = cond_prior(sd, y1, t)
    else:
        p_loc = torch.zeros(y.shape[0], sd["decoder.p_loc"].shape[1])
        p_ls = torch.zeros_like(p_loc) + (math.log(t) if t is not None else 0.0)
    if z is None:
        z = p_loc + p_ls.exp() * eps
    if cond:
        z = torch.cat((p_feat, z), dim=-1)
    h = F.relu(_lin(sd, "decoder.fc.0", torch.cat((z, y2), dim=-1)))
    h = F.relu(_lin(sd, "decoder.fc.2", h)).reshape(y.shape[0], -1, 4, 4)
    up = lambda v: F.interpolate(v, scale_factor=2, mode="nearest")  # noqa: E731
    h = F.relu(_conv(sd, "decoder.conv.1", up(h), 1, 1))
    h = F.relu(_conv(sd, "decoder.conv.4", up(h), 1, 1))
    h = F.relu(_conv(sd, "decoder.conv.7", up(h), 1, 2))
    return h, (p_loc, p_ls)


def dgauss_params(sd, h, t=None):  # src/simple_vae.py:130-134
    loc, ls = _conv(sd, "likelihood.x_loc", h), _conv(sd, "likelihood.x_logscale", h).clamp(min=EPS)
    if t is not None:
        ls = ls + math.log(t)
    return loc, ls


def dgauss_nll(sd, h, x):  # src/simple_vae.py:136-159 (no channel coefficients here, unlike vae.DGaussNet)
    cdf = lambda v: 0.5 * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (v + 0.044715 * v ** 3)))  # noqa: E731
    loc, ls = dgauss_params(sd, h)
    inv = torch.exp(-ls)
    cp, cm = cdf(inv * (x - loc + 1.0 / 255.0)), cdf(inv * (x - loc - 1.0 / 255.0))
    lp = torch.where(x < -0.999, torch.log(cp.clamp(min=1e-12)),
                     torch.where(x > 0.999, torch.log((1.0 - cm).clamp(min=1e-12)), torch.log((cp - cm).clamp(min=1e-12))))
    return -lp.mean(dim=(1, 2, 3))


def gaussian_kl(q_loc, q_ls, p_loc, p_ls):  # src/simple_vae.py:17-27
    return -0.5 + p_ls - q_ls + 0.5 * (q_ls.exp() ** 2 + (q_loc - p_loc) ** 2) / p_ls.exp() ** 2


def forward(sd, cond: bool, x, parents, eps, beta: float = 1.0, drop=(1, 1)) -> Dict[str, Tensor]:  # :343-352
    q_loc, q_ls = encoder(sd, x, parents)
    z = q_loc + q_ls.exp() * eps
    h, (p_loc, p_ls) = decoder(sd, cond, parents, z=z, drop=drop)
    nll = dgauss_nll(sd, h, x)
    kl = gaussian_kl(q_loc, q_ls, p_loc, p_ls).sum(dim=-1) / float(x[0].numel())
    return dict(elbo=nll.mean() + beta * kl.mean(), nll=nll.mean(), kl=kl.mean())


def abduct(sd, cond: bool, x, parents, eps, cf_parents=None, alpha=0.5, t=None):  # src/simple_vae.py:360-405
    q_loc, q_ls = encoder(sd, x, parents)
    z = q_loc + q_ls.exp() * eps
    if not cond:
        return [z]
    if cf_parents is None:
        return [dict(z=z, q_loc=q_loc, q_logscale=q_ls)]
    p_loc, p_ls, _ = cond_prior(sd, cf_parents, t)
    q_scale = q_ls.exp()
    u = (z - q_loc) / q_scale
    r_loc = alpha * q_loc + (1 - alpha) * p_loc
    r_var = alpha * q_scale ** 2 + (1 - alpha) * p_ls.exp() ** 2  # alpha, NOT alpha^2 as in HVAE.abduct (:389)
    r_scale = r_var.sqrt() * (t if t is not None else 1.0)
    return [r_loc + r_scale * u]


def forward_latents(sd, cond: bool, latents, parents, t=None):  # src/simple_vae.py:407-415, return_loc=True
    h, _ = decoder(sd, cond, parents, z=latents[0], t=t)
    loc, ls = dgauss_params(sd, h)
    return loc.clamp(-1.0, 1.0), ls.exp()


def sample(sd, cond: bool, parents, eps, t=None):  # src/simple_vae.py:354-358, return_loc=True
    h, _ = decoder(sd, cond, parents, t=t, eps=eps)
    loc, ls = dgauss_params(sd, h)
    return loc.clamp(-1.0, 1.0), ls.exp()
