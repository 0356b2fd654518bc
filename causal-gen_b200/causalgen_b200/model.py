"""Parameter containers of the drop-in HVAE.

The module tree only exists to own the fp32 master weights under the reference's state_dict keys
and OIHW shapes (``encoder.stem.weight``, ``decoder.blocks.3.prior.conv.1.weight``,
``decoder.bias.0`` ...) so reference checkpoints load and save unchanged
(reference src/trainer.py:154-168, src/pgm/train_cf.py:363-364).  No module here computes anything:
all arithmetic is launched by ``engine.Engine`` through the C ABI.

Modules are created in the reference's construction order so that, under the same torch seed, the
default nn.Conv2d initialisation draws the same numbers (reference src/vae.py:87-123,137-170,195-220,
322-350; init scaling src/vae.py:121-122,303-308).
"""
from __future__ import annotations

import math
from typing import List

import torch
from torch import nn

from .arch import bias_resolutions, decoder_plan, encoder_plan


class _Slot(nn.Module):
    """parameter-free placeholder keeping the conv indices 1,3,5,7 of the reference's nn.Sequential"""

    def forward(self, x):  # pragma: no cover - never called
        raise RuntimeError("causalgen_b200 modules are parameter containers; call the HVAE surface instead")


def _owner_of(child: nn.Module):
    """the HVAE a sub-module belongs to (weak back-reference bound by HVAE._bind_children)"""
    ref = child.__dict__.get("_owner")
    owner = ref() if ref is not None else None
    if owner is None or not any(c is child for c in (owner.encoder, owner.decoder, owner.likelihood)):
        raise RuntimeError("this sub-module is not bound to a causalgen_b200.HVAE (stand-alone or copied container): call it "
                           "through its HVAE, e.g. model.encoder(x) after model.engine()")
    return owner


class _Bound(nn.Module):
    """sub-module with a weak back-reference to its HVAE; the reference is not part of its pickled / copied state"""

    def __getstate__(self):
        st = self.__dict__.copy()
        st.pop("_owner", None)
        return st


class Block(nn.Module):
    def __init__(self, cin, cmid, cout, ksize=3, residual=True, down=None, light=False):
        super().__init__()
        self.cin, self.cmid, self.cout = cin, cmid, cout
        self.ksize, self.residual, self.d, self.light = ksize, residual, down, light
        pad = 0 if ksize == 1 else 1
        if light:
            convs = [nn.Conv2d(cin, cmid, ksize, 1, pad), nn.Conv2d(cmid, cout, ksize, 1, pad)]
        else:
            convs = [nn.Conv2d(cin, cmid, 1, 1), nn.Conv2d(cmid, cmid, ksize, 1, pad),
                     nn.Conv2d(cmid, cmid, ksize, 1, pad), nn.Conv2d(cmid, cout, 1, 1)]
        seq: List[nn.Module] = []
        for c in convs:
            seq += [_Slot(), c]
        self.conv = nn.Sequential(*seq)
        if residual and (down or cin > cout):
            self.width_proj = nn.Conv2d(cin, cout, 1, 1)

    @property
    def convs(self) -> List[nn.Conv2d]:
        return [m for m in self.conv if isinstance(m, nn.Conv2d)]


class Encoder(_Bound):
    def __init__(self, args):
        super().__init__()
        self.plan = encoder_plan(args)
        self.stem = nn.Conv2d(args.input_channels, args.widths[0], kernel_size=7, stride=1, padding=3)
        light = args.vr == "light"
        blocks = [Block(s.cin, s.cmid, s.cout, down=s.down, light=light) for s in self.plan]
        for b in blocks:
            b.convs[-1].weight.data *= math.sqrt(1.0 / len(blocks))
        self.blocks = nn.ModuleList(blocks)

    def forward(self, x):
        """reference Encoder.forward (src/vae.py:124-134): {resolution: activations}, fp32 NCHW; inference only"""
        return _owner_of(self)._call_encoder(x)


class DecoderBlock(nn.Module):
    def __init__(self, args, st):
        super().__init__()
        self.res, self.stochastic = st.res, st.stochastic
        self.z_dim, self.cond_prior, self.q_correction = args.z_dim, args.cond_prior, args.q_correction
        light = args.vr == "light"
        ctx = args.context_dim
        self.prior = Block(st.cin + (ctx if args.cond_prior else 0), st.cmid, 2 * args.z_dim + st.cin,
                           ksize=st.ksize, residual=False, light=light)
        if st.stochastic:
            self.posterior = Block(2 * st.cin + ctx, st.cmid, 2 * args.z_dim, ksize=st.ksize, residual=False,
                                   light=light)
        self.z_proj = nn.Conv2d(args.z_dim + ctx, st.cin, 1)
        if not args.q_correction:
            self.z_feat_proj = nn.Conv2d(args.z_dim + st.cin, st.cout, 1)
        self.conv = Block(st.cin, st.cmid, st.cout, ksize=st.ksize, light=light)


class Decoder(_Bound):
    def __init__(self, args):
        super().__init__()
        self.plan = decoder_plan(args)
        blocks = [DecoderBlock(args, st) for st in self.plan]
        scale = math.sqrt(1.0 / len(blocks))
        for b in blocks:
            b.z_proj.weight.data *= scale
            b.conv.convs[-1].weight.data *= scale
            b.prior.convs[-1].weight.data *= 0.0
        self.blocks = nn.ModuleList(blocks)
        self.bias_res = bias_resolutions(args)
        self.bias = nn.ParameterList([nn.Parameter(torch.zeros(1, w, r, r)) for r, w in self.bias_res])
        self.cond_prior = args.cond_prior
        self.is_drop_cond = "morphomnist" in args.hps  # src/vae.py:220

    def forward(self, parents, x=None, t=None, abduct=False, latents=(), eps=None):
        """reference Decoder.forward (src/vae.py:222-301) -> (h, stats); fp32 NCHW; inference only.  `eps` (extension):
        explicit N(0,1) draws, one per sampled block in block order, instead of torch's device RNG"""
        return _owner_of(self)._call_decoder(parents, x=x, t=t, abduct=abduct, latents=latents, eps=eps)


class DGaussNet(_Bound):
    """discretised-Gaussian likelihood head (reference src/vae.py:322-350)"""

    def __init__(self, args):
        super().__init__()
        w0, C = args.widths[0], args.input_channels
        self.x_loc = nn.Conv2d(w0, C, kernel_size=1, stride=1)
        self.x_logscale = nn.Conv2d(w0, C, kernel_size=1, stride=1)
        if C == 3:
            self.channel_coeffs = nn.Conv2d(w0, 3, kernel_size=1, stride=1)
        if args.std_init > 0:
            nn.init.zeros_(self.x_logscale.weight)
            nn.init.constant_(self.x_logscale.bias, math.log(args.std_init))
            cov = args.x_like.split("_")[0]
            if cov == "fixed":
                self.x_logscale.weight.requires_grad = False
                self.x_logscale.bias.requires_grad = False
            elif cov == "shared":
                self.x_logscale.weight.requires_grad = False
                self.x_logscale.bias.requires_grad = True
            elif cov != "diag":
                raise NotImplementedError(f"{args.x_like} not implemented.")

    def nll(self, h, x):
        """src/vae.py:371-411: per-sample negative log-likelihood (N,) of x under the heads applied to h; inference only"""
        return _owner_of(self)._call_likelihood("nll", h, x)

    def sample(self, h, return_loc: bool = True, t=None):
        """src/vae.py:413-422 with return_loc=True (every reference caller): (clamped loc, scale)"""
        if not return_loc:
            raise NotImplementedError("return_loc=False hits a reference bug (t passed as x, src/vae.py:419 vs :352)")
        return _owner_of(self)._call_likelihood("sample", h)


class DmolNet(_Bound):
    """mixture-of-logistics head (reference src/dmol.py:218-226): 1x1 conv to 10 * 10 channels"""

    def __init__(self, args):
        super().__init__()
        self.width = args.widths[0]
        self.num_mixtures = 10
        self.conv = nn.Conv2d(self.width, 100, kernel_size=1, stride=1, padding=0)
        self.mask = "soft"

    def nll(self, h, x):
        """src/dmol.py:228-229: per-sample mixture-of-logistics loss (N,); inference only"""
        return _owner_of(self)._call_likelihood("nll", h, x)

    def sample(self, h, return_loc: bool = True, t=None):
        """src/dmol.py:231-245 with return_loc=True: mean under `self.mask` (soft / hard / top-k) and scale"""
        if not return_loc:
            raise NotImplementedError("stochastic DMoL samples need caller-drawn uniforms: use cg_dmol_predict mode 2")
        return _owner_of(self)._call_likelihood("sample", h)
