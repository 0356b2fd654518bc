"""CPU oracle for the HVAE / likelihood / counterfactual hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``causal-gen_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the CPU arm that is timed next to the GPU path.

This is a *functional restatement* (plain ``torch`` CPU ops driven by a
``state_dict``) of the reference algorithm, written from the reference's
behaviour, not a copy of its modules.  Each function cites the reference
lines it follows (paths relative to ``/root/reference``).

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real
reference (``src/vae.py``, ``src/dmol.py``) in the build container, loads the
same seeded weights into it and records its outputs; ``tests/test_oracle.py``
checks this restatement against those committed vectors.

Noise is explicit: every place the reference draws ``randn_like`` /
``uniform_`` takes the next tensor from a caller-provided ``NoiseTape`` so
the oracle, the reference (patched to consume the same tape) and the CUDA
path all see the same eps.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
MIN_LOGSCALE = -9.0  # src/vae.py:11


# --------------------------------------------------------------------------
# hyper-parameters (src/hps.py:12-78 presets + src/run_local.sh / run_slurm.sh flags)
# --------------------------------------------------------------------------
def make_cfg(name: str, **over) -> SimpleNamespace:
    base = dict(
        hps=name, input_channels=1, bottleneck=4, z_dim=16, z_max_res=192,
        bias_max_res=64, cond_prior=False, q_correction=False,
        x_like="diag_dgauss", std_init=0.0, kl_free_bits=0.0, context_dim=4,
        beta=1.0,
    )
    mnist = dict(input_res=32, enc_arch="32b3d2,16b3d2,8b3d2,4b3d4,1b4",
                 dec_arch="1b4,4b4,8b4,16b4,32b4", widths=[16, 32, 64, 128, 256])
    deep = dict(input_res=192, enc_arch="192b1d2,96b3d2,48b7d2,24b11d2,12b7d2,6b3d6,1b2",
                dec_arch="1b2,6b4,12b8,24b12,48b8,96b4,192b2",
                widths=[32, 64, 96, 128, 160, 192, 512], z_max_res=96)
    presets = {
        # src/run_local.sh:3-15
        "morphomnist": dict(mnist, context_dim=12, cond_prior=True),
        "cmnist": dict(mnist, context_dim=20, input_channels=3),
        # src/run_slurm.sh:23-36
        "ukbb192": dict(deep, context_dim=4, beta=5.0),
        # src/run_slurm.sh:38-52
        "mimic192": dict(deep, context_dim=6, beta=9.0),
        # the 224x224 arch that actually runs
        "mimic224": dict(deep, input_res=224, context_dim=6, beta=9.0, z_max_res=112,
                         enc_arch="224b1d2,112b3d2,56b7d2,28b11d2,14b7d2,7b3d7,1b2",
                         dec_arch="1b2,8b4,14b8,28b12,56b8,112b4,224b2"),
        # small synthetic archs for fast tests
        "tiny_ukbb": dict(input_res=16, enc_arch="16b1d2,8b2d2,4b1d4,1b1",
                          dec_arch="1b1,4b2,8b2,16b1", widths=[16, 32, 48, 64],
                          context_dim=4, z_max_res=8, beta=5.0),
        "tiny_morphomnist": dict(input_res=16, enc_arch="16b1d2,8b1d2,4b1d4,1b1",
                                 dec_arch="1b1,4b1,8b1,16b1", widths=[16, 32, 48, 64],
                                 context_dim=12, cond_prior=True),
        "tiny_cmnist": dict(input_res=16, enc_arch="16b1d2,8b1d2,4b1d4,1b1",
                            dec_arch="1b1,4b1,8b1,16b1", widths=[16, 32, 48, 64],
                            context_dim=20, input_channels=3),
    }
    cfg = dict(base)
    cfg.update(presets[name])
    cfg.update(over)
    ns = SimpleNamespace(**cfg)
    ns.vr = "light" if "ukbb" in ns.hps else None  # src/vae.py:428
    return ns


# --------------------------------------------------------------------------
# architecture strings -> layer tables  (src/vae.py:90-120, 198-218)
# --------------------------------------------------------------------------
@dataclass
class BlockSpec:
    prefix: str
    cin: int
    cmid: int
    cout: int
    ksize: int = 3
    residual: bool = True
    down: Optional[int] = None
    light: bool = False

    @property
    def has_proj(self) -> bool:  # src/vae.py:70
        return self.residual and (bool(self.down) or self.cin > self.cout)


@dataclass
class DecSpec:
    idx: int
    res: int
    cin: int
    cout: int
    stochastic: bool
    prior: BlockSpec = None
    posterior: Optional[BlockSpec] = None
    conv: BlockSpec = None


@dataclass
class Arch:
    enc: List[BlockSpec] = field(default_factory=list)
    dec: List[DecSpec] = field(default_factory=list)
    bias_res: List[int] = field(default_factory=list)


def build_arch(cfg) -> Arch:
    light = cfg.vr == "light"
    arch = Arch()
    # encoder: "<res>b<n>[d<rate>]" per stage  (src/vae.py:92-120)
    plan: List[Tuple[int, Optional[int]]] = []
    for si, tok in enumerate(cfg.enc_arch.split(",")):
        body = tok.split("b")[1]
        n_plain, _, rate = body.partition("d")
        plan += [(cfg.widths[si], None)] * int(n_plain)
        if rate:
            plan.append((cfg.widths[si + 1], int(rate[0])))
    for i, (w, d) in enumerate(plan):
        w_prev = plan[max(i - 1, 0)][0]
        arch.enc.append(BlockSpec(f"encoder.blocks.{i}", w_prev, int(w_prev / cfg.bottleneck), w,
                                  down=d, light=light))
    # decoder: "<res>b<n>" per stage, widths reversed  (src/vae.py:199-207)
    rev = cfg.widths[::-1]
    dplan: List[Tuple[int, int]] = []
    for si, tok in enumerate(cfg.dec_arch.split(",")):
        r, n = tok.split("b")
        dplan += [(int(r), rev[si])] * int(n)
    for i, (r, w) in enumerate(dplan):
        w_next = dplan[min(i + 1, len(dplan) - 1)][1]
        mid = int(w / cfg.bottleneck)
        k = 3 if r > 2 else 1  # src/vae.py:146
        pre = f"decoder.blocks.{i}"
        d = DecSpec(i, r, w, w_next, stochastic=r <= cfg.z_max_res)
        d.prior = BlockSpec(pre + ".prior", w + (cfg.context_dim if cfg.cond_prior else 0), mid,
                            2 * cfg.z_dim + w, k, residual=False, light=light)
        if d.stochastic:
            d.posterior = BlockSpec(pre + ".posterior", 2 * w + cfg.context_dim, mid,
                                    2 * cfg.z_dim, k, residual=False, light=light)
        d.conv = BlockSpec(pre + ".conv", w, mid, w_next, k, light=light)
        arch.dec.append(d)
    all_res = sorted({r for r, _ in dplan})
    arch.bias_res = [r for r in all_res if r <= cfg.bias_max_res]  # src/vae.py:211-218
    return arch


def param_shapes(cfg) -> Dict[str, Tuple[int, ...]]:
    """state_dict key -> shape, in the reference's registration order."""
    arch = build_arch(cfg)
    out: Dict[str, Tuple[int, ...]] = {}

    def conv(name, ci, co, k):
        out[name + ".weight"] = (co, ci, k, k)
        out[name + ".bias"] = (co,)

    def block(b: BlockSpec):
        p = b.prefix + ".conv."
        if b.light:  # src/vae.py:49-56
            conv(p + "1", b.cin, b.cmid, b.ksize)
            conv(p + "3", b.cmid, b.cout, b.ksize)
        else:  # src/vae.py:57-68
            conv(p + "1", b.cin, b.cmid, 1)
            conv(p + "3", b.cmid, b.cmid, b.ksize)
            conv(p + "5", b.cmid, b.cmid, b.ksize)
            conv(p + "7", b.cmid, b.cout, 1)
        if b.has_proj:
            conv(b.prefix + ".width_proj", b.cin, b.cout, 1)

    conv("encoder.stem", cfg.input_channels, cfg.widths[0], 7)
    for b in arch.enc:
        block(b)
    rev = cfg.widths[::-1]
    all_res = sorted({d.res for d in arch.dec})
    # nn.ParameterList registered after blocks but ModuleList assigned first; state_dict
    # order follows registration: blocks, then bias (src/vae.py:209-218)
    for d in arch.dec:
        block(d.prior)
        if d.stochastic:
            block(d.posterior)
        conv(f"decoder.blocks.{d.idx}.z_proj", cfg.z_dim + cfg.context_dim, d.cin, 1)
        if not cfg.q_correction:
            conv(f"decoder.blocks.{d.idx}.z_feat_proj", cfg.z_dim + d.cin, d.cout, 1)
        block(d.conv)
    n_bias = 0
    for j, r in enumerate(all_res):
        if r <= cfg.bias_max_res:
            out[f"decoder.bias.{n_bias}"] = (1, rev[j], r, r)
            n_bias += 1
    if cfg.x_like.endswith("dmol"):
        conv("likelihood.conv", cfg.widths[0], 100, 1)  # src/dmol.py:223-225
    else:
        conv("likelihood.x_loc", cfg.widths[0], cfg.input_channels, 1)
        conv("likelihood.x_logscale", cfg.widths[0], cfg.input_channels, 1)
        if cfg.input_channels == 3:
            conv("likelihood.channel_coeffs", cfg.widths[0], 3, 1)
    return out


def seeded_state_dict(cfg, seed: int = 7, dtype=torch.float32) -> Dict[str, Tensor]:
    """Deterministic weights from numpy PCG64 (platform-stable), fan-in scaled so
    activations stay O(1).  Unlike the reference init the prior's last conv and
    all biases are non-zero so every term of the path is exercised."""
    rng = np.random.default_rng(seed)
    sd: Dict[str, Tensor] = {}
    n_enc = len(build_arch(cfg).enc)
    n_dec = len(build_arch(cfg).dec)
    for name, shape in param_shapes(cfg).items():
        if name.startswith("decoder.bias."):
            v = 0.1 * rng.standard_normal(shape)
        elif name.endswith(".bias"):
            v = 0.05 * rng.standard_normal(shape)
        else:
            fan_in = shape[1] * shape[2] * shape[3]
            v = rng.uniform(-1.0, 1.0, shape) * math.sqrt(3.0 / fan_in)
            # keep the residual streams bounded like src/vae.py:121-122,303-308
            last = ".conv.3.weight" if cfg.vr == "light" else ".conv.7.weight"
            if name.startswith("encoder.blocks") and name.endswith(last):
                v = v * math.sqrt(1.0 / n_enc)
            if ".conv.conv." in name and name.endswith(last):
                v = v * math.sqrt(1.0 / n_dec)
            if name.endswith("z_proj.weight"):
                v = v * math.sqrt(1.0 / n_dec)
            if ".prior.conv." in name and name.endswith(last):
                v = v * 0.5
            if name.startswith("likelihood.x_logscale"):
                v = v * 0.25
        sd[name] = torch.from_numpy(np.asarray(v)).to(dtype)
    if not cfg.x_like.endswith("dmol"):
        sd["likelihood.x_logscale.bias"] = sd["likelihood.x_logscale.bias"] - 1.0
    return sd


# --------------------------------------------------------------------------
# explicit noise
# --------------------------------------------------------------------------
class NoiseTape:
    """Pre-drawn noise consumed in call order (one tensor per reference RNG call)."""

    def __init__(self, tensors: Optional[Sequence[Tensor]] = None, seed: Optional[int] = None):
        self.tensors = list(tensors) if tensors is not None else None
        self.gen = np.random.default_rng(seed) if seed is not None else None
        self.pos = 0
        self.drawn: List[Tensor] = []

    def normal(self, like: Tensor) -> Tensor:
        if self.tensors is not None:
            e = self.tensors[self.pos].to(like.dtype)
            assert e.shape == like.shape, (e.shape, like.shape)
        else:
            e = torch.from_numpy(self.gen.standard_normal(tuple(like.shape)).astype(np.float32)).to(like.dtype)
        self.pos += 1
        self.drawn.append(e)
        return e

    def uniform(self, shape, lo, hi, dtype=torch.float32) -> Tensor:
        if self.tensors is not None:
            e = self.tensors[self.pos].to(dtype)
        else:
            e = torch.from_numpy(self.gen.uniform(lo, hi, tuple(shape)).astype(np.float32)).to(dtype)
        self.pos += 1
        self.drawn.append(e)
        return e


# --------------------------------------------------------------------------
# elementwise pieces
# --------------------------------------------------------------------------
def gaussian_kl(q_loc, q_ls, p_loc, p_ls):  # src/vae.py:14-25
    var_ratio = torch.exp(2.0 * (q_ls - p_ls))
    maha = (q_loc - p_loc) ** 2 * torch.exp(-2.0 * p_ls)
    return p_ls - q_ls - 0.5 + 0.5 * (var_ratio + maha)


def gaussian_kl_ref_form(q_loc, q_ls, p_loc, p_ls):
    """Same op order as the reference expression (for the fp32 bit-level check)."""
    return -0.5 + p_ls - q_ls + 0.5 * (q_ls.exp().pow(2) + (q_loc - p_loc).pow(2)) / p_ls.exp().pow(2)


# Optional emulation of the CUDA path's storage precision (activations and conv weights held in bf16,
# fp32 accumulation, fp32 latent statistics).  Off by default: the oracle proper is fp32.  The GPU parity
# tests use it to separate "implementation differs" from "bf16 storage differs".
EMULATE_BF16 = False


def _q(t: Tensor) -> Tensor:
    return t.to(torch.bfloat16).to(t.dtype) if EMULATE_BF16 else t


def _conv(sd, name, x, pad=0):
    return F.conv2d(_q(x), _q(sd[name + ".weight"]), sd[name + ".bias"], stride=1, padding=pad)


def run_block(sd, b: BlockSpec, x: Tensor) -> Tensor:  # src/vae.py:73-84
    pad = 0 if b.ksize == 1 else 1
    p = b.prefix + ".conv."
    if b.light:
        y = _q(_conv(sd, p + "1", F.relu(x), pad))
        y = _conv(sd, p + "3", F.relu(y), pad)
    else:
        y = _q(_conv(sd, p + "1", F.gelu(x)))
        y = _q(_conv(sd, p + "3", F.gelu(y), pad))
        y = _q(_conv(sd, p + "5", F.gelu(y), pad))
        y = _conv(sd, p + "7", F.gelu(y))
    if b.residual:
        skip = _q(_conv(sd, b.prefix + ".width_proj", x)) if x.shape[1] != y.shape[1] else x
        y = skip + y
    if b.down:
        y = F.avg_pool2d(_q(y), b.down, b.down)
    return y


def encoder(sd, cfg, arch: Arch, x: Tensor) -> Dict[int, Tensor]:  # src/vae.py:125-134
    h = _q(F.conv2d(x, sd["encoder.stem.weight"], sd["encoder.stem.bias"], padding=3))
    acts: Dict[int, Tensor] = {}
    for b in arch.enc:
        h = _q(run_block(sd, b, h))
        r = h.shape[2]
        if r % 2 == 1 and r > 1:
            h = F.pad(h, [0, 1, 0, 1])
        acts[h.shape[-1]] = h
    return acts


def _nearest_up(x: Tensor, res: int) -> Tensor:
    # F.interpolate(scale_factor=res/cur) nearest; integer for every preset except 7->8 style
    cur = x.shape[-1]
    if res % cur == 0:
        return x.repeat_interleave(res // cur, 2).repeat_interleave(res // cur, 3)
    return F.interpolate(x, scale_factor=res / cur)


def decoder(sd, cfg, arch: Arch, parents: Tensor, noise: NoiseTape,
            acts: Optional[Dict[int, Tensor]] = None, t: Optional[float] = None,
            abduct: bool = False, latents: Optional[Sequence[Optional[Tensor]]] = None,
            drop: Tuple[float, float] = (1.0, 1.0)):
    """src/vae.py:222-301.  ``drop`` = (p_sto, p_det) of the morphomnist conditioning
    dropout (src/vae.py:234-249); the caller supplies the draw."""
    zd = cfg.z_dim
    latents = list(latents) if latents is not None else []
    biases = {}
    for k in sd:
        if k.startswith("decoder.bias."):
            biases[sd[k].shape[2]] = sd[k]
    B = parents.shape[0]
    h = z = _q(biases[1].repeat(B, 1, 1, 1))
    stats: List[Dict] = []
    is_drop = "morphomnist" in cfg.hps  # src/vae.py:220
    log_t = math.log(t) if t is not None else None
    up_bias = 0
    for d in arch.dec:
        pa = parents[..., : d.res, : d.res]
        if is_drop:
            pa_sto, pa_det = pa.clone(), pa.clone()
            pa_sto[:, 2:] = pa_sto[:, 2:] * drop[0]
            pa_det[:, 2:] = pa_det[:, 2:] * drop[1]
        else:
            pa_sto = pa_det = pa
        if h.shape[-1] < d.res:
            up_bias = biases.get(d.res, 0)
            h = _q(up_bias + _nearest_up(h, d.res))
        if cfg.q_correction:
            p_in = h
        else:
            p_in = _q(up_bias + _nearest_up(z, d.res)) if z.shape[-1] < d.res else z
        if cfg.cond_prior:
            p_in = torch.cat([p_in, pa_sto], 1)
        pr = run_block(sd, d.prior, p_in)
        p_loc, p_ls, p_feat = pr[:, :zd], pr[:, zd: 2 * zd], pr[:, 2 * zd:]
        p_feat = _q(p_feat)
        if log_t is not None:
            p_ls = p_ls + log_t
        if d.stochastic:
            if acts is not None:
                q = run_block(sd, d.posterior, torch.cat([h, pa, acts[d.res]], 1))
                q_loc, q_ls = q[:, :zd], q[:, zd:]
                if log_t is not None:
                    q_ls = q_ls + log_t
                z = q_loc + q_ls.exp() * noise.normal(q_loc)
                st = {"kl": gaussian_kl_ref_form(q_loc, q_ls, p_loc, p_ls)}
                if abduct:
                    st["z"] = {"z": z, "q_loc": q_loc, "q_logscale": q_ls} if cfg.cond_prior else z
                stats.append(st)
            else:
                given = latents[d.idx] if d.idx < len(latents) else None
                if given is not None:
                    z = given
                else:
                    z = p_loc + p_ls.exp() * noise.normal(p_loc)
                    # quirk: only the out-of-range (except:) branch records stats (src/vae.py:284-289)
                    if d.idx >= len(latents) and abduct and cfg.cond_prior:
                        stats.append({"z": {"p_loc": p_loc, "p_logscale": p_ls}})
        else:
            z = p_loc
        h = h + p_feat
        # quirk (pinned by golden elbo_drop1): the reference builds pa_det but feeds the
        # un-dropped pa to z_proj (src/vae.py:294), so p_det has no effect.
        h = _q(h + _conv(sd, f"decoder.blocks.{d.idx}.z_proj", torch.cat([z, pa], 1)))
        h = _q(run_block(sd, d.conv, h))
        if not cfg.q_correction and d.idx + 1 < len(arch.dec):
            z = _q(_conv(sd, f"decoder.blocks.{d.idx}.z_feat_proj", torch.cat([z, p_feat], 1)))
    return h, stats


# --------------------------------------------------------------------------
# likelihoods
# --------------------------------------------------------------------------
def dgauss_params(sd, cfg, h: Tensor, x: Optional[Tensor] = None, t: Optional[float] = None):
    """src/vae.py:352-386"""
    loc = _conv(sd, "likelihood.x_loc", h)
    ls = _conv(sd, "likelihood.x_logscale", h).clamp(min=MIN_LOGSCALE)
    if cfg.input_channels == 3:
        c = torch.tanh(_conv(sd, "likelihood.channel_coeffs", h))
        if x is None:
            r = loc[:, 0].clamp(-1, 1)
            g = (loc[:, 1] + c[:, 0] * r).clamp(-1, 1)
            b = (loc[:, 2] + c[:, 1] * r + c[:, 2] * g).clamp(-1, 1)
        else:
            r = loc[:, 0]
            g = loc[:, 1] + c[:, 0] * x[:, 0]
            b = loc[:, 2] + c[:, 1] * x[:, 0] + c[:, 2] * x[:, 1]
        loc = torch.stack([r, g, b], 1)
    if t is not None:
        ls = ls + math.log(t)
    return loc, ls


def _tanh_cdf(v):  # src/vae.py:388-391
    return 0.5 * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (v + 0.044715 * v ** 3)))


def dgauss_nll(sd, cfg, h: Tensor, x: Tensor) -> Tensor:  # src/vae.py:393-411
    loc, ls = dgauss_params(sd, cfg, h, x)
    inv = torch.exp(-ls)
    d = x - loc
    hi = _tanh_cdf(inv * (d + 1.0 / 255.0))
    lo = _tanh_cdf(inv * (d - 1.0 / 255.0))
    lp_mid = torch.log((hi - lo).clamp(min=1e-12))
    lp_left = torch.log(hi.clamp(min=1e-12))
    lp_right = torch.log((1.0 - lo).clamp(min=1e-12))
    lp = torch.where(x < -0.999, lp_left, torch.where(x > 0.999, lp_right, lp_mid))
    return -lp.mean(dim=(1, 2, 3))


def dgauss_sample(sd, cfg, h: Tensor, noise: Optional[NoiseTape] = None, return_loc: bool = True,
                  t: Optional[float] = None):
    """src/vae.py:413-422.  ``return_loc=False`` follows the intended semantics
    (temperature on the scale); the reference passes ``t`` in the ``x`` slot there
    (SURVEY Q1), so that branch is fenced, not pinned."""
    if return_loc:
        x, ls = dgauss_params(sd, cfg, h)
    else:
        loc, ls = dgauss_params(sd, cfg, h, None, t)
        x = loc + ls.exp() * noise.normal(loc)
    return x.clamp(-1, 1), ls.exp()


def _log_softmax_last(v):  # src/dmol.py:7-11
    m = v.max(dim=-1, keepdim=True)[0]
    return v - m - torch.log(torch.exp(v - m).sum(dim=-1, keepdim=True))


def dmol_unpack(l: Tensor, nmix: int = 10):
    """l (B,H,W,10*nmix) -> logits (B,H,W,M), means/log_scales/coeffs (B,H,W,3,M)  src/dmol.py:31-38"""
    B, H, W, _ = l.shape
    logits = l[..., :nmix]
    rest = l[..., nmix:].reshape(B, H, W, 3, 3 * nmix)
    return logits, rest[..., :nmix], rest[..., nmix: 2 * nmix], rest[..., 2 * nmix:]


def dmol_loss(x_nhwc: Tensor, l: Tensor, nmix: int = 10) -> Tensor:
    """src/dmol.py:24-118 (8-bit branch).  x (B,H,W,3) in [-1,1], l (B,H,W,100) -> (B,)"""
    logits, means, ls, co = dmol_unpack(l, nmix)
    ls = ls.clamp(min=-7.0)
    co = torch.tanh(co)
    xr, xg, xb = (x_nhwc[..., c: c + 1] for c in range(3))
    mu = torch.stack([means[..., 0, :],
                      means[..., 1, :] + co[..., 0, :] * xr,
                      means[..., 2, :] + co[..., 1, :] * xr + co[..., 2, :] * xg], dim=3)
    xe = x_nhwc.unsqueeze(-1)
    d = xe - mu
    inv = torch.exp(-ls)
    up = inv * (d + 1.0 / 255.0)
    dn = inv * (d - 1.0 / 255.0)
    mid = inv * d
    delta = torch.sigmoid(up) - torch.sigmoid(dn)
    lp_left = up - F.softplus(up)
    lp_right = -F.softplus(dn)
    lp_pdf = mid - ls - 2.0 * F.softplus(mid) - math.log(127.5)
    inner = torch.where(delta > 1e-5, torch.log(delta.clamp(min=1e-12)), lp_pdf)
    xe = xe.expand_as(mu)
    lp = torch.where(xe < -0.999, lp_left, torch.where(xe > 0.999, lp_right, inner))
    lp = lp.sum(dim=3) + _log_softmax_last(logits)
    return -torch.logsumexp(lp, -1).sum(dim=(1, 2)) / float(np.prod(x_nhwc.shape[1:]))


def _dmol_ar_clamp(x3: Tensor, co: Tensor) -> Tensor:  # src/dmol.py:142-157, 196-211
    x0 = x3[..., 0].clamp(-1, 1)
    x1 = (x3[..., 1] + co[..., 0] * x0).clamp(-1, 1)
    x2 = (x3[..., 2] + co[..., 1] * x0 + co[..., 2] * x1).clamp(-1, 1)
    return torch.stack([x0, x1, x2], -1)


def dmol_mean(l: Tensor, nmix: int = 10, mask: str = "soft"):
    """src/dmol.py:164-215 -> (x (B,H,W,3), scale (B,H,W,3))"""
    logits, means, ls, co = dmol_unpack(l, nmix)
    if mask == "soft":
        sel = _log_softmax_last(logits).exp().unsqueeze(-2)
    elif mask == "hard":
        sel = F.one_hot(logits.argmax(-1), nmix).to(l.dtype).unsqueeze(-2)
    elif mask.startswith("top"):
        k = int(mask[-1])
        kth = torch.sort(logits, descending=True, dim=-1)[0][..., k - 1: k]
        sel = _log_softmax_last(logits.masked_fill(logits < kth, -float("inf"))).exp().unsqueeze(-2)
    else:
        raise ValueError(mask)
    m = (means * sel).sum(-1)
    s = (ls * sel).sum(-1).clamp(min=-7.0)
    c = (torch.tanh(co) * sel).sum(-1)
    return _dmol_ar_clamp(m, c), s.exp()


def dmol_sample(l: Tensor, noise: NoiseTape, nmix: int = 10, t: Optional[float] = None):
    """src/dmol.py:121-161 -> (x, scale); two uniform(1e-5, 1-1e-5) draws"""
    logits, means, ls, co = dmol_unpack(l, nmix)
    g = noise.uniform(logits.shape, 1e-5, 1 - 1e-5, l.dtype)
    sel = F.one_hot((logits - torch.log(-torch.log(g))).argmax(-1), nmix).to(l.dtype).unsqueeze(-2)
    m = (means * sel).sum(-1)
    s = (ls * sel).sum(-1).clamp(min=-7.0)
    c = (torch.tanh(co) * sel).sum(-1)
    u = noise.uniform(m.shape, 1e-5, 1 - 1e-5, l.dtype)
    if t is not None:
        s = s + math.log(t)
    x = m + s.exp() * (torch.log(u) - torch.log(1.0 - u))
    return _dmol_ar_clamp(x, c), s.exp()


def dmolnet_params(sd, h: Tensor) -> Tensor:  # src/dmol.py:228-229
    return _conv(sd, "likelihood.conv", h).permute(0, 2, 3, 1)


def likelihood_nll(sd, cfg, h, x):
    if cfg.x_like.endswith("dmol"):  # src/dmol.py:231-232
        return dmol_loss(x.permute(0, 2, 3, 1), dmolnet_params(sd, h))
    return dgauss_nll(sd, cfg, h, x)


def likelihood_sample(sd, cfg, h, noise=None, return_loc=True, t=None):
    if cfg.x_like.endswith("dmol"):  # src/dmol.py:234-245
        l = dmolnet_params(sd, h)
        x, s = dmol_mean(l, mask=getattr(cfg, "dmol_mask", "soft")) if return_loc else dmol_sample(l, noise, t=t)  # DmolNet.mask
        return x.clamp(-1, 1).permute(0, 3, 1, 2), s.permute(0, 3, 1, 2)
    return dgauss_sample(sd, cfg, h, noise, return_loc, t)


# --------------------------------------------------------------------------
# HVAE surface  (src/vae.py:439-522)
# --------------------------------------------------------------------------
def hvae_forward(sd, cfg, x, parents, noise: NoiseTape, beta: float = 1.0, drop=(1.0, 1.0),
                 detail: bool = False):
    arch = build_arch(cfg)
    acts = encoder(sd, cfg, arch, x)
    h, stats = decoder(sd, cfg, arch, parents, noise, acts=acts, drop=drop)
    nll = likelihood_nll(sd, cfg, h, x)
    npix = float(np.prod(x.shape[1:]))
    if cfg.kl_free_bits > 0:  # src/vae.py:443-449
        fb = torch.tensor(cfg.kl_free_bits, dtype=nll.dtype)
        kl = sum(torch.maximum(fb, s["kl"].sum(dim=(2, 3)).mean(0)).sum() for s in stats)
        kl = kl / npix
        kl_mean = kl
    else:
        kl = sum(s["kl"].sum(dim=(1, 2, 3)) for s in stats) / npix
        kl_mean = kl.mean()
    out = dict(elbo=nll.mean() + beta * kl_mean, nll=nll.mean(), kl=kl_mean)
    if detail:
        out["block_kl"] = torch.stack([s["kl"].sum(dim=(1, 2, 3)) for s in stats], 1)  # (B, nsto)
        out["h"] = h
        out["nll_per_sample"] = nll
    return out


def hvae_abduct(sd, cfg, x, parents, noise: NoiseTape, cf_parents=None, alpha=0.5, t=None):
    arch = build_arch(cfg)
    acts = encoder(sd, cfg, arch, x)
    _, q = decoder(sd, cfg, arch, parents, noise, acts=acts, abduct=True, t=t)
    q = [s["z"] for s in q]
    if not (cfg.cond_prior and cf_parents is not None):
        return q
    _, p = decoder(sd, cfg, arch, cf_parents, noise, abduct=True, t=t)
    p = [s["z"] for s in p]
    out = []
    for qs, ps in zip(q, p):  # src/vae.py:485-513
        q_scale = qs["q_logscale"].exp()
        u = (qs["z"] - qs["q_loc"]) / q_scale
        p_var = ps["p_logscale"].exp() ** 2
        r_loc = alpha * qs["q_loc"] + (1 - alpha) * ps["p_loc"]
        r_scale = (alpha ** 2 * q_scale ** 2 + (1 - alpha) ** 2 * p_var).sqrt()
        if t is not None:
            r_scale = r_scale * t
        out.append(r_loc + r_scale * u)
    return out


def hvae_forward_latents(sd, cfg, latents, parents, noise: Optional[NoiseTape] = None, t=None):
    arch = build_arch(cfg)
    h, _ = decoder(sd, cfg, arch, parents, noise or NoiseTape(seed=0), latents=latents, t=t)
    return likelihood_sample(sd, cfg, h, noise, True, t)


def hvae_sample(sd, cfg, parents, noise: NoiseTape, return_loc=True, t=None):
    arch = build_arch(cfg)
    h, _ = decoder(sd, cfg, arch, parents, noise, t=t)
    return likelihood_sample(sd, cfg, h, noise, return_loc, t)


# --------------------------------------------------------------------------
# DSCM hot lines  (src/pgm/dscm.py:47-72, 121-132; notebook cell 9 for cond_prior)
# --------------------------------------------------------------------------
def expand_parents(pa: Tensor, res: int) -> Tensor:  # src/pgm/dscm.py:129-131, src/trainer.py:20
    return pa[..., None, None].repeat(1, 1, res, res).float()


def counterfactual(sd, cfg, x, pa, cf_pa, noise: NoiseTape, t_abduct: float = 1.0, particles: int = 1):
    acc = torch.zeros_like(x)
    acc2 = torch.zeros_like(x)
    cf_x = None
    for _ in range(particles):
        zs = hvae_abduct(sd, cfg, x, pa, noise, t=t_abduct)
        if cfg.cond_prior:
            zs = [z["z"] for z in zs]
        cf_loc, cf_scale = hvae_forward_latents(sd, cfg, zs, cf_pa, noise)
        rec_loc, rec_scale = hvae_forward_latents(sd, cfg, zs, pa, noise)
        u = (x - rec_loc) / rec_scale.clamp(min=1e-12)
        cf_x = torch.clamp(cf_loc + cf_scale * u, min=-1, max=1)
        acc += cf_x
        acc2 += cf_x ** 2
    if particles > 1:
        var = (acc2 - acc ** 2 / particles) / particles
        return acc / particles, var
    return cf_x, None


# --------------------------------------------------------------------------
# synthetic inputs and the CPU training step used as baseline
# --------------------------------------------------------------------------
def synthetic_batch(cfg, batch: int, seed: int = 7):
    rng = np.random.default_rng(seed)
    C, R = cfg.input_channels, cfg.input_res
    x8 = rng.integers(0, 256, (batch, C, R, R), dtype=np.uint8)
    x8[rng.random((batch, C, R, R)) < 0.4] = 0
    x8[rng.random((batch, C, R, R)) < 0.02] = 255
    ctx = cfg.context_dim
    pa = rng.standard_normal((batch, ctx)).astype(np.float32)
    cf = pa.copy()
    cf[:, 0] = rng.standard_normal(batch).astype(np.float32)
    if "mnist" in cfg.hps:  # continuous attrs then one-hot digit
        pa[:, 2:] = 0
        pa[np.arange(batch), 2 + rng.integers(0, ctx - 2, batch)] = 1
        pa[:, :2] = rng.uniform(-1, 1, (batch, 2))
        cf = pa.copy()
        cf[:, 0] = rng.uniform(-1, 1, batch)
    return torch.from_numpy(x8), torch.from_numpy(pa), torch.from_numpy(cf)


def normalise_x(x8: Tensor) -> Tensor:  # src/trainer.py:17
    return (x8.float() - 127.5) / 127.5


def ema_decay(call_index: int, beta: float = 0.999, update_after: int = 100) -> float:
    """Decay the reference's EMA applies on its `call_index`-th update() (0-based), src/utils.py:169-193:
    calls 0..update_after copy the weights, the first call after that copies once more (`initted`), then the decay
    is 1 - 1/(1 + k) with k = call_index - update_after (get_current_decay reads the ALREADY incremented step),
    clamped to [0, beta].  0.0 means "ema = online weights"."""
    if call_index <= update_after + 1:
        return 0.0
    k = call_index - update_after
    return min(max(1.0 - 1.0 / (1.0 + k), 0.0), beta)


def warmup_lr(base_lr: float, step: int, warmup: int) -> float:
    """lr used by optimizer.step() number `step` (1-based) under LambdaLR(linear_warmup(warmup)):
    src/train_setup.py:47-50, src/utils.py:32-36 -- the scheduler has been stepped step-1 times"""
    it = step - 1
    return base_lr * (1.0 if (it > warmup or warmup <= 0) else it / warmup)


def train_step_cpu(sd, cfg, x, pa_full, noise, opt_state, lr=1e-3, wd=0.01, betas=(0.9, 0.9),
                   grad_clip=350.0, grad_skip=500.0, step=1, ema=None, ema_update_after=100, beta=None,
                   drop=(1.0, 1.0)):
    """One reference training step on CPU: src/trainer.py:62-87 + AdamW (src/train_setup.py:42-53).
    ``sd`` tensors must have requires_grad=True (frozen ones False).  Returns (out, grad_norm)."""
    for p in sd.values():
        p.grad = None
    out = hvae_forward(sd, cfg, x, pa_full, noise, beta=cfg.beta if beta is None else beta, drop=drop)
    out["elbo"].backward()
    params = [p for p in sd.values() if p.grad is not None]
    gn = torch.nn.utils.clip_grad_norm_(params, grad_clip)
    if gn < grad_skip and not torch.isnan(out["nll"]) and not torch.isnan(out["kl"]):
        with torch.no_grad():
            for k, p in sd.items():
                if p.grad is None:
                    continue
                m, v = opt_state.setdefault(k, (torch.zeros_like(p), torch.zeros_like(p)))
                p.mul_(1 - lr * wd)
                m.mul_(betas[0]).add_(p.grad, alpha=1 - betas[0])
                v.mul_(betas[1]).addcmul_(p.grad, p.grad, value=1 - betas[1])
                mh = m / (1 - betas[0] ** step)
                vh = v / (1 - betas[1] ** step)
                p.addcdiv_(mh, vh.sqrt().add_(1e-8), value=-lr)
                if ema is not None:  # ema.update() of the same (non-skipped) step, src/trainer.py:74-77
                    ema[k].lerp_(p, 1.0 - ema_decay(step - 1, update_after=ema_update_after))
    return out, gn
