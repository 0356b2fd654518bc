"""Drop-in surface of the reference's image mechanism: ``HVAE`` (src/vae.py:425-522), its likelihood heads
and the DSCM abduction -> action -> prediction lines (src/pgm/dscm.py:47-72,121-132).

Same constructor (an ``args`` bag built by the reference's own ``hps.py``), same state_dict keys, same
methods and return types.  Differences are documented fences, not silent changes:
  * runs only on an sm_100 CUDA device through libcausalgen_b200.so (no CPU/eager fallback);
  * activations are bf16 with fp32 accumulation / fp32 latent + likelihood math (reference: fp32/TF32);
  * ``forward`` accepts an optional ``eps`` list (one (B,16,r,r) tensor per stochastic block, reference RNG
    order) so tests can feed identical noise; without it noise comes from an in-kernel Philox stream;
  * gradients flow into the parameters for ``forward`` (the training path).  Gradients through
    ``abduct``/``forward_latents`` (counterfactual fine-tuning) are not implemented yet.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor, nn

from . import _lib as L
from .engine import Engine, TRACE_ONLY, no_grad_param_ids
from .model import Decoder, DGaussNet, DmolNet, Encoder


def _stream() -> int:
    return 0 if TRACE_ONLY else torch.cuda.current_stream().cuda_stream


def _pa_vector(parents: Tensor) -> Tensor:
    """(B,ctx,R,R) spatially-constant parents (src/trainer.py:20, src/pgm/dscm.py:129-131) or (B,ctx)"""
    if parents.dim() == 4:
        return parents[:, :, 0, 0]
    return parents


class _ElboFn(torch.autograd.Function):
    """One autograd node for the whole hand-written forward+backward."""

    @staticmethod
    def forward(ctx, model, x, parents, beta, eps, *params):
        out3 = model._run_elbo(x, parents, beta, eps, train=True)
        ctx.model = model
        eng = model._engine_or_none()
        # the engine keeps ONE gradient bucket that the next training forward (or a counterfactual backward) overwrites:
        # this node keeps its own copy, so forwards and backwards may interleave freely (DSCM.forward runs the ELBO and the
        # counterfactual before a single backward, src/pgm/dscm.py:41-88)
        ctx.grads = eng.flat_grad.clone()
        ctx.param_ids = [id(p) for p in params]
        return out3[0].clone()

    @staticmethod
    def backward(ctx, g):
        eng = ctx.model._engine_or_none()
        flat = ctx.grads * g
        by_id, off = {}, 0
        dead = no_grad_param_ids(ctx.model)  # never touched by the forward: grad stays None like in the reference
        for p in eng.params:
            by_id[id(p)] = flat[off: off + p.numel()].view_as(p) if (p.requires_grad and id(p) not in dead) else None
            off += p.numel()
        return (None, None, None, None, None) + tuple(by_id[i] for i in ctx.param_ids)


class _CfFn(torch.autograd.Function):
    """One autograd node for the fused counterfactual pass: forward = abduct -> 2 x forward_latents -> combine
    (src/pgm/dscm.py:52-56); backward = the hand-derived backward program of engine.build_counterfactual(train=True)."""

    @staticmethod
    def forward(ctx, model, x, pa, cf_pa, t_abduct, *params):
        eng = model.engine()
        N = x.shape[0]
        prog = model._program(("cf_train", N), lambda: eng.build_counterfactual(N, train=True))
        if model.__dict__.get("export_eps") and not hasattr(prog, "eps_out"):
            # parity hook: the abduction's latent kernels also write the eps they drew (fp32 NCHW, block order)
            prog.eps_out = []
            for la in prog.D.latent_args:
                if la.mode in (0, 1):
                    r = int(round(la.HW ** 0.5))
                    t = torch.zeros(N, la.zdim, r, r, device=eng.device, dtype=torch.float32)
                    la.eps_out = t.data_ptr()
                    prog.eps_out.append(t)
        prog.io.x.copy_(x)
        model._load_parents(prog, prog.io, [pa, cf_pa])
        model._set_noise(prog.D, None, math.log(t_abduct) if t_abduct is not None else 0.0, prog.D.latent_bwd_args)
        eng.pack_weights()
        prog.run()
        prog.seed_ctr.add_(1)
        ctx.model, ctx.prog = model, prog
        ctx.seed_at_fwd = int(prog.seed_ctr.item()) - 1
        ctx.param_ids = [id(p) for p in params]
        return prog.cf_x.clone()

    @staticmethod
    def backward(ctx, dcf):
        model, prog = ctx.model, ctx.prog
        eng = model.engine()
        if int(prog.seed_ctr.item()) - 1 != ctx.seed_at_fwd:
            raise RuntimeError("causalgen_b200: another counterfactual forward of the same batch size ran before this backward; "
                               "its saved activations were overwritten (call backward() first)")
        prog.seed_ctr.sub_(1)  # the backward regenerates the abduction noise of THIS forward
        prog.dcf.copy_(dcf)
        eng.flat_grad.zero_()
        prog.bwd.run()
        prog.seed_ctr.add_(1)
        flat = eng.flat_grad.clone()
        by_id, off = {}, 0
        dead = no_grad_param_ids(model)
        for p in eng.params:
            by_id[id(p)] = flat[off: off + p.numel()].view_as(p) if (p.requires_grad and id(p) not in dead) else None
            off += p.numel()
        return (None, None, None, None, None) + tuple(by_id[i] for i in ctx.param_ids)


class HVAE(nn.Module):
    def __init__(self, args):
        super().__init__()
        args.vr = "light" if "ukbb" in args.hps else None  # src/vae.py:428 (hidden state on args, kept)
        self.encoder = Encoder(args)
        self.decoder = Decoder(args)
        if args.x_like.split("_")[1] == "dgauss":
            self.likelihood = DGaussNet(args)
        elif args.x_like.split("_")[1] == "dmol":
            self.likelihood = DmolNet(args)
        else:
            raise NotImplementedError(f"{args.x_like} not implemented.")
        self.cond_prior = args.cond_prior
        self.free_bits = args.kl_free_bits
        keys = ("hps", "enc_arch", "dec_arch", "widths", "bottleneck", "z_dim", "z_max_res", "bias_max_res",
                "input_channels", "input_res", "context_dim", "cond_prior", "q_correction", "x_like", "std_init",
                "kl_free_bits", "vr")
        self._hp = {k: getattr(args, k) for k in keys}
        self.__dict__["_engine"] = None
        self.__dict__["_noise_calls"] = 0
        self._bind_children()

    # ------------------------------------------------------------------ engine plumbing
    def __getstate__(self):  # copy.deepcopy (EMA, src/utils.py:125) must not drag device programs along
        st = self.__dict__.copy()
        st["_engine"] = None
        return st

    def __setstate__(self, state):
        super().__setstate__(state)
        self._bind_children()   # a deep copy (EMA, src/utils.py:125) binds its own sub-modules

    def _bind_children(self):
        """encoder / decoder / likelihood are callable like the reference's sub-modules (src/vae.py:440-442); they reach the
        engine through a weak back-reference that is re-bound whenever an engine is (re)built, e.g. after copy.deepcopy"""
        import weakref
        for child in (self.encoder, self.decoder, self.likelihood):
            child.__dict__["_owner"] = weakref.ref(self)

    def _engine_or_none(self) -> Optional[Engine]:
        return self.__dict__.get("_engine")

    def engine(self) -> Engine:
        eng = self.__dict__.get("_engine")
        if eng is None or eng.signature != Engine.param_signature(self):
            from types import SimpleNamespace
            eng = Engine(self, SimpleNamespace(**self._hp))
            self.__dict__["_engine"] = eng
            self._bind_children()
        return eng

    def _seed(self) -> int:
        self.__dict__["_noise_calls"] = self.__dict__.get("_noise_calls", 0) + 1
        return (torch.initial_seed() * 0x9E3779B97F4A7C15 + self._noise_calls * 0xD1B54A32D192ED03) % (1 << 64)

    def _program(self, key, build):
        eng = self.engine()
        if key not in eng.programs:
            if eng.device.type == "cuda":
                with torch.cuda.device(eng.device):
                    eng.programs[key] = build()
            else:
                eng.programs[key] = build()
        prog = eng.programs[key]
        for m in getattr(prog, "dmol_modes", ()):  # DmolNet.mask (src/dmol.py:226) is read at call time, like the reference
            m.value = eng.dmol_mode()
        return prog

    def _load_parents(self, prog, io, plist: Sequence[Tensor], training_drop: Optional[Tuple[float, float]] = None):
        for buf, pa in zip(io.pa_in, plist):
            buf.copy_(_pa_vector(pa).to(torch.float32))
        Engine.set_hyper(prog, drop_sto=1.0 if training_drop is None else float(training_drop[0]))

    def _set_noise(self, D, eps: Optional[Sequence[Tensor]], log_t: float, bwd=None):
        seed = self._seed()
        k = 0
        for la in D.latent_args:
            la.log_t = log_t
            if la.mode in (0, 1):
                la.seed = seed
        if eps is not None:
            bufs = [b for b in D.eps if b is not None]
            assert len(eps) >= len(bufs), f"need {len(bufs)} eps tensors, got {len(eps)}"
            for b, e in zip(bufs, eps):
                b.copy_(e)
        if bwd is not None:
            for lb in bwd:
                lb.seed, lb.log_t = seed, log_t
        return seed

    # ------------------------------------------------------------------ ELBO
    def drop_cond(self) -> Tuple[int, int]:
        """conditioning dropout draw, CPU RNG like the reference (src/vae.py:310-319)"""
        opt = int(torch.distributions.Categorical(1 / 3 * torch.ones(3)).sample())
        return [(0, 1), (1, 0), (1, 1)][opt]

    def _run_elbo(self, x: Tensor, parents: Tensor, beta, eps, train: bool) -> Tensor:
        eng = self.engine()
        N = x.shape[0]
        explicit = eps is not None
        prog = self._program(("elbo", N, train, explicit), lambda: eng.build_elbo(N, train, explicit))
        prog.io.x.copy_(x)
        drop = None
        if self.training and self.cond_prior and self.decoder.is_drop_cond:
            drop = self.drop_cond()
        self._load_parents(prog, prog.io, [parents], drop)
        self._set_noise(prog.D, eps, 0.0, getattr(prog.D, "latent_bwd_args", None))
        eng.set_hyper(prog, beta=float(beta))
        for t in prog.zero:
            t.zero_()
        if train:
            eng.flat_grad.zero_()
            eng.generation += 1
        eng.pack_weights()
        prog.run()
        return prog.out3

    def forward(self, x: Tensor, parents: Tensor, beta: float = 1, eps: Optional[Sequence[Tensor]] = None
                ) -> Dict[str, Tensor]:
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if need_grad:
            params = list(self.parameters())
            elbo = _ElboFn.apply(self, x, parents, beta, eps, *params)
            out3 = self.engine().programs[("elbo", x.shape[0], True, eps is not None)].out3
        else:
            out3 = self._run_elbo(x, parents, beta, eps, train=False)
            elbo = out3[0].clone()
        return dict(elbo=elbo, nll=out3[1].clone(), kl=out3[2].clone())

    def block_kl(self) -> Tensor:
        """(B, n_stochastic) per-block KL sums of the last forward() -- parity tests / TensorBoard"""
        eng = self.engine()
        progs = [p for k, p in eng.programs.items() if k[0] == "elbo"]
        return progs[-1].kl_rows.t().clone()

    # ------------------------------------------------------------------ inference surface
    @torch.no_grad()
    def abduct(self, x: Tensor, parents: Tensor, cf_parents: Optional[Tensor] = None, alpha: float = 0.5,
               t: Optional[float] = None, eps: Optional[Sequence[Tensor]] = None):
        eng = self.engine()
        N = x.shape[0]
        prog = self._program(("abduct", N, self.cond_prior),
                             lambda: eng.build_decode(N, "abduct", want_stats=self.cond_prior))
        prog.io.x.copy_(x)
        self._load_parents(prog, prog.io, [parents])
        log_t = math.log(t) if t is not None else 0.0
        n_eps = len([b for b in prog.D.eps if b is not None])
        self._set_noise(prog.D, eps[:n_eps] if eps is not None else self._host_eps(prog.D), log_t)
        eng.pack_weights()
        prog.run()
        zs = [prog.D.z_out[k].clone() for k in sorted(prog.D.z_out)]
        if not self.cond_prior:
            return zs
        lib = L.load()
        s = _stream()
        q_stats = []
        for k in sorted(prog.D.z_out):
            qstat, _ = prog.D.stats_out[k]
            q_stats.append({"z": zs[k], "q_loc": self._stat_nchw(qstat, 0, 0.0), "q_logscale": self._stat_nchw(qstat, 16, log_t)})
        if cf_parents is None:
            return q_stats
        # second decoder pass: prior only under the counterfactual parents (src/vae.py:480-482)
        pprog = self._program(("prior_stats", N), lambda: eng.build_decode(N, "latents", given=None, want_stats=True))
        self._load_parents(pprog, pprog.io, [cf_parents])
        D = pprog.Ds[0]
        self._set_noise(D, eps[n_eps:] if eps is not None else self._host_eps(D), log_t)
        pprog.run()
        out = []
        for k, qs in enumerate(q_stats):
            _, pstat = D.stats_out[k]
            p_loc, p_ls = self._stat_nchw(pstat, 0, 0.0), self._stat_nchw(pstat, 16, log_t)
            z = torch.empty_like(qs["z"])
            L.check(lib.cg_latent_mix(qs["z"].data_ptr(), qs["q_loc"].data_ptr(), qs["q_logscale"].data_ptr(),
                                      p_loc.data_ptr(), p_ls.data_ptr(), z.data_ptr(), z.numel(), float(alpha),
                                      float(t) if t is not None else 1.0, int(t is not None), s), "cg_latent_mix")
            out.append(z)
        return out

    # ------------------------------------------------------------------ stand-alone sub-module calls (inference)
    @torch.no_grad()
    def _call_encoder(self, x: Tensor) -> Dict[int, Tensor]:
        eng = self.engine()
        N = x.shape[0]
        prog = self._program(("encoder", N), lambda: eng.build_encoder_call(N))
        prog.x.copy_(x)
        eng.pack_weights()
        prog.run()
        return {res: t.clone() for res, t in prog.acts_out.items()}

    @torch.no_grad()
    def _call_decoder(self, parents: Tensor, x: Optional[Dict[int, Tensor]] = None, t: Optional[float] = None,
                      abduct: bool = False, latents: Sequence[Optional[Tensor]] = (),
                      eps: Optional[Sequence[Tensor]] = None):
        """Decoder.forward (src/vae.py:222-301) -> (h, stats) with the reference's stats layout"""
        eng = self.engine()
        N = parents.shape[0]
        nsto = sum(1 for d in eng.dec_layers if d.st.stochastic)
        has_acts = x is not None
        given = None if has_acts else tuple(i < len(latents) and latents[i] is not None for i in range(nsto))
        prog = self._program(("decoder", N, has_acts, given), lambda: eng.build_decoder_call(N, has_acts, given))
        drop = self.drop_cond() if (self.training and self.cond_prior and self.decoder.is_drop_cond) else None
        self._load_parents(prog, prog.io, [parents], drop)
        if has_acts:
            for res, buf in prog.acts_in.items():
                buf.copy_(x[res])
        D = prog.D
        for k, buf in D.z_in.items():
            buf.copy_(latents[k])
        log_t = math.log(t) if t is not None else 0.0
        self._set_noise(D, eps if eps is not None else self._host_eps(D), log_t)
        eng.pack_weights()
        prog.run()
        stats = []
        for k in range(nsto):
            if has_acts:  # src/vae.py:265-281
                st = {"kl": D.kl_elem[k].clone()}
                if abduct:
                    z = D.z_out[k].clone()
                    if self.cond_prior:
                        qstat, _ = D.stats_out[k]
                        z = {"z": z, "q_loc": self._stat_nchw(qstat, 0, 0.0), "q_logscale": self._stat_nchw(qstat, 16, log_t)}
                    st["z"] = z
                stats.append(st)
            elif abduct and self.cond_prior and k >= len(latents):  # src/vae.py:284-290: only the out-of-range branch records
                _, pstat = D.stats_out[k]
                stats.append({"z": {"p_loc": self._stat_nchw(pstat, 0, 0.0), "p_logscale": self._stat_nchw(pstat, 16, log_t)}})
        return prog.h_out.clone(), stats

    @torch.no_grad()
    def _call_likelihood(self, kind: str, h: Tensor, x: Optional[Tensor] = None):
        eng = self.engine()
        N = h.shape[0]
        prog = self._program(("likelihood", kind, N), lambda: eng.build_likelihood_call(N, kind))
        prog.h_in.copy_(h)
        if kind == "nll":
            prog.x.copy_(x)
            prog.nll.zero_()
        prog.run()
        return prog.nll.clone() if kind == "nll" else (prog.x_out.clone(), prog.scale_out.clone())

    def _stat_nchw(self, stat, c0: int, add: float) -> Tensor:
        N, H, W, _ = stat.t.shape
        out = torch.empty(N, 16, H, W, device=stat.t.device, dtype=torch.float32)
        L.check(L.load().cg_stats_to_nchw(stat.ptr, stat.ns, c0, float(add), out.data_ptr(), N, 16, H * W,
                                          _stream()), "cg_stats_to_nchw")
        return out

    def _host_eps(self, D):
        """explicit-eps programs without caller noise: draw with torch's device RNG like the reference
        (randn_like, src/vae.py:30)"""
        return [torch.randn_like(b) for b in D.eps if b is not None]

    @torch.no_grad()
    def _decode(self, latents: Sequence[Optional[Tensor]], plist: Sequence[Tensor], t: Optional[float],
                eps: Optional[Sequence[Tensor]]):
        eng = self.engine()
        N = plist[0].shape[0]
        nsto = sum(1 for d in eng.dec_layers if d.st.stochastic)
        given = tuple(i < len(latents) and latents[i] is not None for i in range(nsto))
        prog = self._program(("latents", N, given, len(plist)),
                             lambda: eng.build_decode(N, "latents", given=given, n_pa=len(plist)))
        self._load_parents(prog, prog.io, plist)
        log_t = math.log(t) if t is not None else 0.0
        ofs = 0
        for D in prog.Ds:
            for k, buf in D.z_in.items():
                buf.copy_(latents[k])
            n_eps = len([b for b in D.eps if b is not None])
            self._set_noise(D, eps[ofs: ofs + n_eps] if eps is not None else self._host_eps(D), log_t)
            ofs += n_eps
        eng.pack_weights()
        prog.run()
        return prog

    def forward_latents(self, latents: List[Tensor], parents: Tensor, t: Optional[float] = None,
                        eps: Optional[Sequence[Tensor]] = None) -> Tuple[Tensor, Tensor]:
        latents = [z["z"] if isinstance(z, dict) else z for z in latents]
        prog = self._decode(latents, [parents], t, eps)
        return prog.x_out[0].clone(), prog.scale_out[0].clone()

    def sample(self, parents: Tensor, return_loc: bool = True, t: Optional[float] = None,
               eps: Optional[Sequence[Tensor]] = None) -> Tuple[Tensor, Tensor]:
        if not return_loc:
            raise NotImplementedError("return_loc=False hits a reference bug (t passed as x, src/vae.py:419 vs "
                                      ":352); use return_loc=True as every reference caller does")
        prog = self._decode([], [parents], t, eps)
        return prog.x_out[0].clone(), prog.scale_out[0].clone()


# ---------------------------------------------------------------------------------------------
# DSCM hot lines (src/pgm/dscm.py)
# ---------------------------------------------------------------------------------------------
# UKBB attribute ranges (max, min) of the PGM's [-1,1] normalisation (src/datasets.py:89-98) and the log-standardisation
# constants the released UKBB HVAE was trained with (src/pgm/dscm.py:112-117)
_UKBB_MAX_MIN = {"age": (73.0, 44.0), "brain_volume": (1629520.0, 841919.0),
                 "ventricle_volume": (157075.0, 7613.27001953125)}
_UKBB_LOG_STD = {"age": (4.112339973449707, 0.11769197136163712),
                 "brain_volume": (13.965583801269531, 0.09537758678197861),
                 "ventricle_volume": (10.345998764038086, 0.43127763271331787)}


def ukbb_preprocess(pa: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """src/pgm/dscm.py:98-118: undo the PGM's [-1,1] normalisation of the continuous UKBB parents, then
    log-standardise them the way the UKBB HVAE saw them in training.  Host-side control plane (B scalars per
    attribute).  Like the reference, an attribute other than mri_seq / sex / the three known ones cannot be mapped
    back (the reference fails on `_max, _min = None`); here that is a KeyError naming the attribute."""
    out = dict(pa)
    for k, v in pa.items():
        if k in ("mri_seq", "sex"):
            continue
        if k not in _UKBB_MAX_MIN:
            raise KeyError(f"ukbb_preprocess: no (max, min) statistics for parent '{k}' (src/datasets.py:89-98)")
        mx, mn = _UKBB_MAX_MIN[k]
        out[k] = ((v + 1) / 2) * (mx - mn) + mn
    for k, v in out.items():
        if k in _UKBB_LOG_STD:
            mu, sd = _UKBB_LOG_STD[k]
            out[k] = (torch.log(v.clamp(min=1e-12)) - mu) / sd
    return out


def vae_preprocess(args, pa: Dict[str, Tensor], expand: bool = False) -> Tensor:
    """src/pgm/dscm.py:121-132.  Returns (B, ctx) on the GPU -- every entry point of ``HVAE`` accepts that or the
    reference's materialised (B, ctx, R, R) form (``expand=True`` produces it)."""
    if "ukbb" in getattr(args, "dataset", ""):
        pa = ukbb_preprocess(pa)
    cols = [pa[k] if pa[k].dim() > 1 else pa[k][..., None] for k in args.parents_x]
    out = torch.cat(cols, dim=1)
    if expand:
        out = out[..., None, None].repeat(1, 1, args.input_res, args.input_res)
    return (out.cuda() if torch.cuda.is_available() else out).float()  # host glue; every HVAE entry needs the GPU


def counterfactual(vae: HVAE, x: Tensor, pa: Tensor, cf_pa: Tensor, t_abduct: float = 1.0, particles: int = 1,
                   eps: Optional[Sequence[Sequence[Tensor]]] = None):
    """abduct -> forward_latents(cf_pa) & forward_latents(pa) (one 2-parent-set pass) -> combine
    (src/pgm/dscm.py:47-72).  Returns (cf_x, var_cf_x or None).

    With ``vae.train()``, autograd enabled and trainable HVAE parameters (counterfactual fine-tuning, src/pgm/train_cf.py:159-180)
    ``cf_x`` is differentiable w.r.t. the HVAE parameters: ``aux_loss(cf_x).backward()`` reaches encoder, posterior,
    prior, decoder and likelihood weights exactly as the reference's autograd graph does (src/pgm/dscm.py:52-56,78-88).
    Gradient support covers particles == 1 (the reference's training setting) and in-kernel noise."""
    # the reference fine-tunes with vae.train() and autograd on, and evaluates under vae.eval() + no_grad
    # (src/pgm/train_cf.py:122,144,182): the differentiable path is taken in exactly that training situation
    need_grad = (torch.is_grad_enabled() and vae.training and eps is None and not TRACE_ONLY
                 and any(p.requires_grad for p in vae.parameters()))
    if need_grad:
        if particles != 1:
            raise NotImplementedError("gradients through the counterfactual are implemented for cf_particles == 1 (the "
                                      "reference's training setting, src/pgm/train_cf.py); use torch.no_grad() for "
                                      "multi-particle uncertainty estimates")
        return _CfFn.apply(vae, x, _pa_vector(pa), _pa_vector(cf_pa), t_abduct, *list(vae.parameters())), None
    with torch.no_grad():
        return _counterfactual_nograd(vae, x, pa, cf_pa, t_abduct, particles, eps)


def _counterfactual_nograd(vae: HVAE, x: Tensor, pa: Tensor, cf_pa: Tensor, t_abduct, particles, eps):
    if eps is None and not TRACE_ONLY:
        # serving path: one fused program (abduction with in-kernel Philox noise, both decodes on the bf16 latents in
        # place, combine); the explicit-eps path below keeps the reference's fp32 latent interface for parity tests
        eng = vae.engine()
        N = x.shape[0]
        prog = vae._program(("cf", N), lambda: eng.build_counterfactual(N))
        prog.io.x.copy_(x)
        vae._load_parents(prog, prog.io, [pa, cf_pa])
        vae._set_noise(prog.D, None, math.log(t_abduct) if t_abduct is not None else 0.0)
        if particles > 1:
            prog.acc.zero_()
            prog.acc2.zero_()
        eng.pack_weights()
        for _ in range(particles):
            prog.run()
            prog.seed_ctr.add_(1)
        if particles > 1:
            mean = prog.acc / particles
            return mean, (prog.acc2 - prog.acc ** 2 / particles) / particles  # src/pgm/dscm.py:68
        return prog.cf_x.clone(), None
    lib = L.load()
    n = x.numel()
    acc = torch.zeros_like(x) if particles > 1 else None
    acc2 = torch.zeros_like(x) if particles > 1 else None
    cf_x = torch.empty_like(x)
    for i in range(particles):
        zs = vae.abduct(x, pa, t=t_abduct, eps=None if eps is None else eps[i])
        zs = [z["z"] if isinstance(z, dict) else z for z in zs]
        prog = vae._decode(zs, [cf_pa, pa], None, None)
        s = _stream()
        L.check(lib.cg_cf_combine(x.data_ptr(), prog.x_out[1].data_ptr(), prog.scale_out[1].data_ptr(),
                                  prog.x_out[0].data_ptr(), prog.scale_out[0].data_ptr(), cf_x.data_ptr(),
                                  acc.data_ptr() if acc is not None else None,
                                  acc2.data_ptr() if acc2 is not None else None, n, s), "cg_cf_combine")
    if particles > 1:
        mean = acc / particles
        var = (acc2 - acc ** 2 / particles) / particles  # src/pgm/dscm.py:68
        return mean, var
    return cf_x, None


class CounterfactualGraph:
    """``counterfactual`` for a fixed batch size replayed from one CUDA graph (inference serving path).

    The eager helper issues ~1500 small launches per call and is CPU-bound at small batches; the graph replays the
    identical launch sequence (the fused program of ``Engine.build_counterfactual``: abduction, two prior-only decodes on
    the shared bf16 latents, combine) from static buffers.  Noise (the reference's ``randn_like``, src/vae.py:30) is drawn
    inside the latent kernels from a Philox stream keyed by a host seed captured at build time plus the device-side
    ``seed_ctr`` that the captured ``add_`` advances, so every replay sees fresh eps without host involvement."""

    def __init__(self, vae: HVAE, batch: int, t_abduct: float = 1.0, particles: int = 1):
        eng = vae.engine()
        dev = eng.device
        self.vae, self.t, self.particles = vae, t_abduct, particles
        self.x = torch.zeros(batch, eng.C, eng.R, eng.R, device=dev)
        self.pa = torch.zeros(batch, eng.ctx, device=dev)
        self.cf_pa = torch.zeros(batch, eng.ctx, device=dev)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # builds the programs and sets kernel attributes outside the capture
            counterfactual(vae, self.x, self.pa, self.cf_pa, t_abduct, particles)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.cf_x, self.var = counterfactual(vae, self.x, self.pa, self.cf_pa, t_abduct, particles)

    @torch.no_grad()
    def __call__(self, x: Tensor, pa: Tensor, cf_pa: Tensor):
        self.x.copy_(x, non_blocking=True)
        self.pa.copy_(_pa_vector(pa), non_blocking=True)
        self.cf_pa.copy_(_pa_vector(cf_pa), non_blocking=True)
        self.graph.replay()
        return self.cf_x, self.var
