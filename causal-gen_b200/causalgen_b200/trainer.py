"""Fused ELBO training step: the public training API of this package.

One `Trainer.step(x_uint8, parents)` call = reference `trainer.run_epoch` body for one batch
(src/trainer.py:50-87): preprocess (uint8 -> [-1,1]), HVAE forward, hand-written backward, data-parallel
all-reduce of ONE flat gradient bucket (NCCL, sum then 1/world inside the optimiser kernel), global-norm
clip + NaN/grad-skip test, AdamW with linear warm-up, EMA -- all on the device, no host sync inside the
step.  Everything between the H2D copy and the loss read-back is kernels of libcausalgen_b200.so plus one
NCCL collective; the steady-state step is replayed from CUDA graphs.

Reference behaviours kept (citations relative to /root/reference):
  * beta warm-up (src/trainer.py:52-57): beta is a DEVICE scalar read by the kernels, so it changes under graph replay;
  * gradient accumulation (src/trainer.py:62-66): the loss is scaled by 1/accu_steps and the optimiser runs when
    ``i % accu_steps == 0`` (i = 0-based iteration of the epoch; quirk Q8: i = 0 updates too);
  * frozen likelihood parameters (src/vae.py:340-349, x_like fixed_/shared_ with std_init > 0) get no gradient, no weight
    decay and no Adam state, exactly like AdamW over ``grad is None`` parameters: they sit at the tail of the flat
    buffers and the optimiser kernels stop before them;
  * conditioning dropout (src/vae.py:234-249, morphomnist only): drawn on the host CPU RNG per step, delivered as a
    device scalar, so that config replays from the graph as well.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import _lib as L
from . import dp
from .engine import TRACE_ONLY, Engine, lane_stream, no_grad_param_ids, ordered_params
from .hvae import HVAE, _stream


class Trainer:
    def __init__(self, model: HVAE, batch_size: int, lr: float = 1e-3, wd: float = 0.01, betas=(0.9, 0.9),
                 lr_warmup_steps: int = 100, grad_clip: float = 350.0, grad_skip: float = 500.0,
                 ema_rate: float = 0.999, beta: float = 1.0, use_graph: bool = True, noise_seed: int = 7,
                 ema_update_after: int = 100, accu_steps: int = 1, beta_warmup_steps: int = 0,
                 export_eps: bool = False):
        self.model = model
        self.N = batch_size
        self.hp = dict(lr=lr, wd=wd, b1=betas[0], b2=betas[1], warmup=lr_warmup_steps, clip=grad_clip,
                       skip=grad_skip, ema=ema_rate, ema_after=ema_update_after)
        self.beta_target = float(beta)
        self.beta = float(beta)
        self.beta_warmup_steps = int(beta_warmup_steps)
        self.accu_steps = max(1, int(accu_steps))
        self.world, self.rank = dp.world_info()
        dev = next(model.parameters()).device
        self.device = dev
        # flatten parameters (trainable first, frozen last): AdamW / EMA run over one buffer; nn.Parameters become
        # views of it.  The engine lays its gradient bucket out in the same order.
        self.params: List[torch.nn.Parameter] = ordered_params(model)
        n = sum(p.numel() for p in self.params)
        dead = no_grad_param_ids(model)
        self.n_train = sum(p.numel() for p in self.params if p.requires_grad and id(p) not in dead)
        self.flat_p = torch.empty(n, device=dev, dtype=torch.float32)
        off = 0
        for p in self.params:
            self.flat_p[off: off + p.numel()].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off: off + p.numel()].view_as(p)
            off += p.numel()
        dp.broadcast_params_(self.flat_p)  # identical initial weights on every rank
        self.m = torch.zeros(self.n_train, device=dev, dtype=torch.float32)
        self.v = torch.zeros(self.n_train, device=dev, dtype=torch.float32)
        self.ema = self.flat_p.clone()
        self.state = torch.zeros(4, dtype=torch.int32, device=dev)
        self.dyn = torch.zeros(6, dtype=torch.float32, device=dev)
        self.gsumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.gscratch = torch.zeros(148 * 4, dtype=torch.float32, device=dev)  # per-block partials (deterministic norm)
        self.seed_ctr = torch.zeros(1, dtype=torch.int64, device=dev)
        model.train()
        self.eng: Engine = model.engine()
        assert [id(p) for p in self.eng.params] == [id(p) for p in self.params]
        self.prog = model._program(("elbo", self.N, True, False), lambda: self.eng.build_elbo(self.N, True, False))
        self.eng.set_hyper(self.prog, beta=self.beta, drop_sto=1.0)
        self.grad = self.eng.flat_grad[: self.n_train]  # the bucket that is all-reduced / clipped / applied
        # noise: Philox keyed by (seed + device step counter, rank-disjoint stream)
        base = dp.rank_noise_seed(noise_seed, self.rank)
        for la in self.prog.D.latent_args:
            la.seed, la.seed_dev = base, self.seed_ctr.data_ptr()
        for lb in self.prog.D.latent_bwd_args:
            lb.seed, lb.seed_dev = base, self.seed_ctr.data_ptr()
        # parity hook: every latent kernel also writes the eps it drew (fp32 NCHW, reference RNG order) so a CPU oracle
        # can be fed exactly the same noise
        self.eps_out: Optional[List[torch.Tensor]] = None
        if export_eps:
            self.eps_out = []
            for la in self.prog.D.latent_args:
                if la.mode in (0, 1):
                    r = int(round(la.HW ** 0.5))
                    t = torch.zeros(self.N, la.zdim, r, r, device=dev, dtype=torch.float32)
                    la.eps_out = t.data_ptr()
                    self.eps_out.append(t)
        self.x8 = torch.zeros(self.N, self.eng.C, self.eng.R, self.eng.R, dtype=torch.uint8, device=dev)
        self.loss_host = torch.zeros(3, dtype=torch.float32).pin_memory() if dev.type == "cuda" else torch.zeros(3)
        self.use_graph = use_graph and not TRACE_ONLY and not getattr(self.prog, "no_graph", False)
        self.g_fb: Optional[torch.cuda.CUDAGraph] = None
        self.g_opt: Optional[torch.cuda.CUDAGraph] = None
        self.steps_done = 0        # micro-batches seen
        self.iter_in_epoch = 0     # `i` of src/trainer.py:50
        self.kernels_per_step = self.prog.n_kernels + 1 + 1 + 4  # + pack, normalise, sumsq x2 / advance / adamw

    # ------------------------------------------------------------------ pieces
    def _fwd_bwd(self):
        """normalise -> ELBO forward -> backward; gradients ACCUMULATE into the flat bucket (zeroed by `_optim`)"""
        lib = L.load()
        s = _stream()
        prog = self.prog
        L.check(lib.cg_normalise_u8(self.x8.data_ptr(), prog.io.x.data_ptr(), self.x8.numel(), s), "cg_normalise_u8")
        for t in prog.zero:
            t.zero_()
        self.eng.pack_weights(s)
        prog.run(s)
        self.seed_ctr.add_(1)

    def _optim(self):
        lib = L.load()
        s = _stream()
        hp = self.hp
        g = self.grad
        L.check(lib.cg_sumsq(g.data_ptr(), self.gsumsq.data_ptr(), g.numel(), self.gscratch.data_ptr(),
                             self.gscratch.numel(), s), "cg_sumsq")
        # 1/world (data parallel mean) and 1/accu_steps (src/trainer.py:63) are folded into the clip coefficient
        scale = 1.0 / (self.world * self.accu_steps)
        L.check(lib.cg_optim_advance(self.state.data_ptr(), self.dyn.data_ptr(), self.gsumsq.data_ptr(),
                                     self.prog.out3.data_ptr(), hp["lr"], hp["warmup"], hp["b1"], hp["b2"], hp["clip"],
                                     hp["skip"], scale, hp["ema"], hp["ema_after"], s), "cg_optim_advance")
        L.check(lib.cg_adamw_ema_step(self.flat_p.data_ptr(), g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                      self.ema.data_ptr(), g.numel(), self.state.data_ptr(), self.dyn.data_ptr(),
                                      hp["b1"], hp["b2"], 1e-8, hp["wd"], s), "cg_adamw_ema_step")
        self.eng.flat_grad.zero_()  # model.zero_grad() of src/trainer.py:87 (also after a skipped update)

    def _load_inputs(self, x8: torch.Tensor, pa: torch.Tensor):
        self.x8.copy_(x8, non_blocking=True)
        self.prog.io.pa_in[0].copy_(pa if pa.dim() == 2 else pa[:, :, 0, 0], non_blocking=True)
        drop = None
        if self.model.decoder.is_drop_cond and self.model.cond_prior:
            drop = float(self.model.drop_cond()[0])
        if self.beta_warmup_steps > 0:  # src/trainer.py:52-57 (args.iter is 1-based)
            it = self.steps_done + 1
            self.beta = self.beta_target * (1.0 if it > self.beta_warmup_steps else it / self.beta_warmup_steps)
        self.eng.set_hyper(self.prog, beta=self.beta, drop_sto=drop)

    def _capture(self):
        """Capture fwd+bwd and the optimiser tail as two graphs.  No warm-up pass here: the eager first step already
        set every kernel attribute, and a warm-up would apply an extra (un-reduced) update."""
        torch.cuda.synchronize()
        saved = (self.seed_ctr.clone(), self.eng.flat_grad.clone())
        self.g_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_fb, stream=lane_stream()):  # lane 0 of the program = the capture stream
            self._fwd_bwd()
        self.g_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_opt):
            self._optim()
        torch.cuda.synchronize()
        # capture does not execute, but keep the invariant explicit: state is exactly what it was before
        self.seed_ctr.copy_(saved[0])
        self.eng.flat_grad.copy_(saved[1])

    # ------------------------------------------------------------------ public
    def set_beta(self, beta: float):
        """change beta (KL weight) for the following steps; works under graph replay (device scalar)"""
        self.beta_target = self.beta = float(beta)

    def start_epoch(self):
        self.iter_in_epoch = 0

    def invalidate_graphs(self):
        """drop the captured graphs (optimiser hyper-parameters are baked into them); the next step re-captures"""
        self.g_fb = self.g_opt = None

    def step_device(self, x8_dev: torch.Tensor, pa_dev: torch.Tensor) -> torch.Tensor:
        """inputs already resident on the device; returns the device tensor {elbo, nll, kl} (overwritten by the next
        step)"""
        self._load_inputs(x8_dev, pa_dev)
        if self.use_graph and self.g_fb is None and self.steps_done >= 1:
            self._capture()
        if self.g_fb is not None:
            self.g_fb.replay()
        else:
            self._fwd_bwd()
        if self.iter_in_epoch % self.accu_steps == 0:  # src/trainer.py:66 (quirk Q8)
            dp.reduce_gradients_(self.grad)  # the one exchange step of the path (NCCL over NVLink)
            if self.g_opt is not None:
                self.g_opt.replay()
            else:
                self._optim()
        self.steps_done += 1
        self.iter_in_epoch += 1
        return self.prog.out3

    def step(self, x8_host: torch.Tensor, pa_host: torch.Tensor) -> torch.Tensor:
        """x8_host (B,C,R,R) uint8 and pa_host (B,ctx) fp32 in (pinned) host memory -> host {elbo,nll,kl} (a fresh
        tensor per call).  H2D copy of the batch and D2H read of the loss are part of the call."""
        out = self.step_device(x8_host, pa_host)
        self.loss_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.loss_host.clone()

    def grad_norm(self) -> float:
        return float(self.dyn[5])

    def skipped_updates(self) -> int:
        return int(self.state[2])

    def _views(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        """flat buffer (trainer order) -> {reference state_dict key: tensor}; shorter buffers cover trainable params only"""
        name_of = {id(p): k for k, p in self.model.named_parameters()}
        out, off = {}, 0
        for p in self.params:
            if off + p.numel() > flat.numel():
                break
            out[name_of[id(p)]] = flat[off: off + p.numel()].view_as(p)
            off += p.numel()
        return out

    def ema_state_dict(self):
        """EMA weights under the reference's state_dict keys, in the reference's key order (src/trainer.py:161)"""
        v = self._views(self.ema)
        return {k: v[k].clone() for k, _ in self.model.named_parameters()}

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """everything a resumed run needs beyond the model weights (Adam moments, EMA, counters, noise counter)"""
        return dict(m=self.m.clone(), v=self.v.clone(), ema=self.ema.clone(), state=self.state.clone(),
                    dyn=self.dyn.clone(), seed_ctr=self.seed_ctr.clone(),
                    steps_done=torch.tensor(self.steps_done), iter_in_epoch=torch.tensor(self.iter_in_epoch),
                    beta=torch.tensor(self.beta), beta_target=torch.tensor(self.beta_target))

    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        for k in ("m", "v", "ema", "state", "dyn", "seed_ctr"):
            getattr(self, k).copy_(sd[k])
        self.steps_done = int(sd["steps_done"])
        self.iter_in_epoch = int(sd["iter_in_epoch"])
        self.beta, self.beta_target = float(sd["beta"]), float(sd["beta_target"])
