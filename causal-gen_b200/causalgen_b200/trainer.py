"""Fused ELBO training step: the public training API of this package.

One `Trainer.step(x_uint8, parents)` call = reference `trainer.run_epoch` body for one batch
(src/trainer.py:50-87): preprocess (uint8 -> [-1,1]), HVAE forward, hand-written backward, data-parallel
all-reduce of ONE flat gradient bucket (NCCL, sum then 1/world inside the optimiser kernel), global-norm
clip + NaN/grad-skip test, AdamW with linear warm-up, EMA -- all on the device, no host sync inside the
step.  Everything between the H2D copy and the loss read-back is kernels of libcausalgen_b200.so plus one
NCCL collective; the steady-state step is replayed from CUDA graphs.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib as L
from . import dp
from .engine import TRACE_ONLY
from .hvae import HVAE, _stream


class Trainer:
    def __init__(self, model: HVAE, batch_size: int, lr: float = 1e-3, wd: float = 0.01, betas=(0.9, 0.9),
                 lr_warmup_steps: int = 100, grad_clip: float = 350.0, grad_skip: float = 500.0,
                 ema_rate: float = 0.999, beta: float = 1.0, use_graph: bool = True, noise_seed: int = 7,
                 ema_update_after: int = 100):
        self.model = model
        self.N = batch_size
        self.hp = dict(lr=lr, wd=wd, b1=betas[0], b2=betas[1], warmup=lr_warmup_steps, clip=grad_clip,
                       skip=grad_skip, ema=ema_rate, ema_after=ema_update_after)
        self.beta = float(beta)
        self.world, self.rank = dp.world_info()
        dev = next(model.parameters()).device
        self.device = dev
        # flatten parameters: AdamW / EMA run over one buffer; nn.Parameters become views of it
        params = list(model.parameters())
        n = sum(p.numel() for p in params)
        self.flat_p = torch.empty(n, device=dev, dtype=torch.float32)
        off = 0
        for p in params:
            self.flat_p[off: off + p.numel()].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off: off + p.numel()].view_as(p)
            off += p.numel()
        dp.broadcast_params_(self.flat_p)  # identical initial weights on every rank
        self.m = torch.zeros_like(self.flat_p)
        self.v = torch.zeros_like(self.flat_p)
        self.ema = self.flat_p.clone()
        self.state = torch.zeros(4, dtype=torch.int32, device=dev)
        self.dyn = torch.zeros(6, dtype=torch.float32, device=dev)
        self.gsumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.seed_ctr = torch.zeros(1, dtype=torch.int64, device=dev)
        model.train()
        self.eng = model.engine()
        self.prog = model._program(("elbo", self.N, True, False), lambda: self.eng.build_elbo(self.N, True, False))
        self.eng.set_beta(self.prog, self.beta, self.N)
        # noise: Philox keyed by (seed + device step counter, rank-disjoint stream)
        base = dp.rank_noise_seed(noise_seed, self.rank)
        for la in self.prog.D.latent_args:
            la.seed, la.seed_dev = base, self.seed_ctr.data_ptr()
        for lb in self.prog.D.latent_bwd_args:
            lb.seed, lb.seed_dev = base, self.seed_ctr.data_ptr()
        self.x8 = torch.zeros(self.N, self.eng.C, self.eng.R, self.eng.R, dtype=torch.uint8, device=dev)
        self.loss_host = torch.zeros(3, dtype=torch.float32).pin_memory() if dev.type == "cuda" else torch.zeros(3)
        self.use_graph = use_graph and not TRACE_ONLY and not model.decoder.is_drop_cond
        self.g_fb: Optional[torch.cuda.CUDAGraph] = None
        self.g_opt: Optional[torch.cuda.CUDAGraph] = None
        self.steps_done = 0
        self.kernels_per_step = self.prog.n_kernels + 1 + 1 + 3  # + pack, normalise, sumsq/advance/adamw

    # ------------------------------------------------------------------ pieces
    def _fwd_bwd(self):
        lib = L.load()
        s = _stream()
        prog = self.prog
        L.check(lib.cg_normalise_u8(self.x8.data_ptr(), prog.io.x.data_ptr(), self.x8.numel(), s), "cg_normalise_u8")
        for t in prog.zero:
            t.zero_()
        self.eng.flat_grad.zero_()
        self.eng.pack_weights(s)
        prog.run(s)
        self.seed_ctr.add_(1)

    def _optim(self):
        lib = L.load()
        s = _stream()
        hp = self.hp
        g = self.eng.flat_grad
        self.gsumsq.zero_()
        L.check(lib.cg_sumsq(g.data_ptr(), self.gsumsq.data_ptr(), g.numel(), s), "cg_sumsq")
        L.check(lib.cg_optim_advance(self.state.data_ptr(), self.dyn.data_ptr(), self.gsumsq.data_ptr(),
                                     self.prog.out3.data_ptr(), hp["lr"], hp["warmup"], hp["b1"], hp["b2"], hp["clip"],
                                     hp["skip"], 1.0 / self.world, hp["ema"], hp["ema_after"], s), "cg_optim_advance")
        L.check(lib.cg_adamw_ema_step(self.flat_p.data_ptr(), g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                      self.ema.data_ptr(), g.numel(), self.state.data_ptr(), self.dyn.data_ptr(),
                                      hp["b1"], hp["b2"], 1e-8, hp["wd"], s), "cg_adamw_ema_step")

    def _load_parents(self, pa: torch.Tensor):
        self.prog.io.pa_in[0].copy_(pa if pa.dim() == 2 else pa[:, :, 0, 0], non_blocking=True)
        if self.model.decoder.is_drop_cond and self.model.cond_prior:
            drop = self.model.drop_cond()
            for ln in self.prog.io.drop_launch:
                a = list(ln.args)
                a[-1] = C.c_float(float(drop[0]))
                ln.args = tuple(a)

    def _capture(self):
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._fwd_bwd()  # warm-up on the capture stream (cudaFuncSetAttribute etc. happen here)
            self._optim()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.g_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_fb):
            self._fwd_bwd()
        self.g_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_opt):
            self._optim()

    # ------------------------------------------------------------------ public
    def step_device(self, x8_dev: torch.Tensor, pa_dev: torch.Tensor) -> torch.Tensor:
        """inputs already resident on the device; returns the device tensor {elbo, nll, kl}"""
        self.x8.copy_(x8_dev, non_blocking=True)
        self._load_parents(pa_dev)
        if self.use_graph and self.g_fb is None and self.steps_done >= 1:
            self._capture()
        if self.g_fb is not None:
            self.g_fb.replay()
        else:
            self._fwd_bwd()
        dp.reduce_gradients_(self.eng.flat_grad)  # the one exchange step of the path (NCCL over NVLink)
        if self.g_opt is not None:
            self.g_opt.replay()
        else:
            self._optim()
        self.steps_done += 1
        return self.prog.out3

    def step(self, x8_host: torch.Tensor, pa_host: torch.Tensor) -> torch.Tensor:
        """x8_host (B,C,R,R) uint8 and pa_host (B,ctx) fp32 in (pinned) host memory -> host {elbo,nll,kl}.
        H2D copy of the batch and D2H read of the loss are part of the call."""
        out = self.step_device(x8_host, pa_host)
        self.loss_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.loss_host

    def grad_norm(self) -> float:
        return float(self.dyn[5])

    def skipped_updates(self) -> int:
        return int(self.state[2])

    def ema_state_dict(self):
        """EMA weights under the reference's state_dict keys (src/trainer.py:161)"""
        out, off = {}, 0
        for (k, p) in self.model.named_parameters():
            out[k] = self.ema[off: off + p.numel()].view_as(p).clone()
            off += p.numel()
        return out
