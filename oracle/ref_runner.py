"""Runs the UNMODIFIED reference (staged under baseline/_ref by oracle/stage_reference.py) for bench.py's reference
arms.  TEST / BENCH INFRASTRUCTURE ONLY -- never imported by the product package.

What it does with the reference is what the reference's own entry points do:
  * model construction: hps.add_arguments / HPARAMS_REGISTRY + the launcher flag sets (src/run_local.sh:3-15,
    src/run_slurm.sh:23-52), `model.apply(init_bias)` (src/main.py:51-55);
  * train step: src/trainer.py:62-87 with AdamW + LambdaLR (src/train_setup.py:42-53) and utils.EMA;
  * counterfactual pass: src/pgm/dscm.py:52-56 restated over the imported HVAE (dscm.py itself needs Pyro).
"""
import argparse
import os
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = os.path.join(ROOT, "baseline", "_ref", "src")

FLAGS = {
    "morphomnist": ("morphomnist", ["--context_dim", "12", "--cond_prior"]),
    "cmnist": ("cmnist", ["--context_dim", "20"]),
    "ukbb192": ("ukbb192", ["--context_dim", "4", "--z_max_res", "96", "--beta", "5"]),
    "mimic192": ("mimic192", ["--context_dim", "6", "--z_max_res", "96", "--beta", "9"]),
    "mimic224": ("mimic192", ["--input_res", "224", "--enc_arch", "224b1d2,112b3d2,56b7d2,28b11d2,14b7d2,7b3d7,1b2",
                              "--dec_arch", "1b2,8b4,14b8,28b12,56b8,112b4,224b2", "--context_dim", "6",
                              "--z_max_res", "112", "--beta", "9"]),
}


def available() -> bool:
    return os.path.exists(os.path.join(REF_SRC, "vae.py"))


def _import():
    if not available():
        raise RuntimeError("reference not staged (baseline/_ref/src): run __graft_entry__.build() in the build container")
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    for stub in ("imageio", "send2trash"):  # only needed by the image-grid / directory helpers, never on this path
        sys.modules.setdefault(stub, types.ModuleType(stub))
    import hps
    import utils
    import vae
    return hps, vae, utils


def build(name: str, device="cpu", x_like=None):
    hps, vae, utils = _import()
    hps_name, extra = FLAGS[name]
    p = argparse.ArgumentParser()
    hps.add_arguments(p)
    p.set_defaults(**hps.HPARAMS_REGISTRY[hps_name].__dict__)
    a = hps.Hparams()
    a.update(p.parse_args(["--hps", hps_name] + extra + (["--x_like", x_like] if x_like else [])).__dict__)
    torch.manual_seed(7)
    model = vae.HVAE(a)
    if x_like is not None and x_like.endswith("dmol"):
        import dmol
        model.likelihood = dmol.DmolNet(a)
    for m in model.modules():  # init_bias, src/main.py:51-55
        if isinstance(m, torch.nn.Conv2d):
            torch.nn.init.zeros_(m.bias)
    return a, model.to(device), utils


class RefTrainStep:
    """src/trainer.py:62-87 on `device` (cpu: all host threads; cuda: eager PyTorch/cuDNN, optional bf16 autocast)"""

    def __init__(self, name, device="cpu", autocast_bf16=False, tf32=True, x_like=None):
        self.a, self.model, utils = build(name, device, x_like)
        self.device, self.autocast = device, autocast_bf16
        if device != "cpu":
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.benchmark = True
        self.ema = utils.EMA(self.model, beta=self.a.ema_rate)
        self.opt = torch.optim.AdamW(self.model.parameters(), lr=self.a.lr, weight_decay=self.a.wd, betas=self.a.betas)
        self.sched = torch.optim.lr_scheduler.LambdaLR(self.opt, lr_lambda=utils.linear_warmup(self.a.lr_warmup_steps))
        self.model.train()

    def __call__(self, x8, pa):
        a = self.a
        x = (x8.to(self.device).float() - 127.5) / 127.5
        pa = pa.to(self.device).float()[..., None, None].repeat(1, 1, a.input_res, a.input_res)
        self.model.zero_grad(set_to_none=True)
        if self.autocast:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = self.model(x, pa, beta=a.beta)
        else:
            out = self.model(x, pa, beta=a.beta)
        out["elbo"].backward()
        gn = torch.nn.utils.clip_grad_norm_(self.model.parameters(), a.grad_clip)
        if gn < a.grad_skip and not torch.isnan(out["nll"]) and not torch.isnan(out["kl"]):
            self.opt.step()
            self.sched.step()
            self.ema.update()
        return out


@torch.no_grad()
def ref_counterfactual(model, x, pa_full, cf_full, t_abduct=1.0):
    """src/pgm/dscm.py:52-56 over the imported reference HVAE"""
    zs = model.abduct(x, parents=pa_full, t=t_abduct)
    if model.cond_prior:
        zs = [z["z"] for z in zs]
    cf_loc, cf_scale = model.forward_latents(zs, parents=cf_full)
    rec_loc, rec_scale = model.forward_latents(zs, parents=pa_full)
    u = (x - rec_loc) / rec_scale.clamp(min=1e-12)
    return torch.clamp(cf_loc + cf_scale * u, min=-1, max=1)


def time_steps(fn, sync, budget_s, min_steps=2, max_steps=8, warmup=1):
    for _ in range(warmup):
        fn()
    sync()
    times, t_start = [], time.time()
    while True:
        t0 = time.time()
        fn()
        sync()
        times.append(time.time() - t0)
        if (time.time() - t_start > budget_s and len(times) >= min_steps) or len(times) >= max_steps:
            break
    return float(np.median(times)), len(times)
