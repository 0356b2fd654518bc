"""Checkpoint wire format of the reference, both ways (src/trainer.py:154-168 writes it, src/main.py:26-38,74-90
reads it): a ``torch.save``-d dict with

    epoch, step, best_loss, model_state_dict, ema_model_state_dict, optimizer_state_dict, scheduler_state_dict, hparams

``optimizer_state_dict`` / ``scheduler_state_dict`` are produced and consumed THROUGH real ``torch.optim.AdamW`` /
``LambdaLR`` objects built the way src/train_setup.py:42-53 builds them, so the format is whatever the installed torch
writes -- a checkpoint saved here loads into the reference's ``optimizer.load_state_dict`` and a reference checkpoint
loads here.  The Adam moments live in the trainer's flat fp32 buffers; this module only maps them to per-parameter
tensors (parameter index = position in ``model.parameters()``, which is what torch uses).

One extra key, ``causalgen_b200_state`` (ignored by the reference, which reads known keys only), carries what the
reference format cannot: EMA call counter, noise counter, skipped-update count -- for an exact resume.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch


def _linear_warmup(warmup_iters):  # src/utils.py:32-36
    def f(it):
        return 1.0 if it > warmup_iters else it / warmup_iters
    return f


def _torch_optimizer(trainer, lr: Optional[float] = None):
    """AdamW + LambdaLR over the model's parameters exactly as src/train_setup.py:42-53 sets them up"""
    hp = trainer.hp
    opt = torch.optim.AdamW(trainer.model.parameters(), lr=hp["lr"] if lr is None else lr, weight_decay=hp["wd"],
                            betas=(hp["b1"], hp["b2"]))
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=_linear_warmup(hp["warmup"]) if hp["warmup"] > 0
                                              else (lambda it: 1.0))
    return opt, sched


def to_checkpoint(trainer, hparams: Optional[Dict[str, Any]] = None, epoch: int = 0, step: Optional[int] = None,
                  best_loss: float = float("inf")) -> Dict[str, Any]:
    """the dict src/trainer.py:154-165 saves, from a causalgen_b200 Trainer"""
    model = trainer.model
    t = int(trainer.state[0])  # optimizer.step() calls so far
    opt, sched = _torch_optimizer(trainer)
    m, v = trainer._views(trainer.m), trainer._views(trainer.v)
    name_of = {id(p): k for k, p in model.named_parameters()}
    if t > 0:
        for p in model.parameters():
            k = name_of[id(p)]
            if k in m:  # parameters that receive gradients; frozen / never-used ones have no state in torch either
                opt.state[p] = {"step": torch.tensor(float(t)), "exp_avg": m[k].detach().clone(),
                                "exp_avg_sq": v[k].detach().clone()}
    # scheduler.step() has been called t times
    sched.last_epoch = t
    sched._step_count = t + 1
    lr_now = [base * fn(t) for base, fn in zip(sched.base_lrs, sched.lr_lambdas)]
    for g, lr in zip(opt.param_groups, lr_now):
        g["lr"] = lr
    sched._last_lr = lr_now
    return {
        "epoch": int(epoch),
        "step": int(trainer.steps_done if step is None else step),
        "best_loss": float(best_loss),
        "model_state_dict": {k: w.detach().clone() for k, w in model.state_dict().items()},
        "ema_model_state_dict": trainer.ema_state_dict(),
        "optimizer_state_dict": opt.state_dict(),
        "scheduler_state_dict": sched.state_dict(),
        "hparams": dict(hparams or {}),
        "causalgen_b200_state": {k: w.detach().cpu().clone() for k, w in trainer.state_dict().items()
                                 if k in ("state", "dyn", "seed_ctr", "steps_done", "iter_in_epoch", "beta", "beta_target")},
    }


def from_checkpoint(trainer, ckpt: Dict[str, Any], mode: str = "exact") -> None:
    """restore a Trainer from a checkpoint dict of the reference format.

    mode "exact": continue as if never interrupted (LR warm-up position, EMA schedule position, noise counter); for
        a checkpoint written by the reference itself the EMA position is taken as the optimiser step count.
    mode "reference": the reference's own resume semantics (src/main.py:74-90): constant learning rate from then on (the
        scheduler is replaced by ``lambda x: 1``) and a FRESH EMA counter, i.e. the first 100 updates copy the online
        weights over the loaded EMA weights."""
    if mode not in ("exact", "reference"):
        raise ValueError(mode)
    model = trainer.model
    with torch.no_grad():
        sd = ckpt["model_state_dict"]
        missing = [k for k, _ in model.named_parameters() if k not in sd]
        if missing:
            raise KeyError(f"checkpoint lacks parameters {missing[:4]}...")
        for k, p in model.named_parameters():
            p.copy_(sd[k])  # parameters are views of the trainer's flat buffer: this writes through
        ema_views = trainer._views(trainer.ema)
        for k, w in ckpt["ema_model_state_dict"].items():
            if k in ema_views:
                ema_views[k].copy_(w)
        # Adam moments through a real torch optimizer (validates the format the same way the reference does)
        opt, _ = _torch_optimizer(trainer)
        opt.load_state_dict(ckpt["optimizer_state_dict"])
        m, v = trainer._views(trainer.m), trainer._views(trainer.v)
        name_of = {id(p): k for k, p in model.named_parameters()}
        t = 0
        trainer.m.zero_()
        trainer.v.zero_()
        for p, st in opt.state.items():
            k = name_of[id(p)]
            if k in m:
                m[k].copy_(st["exp_avg"])
                v[k].copy_(st["exp_avg_sq"])
                t = max(t, int(float(st["step"])))
        extra = ckpt.get("causalgen_b200_state")
        state = trainer.state.clone()
        state[0] = t
        state[1] = t  # EMA.update() is called once per optimiser step (src/trainer.py:74-77)
        state[2] = 0
        state[3] = 0
        if extra is not None and mode == "exact":
            state.copy_(extra["state"].to(state.device))
            trainer.seed_ctr.copy_(extra["seed_ctr"].to(trainer.seed_ctr.device))
            trainer.steps_done = int(extra["steps_done"])
            trainer.iter_in_epoch = int(extra["iter_in_epoch"])
            trainer.beta, trainer.beta_target = float(extra["beta"]), float(extra["beta_target"])
        else:
            trainer.steps_done = int(ckpt.get("step", t))
        if mode == "reference":
            state[1] = 0                     # fresh EMA object (src/main.py:57)
            trainer.hp["warmup"] = 0         # LambdaLR(lambda x: 1) (src/main.py:83-85)
            lr = ckpt.get("hparams", {}).get("lr")
            if lr is not None:
                trainer.hp["lr"] = min(trainer.hp["lr"], float(lr))  # src/main.py:33-34
            trainer.invalidate_graphs()  # hyper-parameters are kernel arguments: re-capture with the new ones
        trainer.state.copy_(state)


def save(path: str, trainer, **kw) -> None:
    torch.save(to_checkpoint(trainer, **kw), path)


def load(path: str, trainer, mode: str = "exact", map_location="cpu") -> Dict[str, Any]:
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    from_checkpoint(trainer, ckpt, mode)
    return ckpt
