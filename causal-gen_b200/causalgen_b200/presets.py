"""Hyper-parameter bags for the named configurations (what the reference builds with ``hps.py`` presets,
src/hps.py:12-78, plus the launcher flags of src/run_local.sh:3-15 and src/run_slurm.sh:23-52).
An ``args`` object produced by the reference's own ``hps.setup_hparams`` works just as well."""
from __future__ import annotations

from types import SimpleNamespace

_BASE = dict(input_channels=1, bottleneck=4, z_dim=16, z_max_res=192, bias_max_res=64, cond_prior=False,
             q_correction=False, x_like="diag_dgauss", std_init=0.0, kl_free_bits=0.0, context_dim=4, beta=1.0,
             lr=1e-3, wd=0.01, betas=(0.9, 0.9), lr_warmup_steps=100, grad_clip=350.0, grad_skip=500.0,
             ema_rate=0.999, bs=32, seed=7, dataset="none")
_MNIST = dict(input_res=32, enc_arch="32b3d2,16b3d2,8b3d2,4b3d4,1b4", dec_arch="1b4,4b4,8b4,16b4,32b4",
              widths=[16, 32, 64, 128, 256])
_DEEP = dict(input_res=192, enc_arch="192b1d2,96b3d2,48b7d2,24b11d2,12b7d2,6b3d6,1b2",
             dec_arch="1b2,6b4,12b8,24b12,48b8,96b4,192b2", widths=[32, 64, 96, 128, 160, 192, 512], z_max_res=96)

PRESETS = {
    "morphomnist": dict(_MNIST, context_dim=12, cond_prior=True, parents_x=["thickness", "intensity", "digit"],
                        dataset="morphomnist"),
    "cmnist": dict(_MNIST, context_dim=20, input_channels=3, parents_x=["digit", "colour"], dataset="cmnist"),
    "ukbb192": dict(_DEEP, context_dim=4, beta=5.0, wd=0.05, dataset="ukbb",
                    parents_x=["mri_seq", "brain_volume", "ventricle_volume", "sex"]),
    "mimic192": dict(_DEEP, context_dim=6, beta=9.0, wd=0.05, bs=16, parents_x=["age", "race", "sex", "finding"],
                     dataset="mimic"),
    "mimic224": dict(_DEEP, input_res=224, context_dim=6, beta=9.0, wd=0.05, bs=16, z_max_res=112, dataset="mimic",
                     enc_arch="224b1d2,112b3d2,56b7d2,28b11d2,14b7d2,7b3d7,1b2",
                     dec_arch="1b2,8b4,14b8,28b12,56b8,112b4,224b2", parents_x=["age", "race", "sex", "finding"]),
    "tiny_ukbb": dict(input_res=16, enc_arch="16b1d2,8b2d2,4b1d4,1b1", dec_arch="1b1,4b2,8b2,16b1",
                      widths=[16, 32, 48, 64], context_dim=4, z_max_res=8, beta=5.0,
                      parents_x=["mri_seq", "brain_volume", "ventricle_volume", "sex"]),
}


def make_args(name: str, **over) -> SimpleNamespace:
    cfg = dict(_BASE, hps=name)
    cfg.update(PRESETS[name])
    cfg.update(over)
    return SimpleNamespace(**cfg)


def init_like_reference_main(model):
    """model.apply(init_bias) of src/main.py:51-55: every nn.Conv2d bias starts at zero"""
    import torch
    for m in model.modules():
        if isinstance(m, torch.nn.Conv2d):
            torch.nn.init.zeros_(m.bias)
    return model
