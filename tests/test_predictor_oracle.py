"""Pins oracle/predictor_oracle.py (anticausal predictors) against outputs of the real reference classes
(tests/golden/make_golden_predictors.py).  No access to /root/reference here."""
import os

import numpy as np
import pytest
import torch

import predictor_oracle as PO

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "predictors.npz"))
CASES = {
    # name: (kind, in_shape, context_dim, batch)   -- same table as make_golden_predictors.py
    "cnn_ukbb": ("cnn", (1, 192, 192), 1, 2),
    "cnn_morpho": ("cnn", (1, 32, 32), 0, 3),
    "cnn_cmnist": ("cnn", (3, 32, 32), 0, 3),
    "cnn_mid": ("cnn", (1, 64, 64), 1, 2),
    "resnet_mimic": ("resnet", (1, 224, 224), 1, 2),
    "resnet_small": ("resnet", (1, 64, 64), 0, 2),
}


def case_state(name):
    i = list(CASES).index(name)
    shapes = {k: tuple(int(d) for d in s.split(",") if d) for k, s in zip(GOLD[name + "::keys"], GOLD[name + "::shapes"])}
    return PO.seeded_predictor_state(shapes, seed=100 + i)


def case_inputs(name):
    i = list(CASES).index(name)
    kind, in_shape, ctx, B = CASES[name]
    g = torch.Generator().manual_seed(200 + i)
    x = torch.rand(B, *in_shape, generator=g) * 2 - 1
    y = torch.randn(B, ctx, generator=g) if ctx else None
    return x, y


@pytest.mark.parametrize("name", list(CASES))
def test_predictor_oracle_matches_reference(name):
    sd = case_state(name)
    x, y = case_inputs(name)
    fwd = PO.cnn_forward if CASES[name][0] == "cnn" else PO.resnet18_forward
    out = fwd(sd, x, y)
    np.testing.assert_allclose(out.numpy(), GOLD[name], rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("name", list(CASES))
def test_predictor_containers_have_reference_keys_and_shapes(name):
    """the product's parameter containers load a reference checkpoint with strict=True (keys / shapes recorded from the real
    classes by make_golden_predictors.py); no compute, no GPU"""
    from causalgen_b200.predictors import CNN, ResNet18
    kind, in_shape, ctx, _ = CASES[name]
    n_out = GOLD[name].shape[1]
    if kind == "cnn":
        m = CNN(in_shape=in_shape, width=GOLD[name + "::feat"].shape[1] // 8, num_outputs=n_out, context_dim=ctx)
    else:
        m = ResNet18(in_shape=in_shape, num_outputs=n_out, context_dim=ctx)
    sd = m.state_dict()
    assert list(sd) == list(GOLD[name + "::keys"])
    assert [",".join(map(str, v.shape)) for v in sd.values()] == list(GOLD[name + "::shapes"])
    m.load_state_dict(case_state(name), strict=True)
    with pytest.raises(RuntimeError):   # no CPU path: fails loudly instead of falling back
        m.eval()(torch.zeros(1, *in_shape), torch.zeros(1, ctx) if ctx else None)
