"""Golden vectors for the anticausal predictors from the REAL reference classes (build container only):

    python tests/golden/make_golden_predictors.py

Imports ``CNN`` from /root/reference/src/pgm/layers.py and ``ResNet18`` from /root/reference/src/pgm/resnet.py (pyro, which
layers.py imports for unrelated classes, is absent from this image and is stubbed with empty placeholder modules), loads the
deterministic parameters + running statistics of ``oracle.predictor_oracle.seeded_predictor_state`` (strict=True: also pins
the key names / shapes), runs eval-mode forwards and stores inputs' seeds and outputs in tests/golden/predictors.npz."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, "/root/reference/src/pgm")
sys.path.insert(0, "/root/reference/src")


def _stub_pyro():
    class _Any:
        def __init__(self, *a, **k):
            pass

    names = ["pyro", "pyro.infer", "pyro.distributions", "pyro.distributions.conditional",
             "pyro.distributions.torch_distribution", "pyro.nn", "pyro.infer.reparam.transform", "pyro.infer.reparam",
             "pyro.distributions.transforms"]
    for n in names:
        m = types.ModuleType(n)
        def _attr(attr, _A=_Any):  # any public attribute is a fresh placeholder class
            if attr.startswith("__"):
                raise AttributeError(attr)
            return type(attr, (_A,), {})
        m.__getattr__ = _attr
        m.__file__ = "<stub>" 
        sys.modules[n] = m
    sys.modules["pyro"].infer = sys.modules["pyro.infer"]
    sys.modules["pyro"].distributions = sys.modules["pyro.distributions"]


if "pyro" not in sys.modules:
    try:
        import pyro  # noqa: F401
    except Exception:
        _stub_pyro()

import predictor_oracle as PO  # noqa: E402
from layers import CNN  # noqa: E402
from resnet import ResNet18  # noqa: E402

CASES = {
    # name: (class, kwargs, batch, has context)
    "cnn_ukbb": (CNN, dict(in_shape=(1, 192, 192), width=16, num_outputs=2, context_dim=1), 2, True),
    "cnn_morpho": (CNN, dict(in_shape=(1, 32, 32), width=8, num_outputs=10, context_dim=0), 3, False),
    "cnn_cmnist": (CNN, dict(in_shape=(3, 32, 32), width=8, num_outputs=10, context_dim=0), 3, False),
    "cnn_mid": (CNN, dict(in_shape=(1, 64, 64), width=16, num_outputs=1, context_dim=1), 2, True),
    "resnet_mimic": (ResNet18, dict(in_shape=(1, 224, 224), num_outputs=2, context_dim=1), 2, True),
    "resnet_small": (ResNet18, dict(in_shape=(1, 64, 64), num_outputs=3, context_dim=0), 2, False),
}


def inputs(kw, B, ctx, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, *kw["in_shape"], generator=g) * 2 - 1
    y = torch.randn(B, kw["context_dim"], generator=g) if ctx else None
    return x, y


def main():
    out = {}
    for i, (name, (cls, kw, B, ctx)) in enumerate(CASES.items()):
        torch.manual_seed(0)
        m = cls(**kw).eval()
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        sd = PO.seeded_predictor_state(shapes, seed=100 + i)
        m.load_state_dict(sd, strict=True)
        x, y = inputs(kw, B, ctx, seed=200 + i)
        with torch.no_grad():
            o = m(x, y=y) if ctx else m(x)
            feat = m.cnn(x).mean(dim=(-2, -1)) if cls is CNN else m.resnet(x).flatten(1)
        out[name] = o.numpy()
        out[name + "::feat"] = feat.numpy()   # pooled features in front of the head (N, 8*width | 512)
        out[name + "::keys"] = np.array(list(shapes))
        out[name + "::shapes"] = np.array([",".join(map(str, s)) for s in shapes.values()])
        print(name, tuple(o.shape), o.flatten()[:4].tolist())
    np.savez_compressed(os.path.join(HERE, "predictors.npz"), **out)


if __name__ == "__main__":
    main()
