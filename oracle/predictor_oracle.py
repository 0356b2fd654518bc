"""CPU oracle (test infrastructure, never imported by the product) for the anticausal predictors that consume the
counterfactual image right after the hot path (call site /root/reference/src/pgm/dscm.py:78-83):

  * ``cnn_forward``      -- ``CNN`` of src/pgm/layers.py:64-104 (7x7 stem, BatchNorm2d, LeakyReLU, strided 3x3 convs, global
                            average pool, optional context concat, Linear-BatchNorm1d-LeakyReLU-Linear head), eval mode.
  * ``resnet18_forward`` -- ``ResNet18`` of src/pgm/resnet.py:212-239 over ``ResNet(CustomBlock, [2,2,2,2],
                            [64,128,256,512], GroupNorm(min(32, c//4), c))`` (src/pgm/resnet.py:9-209), eval mode
                            (dropout inactive).

Both are functional restatements over the reference's state_dict keys, pinned against outputs of the real classes by
tests/golden/make_golden_predictors.py -> tests/golden/predictors.npz (tests/test_predictor_oracle.py)."""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5      # nn.BatchNorm2d / BatchNorm1d / GroupNorm default
LRELU = 0.01       # nn.LeakyReLU default negative slope


def _bn(x, sd, key):
    shape = (1, -1) + (1,) * (x.dim() - 2)
    inv = torch.rsqrt(sd[key + ".running_var"] + BN_EPS) * sd[key + ".weight"]
    return (x - sd[key + ".running_mean"].view(shape)) * inv.view(shape) + sd[key + ".bias"].view(shape)


def cnn_forward(sd, x, y=None):
    """src/pgm/layers.py:64-104; x (N, C, R, R) in [-1, 1], y optional (N, context_dim)"""
    res = x.shape[-1]
    s = 2 if res > 64 else 1
    h = F.leaky_relu(_bn(F.conv2d(x, sd["cnn.0.weight"], stride=s, padding=3), sd, "cnn.1"), LRELU)
    if res > 32:
        h = F.max_pool2d(h, 2, 2)
    for idx, stride in ((4, 2), (7, 1), (10, 2), (13, 1), (16, 2)):
        h = F.leaky_relu(_bn(F.conv2d(h, sd[f"cnn.{idx}.weight"], stride=stride, padding=1), sd, f"cnn.{idx + 1}"), LRELU)
    h = h.mean(dim=(-2, -1))
    if y is not None:
        h = torch.cat([h, y], dim=-1)
    h = F.leaky_relu(_bn(F.linear(h, sd["fc.0.weight"]), sd, "fc.1"), LRELU)
    return F.linear(h, sd["fc.3.weight"], sd["fc.3.bias"])


def _gn(x, sd, key):
    c = x.shape[1]
    return F.group_norm(x, min(32, c // 4), sd[key + ".weight"], sd[key + ".bias"], BN_EPS)


def _block(sd, x, key, stride):
    """CustomBlock.forward, src/pgm/resnet.py:41-61"""
    out = F.relu(_gn(F.conv2d(x, sd[key + ".conv1.weight"], stride=stride, padding=1), sd, key + ".bn1"))
    out = _gn(F.conv2d(out, sd[key + ".conv2.weight"], padding=1), sd, key + ".bn2")
    idn = x
    if key + ".downsample.0.weight" in sd:
        idn = _gn(F.conv2d(x, sd[key + ".downsample.0.weight"], stride=stride), sd, key + ".downsample.1")
    return F.relu(out + idn)


def resnet18_forward(sd, x, y=None):
    """src/pgm/resnet.py:212-239 (children()[:-1] of the base model: conv1, bn1, relu, maxpool, layer1-4, avgpool)"""
    h = F.relu(_gn(F.conv2d(x, sd["resnet.0.weight"], stride=2, padding=3), sd, "resnet.1"))
    h = F.max_pool2d(h, 3, 2, 1)
    for li, stride in ((4, 1), (5, 2), (6, 2), (7, 2)):
        h = _block(sd, h, f"resnet.{li}.0", stride)
        h = _block(sd, h, f"resnet.{li}.1", 1)
    h = h.mean(dim=(-2, -1))
    if y is not None:
        h = torch.cat([h, y], dim=-1)
    return F.linear(h, sd["fc.weight"], sd["fc.bias"])


def seeded_predictor_state(shapes, seed):
    """deterministic, non-trivial parameters AND running statistics for a predictor given {key: shape}"""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in shapes.items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.tensor(7, dtype=torch.long)
        elif k.endswith("running_var"):
            sd[k] = 0.5 + torch.rand(shp, generator=g)
        elif k.endswith("running_mean"):
            sd[k] = 0.2 * torch.randn(shp, generator=g)
        elif len(shp) >= 2:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            sd[k] = torch.randn(shp, generator=g) * (1.6 / fan_in) ** 0.5
        elif k.endswith(".weight"):      # norm scales
            sd[k] = 1.0 + 0.2 * torch.randn(shp, generator=g)
        else:                            # biases / norm shifts
            sd[k] = 0.1 * torch.randn(shp, generator=g)
    return sd
