"""Anticausal predictors evaluated on counterfactual images -- the step right after the hot path in
counterfactual training / evaluation (reference call site src/pgm/dscm.py:78-83, built in src/pgm/flow_pgm.py:155-163,
351-355, 591-597):

  * ``CNN``      -- src/pgm/layers.py:64-104 (7x7 stem, BatchNorm2d + LeakyReLU, strided 3x3 convolutions, global average
                    pooling, optional context concat, Linear-BatchNorm1d-LeakyReLU-Linear head)
  * ``ResNet18`` -- src/pgm/resnet.py:212-239 over ResNet(CustomBlock, [2,2,2,2], [64,128,256,512], GroupNorm(min(32,c//4),c))

Same constructors, same state_dict keys / shapes (a reference checkpoint loads with strict=True), same
``forward(x, y=None)``.  INFERENCE (eval mode) only: the predictors are frozen when counterfactuals are scored; calling them
in train mode or under autograd raises.  Every arithmetic op is a kernel of libcausalgen_b200.so: eval-mode BatchNorm is
folded into the packed tcgen05 weights + bias (cg_bn_fold, cg_pack_desc.n_scale), LeakyReLU runs in the conv epilogue, the
stride-2 3x3 convolutions are stride-1 tensor-core convolutions followed by a strided pick (elementwise ops commute with the
pick), GroupNorm + residual + ReLU is one two-kernel call.  The nn.Modules below hold parameters only."""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import nn

from . import _lib as L
from .hvae import _stream
from .ops import ConvLayer, PackTable, SegSpec, View, new_act

BN_EPS = 1e-5


class _Plan:
    """buffers + recorded launches for one (batch size, parameter storage) pair"""

    def __init__(self):
        self.launches: List = []
        self.keep: List = []
        self.table: Optional[PackTable] = None
        self.x = self.y = self.feat = self.out = None

    def call(self, name, *args):
        self.launches.append(L.Launch(name, *args))

    def run(self):
        s = _stream()
        for ln in self.launches[: self.n_fold]:  # BatchNorm folds first: the weight pack reads their scales
            ln(s)
        self.table.launch(s)
        for ln in self.launches[self.n_fold:]:
            ln(s)


class _Predictor(nn.Module):
    def __init__(self):
        super().__init__()
        self._plans = {}

    def _sig(self, N):
        return (N,) + tuple(p.data_ptr() for p in self.state_dict().values())

    def _check(self, x, y):
        if self.training:
            raise RuntimeError("causalgen_b200 predictors are inference-only: call .eval() (the reference scores "
                               "counterfactuals with frozen predictors, src/pgm/train_cf.py)")
        if not x.is_cuda:
            raise RuntimeError("causalgen_b200 has no CPU path: predictor inputs must live on a B200")
        if (y is None) != (self.context_dim == 0):
            raise ValueError(f"context_dim={self.context_dim} but y is {'missing' if y is None else 'given'}")

    @torch.no_grad()
    def forward(self, x: torch.Tensor, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        self._check(x, y)
        N = x.shape[0]
        sig = self._sig(N)
        plan = self._plans.get(N)
        if plan is None or plan.sig != sig:
            plan = self._build(N, x.device)
            plan.sig = sig
            self._plans[N] = plan
        plan.x.copy_(x)
        if y is not None:
            plan.feat[:, self.feat_dim:].copy_(y.reshape(N, -1))
        plan.run()
        return plan.out.clone()

    # helpers ------------------------------------------------------------------------------------------------------
    @staticmethod
    def _fold_bn(plan: _Plan, bn: nn.Module, dev):
        C = bn.num_features
        scale = torch.zeros(C, device=dev, dtype=torch.float32)
        shift = torch.zeros(C, device=dev, dtype=torch.float32)
        plan.call("cg_bn_fold", bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(),
                  bn.running_var.data_ptr(), float(bn.eps), scale.data_ptr(), shift.data_ptr(), C)
        plan.keep += [scale, shift]
        return scale, shift

    @staticmethod
    def _pick2(plan: _Plan, v: View, N, H, W, dev) -> View:
        """out[h, w] = v[2h, 2w]: what a stride-2 convolution keeps of its stride-1 result"""
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        o = new_act(N, Ho, Wo, v.logical, dev)
        plan.call("cg_pool_max_fwd", v.ptr, o.ptr, N, v.C, H, W, 1, 2, 0, v.ns, o.ns)
        plan.keep.append(o)
        return o


class CNN(_Predictor):
    """src/pgm/layers.py:64-104"""

    def __init__(self, in_shape=(1, 192, 192), width=16, num_outputs=1, context_dim=0):
        super().__init__()
        in_channels, res = in_shape[0], in_shape[1]
        self.in_shape, self.width, self.num_outputs, self.context_dim = tuple(in_shape), width, num_outputs, context_dim
        self.feat_dim = 8 * width
        s = 2 if res > 64 else 1
        act = nn.LeakyReLU()
        self.cnn = nn.Sequential(
            nn.Conv2d(in_channels, width, 7, s, 3, bias=False), nn.BatchNorm2d(width), act,
            (nn.MaxPool2d(2, 2) if res > 32 else nn.Identity()),
            nn.Conv2d(width, 2 * width, 3, 2, 1, bias=False), nn.BatchNorm2d(2 * width), act,
            nn.Conv2d(2 * width, 2 * width, 3, 1, 1, bias=False), nn.BatchNorm2d(2 * width), act,
            nn.Conv2d(2 * width, 4 * width, 3, 2, 1, bias=False), nn.BatchNorm2d(4 * width), act,
            nn.Conv2d(4 * width, 4 * width, 3, 1, 1, bias=False), nn.BatchNorm2d(4 * width), act,
            nn.Conv2d(4 * width, 8 * width, 3, 2, 1, bias=False), nn.BatchNorm2d(8 * width), act,
        )
        self.fc = nn.Sequential(
            nn.Linear(8 * width + context_dim, 8 * width, bias=False), nn.BatchNorm1d(8 * width), act,
            nn.Linear(8 * width, num_outputs),
        )

    def _build(self, N, dev) -> _Plan:
        plan = _Plan()
        plan.table = PackTable(dev)
        Cin, R = self.in_shape[0], self.in_shape[1]
        w = self.width
        plan.x = torch.zeros(N, Cin, R, R, device=dev, dtype=torch.float32)
        # BatchNorm folds (recomputed on every call: cheap, and updated running statistics are honoured)
        folds = {i: self._fold_bn(plan, self.cnn[i], dev) for i in (1, 5, 8, 11, 14, 17)}
        fc_scale, fc_shift = self._fold_bn(plan, self.fc[1], dev)
        plan.n_fold = len(plan.launches)
        stem = self.cnn[0]
        s = stem.stride[0]
        H = (R + 6 - 7) // s + 1
        a = new_act(N, H, H, w, dev)
        plan.call("cg_conv_direct_fwd", plan.x.data_ptr(), stem.weight.data_ptr(), folds[1][0].data_ptr(),
                  folds[1][1].data_ptr(), a.ptr, N, Cin, R, R, w, 7, s, 3, L.ACT_LRELU, a.ns)
        if R > 32:
            Ho = (H - 2) // 2 + 1
            o = new_act(N, Ho, Ho, w, dev)
            plan.call("cg_pool_max_fwd", a.ptr, o.ptr, N, a.C, H, H, 2, 2, 0, a.ns, o.ns)
            plan.keep.append(a)
            a, H = o, Ho
        for idx in (4, 7, 10, 13, 16):
            conv = self.cnn[idx]
            scale, shift = folds[idx + 1]
            layer = ConvLayer(plan.table, conv.weight, shift, [conv.in_channels], L.ACT_NONE, grad_srcs=[False],
                              fwd_operands=False, n_scale=scale)
            o = new_act(N, H, H, conv.out_channels, dev)
            plan.launches.append(layer.forward([a], [SegSpec(o, 0, out_act=L.ACT_LRELU)], N, H, H))
            plan.keep += [layer, a]
            a = o
            if conv.stride[0] == 2:
                a = self._pick2(plan, a, N, H, H, dev)
                H = (H - 1) // 2 + 1
        F_ = self.feat_dim
        plan.feat = torch.zeros(N, F_ + self.context_dim, device=dev, dtype=torch.float32)
        plan.hid = torch.zeros(N, F_, device=dev, dtype=torch.float32)
        plan.out = torch.zeros(N, self.num_outputs, device=dev, dtype=torch.float32)
        plan.call("cg_global_avgpool", a.ptr, plan.feat.data_ptr(), N, F_, H * H, a.ns, F_ + self.context_dim)
        plan.call("cg_linear", plan.feat.data_ptr(), F_ + self.context_dim, self.fc[0].weight.data_ptr(), None,
                  fc_scale.data_ptr(), fc_shift.data_ptr(), L.ACT_LRELU, plan.hid.data_ptr(), F_, N, F_ + self.context_dim, F_)
        plan.call("cg_linear", plan.hid.data_ptr(), F_, self.fc[3].weight.data_ptr(), self.fc[3].bias.data_ptr(), None, None,
                  L.ACT_NONE, plan.out.data_ptr(), self.num_outputs, N, F_, self.num_outputs)
        plan.keep.append(a)
        return plan


class _BlockParams(nn.Module):
    """parameter container of CustomBlock (src/pgm/resnet.py:9-61)"""

    def __init__(self, inplanes, planes, stride, downsample):
        super().__init__()
        gn = lambda c: nn.GroupNorm(min(32, c // 4), c)  # noqa: E731  (src/pgm/resnet.py:228)
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = gn(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = gn(planes)
        if downsample:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes, 1, stride, bias=False), gn(planes))
        else:
            self.downsample = None
        self.stride = stride


class ResNet18(_Predictor):
    """src/pgm/resnet.py:212-239 (base model: src/pgm/resnet.py:64-209 with CustomBlock, layers [2,2,2,2])"""

    def __init__(self, in_shape=(1, 224, 224), num_outputs=1, context_dim=0, base_model=None):
        super().__init__()
        if base_model is not None:
            raise NotImplementedError("a custom base_model is not supported: only the reference's GroupNorm ResNet-18")
        self.in_shape, self.num_outputs, self.context_dim = tuple(in_shape), num_outputs, context_dim
        self.feat_dim = 512
        widths = [64, 128, 256, 512]
        layers = []
        inplanes = widths[0]
        for li, planes in enumerate(widths):
            stride = 1 if li == 0 else 2
            blocks = [_BlockParams(inplanes, planes, stride, stride != 1 or inplanes != planes),
                      _BlockParams(planes, planes, 1, False)]
            inplanes = planes
            layers.append(nn.Sequential(*blocks))
        # children()[:-1] of the reference base model: conv1, bn1, relu, maxpool, layer1..4, avgpool
        self.resnet = nn.Sequential(
            nn.Conv2d(in_shape[0], 64, kernel_size=7, stride=2, padding=3, bias=False), nn.GroupNorm(16, 64),
            nn.ReLU(inplace=True), nn.MaxPool2d(kernel_size=3, stride=2, padding=1), *layers, nn.AdaptiveAvgPool2d((1, 1)))
        self.fc = nn.Linear(512 + context_dim, num_outputs)
        for m in self.modules():  # src/pgm/resnet.py:133-138
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def _gn(self, plan, gnm: nn.GroupNorm, v: View, N, H, dev, act, add: Optional[View] = None) -> View:
        C = gnm.num_channels
        o = new_act(N, H, H, C, dev)
        stats = torch.zeros(N * C * 2, device=dev, dtype=torch.float32)
        plan.call("cg_groupnorm_fwd", v.ptr, stats.data_ptr(), gnm.weight.data_ptr(), gnm.bias.data_ptr(), gnm.num_groups,
                  float(gnm.eps), add.ptr if add is not None else None, add.ns if add is not None else 0, act, o.ptr, N, C,
                  H * H, v.ns, o.ns)
        plan.keep += [stats, v, o, add]
        return o

    def _conv(self, plan, conv: nn.Conv2d, v: View, N, H, dev) -> View:
        layer = ConvLayer(plan.table, conv.weight, None, [conv.in_channels], L.ACT_NONE, grad_srcs=[False], fwd_operands=False)
        o = new_act(N, H, H, conv.out_channels, dev)
        plan.launches.append(layer.forward([v], [SegSpec(o, 0)], N, H, H))
        plan.keep += [layer, v, o]
        return o

    def _build(self, N, dev) -> _Plan:
        plan = _Plan()
        plan.table = PackTable(dev)
        plan.n_fold = 0
        Cin, R = self.in_shape[0], self.in_shape[1]
        plan.x = torch.zeros(N, Cin, R, R, device=dev, dtype=torch.float32)
        H = (R + 6 - 7) // 2 + 1
        a = new_act(N, H, H, 64, dev)
        plan.call("cg_conv_direct_fwd", plan.x.data_ptr(), self.resnet[0].weight.data_ptr(), None, None, a.ptr, N, Cin, R, R,
                  64, 7, 2, 3, L.ACT_NONE, a.ns)
        a = self._gn(plan, self.resnet[1], a, N, H, dev, L.ACT_RELU)
        Ho = (H + 2 - 3) // 2 + 1
        o = new_act(N, Ho, Ho, 64, dev)
        plan.call("cg_pool_max_fwd", a.ptr, o.ptr, N, a.C, H, H, 3, 2, 1, a.ns, o.ns)
        a, H = o, Ho
        for li in (4, 5, 6, 7):
            for blk in self.resnet[li]:
                x_in, H_in = a, H
                c1 = self._conv(plan, blk.conv1, x_in, N, H_in, dev)
                if blk.stride == 2:
                    c1 = self._pick2(plan, c1, N, H_in, H_in, dev)
                    H = (H_in - 1) // 2 + 1
                g1 = self._gn(plan, blk.bn1, c1, N, H, dev, L.ACT_RELU)
                c2 = self._conv(plan, blk.conv2, g1, N, H, dev)
                idn = x_in
                if blk.downsample is not None:
                    xs = self._pick2(plan, x_in, N, H_in, H_in, dev) if blk.stride == 2 else x_in
                    d = self._conv(plan, blk.downsample[0], xs, N, H, dev)
                    idn = self._gn(plan, blk.downsample[1], d, N, H, dev, L.ACT_NONE)
                a = self._gn(plan, blk.bn2, c2, N, H, dev, L.ACT_RELU, add=idn)
        plan.feat = torch.zeros(N, 512 + self.context_dim, device=dev, dtype=torch.float32)
        plan.out = torch.zeros(N, self.num_outputs, device=dev, dtype=torch.float32)
        plan.call("cg_global_avgpool", a.ptr, plan.feat.data_ptr(), N, 512, H * H, a.ns, 512 + self.context_dim)
        plan.call("cg_linear", plan.feat.data_ptr(), 512 + self.context_dim, self.fc.weight.data_ptr(), self.fc.bias.data_ptr(),
                  None, None, L.ACT_NONE, plan.out.data_ptr(), self.num_outputs, N, 512 + self.context_dim, self.num_outputs)
        plan.keep.append(a)
        return plan
