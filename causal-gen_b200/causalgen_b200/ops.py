"""Host-side operator objects over the C ABI: activation views, conv layers (forward, data-gradient,
weight-gradient launches and their packed-weight images).  Pure plumbing: every arithmetic op is a
kernel of libcausalgen_b200.so."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _lib as L


# Column-folded 3x3 convolutions (cg_conv_args.fold): wide input, cout <= 32.  Built in round 2, slower THEN because the
# kernel was bound by its epilogue / transform instruction chains; once those were shortened the wide-input layers became
# bound by the tensor cores' shared-memory operand reads (profiles/r3q_conv_ablation_final_kernel.txt), which folding cuts
# to a third.  CAUSALGEN_B200_FOLD=0 switches it off (A/B measurements).
# A folded tile is 8 rows x 14 valid pixels against 16 x 8 of the nine-tap walk: where that costs many more tiles (48^2: 24
# against 18 per image) or the input is narrow (32 -> 8 @192^2: two K-blocks, nothing to save) the nine-tap kernel stays
# (profiles/r3r_conv_ab_folded_final_kernel.txt).  CAUSALGEN_B200_FOLD=0 switches folding off, =2 folds wherever it fits
# (A/B measurements).
_FOLD_ENV = os.environ.get("CAUSALGEN_B200_FOLD", "1")
FOLD = int(_FOLD_ENV) if _FOLD_ENV.isdigit() else 3
# debugging aid: CAUSALGEN_B200_FOLD="res=24+96;kmin=64;dir=fb" folds exactly the layers at these resolutions with at least
# kmin K channels, forward (f) and / or data-gradient (b) launches
_FOLD_SPEC = dict(kv.split("=") for kv in _FOLD_ENV.split(";")) if FOLD == 3 else {}


def fold_pays(res: Optional[int], k_channels: int, bwd: bool = False) -> bool:
    """policy half of the fold decision (the kernel's half is cg_conv_fold_ok): `res` = image side, k_channels = padded
    channels on the GEMM-K axis"""
    if FOLD == 2:
        return True
    if FOLD == 3:
        return (("b" if bwd else "f") in _FOLD_SPEC.get("dir", "fb") and k_channels >= int(_FOLD_SPEC.get("kmin", "0")) and
                ("res" not in _FOLD_SPEC or str(res) in _FOLD_SPEC["res"].split("+")))
    if not FOLD or res is None or k_channels < 64:
        return False
    plain = -(-res // 8) * -(-res // 16)
    folded = -(-res // 14) * -(-res // 8)
    return folded <= 1.2 * plain


def wgrad_min_tiles(batch: int, light: bool) -> int:
    """cg_wgrad_args.min_tiles: 128-pixel tiles per weight-gradient CTA (lower bound).  Measured on UKBB-192, whose "light"
    two-conv ReLU blocks have cheap tiles (profiles/r4g_wgrad_grid_and_stem.txt: images/s at 128 | 32 images per GPU for
    24 / 48 / 96 / 144 / 192 / 384 tiles = 3101 | 2260, 3141 | 2287, 3186 | 2306, 3201 | 2297, 3220 | 2141, 3121 | 2007):
    best at 192 resp. 96 = 17 * sqrt(batch); flat around it at 64 images per GPU).  The same rule LOST 3-4 % on the four-conv
    GELU configs (their expensive tiles make long launches that hold back the pool streams); forced through
    CG_WGRAD_MIN_TILES they peak at 48 (Morpho-MNIST at batch 1024: 45.5 / 46.2 / 46.2 k images/s for 24 / 48 / 96 tiles,
    MIMIC-192 at 64: 1742 / 1741 / 1717)."""
    if not light:
        return 48
    return max(24, min(256, int(17.0 * batch ** 0.5 + 0.5)))


def round16(c: int) -> int:
    return (c + 15) // 16 * 16


def round8(c: int) -> int:
    return (c + 7) // 8 * 8


# Channel octets stored per activation: PAD16=1 restores the round-1 layout (channels padded to a multiple of 16 in HBM;
# A/B measurements).  Default: multiples of 8 -- a tensor with 8 / 24 / 40 channels keeps 1 / 3 / 5 octet planes and the
# kernels let the TMA box of the last 16-channel K-block run past them (zero fill, cg_src.c8).
PAD16 = os.environ.get("CAUSALGEN_B200_PAD16", "0") == "1"


def phys(c: int) -> int:
    """channels physically stored for a bf16 planar tensor with `c` logical channels"""
    return round16(c) if PAD16 else round8(c)


class View:
    """Channel slice [c0, c0+C) of an activation tensor.

    bf16 activations are channel-octet planar: t has shape (N, Ctot/8, H, W, 8); element (n,c,h,w) lives at
    n*ns + (c//8)*H*W*8 + (h*W+w)*8 + c%8.  A slice with c0 % 8 == 0 keeps the same sample stride ``ns`` and
    only advances the base pointer, so it is a valid kernel operand without a copy.
    fp32 statistics tensors keep a row layout (N, H, W, ld); for those ``ns`` is the row pitch."""
    __slots__ = ("t", "c0", "C", "logical")

    def __init__(self, t: torch.Tensor, C: Optional[int] = None, c0: int = 0, logical: Optional[int] = None):
        self.t = t
        self.c0 = c0
        full = t.shape[1] * 8 if self.planar_dtype(t) else t.shape[-1]
        self.C = full - c0 if C is None else C
        self.logical = self.C if logical is None else logical
        assert c0 % 8 == 0 and self.C % 8 == 0

    @staticmethod
    def planar_dtype(t) -> bool:
        return t.dtype == torch.bfloat16

    @property
    def planar(self) -> bool:
        return self.planar_dtype(self.t)

    @property
    def HW(self) -> int:
        return self.t.shape[2] * self.t.shape[3] if self.planar else self.t.shape[1] * self.t.shape[2]

    @property
    def ns(self) -> int:
        return self.t.stride(0) if self.planar else self.t.shape[-1]

    @property
    def ptr(self) -> int:
        if self.planar:
            return self.t.data_ptr() + (self.c0 // 8) * self.HW * 8 * 2
        return self.t.data_ptr() + self.c0 * self.t.element_size()

    def slice(self, c0: int, C: int, logical: Optional[int] = None) -> "View":
        return View(self.t, C, self.c0 + c0, logical)

    def nchw(self) -> torch.Tensor:
        """(N, C, H, W) float copy of the logical channels (tests / debugging only)"""
        if self.planar:
            N, C8, H, W, _ = self.t.shape
            full = self.t.permute(0, 1, 4, 2, 3).reshape(N, C8 * 8, H, W)
        else:
            full = self.t.permute(0, 3, 1, 2)
        return full[:, self.c0: self.c0 + self.logical].float()


def new_act(N, H, W, C, device, dtype=torch.bfloat16, logical=None) -> View:
    """zero-initialised planar buffer; padded channels stay zero for the tensor-core K loop"""
    Cp = phys(C) if dtype == torch.bfloat16 else round16(C)
    if dtype == torch.bfloat16:
        t = torch.zeros(N, Cp // 8, H, W, 8, device=device, dtype=dtype)
    else:
        t = torch.zeros(N, H, W, Cp, device=device, dtype=dtype)
    return View(t, Cp, 0, C if logical is None else logical)


def planar_from_nchw(x: torch.Tensor, pad_to: Optional[int] = None) -> torch.Tensor:
    """(N,C,H,W) float -> zero padded planar bf16 tensor (tests / debugging only)"""
    N, Cc, H, W = x.shape
    Cp = pad_to or phys(Cc)
    full = torch.zeros(N, Cp, H, W, device=x.device, dtype=torch.bfloat16)
    full[:, :Cc] = x.to(torch.bfloat16)
    return full.reshape(N, Cp // 8, 8, H, W).permute(0, 1, 3, 4, 2).contiguous()


@dataclass
class SegSpec:
    out: View
    c0: int                 # first output channel of the conv this segment takes
    add: Optional[View] = None
    mul: Optional[View] = None
    mul_act: int = L.ACT_NONE
    add2: Optional[View] = None
    out_act: int = L.ACT_NONE   # activation applied to the stored value (consumer-side pre-activation hoisted)
    act_copy: Optional[View] = None  # when set: `out` gets the raw value, act_copy gets out_act(value)


class PackTable:
    """All weight-packing descriptors of a model; one kernel launch repacks every conv."""

    def __init__(self, device):
        self.device = device
        self.descs: List[L.PackDesc] = []
        self.keep = []
        self._dev = None

    def add(self, desc: L.PackDesc):
        self.descs.append(desc)
        self._dev = None

    def launch(self, stream):
        if not self.descs:
            return
        if self._dev is None:
            arr = (L.PackDesc * len(self.descs))(*self.descs)
            raw = bytes(arr)
            self._dev = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.device)
        L.check(L.load().cg_pack_weights(self._dev.data_ptr(), len(self.descs), stream), "cg_pack_weights")


class ConvLayer:
    """One nn.Conv2d of the reference (weight OIHW fp32 master) as tcgen05 launches.

    src_logical: logical channel count of each K-concatenated source (torch.cat order).
    taps: k*k, or 1 for a 3x3 conv evaluated on a 1x1 image (only the centre tap touches data).
    """

    def __init__(self, table: PackTable, weight: torch.Tensor, bias: Optional[torch.Tensor],
                 src_logical: Sequence[int], act: int, centre_only: bool = False,
                 grad_srcs: Optional[Sequence[bool]] = None, fwd_operands: bool = True,
                 n_scale: Optional[torch.Tensor] = None, res: Optional[int] = None, light: bool = False):
        """light: the layer belongs to a model built from "light" Blocks (weight-gradient grid policy, `wgrad_min_tiles`).
        res: side of the (square) image the layer runs on, when known at construction (fold policy, `fold_pays`).
        fwd_operands=False: no forward launch of this layer fuses an epilogue operand (add / add2 / mul), so its
        GEMM-N chunk need not reserve shared memory for the operand ring (first convs of a Block: huge K, narrow N)"""
        lib = L.load()
        self.weight, self.bias, self.light = weight, bias, light
        self.cout_l, self.cin_l, self.k = weight.shape[0], weight.shape[1], weight.shape[2]
        assert sum(src_logical) == self.cin_l, (src_logical, self.cin_l)
        self.src_logical = list(src_logical)
        self.src_pad = [round16(c) for c in src_logical]
        self.src_off = [sum(src_logical[:i]) for i in range(len(src_logical))]
        self.cout_pad = round16(self.cout_l)
        self.act = act
        self.taps = 1 if (centre_only or self.k == 1) else self.k * self.k
        self.ksize = 1 if self.taps == 1 else self.k
        dev = weight.device
        kt = self.taps * sum(self.src_pad) // 16
        # column-folded 3x3 (cg_conv_args.fold): wide input, narrow output -- the case bound by shared-memory operand reads
        self.fold = int(self.taps == 9 and sum(self.src_pad) >= 2 * self.cout_pad and
                        fold_pays(res, sum(self.src_pad)) and lib.cg_conv_fold_ok(kt, self.cout_pad, int(fwd_operands)) == 1)
        self.nc = self.cout_pad if self.fold else lib.cg_conv_nchunk_ex(kt, self.cout_pad, int(fwd_operands))
        nbytes = lib.cg_packed_weight_bytes_nc(kt, self.cout_pad, self.nc)
        if self.nc <= 0 or nbytes <= 0:
            raise RuntimeError(f"conv K={kt * 16} does not fit a resident weight slab")
        self.wpack = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        d = L.PackDesc()
        d.w, d.out = weight.data_ptr(), self.wpack.data_ptr()
        d.cout_l, d.cin_l, d.k = self.cout_l, self.cin_l, self.k
        d.transpose, d.taps, d.n_pad, d.nc, d.n_off, d.n_log = 0, self.taps, self.cout_pad, self.nc, 0, self.cout_l
        d.fold = self.fold
        self.n_scale = n_scale  # per-output-channel multiplier folded into the forward pack (eval-mode BatchNorm)
        if n_scale is not None:
            assert n_scale.dtype == torch.float32 and n_scale.numel() == self.cout_l
            d.n_scale = n_scale.data_ptr()
        d.nsrc = len(src_logical)
        for i in range(d.nsrc):
            d.src_c[i], d.src_log[i], d.src_off[i] = self.src_pad[i], self.src_logical[i], self.src_off[i]
        table.add(d)
        # data-gradient packs (one per source that needs a gradient): K side = dY channels
        self.wpack_bwd: List[Optional[torch.Tensor]] = []
        self.nc_bwd: List[int] = []
        self.fold_bwd: List[int] = []
        grad_srcs = [True] * len(src_logical) if grad_srcs is None else list(grad_srcs)
        ktb = self.taps * self.cout_pad // 16
        for i, need in enumerate(grad_srcs):
            if not need:
                self.wpack_bwd.append(None)
                self.nc_bwd.append(0)
                self.fold_bwd.append(0)
                continue
            fb = int(self.taps == 9 and self.cout_pad >= 2 * self.src_pad[i] and fold_pays(res, self.cout_pad, bwd=True) and
                     lib.cg_conv_fold_ok(ktb, self.src_pad[i], 1) == 1)
            self.fold_bwd.append(fb)
            ncb = self.src_pad[i] if fb else lib.cg_conv_nchunk_ex(ktb, self.src_pad[i], 1)
            nb = lib.cg_packed_weight_bytes_nc(ktb, self.src_pad[i], ncb)
            self.nc_bwd.append(ncb)
            buf = torch.zeros(nb, dtype=torch.uint8, device=dev)
            b = L.PackDesc()
            b.w, b.out = weight.data_ptr(), buf.data_ptr()
            b.cout_l, b.cin_l, b.k = self.cout_l, self.cin_l, self.k
            b.transpose, b.taps, b.n_pad, b.nc = 1, self.taps, self.src_pad[i], ncb
            b.n_off, b.n_log, b.nsrc = self.src_off[i], self.src_logical[i], 1
            b.src_c[0], b.src_log[0], b.src_off[0] = self.cout_pad, self.cout_l, 0
            b.fold = fb
            table.add(b)
            self.wpack_bwd.append(buf)

    # ------------------------------------------------------------------ launches
    @staticmethod
    def _fill_srcs(arr, srcs: Sequence[View]):
        for i, s in enumerate(srcs):
            # C: K-blocks of 16 channels; c8: octets the view physically holds (the last K-block may be half there)
            arr[i].ptr, arr[i].ns, arr[i].C, arr[i].c8 = s.ptr, s.ns, round16(s.C), s.C // 8

    @staticmethod
    def _fits(v: View, logical: int) -> bool:
        return round8(logical) <= v.C <= round16(logical)

    @staticmethod
    def _fill_segs(a, segs: Sequence[SegSpec]):
        a.nseg = len(segs)
        for i, sg in enumerate(segs):
            s = a.seg[i]
            s.ptr, s.c0, s.cn, s.ns = sg.out.ptr, sg.c0, sg.out.C, sg.out.ns
            s.dtype = L.F32 if sg.out.t.dtype == torch.float32 else L.BF16
            s.add = sg.add.ptr if sg.add is not None else None
            s.add_ns = sg.add.ns if sg.add is not None else 0
            s.add2 = sg.add2.ptr if sg.add2 is not None else None
            s.add2_ns = sg.add2.ns if sg.add2 is not None else 0
            s.mul = sg.mul.ptr if sg.mul is not None else None
            s.mul_ns = sg.mul.ns if sg.mul is not None else 0
            s.mul_act = sg.mul_act
            s.out_act = sg.out_act
            s.act_copy = sg.act_copy.ptr if sg.act_copy is not None else None
            s.act_copy_ns = sg.act_copy.ns if sg.act_copy is not None else 0

    def forward(self, srcs: Sequence[View], segs: Sequence[SegSpec], N, H, W) -> L.Launch:
        assert all(self._fits(s, c) for s, c in zip(srcs, self.src_logical)), ([s.C for s in srcs], self.src_logical)
        a = L.ConvArgs()
        a.N, a.H, a.W, a.ksize, a.act = N, H, W, self.ksize, self.act
        a.nsrc, a.cout = len(srcs), self.cout_pad
        self._fill_srcs(a.src, srcs)
        self._fill_segs(a, segs)
        a.wpack, a.nc, a.fold = self.wpack.data_ptr(), self.nc, self.fold
        if self.bias is not None:
            a.bias, a.bias_n = self.bias.data_ptr(), self.cout_l
        ln = L.Launch("cg_conv2d", C.byref(a))
        ln.keep = (a, srcs, segs)
        # bf16 (input + output) activation bytes of the conv, LOGICAL channels (no padding, no fused operands)
        ln.algo_bytes = 2 * N * H * W * (self.cin_l + self.cout_l)
        return ln

    def dgrad(self, i: int, dy: View, seg: SegSpec, N, H, W) -> L.Launch:
        """dX_i = conv^T(dY) [* act'(x_i)] [+ add]  for source i"""
        assert self._fits(dy, self.cout_l) and self._fits(seg.out, self.src_logical[i]), (dy.C, seg.out.C)
        a = L.ConvArgs()
        a.N, a.H, a.W, a.ksize, a.act = N, H, W, self.ksize, L.ACT_NONE
        a.nsrc, a.cout = 1, self.src_pad[i]
        self._fill_srcs(a.src, [dy])
        self._fill_segs(a, [seg])
        a.wpack, a.nc, a.fold = self.wpack_bwd[i].data_ptr(), self.nc_bwd[i], self.fold_bwd[i]
        ln = L.Launch("cg_conv2d", C.byref(a))
        ln.keep = (a, dy, seg)
        ln.algo_bytes = 2 * N * H * W * (self.cout_l + self.src_logical[i])
        return ln

    def wgrad(self, srcs: Sequence[View], dy: View, dw: torch.Tensor, db: Optional[torch.Tensor], N, H, W) -> L.Launch:
        a = L.WgradArgs()
        a.N, a.H, a.W, a.ksize, a.act = N, H, W, self.k, self.act
        a.nsrc = len(srcs)
        self._fill_srcs(a.src, srcs)
        a.dy, a.dy_c, a.dy_c8, a.dy_ns = dy.ptr, round16(dy.C), dy.C // 8, dy.ns
        a.dw = dw.data_ptr()
        a.dbias = db.data_ptr() if db is not None else None
        a.cout_l, a.cin_l, a.taps = self.cout_l, self.cin_l, self.taps
        a.min_tiles = wgrad_min_tiles(N, self.light)
        for i in range(len(srcs)):
            a.src_log[i], a.src_off[i] = self.src_logical[i], self.src_off[i]
        ln = L.Launch("cg_conv2d_wgrad", C.byref(a))
        ln.keep = (a, srcs, dy, dw, db)
        ln.algo_bytes = 2 * N * H * W * (self.cin_l + self.cout_l) + 4 * self.weight.numel()
        return ln
