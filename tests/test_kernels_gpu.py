"""Kernel-level numerics on a real B200: every C-ABI op against a plain PyTorch fp32 reference of the
same op (TF32 off) on identical, bf16-representable inputs.  Tolerances: bf16 outputs are compared at
bf16 resolution (rel 1e-2 of the tensor scale), fp32 outputs at 2e-3."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _setup():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from causalgen_b200 import _lib
    _lib.load()
    yield


def stream():
    return torch.cuda.current_stream().cuda_stream


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def nhwc_bf16(x_nchw, pad_to=None):
    """(N,C,H,W) fp32 -> zero padded channel-octet planar bf16 buffer (N, C/8, H, W, 8)"""
    from causalgen_b200.ops import planar_from_nchw
    return planar_from_nchw(x_nchw, pad_to)


def to_nchw(t, C):
    """planar bf16 (N,C8,H,W,8) or fp32 rows (N,H,W,C) -> (N,C,H,W) float"""
    if t.dtype == torch.bfloat16:
        N, C8, H, W, _ = t.shape
        return t.permute(0, 1, 4, 2, 3).reshape(N, C8 * 8, H, W)[:, :C].float()
    return t[..., :C].permute(0, 3, 1, 2).float()


def ns_of(t):
    return t.stride(0)


def act_fn(a):
    return {0: (lambda v: v), 1: F.relu, 2: F.gelu}[a]


def assert_close(ours, ref, rel, what=""):
    scale = ref.abs().max().item() + 1e-6
    err = (ours - ref).abs().max().item()
    assert err <= rel * scale, f"{what}: max err {err:.4g} vs scale {scale:.4g} (rel {err / scale:.3g} > {rel})"


CONV_CASES = [
    # N, H, W, src channels, bcast ctx, cout, k, act
    (2, 16, 16, [32], 0, 16, 3, 1),
    (3, 12, 12, [64], 0, 24, 3, 1),
    (2, 24, 24, [48, 48], 4, 8, 3, 1),          # posterior-style cat[h, pa, x]
    (2, 6, 6, [40], 0, 176, 3, 2),              # N > 128 columns, odd widths
    (5, 1, 1, [128], 12, 544, 1, 2),            # res-1 1x1, Cout > 256 -> N chunks
    (2, 32, 32, [16], 0, 4, 1, 2),              # tiny bottleneck
    (1, 48, 48, [24], 0, 128, 3, 1),
    (2, 8, 8, [16], 20, 64, 1, 0),              # z_proj-style cat[z, pa], no activation
    (1, 7, 7, [32], 0, 16, 3, 0),               # odd resolution
]


def run_conv(N, H, W, chans, ctx, cout, k, act, seed=0):
    from causalgen_b200 import _lib as L
    from causalgen_b200.ops import ConvLayer, PackTable, SegSpec, View, new_act, phys
    xs = [rnd(N, c, H, W, seed=seed + i) for i, c in enumerate(chans)]
    views = [View(nhwc_bf16(x), phys(c), 0, c) for x, c in zip(xs, chans)]
    logical = list(chans)
    pa = None
    if ctx:  # spatially constant parents, materialised (src/vae.py:241)
        pa = rnd(N, ctx, seed=seed + 9)
        pat = nhwc_bf16(pa[:, :, None, None].expand(N, ctx, H, W))
        views.insert(1, View(pat, phys(ctx), 0, ctx))
        logical.insert(1, ctx)
        chans = list(chans)
    cin = sum(logical)
    w = rnd(cout, cin, k, k, scale=1.0 / math.sqrt(cin * k * k), seed=seed + 20)
    b = rnd(cout, scale=0.1, seed=seed + 21)
    table = PackTable(DEV)
    layer = ConvLayer(table, w, b, logical, act)
    table.launch(stream())
    out = new_act(N, H, W, cout, DEV)
    layer.forward(views, [SegSpec(out, 0)], N, H, W)(stream())
    torch.cuda.synchronize()
    # reference on the same bf16-rounded operands
    parts = [to_nchw(v.t, c) for v, c in zip(views, logical)]
    xin = act_fn(act)(torch.cat(parts, 1))
    ref = F.conv2d(xin, w.to(torch.bfloat16).float(), b, padding=k // 2)
    return layer, views, out, ref, w, b


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_forward(case):
    N, H, W, chans, ctx, cout, k, act = case
    layer, views, out, ref, w, b = run_conv(*case)
    assert_close(to_nchw(out.t, cout), ref, 1e-2, f"conv fwd {case}")
    # padded output channels must be exactly zero
    if out.t.shape[1] * 8 > cout:
        assert to_nchw(out.t, out.t.shape[1] * 8)[:, cout:].abs().max().item() == 0.0


@pytest.mark.parametrize("case", [(2, 30, 45, [64], 0, 32, 3, 1),          # H != W, W not a multiple of 14
                                  (3, 13, 14, [32], 0, 16, 3, 2),          # exactly one 14-pixel tile per row, GELU input
                                  (2, 17, 29, [96, 16], 4, 24, 3, 1),      # three K-concatenated sources, cout padded to 32
                                  (8, 96, 96, [64], 0, 16, 3, 1),          # BASELINE shape, many tiles per CTA (two issuers)
                                  (1, 192, 192, [32], 0, 8, 3, 1)])
def test_conv_column_folded_matches_plain(case):
    """cg_conv_args.fold: kernel columns on the GEMM-N axis + shuffle-add epilogue == the nine-tap kernel == F.conv2d
    (src/vae.py:53-56 first conv of a Block), forward and the data gradient of the mirrored (narrow -> wide) conv."""
    from causalgen_b200 import ops
    from causalgen_b200.ops import SegSpec, View, new_act, phys
    N, H, W, chans, ctx, cout, k, act = case
    default = ops.FOLD
    try:
        ops.FOLD = 0
        layer0, _, out0, _, _, _ = run_conv(*case, seed=13)
        ops.FOLD = 2  # fold wherever the kernel can (the policy of ops.fold_pays is a speed matter, tested on the CPU)
        layer, views, out, ref, w, b = run_conv(*case, seed=13)
        assert layer.fold == 1 and layer0.fold == 0, "the case must exercise both paths"
        _folded_checks(case, layer, views, out, out0, ref)
    finally:
        ops.FOLD = default


def _folded_checks(case, layer, views, out, out0, ref):
    from causalgen_b200 import ops
    from causalgen_b200.ops import SegSpec, View, new_act, phys
    N, H, W, chans, ctx, cout, k, act = case
    got, plain = to_nchw(out.t, cout), to_nchw(out0.t, cout)
    assert_close(got, ref, 1e-2, f"folded fwd {case}")
    d = (got - plain).abs().max().item() / (ref.abs().max().item() + 1e-6)
    print(f"fold[{case}] folded vs nine-tap max rel diff {d:.3g} (bf16 rounding of the same fp32 sums)")
    assert d <= 8e-3  # one bf16 ulp of the largest output: the fp32 partial sums are only re-associated
    # fused epilogue operands on the folded path: ReLU' mask + accumulate (the data-gradient use), in place
    res = View(nhwc_bf16(rnd(N, cout, H, W, seed=60)), phys(cout))
    msk = View(nhwc_bf16(rnd(N, cout, H, W, seed=61)), phys(cout))
    want = ref * (to_nchw(msk.t, cout) > 0) + to_nchw(res.t, cout)
    layer.forward(views, [SegSpec(res, 0, add=res, mul=msk, mul_act=1)], N, H, W)(stream())
    torch.cuda.synchronize()
    assert_close(to_nchw(res.t, cout), want, 1e-2, f"folded mul+add {case}")
    if res.t.shape[1] * 8 > cout:
        assert to_nchw(out.t, out.t.shape[1] * 8)[:, cout:].abs().max().item() == 0.0
    # data gradient of the mirrored conv cout -> sum(chans) is a wide -> narrow conv when run backwards
    if len(chans) == 1 and not ctx:
        table = ops.PackTable(DEV)
        wm = rnd(chans[0], cout, 3, 3, scale=1.0 / math.sqrt(cout * 9), seed=70)   # conv cout -> chans[0]
        lm = ops.ConvLayer(table, wm, None, [cout], 0)
        table.launch(stream())
        if chans[0] >= 2 * phys(cout):
            assert lm.fold_bwd[0] == 1
        dy = View(nhwc_bf16(rnd(N, chans[0], H, W, seed=71)), phys(chans[0]))
        dx = new_act(N, H, W, cout, DEV)
        lm.dgrad(0, dy, SegSpec(dx, 0), N, H, W)(stream())
        torch.cuda.synchronize()
        refd = F.conv_transpose2d(to_nchw(dy.t, chans[0]), wm.to(torch.bfloat16).float(), padding=1)
        assert_close(to_nchw(dx.t, cout), refd, 1e-2, f"folded dgrad {case}")


def test_conv_segments_add_and_fp32_split():
    from causalgen_b200.ops import SegSpec, View, new_act
    N, H, W = 2, 12, 12
    layer, views, out, ref, w, b = run_conv(N, H, W, [32], 0, 32 + 48, 3, 1, seed=3)
    stats = View(torch.zeros(N, H, W, 32, device=DEV, dtype=torch.float32), 32)
    feat = new_act(N, H, W, 48, DEV)
    hsum = new_act(N, H, W, 48, DEV)
    h = View(nhwc_bf16(rnd(N, 48, H, W, seed=77)), 48)
    # segments are disjoint channel ranges of the conv output: [stats | feat], then [stats | feat + h]
    layer.forward(views, [SegSpec(stats, 0), SegSpec(feat, 32)], N, H, W)(stream())
    layer.forward(views, [SegSpec(stats, 0), SegSpec(hsum, 32, add=h)], N, H, W)(stream())
    torch.cuda.synchronize()
    assert_close(stats.t.permute(0, 3, 1, 2), ref[:, :32], 2e-3, "fp32 stats split")
    assert_close(to_nchw(feat.t, 48), ref[:, 32:], 1e-2, "feature split")
    assert_close(to_nchw(hsum.t, 48), ref[:, 32:] + to_nchw(h.t, 48), 1e-2, "residual add")
    # stored activation: the producer writes relu(y) / gelu(y) for consumers that all pre-activate
    for act_id in (1, 2):
        ya = new_act(N, H, W, 48, DEV)
        layer.forward(views, [SegSpec(stats, 0), SegSpec(ya, 32, add=h, out_act=act_id)], N, H, W)(stream())
        torch.cuda.synchronize()
        assert_close(to_nchw(ya.t, 48), act_fn(act_id)(ref[:, 32:] + to_nchw(h.t, 48)), 1e-2, f"out_act {act_id}")
        # dual store: raw value + activated copy (GELU blocks keep both), with and without a fused addend
        for addend in (None, h):
            raw, cp = new_act(N, H, W, 48, DEV), new_act(N, H, W, 48, DEV)
            layer.forward(views, [SegSpec(stats, 0), SegSpec(raw, 32, add=addend, out_act=act_id, act_copy=cp)], N, H, W)(stream())
            torch.cuda.synchronize()
            want = ref[:, 32:] + (to_nchw(h.t, 48) if addend is not None else 0)
            assert_close(to_nchw(raw.t, 48), want, 1e-2, f"act_copy raw {act_id}")
            assert_close(to_nchw(cp.t, 48), act_fn(act_id)(want), 1e-2, f"act_copy copy {act_id}")


@pytest.mark.parametrize("case", [(2, 16, 16, [32], 0, 16, 3, 1), (2, 24, 24, [48, 48], 4, 8, 3, 2),
                                  (3, 6, 6, [40], 0, 176, 3, 2), (2, 8, 8, [16], 20, 64, 1, 0),
                                  (4, 1, 1, [128], 0, 64, 1, 2), (1, 48, 48, [24], 0, 128, 3, 1)])
def test_conv_dgrad_and_wgrad(case):
    from causalgen_b200.ops import SegSpec, View, new_act, phys
    N, H, W, chans, ctx, cout, k, act = case
    layer, views, out, ref, w, b = run_conv(*case, seed=5)
    dy_nchw = rnd(N, cout, H, W, seed=31)
    dy = View(nhwc_bf16(dy_nchw), phys(cout), 0, cout)
    # autograd reference on the bf16-rounded operands
    is_pa = [ctx and i == 1 for i in range(len(views))]
    data_views = [v for v, p in zip(views, is_pa) if not p]
    xs = [to_nchw(v.t, c).requires_grad_(True) for v, c in zip(data_views, chans)]
    parts = list(xs)
    if ctx:
        parts.insert(1, to_nchw(views[1].t, ctx))
    wq = w.to(torch.bfloat16).float().requires_grad_(True)
    bq = b.clone().requires_grad_(True)
    y = F.conv2d(act_fn(act)(torch.cat(parts, 1)), wq, bq, padding=k // 2)
    y.backward(to_nchw(dy.t, cout))
    # data gradients, per source, with act'(x) fused and an accumulate-add
    for i, v in enumerate(views):
        if is_pa[i]:
            continue
        j = data_views.index(v)
        dx = new_act(N, H, W, chans[j], DEV)
        prev = View(nhwc_bf16(rnd(N, chans[j], H, W, seed=40 + j)), phys(chans[j]))
        layer.dgrad(i, dy, SegSpec(dx, 0, add=prev, mul=v, mul_act=act), N, H, W)(stream())
        torch.cuda.synchronize()
        assert_close(to_nchw(dx.t, chans[j]), xs[j].grad + to_nchw(prev.t, chans[j]), 1.5e-2, f"dgrad src{i} {case}")
    # weight / bias gradients (accumulated into existing values)
    dw = torch.full_like(w, 0.5)
    db = torch.full_like(b, -0.25)
    layer.wgrad(views, dy, dw, db, N, H, W)(stream())
    torch.cuda.synchronize()
    assert_close(dw - 0.5, wq.grad, 1e-2, f"wgrad {case}")
    assert_close(db + 0.25, bq.grad, 1e-2, f"bias grad {case}")


@pytest.mark.parametrize("case", [(8, 96, 96, [16], 0, 64, 3, 1), (8, 96, 96, [64], 0, 16, 3, 1),
                                  (16, 48, 48, [32], 0, 96, 3, 2), (32, 24, 24, [128, 128], 4, 32, 3, 1),
                                  (8, 96, 96, [16], 4, 64, 1, 0), (4, 192, 192, [16], 0, 32, 3, 1),
                                  # wide outputs: one CTA covers up to 128 output channels in column groups of <= 64
                                  (32, 24, 24, [32], 0, 128, 3, 1),     # Nc 128 = 64 + 64
                                  (32, 24, 24, [32], 0, 160, 3, 1),     # two CTAs of Nc 80 = 48 + 32
                                  (16, 12, 12, [48], 0, 224, 3, 1),     # two CTAs of Nc 112 = 64 + 48
                                  (16, 48, 48, [32], 0, 96, 3, 1)])     # Nc 96 = 48 + 48
def test_conv_many_tiles_per_cta(case):
    """pipeline rings wrap many times: several tiles per persistent CTA, in-place accumulate, fused operands"""
    from causalgen_b200.ops import SegSpec, View, new_act, phys
    N, H, W, chans, ctx, cout, k, act = case
    layer, views, out, ref, w, b = run_conv(*case, seed=11)
    assert_close(to_nchw(out.t, cout), ref, 1e-2, f"fwd {case}")
    # forward again with residual + second addend, accumulating IN PLACE into the first addend's buffer
    acc = View(nhwc_bf16(rnd(N, cout, H, W, seed=50)), phys(cout))
    acc0 = to_nchw(acc.t, cout).clone()
    other = View(nhwc_bf16(rnd(N, cout, H, W, seed=51)), phys(cout))
    layer.forward(views, [SegSpec(acc, 0, add=acc, add2=other)], N, H, W)(stream())
    torch.cuda.synchronize()
    assert_close(to_nchw(acc.t, cout), ref + acc0 + to_nchw(other.t, cout), 1e-2, f"fwd in-place add {case}")
    # data gradient with act'(x) and in-place accumulate; weight gradient
    dy = View(nhwc_bf16(rnd(N, cout, H, W, seed=52)), phys(cout), 0, cout)
    data_views = [v for i, v in enumerate(views) if not (ctx and i == 1)]
    xs = [to_nchw(v.t, c).requires_grad_(True) for v, c in zip(data_views, chans)]
    parts = list(xs)
    if ctx:
        parts.insert(1, to_nchw(views[1].t, ctx))
    wq = w.to(torch.bfloat16).float().requires_grad_(True)
    y = F.conv2d(act_fn(act)(torch.cat(parts, 1)), wq, None, padding=k // 2)
    y.backward(to_nchw(dy.t, cout))
    i0 = views.index(data_views[0])
    dx = View(nhwc_bf16(rnd(N, chans[0], H, W, seed=53)), phys(chans[0]))
    dx0 = to_nchw(dx.t, chans[0]).clone()
    layer.dgrad(i0, dy, SegSpec(dx, 0, add=dx, mul=data_views[0], mul_act=act), N, H, W)(stream())
    torch.cuda.synchronize()
    assert_close(to_nchw(dx.t, chans[0]), xs[0].grad + dx0, 1.5e-2, f"dgrad in-place {case}")
    dw = torch.zeros_like(w)
    layer.wgrad(views, dy, dw, None, N, H, W)(stream())
    torch.cuda.synchronize()
    assert_close(dw, wq.grad, 1e-2, f"wgrad {case}")


WGRAD_MMA_CASES = [
    # N, H, W, src channels, bcast ctx, cout, k, act -- one per template instance of wgrad_mma.cu
    (3, 20, 12, [64], 0, 16, 3, 1),        # dY narrow (16) x X 64: ragged H and W
    (2, 9, 9, [16], 0, 8, 3, 0),           # both 16, odd size, no activation
    (2, 24, 24, [96, 96], 4, 24, 3, 1),    # posterior-style 3 sources, dY 32
    (2, 8, 8, [16, 16], 0, 32, 3, 1),      # two narrow sources, dY 32
    (2, 12, 12, [160], 0, 40, 3, 1),       # dY 48
    (2, 10, 10, [16], 0, 64, 3, 1),        # X narrow (16), dY 64
    (1, 48, 48, [8], 0, 32, 3, 1),         # X 8 (padded 16), dY 32
    (2, 24, 24, [32], 0, 128, 3, 1),       # X 32, dY 128 -> two dY chunks
    (3, 6, 6, [40], 0, 176, 3, 1),         # X 48, dY 176 -> eleven dY chunks
    (40, 24, 24, [128], 0, 32, 3, 1),      # many tiles per CTA, ring wraps
    (6, 96, 96, [16], 0, 64, 3, 1),
    # 1x1 problems (z_proj / z_feat_proj / width_proj shapes) on the tap-less variant
    (2, 24, 24, [16, 128], 0, 128, 1, 0),  # cat[z, p_feat] -> out: X chunks 16 | 64 | 64, dY chunks 48 | 48 | 32
    (3, 12, 12, [160], 0, 192, 1, 0),      # width_proj
    (2, 48, 48, [16], 4, 96, 1, 0),        # cat[z, pa] -> h: narrow X (NT=2), dY 96 in one chunk
    (9, 20, 12, [32], 0, 40, 1, 1),        # ragged tile edges, ReLU, dY 48 padded
]


@pytest.mark.parametrize("case", WGRAD_MMA_CASES)
def test_wgrad_mma_small_channel_3x3(case):
    """weight + bias gradient of the warp-level mma.sync kernel vs autograd on the same bf16 operands"""
    from causalgen_b200.ops import View, phys
    N, H, W, chans, ctx, cout, k, act = case
    layer, views, out, ref, w, b = run_conv(*case, seed=17)
    dy = View(nhwc_bf16(rnd(N, cout, H, W, seed=61)), phys(cout), 0, cout)
    logical = list(chans)
    if ctx:
        logical.insert(1, ctx)
    parts = [to_nchw(v.t, c) for v, c in zip(views, logical)]
    wq = w.to(torch.bfloat16).float().requires_grad_(True)
    bq = b.clone().requires_grad_(True)
    y = F.conv2d(act_fn(act)(torch.cat(parts, 1)), wq, bq, padding=k // 2)
    y.backward(to_nchw(dy.t, cout))
    dw = torch.full_like(w, 0.5)
    db = torch.full_like(b, -0.25)
    for _ in range(2):  # accumulates: two launches = twice the gradient
        layer.wgrad(views, dy, dw, db, N, H, W)(stream())
    torch.cuda.synchronize()
    assert_close(dw - 0.5, 2 * wq.grad, 1e-2, f"wgrad {case}")
    assert_close(db + 0.25, 2 * bq.grad, 1e-2, f"bias grad {case}")


def test_conv_centre_tap_on_1x1_image():
    from causalgen_b200.ops import ConvLayer, PackTable, SegSpec, View, new_act
    N, Cin, Cout = 6, 64, 48
    x = rnd(N, Cin, 1, 1, seed=1)
    w = rnd(Cout, Cin, 3, 3, scale=0.05, seed=2)
    b = rnd(Cout, scale=0.1, seed=3)
    table = PackTable(DEV)
    layer = ConvLayer(table, w, b, [Cin], 2, centre_only=True)
    table.launch(stream())
    xv = View(nhwc_bf16(x), Cin)
    out = new_act(N, 1, 1, Cout, DEV)
    layer.forward([xv], [SegSpec(out, 0)], N, 1, 1)(stream())
    dy = View(nhwc_bf16(rnd(N, Cout, 1, 1, seed=4)), Cout)
    dw = torch.zeros_like(w)
    layer.wgrad([xv], dy, dw, None, N, 1, 1)(stream())
    torch.cuda.synchronize()
    xq = to_nchw(xv.t, Cin)
    wq = w.to(torch.bfloat16).float().requires_grad_(True)
    ref = F.conv2d(F.gelu(xq), wq, b, padding=1)
    assert_close(to_nchw(out.t, Cout), ref, 1e-2, "centre tap fwd")
    ref.backward(to_nchw(dy.t, Cout))
    assert_close(dw, wq.grad, 1e-2, "centre tap wgrad")
    assert dw[:, :, 0, 0].abs().max().item() == 0.0


@pytest.mark.parametrize("cin,R,cout", [(1, 32, 16), (3, 32, 16), (1, 48, 32), (1, 28, 16)])  # 28: not a multiple of the tile
def test_stem(cin, R, cout):
    from causalgen_b200 import _lib as L
    lib = L.load()
    N = 2
    x = rnd(N, cin, R, R, seed=1)
    w = rnd(cout, cin, 7, 7, scale=0.1, seed=2)
    b = rnd(cout, scale=0.1, seed=3)
    y = torch.zeros(N, cout // 8, R, R, 8, device=DEV, dtype=torch.bfloat16)
    L.check(lib.cg_stem_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), N, cin, R, cout, ns_of(y), stream()))
    wr = w.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    ref = F.conv2d(x, wr, br, padding=3)
    torch.cuda.synchronize()
    assert_close(to_nchw(y, cout), ref, 1e-2, "stem fwd")
    dy = nhwc_bf16(rnd(N, cout, R, R, seed=4), cout)
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    L.check(lib.cg_stem_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), N, cin, R, cout, ns_of(dy), stream()))
    ref.backward(to_nchw(dy, cout))
    torch.cuda.synchronize()
    assert_close(dw, wr.grad, 2e-3, "stem wgrad")
    assert_close(db, br.grad, 2e-3, "stem bias grad")


def test_pool_and_upsample():
    from causalgen_b200 import _lib as L
    lib = L.load()
    N, C, H = 2, 32, 24

    def zeros(res):
        return torch.zeros(N, C // 8, res, res, 8, device=DEV, dtype=torch.bfloat16)

    x = rnd(N, C, H, H, seed=1)
    xb = nhwc_bf16(x)
    for d in (2, 4, 6):
        y = zeros(H // d)
        L.check(lib.cg_avgpool_fwd(xb.data_ptr(), y.data_ptr(), N, H, H, C, d, ns_of(xb), ns_of(y), 0, stream()))
        assert_close(to_nchw(y, C), F.avg_pool2d(to_nchw(xb, C), d, d), 1e-2, f"pool {d}")
        dy = nhwc_bf16(rnd(N, C, H // d, H // d, seed=2))
        dx = torch.zeros_like(xb)
        L.check(lib.cg_avgpool_bwd(dy.data_ptr(), dx.data_ptr(), N, H, H, C, d, ns_of(dy), ns_of(dx), 0, 0, stream()))
        xr = to_nchw(xb, C).requires_grad_(True)
        F.avg_pool2d(xr, d, d).backward(to_nchw(dy, C))
        assert_close(to_nchw(dx, C), xr.grad, 1e-2, f"pool bwd {d}")
    # odd resolution: 14 -> 7 padded to 8 (src/vae.py:130-132)
    x14 = nhwc_bf16(rnd(N, C, 14, 14, seed=3))
    y8 = torch.ones(N, C // 8, 8, 8, 8, device=DEV, dtype=torch.bfloat16)
    L.check(lib.cg_avgpool_fwd(x14.data_ptr(), y8.data_ptr(), N, 14, 14, C, 2, ns_of(x14), ns_of(y8), 8, stream()))
    assert_close(to_nchw(y8, C), F.pad(F.avg_pool2d(to_nchw(x14, C), 2, 2), [0, 1, 0, 1]), 1e-2, "pool+pad")
    # ... and its gradient: the 8x8 gradient plane's last row / column (the padding) reaches no input pixel
    dy8 = nhwc_bf16(rnd(N, C, 8, 8, seed=9))
    dx14 = torch.ones_like(x14)
    L.check(lib.cg_avgpool_bwd(dy8.data_ptr(), dx14.data_ptr(), N, 14, 14, C, 2, ns_of(dy8), ns_of(dx14), 8, 0, stream()))
    xr14 = to_nchw(x14, C).requires_grad_(True)
    F.pad(F.avg_pool2d(xr14, 2, 2), [0, 1, 0, 1]).backward(to_nchw(dy8, C))
    assert_close(to_nchw(dx14, C), xr14.grad, 1e-2, "pool+pad bwd")
    # F.avg_pool2d floors: an 8x8 map (7 zero-padded to 8) pooled by 7 -> 1x1 over the top-left 7x7 window
    x8 = nhwc_bf16(rnd(N, C, 8, 8, seed=7))
    y1 = zeros(1)
    L.check(lib.cg_avgpool_fwd(x8.data_ptr(), y1.data_ptr(), N, 8, 8, C, 7, ns_of(x8), ns_of(y1), 0, stream()))
    xr = to_nchw(x8, C).requires_grad_(True)
    pooled = F.avg_pool2d(xr, 7, 7)
    assert_close(to_nchw(y1, C), pooled, 1e-2, "pool 8 by 7")
    dy1 = nhwc_bf16(rnd(N, C, 1, 1, seed=8))
    dx8 = torch.ones_like(x8)
    L.check(lib.cg_avgpool_bwd(dy1.data_ptr(), dx8.data_ptr(), N, 8, 8, C, 7, ns_of(dy1), ns_of(dx8), 0, 0, stream()))
    pooled.backward(to_nchw(dy1, C))
    assert_close(to_nchw(dx8, C), xr.grad, 1e-2, "pool bwd 8 by 7")
    # nearest upsample + learned bias, integer and 7->8 style factors; 11 samples = one full chunk of 8 + a tail of 3
    for hi, ho, N in ((6, 12, 2), (1, 4, 2), (7, 8, 2), (8, 14, 2), (6, 12, 11)):
        xs = nhwc_bf16(rnd(N, C, hi, hi, seed=4))
        bias = rnd(1, C, ho, ho, seed=5)
        y = torch.zeros(N, C // 8, ho, ho, 8, device=DEV, dtype=torch.bfloat16)
        L.check(lib.cg_upsample_fwd(xs.data_ptr(), bias.data_ptr(), y.data_ptr(), N, hi, ho, C, ns_of(xs), ns_of(y), stream()))
        xr = to_nchw(xs, C).requires_grad_(True)
        br = bias.clone().requires_grad_(True)
        ref = br + F.interpolate(xr, scale_factor=ho / hi)
        assert_close(to_nchw(y, C), ref, 1e-2, f"upsample {hi}->{ho}")
        dy = nhwc_bf16(rnd(N, C, ho, ho, seed=6))
        dx = torch.zeros_like(xs)
        dbias = torch.zeros_like(bias)
        L.check(lib.cg_upsample_bwd(dy.data_ptr(), dx.data_ptr(), dbias.data_ptr(), N, hi, ho, C, ns_of(dy), ns_of(dx), 0,
                                    stream()))
        ref.backward(to_nchw(dy, C))
        assert_close(to_nchw(dx, C), xr.grad, 1e-2, f"upsample bwd {hi}->{ho}")
        assert_close(dbias, br.grad, 2e-3, f"upsample dbias {hi}->{ho}")


def test_latent_forward_backward():
    import ctypes as C
    from causalgen_b200 import _lib as L
    lib = L.load()
    N, H, zd = 3, 12, 16
    HW = H * H
    q = rnd(N, HW, 32, seed=1, scale=0.7).contiguous()
    p = rnd(N, HW, 32, seed=2, scale=0.7).contiguous()
    eps = rnd(N, zd, H, H, seed=3)
    z16 = torch.zeros(N, 2, H, H, 8, device=DEV, dtype=torch.bfloat16)
    z32 = torch.zeros(N, zd, H, H, device=DEV)
    kl = torch.zeros(N, device=DEV)
    a = L.LatentArgs()
    a.q, a.p, a.q_ld, a.p_ld, a.eps = q.data_ptr(), p.data_ptr(), 32, 32, eps.data_ptr()
    a.log_t, a.z_bf16, a.z_ns, a.z_f32, a.kl_out = math.log(0.9), z16.data_ptr(), ns_of(z16), z32.data_ptr(), kl.data_ptr()
    a.N, a.HW, a.zdim, a.mode = N, HW, zd, 0
    L.check(lib.cg_latent_fwd(C.byref(a), stream()))
    qn = q.view(N, H, H, 32).permute(0, 3, 1, 2)
    pn = p.view(N, H, H, 32).permute(0, 3, 1, 2)
    ql, qs, pl, ps = qn[:, :16], qn[:, 16:] + math.log(0.9), pn[:, :16], pn[:, 16:] + math.log(0.9)
    zr = ql + qs.exp() * eps
    klr = (-0.5 + ps - qs + 0.5 * (qs.exp() ** 2 + (ql - pl) ** 2) / ps.exp() ** 2).sum(dim=(1, 2, 3))
    torch.cuda.synchronize()
    assert_close(z32, zr, 1e-5, "z fp32")
    assert_close(to_nchw(z16, 16), zr, 1e-2, "z bf16")
    assert_close(kl, klr, 1e-4, "kl")
    # backward (t = None)
    dz = nhwc_bf16(rnd(N, 16, H, H, seed=4))
    dq = torch.zeros(N, 4, H, H, 8, device=DEV, dtype=torch.bfloat16)
    dp = torch.zeros(N, 6, H, H, 8, device=DEV, dtype=torch.bfloat16)
    b = L.LatentBwdArgs()
    b.q, b.p, b.q_ld, b.p_ld, b.eps = q.data_ptr(), p.data_ptr(), 32, 32, eps.data_ptr()
    b.dz, b.dz_ns, b.g_kl = dz.data_ptr(), ns_of(dz), 0.37
    b.dq, b.dq_ns, b.dp, b.dp_ns = dq.data_ptr(), ns_of(dq), dp.data_ptr(), ns_of(dp)
    b.N, b.HW, b.zdim, b.mode = N, HW, zd, 0
    L.check(lib.cg_latent_bwd(C.byref(b), stream()))
    qr = q.clone().requires_grad_(True)
    pr = p.clone().requires_grad_(True)
    qn = qr.view(N, H, H, 32).permute(0, 3, 1, 2)
    pn = pr.view(N, H, H, 32).permute(0, 3, 1, 2)
    ql, qs, pl, ps = qn[:, :16], qn[:, 16:], pn[:, :16], pn[:, 16:]
    z = ql + qs.exp() * eps
    klv = (-0.5 + ps - qs + 0.5 * (qs.exp() ** 2 + (ql - pl) ** 2) / ps.exp() ** 2).sum()
    loss = 0.37 * klv + (z * to_nchw(dz, 16)).sum()
    loss.backward()
    torch.cuda.synchronize()
    assert_close(to_nchw(dq, 32), qr.grad.view(N, H, H, 32).permute(0, 3, 1, 2), 1e-2, "dq")
    assert_close(to_nchw(dp, 32), pr.grad.view(N, H, H, 32).permute(0, 3, 1, 2), 1e-2, "dp")


def test_latent_philox_stream_kernels_regenerate_forward_noise():
    """in-kernel Philox mode: the streaming forward kernel, the layout-staging forward kernel and both backward
    kernels must agree on the noise of every (sample, channel, pixel) -- backward never stores eps"""
    import ctypes as C
    from causalgen_b200 import _lib as L
    lib = L.load()
    N, H, zd = 3, 20, 16
    HW = H * H
    q = rnd(N, HW, 32, seed=11, scale=0.7).contiguous()
    p = rnd(N, HW, 32, seed=12, scale=0.7).contiguous()

    def fwd(with_f32):
        z16 = torch.zeros(N, 2, H, H, 8, device=DEV, dtype=torch.bfloat16)
        z32 = torch.zeros(N, zd, H, H, device=DEV)
        eo = torch.zeros(N, zd, H, H, device=DEV)
        kl = torch.zeros(N, device=DEV)
        a = L.LatentArgs()
        a.q, a.p, a.q_ld, a.p_ld = q.data_ptr(), p.data_ptr(), 32, 32
        a.seed, a.offset = 0x1234567, 5 << 40
        a.z_bf16, a.z_ns, a.kl_out, a.eps_out = z16.data_ptr(), ns_of(z16), kl.data_ptr(), eo.data_ptr()
        if with_f32:
            a.z_f32 = z32.data_ptr()
        a.N, a.HW, a.zdim, a.mode = N, HW, zd, 0
        L.check(lib.cg_latent_fwd(C.byref(a), stream()))
        torch.cuda.synchronize()
        return z16, eo, kl

    z_s, e_s, kl_s = fwd(False)   # streaming kernel
    z_l, e_l, kl_l = fwd(True)    # layout-staging kernel (abduct-style output)
    assert_close(e_s, e_l, 1e-6, "the two forward kernels must draw identical noise")
    assert abs(e_s.mean().item()) < 0.05 and abs(e_s.std().item() - 1.0) < 0.05
    qn = q.view(N, H, H, 32).permute(0, 3, 1, 2)
    assert_close(to_nchw(z_s, 16), qn[:, :16] + qn[:, 16:].exp() * e_s, 1e-2, "z from streamed noise")
    assert_close(kl_s, kl_l, 1e-5, "kl")
    dz = nhwc_bf16(rnd(N, 16, H, H, seed=14))

    def bwd(eps):
        dq = torch.zeros(N, 4, H, H, 8, device=DEV, dtype=torch.bfloat16)
        dp = torch.zeros(N, 6, H, H, 8, device=DEV, dtype=torch.bfloat16)
        b = L.LatentBwdArgs()
        b.q, b.p, b.q_ld, b.p_ld = q.data_ptr(), p.data_ptr(), 32, 32
        b.seed, b.offset = 0x1234567, 5 << 40
        if eps is not None:
            b.eps = eps.data_ptr()
        b.dz, b.dz_ns, b.g_kl = dz.data_ptr(), ns_of(dz), 0.37
        b.dq, b.dq_ns, b.dp, b.dp_ns = dq.data_ptr(), ns_of(dq), dp.data_ptr(), ns_of(dp)
        b.N, b.HW, b.zdim, b.mode = N, HW, zd, 0
        L.check(lib.cg_latent_bwd(C.byref(b), stream()))
        torch.cuda.synchronize()
        return dq, dp

    dq_s, dp_s = bwd(None)     # regenerates the noise
    dq_e, dp_e = bwd(e_s)      # explicit noise of the forward pass
    assert_close(to_nchw(dq_s, 32), to_nchw(dq_e, 32), 1e-3, "dq with regenerated noise")
    assert_close(to_nchw(dp_s, 32), to_nchw(dp_e, 32), 1e-3, "dp with regenerated noise")


@pytest.mark.parametrize("Cc,Cw", [(1, 32), (3, 16)])
def test_dgauss_forward_backward_sample(Cc, Cw):
    import ctypes as C
    from causalgen_b200 import _lib as L
    lib = L.load()
    N, R = 2, 20
    HW = R * R
    h = nhwc_bf16(rnd(N, Cw, R, R, seed=1))
    x8 = torch.randint(0, 256, (N, Cc, R, R), generator=torch.Generator().manual_seed(2))
    x8[torch.rand(N, Cc, R, R, generator=torch.Generator().manual_seed(3)) < 0.4] = 0
    x8[torch.rand(N, Cc, R, R, generator=torch.Generator().manual_seed(4)) < 0.05] = 255
    x = ((x8.float() - 127.5) / 127.5).to(DEV)
    ws = [rnd(Cc, Cw, seed=10 + i, scale=0.3) for i in range(3)]
    bs = [rnd(Cc, seed=20 + i, scale=0.3) for i in range(3)]
    bs[1] = bs[1] - 2.0
    nll = torch.zeros(N, device=DEV)
    a = L.DGaussArgs()
    a.h, a.h_ns, a.Cw, a.x = h.data_ptr(), ns_of(h), Cw, x.data_ptr()
    a.w_loc, a.b_loc, a.w_ls, a.b_ls = ws[0].data_ptr(), bs[0].data_ptr(), ws[1].data_ptr(), bs[1].data_ptr()
    if Cc == 3:
        a.w_co, a.b_co = ws[2].data_ptr(), bs[2].data_ptr()
    a.N, a.HW, a.C, a.nll = N, HW, Cc, nll.data_ptr()
    L.check(lib.cg_dgauss_nll_fwd(C.byref(a), stream()))

    def ref_nll(hh, w, b):
        loc = F.conv2d(hh, w[0][:, :, None, None], b[0])
        ls = F.conv2d(hh, w[1][:, :, None, None], b[1]).clamp(min=-9)
        if Cc == 3:
            co = torch.tanh(F.conv2d(hh, w[2][:, :, None, None], b[2]))
            loc = torch.stack([loc[:, 0], loc[:, 1] + co[:, 0] * x[:, 0],
                               loc[:, 2] + co[:, 1] * x[:, 0] + co[:, 2] * x[:, 1]], 1)
        cdf = lambda v: 0.5 * (1 + torch.tanh(math.sqrt(2 / math.pi) * (v + 0.044715 * v ** 3)))
        inv = torch.exp(-ls)
        hi, lo = cdf(inv * (x - loc + 1 / 255)), cdf(inv * (x - loc - 1 / 255))
        lp = torch.where(x < -0.999, hi.clamp(min=1e-12).log(),
                         torch.where(x > 0.999, (1 - lo).clamp(min=1e-12).log(), (hi - lo).clamp(min=1e-12).log()))
        return -lp.mean(dim=(1, 2, 3))

    hr = to_nchw(h, Cw).requires_grad_(True)
    wr = [w.clone().requires_grad_(True) for w in ws]
    br = [b.clone().requires_grad_(True) for b in bs]
    ref = ref_nll(hr, wr, br)
    torch.cuda.synchronize()
    assert_close(nll, ref, 2e-4, "dgauss nll")
    # backward
    dh = torch.zeros_like(h)
    dws = [torch.zeros_like(w) for w in ws]
    dbs = [torch.zeros_like(b) for b in bs]
    a.g, a.dh, a.dh_ns = 0.5, dh.data_ptr(), ns_of(dh)
    a.dw_loc, a.db_loc, a.dw_ls, a.db_ls = dws[0].data_ptr(), dbs[0].data_ptr(), dws[1].data_ptr(), dbs[1].data_ptr()
    if Cc == 3:
        a.dw_co, a.db_co = dws[2].data_ptr(), dbs[2].data_ptr()
    L.check(lib.cg_dgauss_nll_bwd(C.byref(a), stream()))
    (0.5 * ref.sum()).backward()
    torch.cuda.synchronize()
    assert_close(to_nchw(dh, Cw), hr.grad, 1e-2, "dgauss dh")
    for i in range(3 if Cc == 3 else 2):
        assert_close(dws[i], wr[i].grad, 2e-3, f"dgauss dw{i}")
        assert_close(dbs[i], br[i].grad, 2e-3, f"dgauss db{i}")
    # sample (return_loc=True)
    xo, so = torch.zeros_like(x), torch.zeros_like(x)
    L.check(lib.cg_dgauss_sample(C.byref(a), xo.data_ptr(), so.data_ptr(), None, 0.0, stream()))
    hh = to_nchw(h, Cw)
    loc = F.conv2d(hh, ws[0][:, :, None, None], bs[0])
    ls = F.conv2d(hh, ws[1][:, :, None, None], bs[1]).clamp(min=-9)
    if Cc == 3:
        co = torch.tanh(F.conv2d(hh, ws[2][:, :, None, None], bs[2]))
        r = loc[:, 0].clamp(-1, 1)
        g = (loc[:, 1] + co[:, 0] * r).clamp(-1, 1)
        bl = (loc[:, 2] + co[:, 1] * r + co[:, 2] * g).clamp(-1, 1)
        loc = torch.stack([r, g, bl], 1)
    torch.cuda.synchronize()
    assert_close(xo, loc.clamp(-1, 1), 1e-4, "dgauss sample loc")
    assert_close(so, ls.exp(), 1e-4, "dgauss sample scale")


def test_dmol_against_oracle_functions():
    """DMoL kernels vs the CPU oracle's restatement of src/dmol.py on the same head outputs."""
    import ctypes as C
    import hvae_oracle as O
    from causalgen_b200 import _lib as L
    lib = L.load()
    N, R, Cw = 2, 12, 16
    HW = R * R
    h = nhwc_bf16(rnd(N, Cw, R, R, seed=1))
    x8 = torch.randint(0, 256, (N, 3, R, R), generator=torch.Generator().manual_seed(2))
    x8[torch.rand(N, 3, R, R, generator=torch.Generator().manual_seed(3)) < 0.3] = 0
    x8[torch.rand(N, 3, R, R, generator=torch.Generator().manual_seed(4)) < 0.1] = 255
    x = ((x8.float() - 127.5) / 127.5).to(DEV)
    w = rnd(100, Cw, seed=5, scale=0.5)
    b = rnd(100, seed=6, scale=0.5)
    nll = torch.zeros(N, device=DEV)
    a = L.DmolArgs()
    a.h, a.h_ns, a.Cw, a.x, a.w, a.b = h.data_ptr(), ns_of(h), Cw, x.data_ptr(), w.data_ptr(), b.data_ptr()
    a.N, a.HW, a.nll = N, HW, nll.data_ptr()
    L.check(lib.cg_dmol_loss_fwd(C.byref(a), stream()))
    hc = to_nchw(h, Cw).cpu().requires_grad_(True)
    wc = w.cpu().clone().requires_grad_(True)
    bc = b.cpu().clone().requires_grad_(True)
    l = F.conv2d(hc, wc[:, :, None, None], bc).permute(0, 2, 3, 1)
    ref = O.dmol_loss(x.cpu().permute(0, 2, 3, 1), l)
    torch.cuda.synchronize()
    assert_close(nll.cpu(), ref.detach(), 2e-4, "dmol loss")
    dh = torch.zeros_like(h)
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    a.g, a.dh, a.dh_ns, a.dw, a.db = 0.5, dh.data_ptr(), ns_of(dh), dw.data_ptr(), db.data_ptr()
    L.check(lib.cg_dmol_loss_bwd(C.byref(a), stream()))
    (0.5 * ref.sum()).backward()
    torch.cuda.synchronize()
    assert_close(to_nchw(dh, Cw).cpu(), hc.grad, 1e-2, "dmol dh")
    assert_close(dw.cpu(), wc.grad, 2e-3, "dmol dw")
    assert_close(db.cpu(), bc.grad, 2e-3, "dmol db")
    # predictions
    xo, so = torch.zeros_like(x), torch.zeros_like(x)
    for mode, mask in ((0, "soft"), (1, "hard"), (13, "top3"), (11, "top1"), (19, "top9")):
        L.check(lib.cg_dmol_predict(C.byref(a), mode, None, None, 0.0, xo.data_ptr(), so.data_ptr(), stream()))
        m, s = O.dmol_mean(l.detach(), mask=mask)
        torch.cuda.synchronize()
        assert_close(xo.cpu(), m.permute(0, 3, 1, 2), 1e-4, f"dmol mean {mask}")
        assert_close(so.cpu(), s.permute(0, 3, 1, 2), 1e-4, f"dmol scale {mask}")
    ug = torch.rand(N, R, R, 10, generator=torch.Generator().manual_seed(7)).clamp(1e-5, 1 - 1e-5)
    ul = torch.rand(N, R, R, 3, generator=torch.Generator().manual_seed(8)).clamp(1e-5, 1 - 1e-5)
    ugd, uld = ug.to(DEV), ul.to(DEV)
    L.check(lib.cg_dmol_predict(C.byref(a), 2, ugd.data_ptr(), uld.data_ptr(), math.log(0.7), xo.data_ptr(),
                                so.data_ptr(), stream()))
    sx, ss = O.dmol_sample(l.detach(), O.NoiseTape([ug, ul]), t=0.7)
    torch.cuda.synchronize()
    # argmax ties aside, identical uniforms give identical component picks
    frac = ((xo.cpu() - sx.permute(0, 3, 1, 2)).abs() < 1e-3).float().mean().item()
    assert frac > 0.995, frac


def test_mix_cf_and_layout_glue():
    from causalgen_b200 import _lib as L
    lib = L.load()
    n = 5000
    z, ql, qs, pl, ps = (rnd(n, seed=i, scale=0.5) for i in range(5))
    out = torch.zeros(n, device=DEV)
    L.check(lib.cg_latent_mix(z.data_ptr(), ql.data_ptr(), qs.data_ptr(), pl.data_ptr(), ps.data_ptr(), out.data_ptr(),
                              n, 0.65, 0.8, 1, stream()))
    u = (z - ql) / qs.exp()
    ref = 0.65 * ql + 0.35 * pl + (0.65 ** 2 * qs.exp() ** 2 + 0.35 ** 2 * ps.exp() ** 2).sqrt() * 0.8 * u
    assert_close(out, ref, 1e-5, "latent mix")
    x, rl, cl = (rnd(n, seed=10 + i, scale=0.5) for i in range(3))
    rs, cs = rnd(n, seed=20).abs() + 0.01, rnd(n, seed=21).abs() + 0.01
    cf = torch.zeros(n, device=DEV)
    s1, s2 = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for _ in range(2):
        L.check(lib.cg_cf_combine(x.data_ptr(), rl.data_ptr(), rs.data_ptr(), cl.data_ptr(), cs.data_ptr(),
                                  cf.data_ptr(), s1.data_ptr(), s2.data_ptr(), n, stream()))
    ref = (cl + cs * (x - rl) / rs.clamp(min=1e-12)).clamp(-1, 1)
    assert_close(cf, ref, 1e-6, "cf combine")
    assert_close(s1, 2 * ref, 1e-6, "cf sum")
    assert_close(s2, 2 * ref ** 2, 1e-6, "cf sum2")
    # layout glue round trip
    t = rnd(2, 20, 9, 9, seed=30)
    nh = torch.ones(2, 4, 9, 9, 8, device=DEV, dtype=torch.bfloat16)
    L.check(lib.cg_nchw_f32_to_planar(t.data_ptr(), nh.data_ptr(), 2, 20, 81, ns_of(nh), stream()))
    assert_close(to_nchw(nh, 20), t.to(torch.bfloat16).float(), 1e-6, "nchw -> planar")
    assert to_nchw(nh, 24)[:, 20:].abs().max().item() == 0 and nh[:, 3].float().abs().max().item() == 1.0
    back = torch.zeros_like(t)
    L.check(lib.cg_planar_to_nchw_f32(nh.data_ptr(), back.data_ptr(), 2, 20, 81, ns_of(nh), stream()))
    assert_close(back, t.to(torch.bfloat16).float(), 1e-6, "layout round trip")
    pa = rnd(3, 12, seed=31)
    pl = torch.zeros(3, 2, 5, 5, 8, device=DEV, dtype=torch.bfloat16)
    L.check(lib.cg_parents_plane(pa.data_ptr(), 12, 1, pl.data_ptr(), 3, 12, 16, 25, ns_of(pl), 2, 1.0,
                                 torch.zeros(1, device=DEV).data_ptr(), stream()))  # device scalar overrides the 1.0
    want = pa.clone()
    want[:, 2:] = 0
    assert_close(to_nchw(pl, 12), want.to(torch.bfloat16).float()[:, :, None, None].expand(3, 12, 5, 5), 1e-6, "parents plane")


@pytest.mark.parametrize("N,HW,C", [(2, 1, 288), (5, 1000, 32), (3, 259, 24), (4, 9216, 16), (3, 1, 544), (7, 13, 8)])
def test_colsum_bias_gradient(N, HW, C):
    from causalgen_b200 import _lib as L
    dy = nhwc_bf16(rnd(N, C, HW, 1, seed=3))
    out = torch.full((C,), 0.5, device=DEV)
    L.check(L.load().cg_colsum(dy.data_ptr(), out.data_ptr(), N, HW, C, ns_of(dy), stream()))
    assert_close(out - 0.5, to_nchw(dy, C).sum(dim=(0, 2, 3)), 2e-3, f"colsum {N}x{HW}x{C}")


def test_optimizer_tail_matches_torch_adamw():
    from causalgen_b200 import _lib as L
    lib = L.load()
    n = 10007
    p0 = rnd(n, seed=1)
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref_p], lr=1e-3, weight_decay=0.05, betas=(0.9, 0.9))
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda it: 1.0 if it > 100 else it / 100)
    p = p0.clone()
    m, v, ema = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV), p0.clone()
    state = torch.zeros(4, dtype=torch.int32, device=DEV)
    dyn = torch.zeros(6, device=DEV)
    ss = torch.zeros(1, device=DEV)
    for it in range(5):
        g = rnd(n, seed=100 + it, scale={2: 30.0, 1: 4.0}.get(it, 1.0))
        ref_p.grad = g.clone()
        gn = torch.nn.utils.clip_grad_norm_([ref_p], 350.0)
        if gn < 500:
            opt.step()
            sched.step()
        ss.zero_()
        L.check(lib.cg_sumsq(g.data_ptr(), ss.data_ptr(), n, None, 0, stream()))
        L.check(lib.cg_optim_advance(state.data_ptr(), dyn.data_ptr(), ss.data_ptr(), None, 1e-3, 100, 0.9, 0.9, 350.0,
                                     500.0, 1.0, 0.999, 100, stream()))
        L.check(lib.cg_adamw_ema_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), ema.data_ptr(), n,
                                      state.data_ptr(), dyn.data_ptr(), 0.9, 0.9, 1e-8, 0.05, stream()))
        torch.cuda.synchronize()
        assert abs(dyn[5].item() - gn.item()) <= 1e-3 * gn.item()
    assert_close(p, ref_p.detach(), 1e-5, "adamw params")
    assert state[2].item() == 1 and state[0].item() == 4  # the 30x gradient step was skipped
    assert_close(ema, p, 1e-6, "ema copies params during its first 100 calls")


def test_ema_decay_schedule_on_device():
    """cg_optim_advance's EMA decay (dyn[3]) against the rule pinned to the reference trajectory
    (oracle ema_decay <- tests/golden/ema_schedule.npz): copy phase, one more copy, inverse-decay ramp"""
    import hvae_oracle as O
    from causalgen_b200 import _lib as L
    lib = L.load()
    state = torch.zeros(4, dtype=torch.int32, device=DEV)
    dyn = torch.zeros(6, device=DEV)
    ss = torch.full((1,), 1e-4, device=DEV)  # sum of squares of a tiny gradient: never skipped
    got = []
    for s in range(130):
        L.check(lib.cg_optim_advance(state.data_ptr(), dyn.data_ptr(), ss.data_ptr(), None, 1e-3, 100, 0.9, 0.9, 350.0,
                                     500.0, 1.0, 0.999, 100, stream()))
        got.append(dyn[3].item())
    want = [O.ema_decay(s, 0.999, 100) for s in range(130)]
    assert max(abs(a - b) for a, b in zip(got, want)) <= 1e-6, list(zip(got, want))[98:106]
    assert state[1].item() == 130 and state[2].item() == 0


def test_normalise_u8_and_deterministic_sumsq():
    """cg_normalise_u8 == trainer.preprocess_batch (src/trainer.py:17) bit for bit on every uint8 value (fixture produced by
    the reference function, tests/golden/preprocess.npz); cg_sumsq with scratch is run-to-run bit-identical"""
    import os
    import numpy as np
    from causalgen_b200 import _lib as L
    lib = L.load()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "preprocess.npz"))
    x8 = torch.from_numpy(g["x8"]).to(DEV)
    out = torch.empty(x8.shape, device=DEV, dtype=torch.float32)
    L.check(lib.cg_normalise_u8(x8.data_ptr(), out.data_ptr(), x8.numel(), stream()))
    assert torch.equal(out.cpu(), torch.from_numpy(g["x_norm"]))
    allv = torch.arange(256, dtype=torch.uint8, device=DEV).repeat(5)[:1279]  # ragged length, every value
    out = torch.empty(allv.numel(), device=DEV, dtype=torch.float32)
    L.check(lib.cg_normalise_u8(allv.data_ptr(), out.data_ptr(), allv.numel(), stream()))
    assert torch.equal(out.cpu(), (allv.cpu().float() - 127.5) / 127.5)
    n = 3_000_003
    v = torch.randn(n, device=DEV)
    scratch = torch.zeros(592, device=DEV)
    res = []
    for _ in range(3):
        o = torch.zeros(1, device=DEV)
        L.check(lib.cg_sumsq(v.data_ptr(), o.data_ptr(), n, scratch.data_ptr(), 592, stream()))
        res.append(o.item())
    assert res[0] == res[1] == res[2]
    assert abs(res[0] - float((v.double() ** 2).sum())) <= 1e-5 * res[0]


@pytest.mark.parametrize("Cc,Cw", [(1, 32), (3, 16)])
def test_cf_combine_and_sample_backward(Cc, Cw):
    """the two kernels the counterfactual backward adds (src/pgm/dscm.py:55-56, src/vae.py:352-369,413-422) against torch
    autograd of the same expressions, fp32: 2e-3 of the tensor scale (h is bf16-rounded on both sides)"""
    import ctypes as C
    from causalgen_b200 import _lib as L
    lib = L.load()
    N, H = 3, 12
    HW = H * H
    n = N * Cc * HW
    g = torch.Generator().manual_seed(4)
    x = (torch.rand(N, Cc, H, H, generator=g) * 1.6 - 0.8).to(DEV)
    t = {k: torch.randn(N, Cc, H, H, generator=g).to(DEV) for k in ("rec_loc", "cf_loc", "dcf")}
    t["rec_loc"] = (t["rec_loc"] * 0.3).requires_grad_(True)
    t["cf_loc"] = (t["cf_loc"] * 0.3).requires_grad_(True)
    rec_scale = (torch.rand(N, Cc, H, H, generator=g) * 0.5 + 0.2).to(DEV).requires_grad_(True)
    cf_scale = (torch.rand(N, Cc, H, H, generator=g) * 0.5 + 0.2).to(DEV).requires_grad_(True)
    u = (x - t["rec_loc"]) / rec_scale.clamp(min=1e-12)
    cf = torch.clamp(t["cf_loc"] + cf_scale * u, -1, 1)
    cf.backward(t["dcf"])
    outs = [torch.zeros(N, Cc, H, H, device=DEV) for _ in range(4)]
    L.check(lib.cg_cf_combine_bwd(x.data_ptr(), t["rec_loc"].data_ptr(), rec_scale.data_ptr(), t["cf_loc"].data_ptr(),
                                  cf_scale.data_ptr(), t["dcf"].data_ptr(), outs[0].data_ptr(), outs[1].data_ptr(),
                                  outs[2].data_ptr(), outs[3].data_ptr(), n, stream()))
    torch.cuda.synchronize()
    assert float((cf.abs() >= 1).float().mean()) > 0.02, "the clamp mask must be exercised"
    for got, want, nm in zip(outs, (t["rec_loc"].grad, rec_scale.grad, t["cf_loc"].grad, cf_scale.grad),
                             ("d rec_loc", "d rec_scale", "d cf_loc", "d cf_scale")):
        assert_close(got, want, 1e-5, f"cf_combine_bwd {nm}")
    # likelihood.sample(h, return_loc=True) backward
    h_nchw = rnd(N, Cw, H, H, seed=8).to(torch.bfloat16).float().requires_grad_(True)
    hp = nhwc_bf16(h_nchw.detach())
    W = {k: (rnd(Cc, Cw, seed=20 + i, scale=0.6 / math.sqrt(Cw))).requires_grad_(True) for i, k in enumerate(("loc", "ls", "co"))}
    Bv = {k: rnd(Cc, seed=30 + i, scale=0.3).requires_grad_(True) for i, k in enumerate(("loc", "ls", "co"))}
    with torch.no_grad():
        Bv["ls"] -= 1.0
        Bv["loc"] += 0.6   # some means beyond +1 so the clamp mask matters

    def head(k):
        return torch.einsum("nkhw,ck->nchw", h_nchw, W[k]) + Bv[k][None, :, None, None]
    loc, ls = head("loc"), head("ls").clamp(min=-9.0)
    if Cc == 3:
        co = torch.tanh(head("co"))
        r = loc[:, 0].clamp(-1, 1)
        gg = (loc[:, 1] + co[:, 0] * r).clamp(-1, 1)
        b = (loc[:, 2] + co[:, 1] * r + co[:, 2] * gg).clamp(-1, 1)
        loc = torch.stack([r, gg, b], 1)
    xo, so = loc.clamp(-1, 1), ls.exp()
    dxo, dso = rnd(N, Cc, H, H, seed=41), rnd(N, Cc, H, H, seed=42)
    (xo * dxo + so * dso).sum().backward()
    a = L.DGaussArgs()
    a.h, a.h_ns, a.Cw = hp.data_ptr(), ns_of(hp), Cw
    a.w_loc, a.b_loc, a.w_ls, a.b_ls = W["loc"].data_ptr(), Bv["loc"].data_ptr(), W["ls"].data_ptr(), Bv["ls"].data_ptr()
    if Cc == 3:
        a.w_co, a.b_co = W["co"].data_ptr(), Bv["co"].data_ptr()
    a.N, a.HW, a.C = N, HW, Cc
    dh = torch.zeros_like(hp)
    gw = {k: torch.zeros(Cc, Cw, device=DEV) for k in W}
    gb = {k: torch.zeros(Cc, device=DEV) for k in W}
    a.dh, a.dh_ns = dh.data_ptr(), ns_of(dh)
    a.dw_loc, a.db_loc, a.dw_ls, a.db_ls = gw["loc"].data_ptr(), gb["loc"].data_ptr(), gw["ls"].data_ptr(), gb["ls"].data_ptr()
    if Cc == 3:
        a.dw_co, a.db_co = gw["co"].data_ptr(), gb["co"].data_ptr()
    L.check(lib.cg_dgauss_sample_bwd(C.byref(a), dxo.data_ptr(), dso.data_ptr(), stream()))
    torch.cuda.synchronize()
    assert_close(to_nchw(dh, Cw), h_nchw.grad, 1e-2, "sample_bwd dh (bf16 store)")
    for k in (("loc", "ls", "co") if Cc == 3 else ("loc", "ls")):
        assert_close(gw[k], W[k].grad, 2e-3, f"sample_bwd dW_{k}")
        assert_close(gb[k], Bv[k].grad, 2e-3, f"sample_bwd db_{k}")
