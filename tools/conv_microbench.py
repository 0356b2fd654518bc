"""Micro-benchmark of single conv / wgrad launches at the UKBB-192 layer shapes (batch 32).
usage: conv_microbench.py [reps]   -- prints us/launch and algorithmic GB/s per case"""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "causal-gen_b200"))
import torch
from bench import conv_bytes, wgrad_bytes
from causalgen_b200 import _lib as L
from causalgen_b200.ops import ConvLayer, PackTable, SegSpec, View, new_act, planar_from_nchw, phys
DEV = "cuda"
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
only = sys.argv[2] if len(sys.argv) > 2 else None
N = int(os.environ.get("MB_N", "32"))
# name, H, cin list, cout, k, act, epilogue: none | add | mul | muladd
CASES = [
    ("fwd 64->16 r96", 96, [64], 16, 3, 1, "none"),
    ("fwd 16->64 r96 +res", 96, [16], 64, 3, 1, "add"),
    ("dgrad 16->64 r96 mul", 96, [16], 64, 3, 0, "mul"),
    ("dgrad 16->64 r96 mul+add", 96, [16], 64, 3, 0, "muladd"),
    ("dgrad 64->16 r96 mul", 96, [64], 16, 3, 0, "mul"),
    ("fwd 32->8 r192", 192, [32], 8, 3, 1, "none"),
    ("fwd 8->32 r192 +res", 192, [8], 32, 3, 1, "add"),
    ("fwd 96->24 r48", 48, [96], 24, 3, 1, "none"),
    ("fwd 24->96 r48 +res", 48, [24], 96, 3, 1, "add"),
    ("dgrad 24->96 r48 muladd", 48, [24], 96, 3, 0, "muladd"),
    ("fwd post 208->24 r48", 48, [96, 4, 96], 24, 3, 1, "none"),
    ("fwd 128->32 r24", 24, [128], 32, 3, 1, "none"),
    ("fwd 32->128 r24 +res", 24, [32], 128, 3, 1, "add"),
    ("fwd 32->160 r24 prior", 24, [32], 160, 3, 1, "none"),
    ("fwd 1x1 zproj 32->64 r96", 96, [16, 16], 64, 1, 0, "add"),
    ("fwd 1x1 zfp 80->64 r96", 96, [16, 64], 64, 1, 0, "none"),
    ("fwd 48->160 r12 +res", 12, [48], 160, 3, 1, "add"),
    ("fwd 48->192 r6 +res", 6, [48], 192, 3, 1, "add"),
    ("dgrad 1x1 192->192 r6 +add", 6, [192], 192, 1, 0, "add"),
    ("dgrad 1x1 192->192 r12 +add", 12, [192], 192, 1, 0, "add"),
]
def s(): return torch.cuda.current_stream().cuda_stream
def timeit(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
print("%-30s %9s %8s | %9s %8s" % ("case", "conv us", "GB/s", "wgrad us", "GB/s"))
for name, H, cins, cout, k, act, epi in CASES:
    if only and only not in name: continue
    g = torch.Generator().manual_seed(0)
    views = []
    for i, c in enumerate(cins):
        t = planar_from_nchw(torch.randn(N, phys(c), H, H, generator=g).to(DEV))
        views.append(View(t, phys(c), 0, c))
    w = (torch.randn(cout, sum(cins), k, k, generator=g) * 0.05).to(DEV); b = torch.zeros(cout, device=DEV)
    table = PackTable(DEV); layer = ConvLayer(table, w, b, cins, act, res=H); table.launch(s())
    out = new_act(N, H, H, cout, DEV)
    x1 = View(planar_from_nchw(torch.randn(N, phys(cout), H, H, generator=g).to(DEV)), phys(cout))
    x2 = View(planar_from_nchw(torch.randn(N, phys(cout), H, H, generator=g).to(DEV)), phys(cout))
    seg = SegSpec(out, 0)
    if epi == "add": seg.add = x1
    if epi in ("mul", "muladd"): seg.mul, seg.mul_act = x1, 1
    if epi == "muladd": seg.add = x2
    ln = layer.forward(views, [seg], N, H, H)
    t = timeit(lambda: ln(s()))
    by = conv_bytes(ln.keep[0])
    dy = x1; dw = torch.zeros_like(w); db = torch.zeros_like(b)
    lw = layer.wgrad(views, dy, dw, None, N, H, H)
    tw = 1.0 if os.environ.get("MB_NOWGRAD") else timeit(lambda: lw(s()))
    bw = wgrad_bytes(lw.keep[0])
    print("%-30s %9.1f %8.0f | %9.1f %8.0f" % (name, t, by / t / 1e3, tw, bw / tw / 1e3))
