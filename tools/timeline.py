"""In-kernel timeline of one conv / wgrad launch (CTA 0) from the -DCG_TIMELINE debug library.
usage: CAUSALGEN_B200_LIB=causal-gen_b200/causalgen_b200/libcausalgen_b200_tl.so python tools/timeline.py [case-substring]
Marks (ns relative to kernel entry of CTA (0,0)):
  conv : 32 entry | 33 prologue done | 34 weight slab landed | 35+2i tile i first A chunk ready | 36+2i tile i MMAs
         issued | 50+i epilogue of tile i done (warp 0) | 60 all warps done | 70+i accumulator free for tile i (MMA warp)
         | 80+i accumulator of tile i complete (epilogue warp 0) | 90+4i producer starts tile i, 91+4i+c: its chunk c issued.  CG_TL_FIRST=k moves the tile window to tiles k.. of CTA 0
  wgrad: 1 entry | 0 prologue done | 2+2i tile i ready | 3+2i tile i MMAs issued | 20 accumulators complete
         | 21 flush done | 22 all warps done"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "causal-gen_b200"))
import torch
from causalgen_b200 import _lib as L
from causalgen_b200.ops import ConvLayer, PackTable, SegSpec, View, new_act, planar_from_nchw, phys
sys.argv, argv = sys.argv[:1], sys.argv[1:]
import importlib.util
spec = importlib.util.spec_from_file_location("mb", os.path.join(ROOT, "tools", "conv_microbench.py"))
src = open(os.path.join(ROOT, "tools", "conv_microbench.py")).read().split("def s():")[0]
ns = {"__file__": os.path.join(ROOT, "tools", "conv_microbench.py")}
exec(compile(src, "mb_cases", "exec"), ns)
CASES = ns["CASES"]
only = argv[0] if argv else None
N = int(argv[1]) if len(argv) > 1 else 32
lib = L.load()
lib.cg_debug_timeline.argtypes = [C.c_void_p]; lib.cg_debug_timeline.restype = None
tl = torch.zeros(128, dtype=torch.int64, device="cuda")
def s(): return torch.cuda.current_stream().cuda_stream
def show(tag, base, keys):
    v = tl.cpu().tolist()
    t0 = v[base]
    print("  %-6s" % tag, " ".join("%d:%.1f" % (k, (v[k] - t0) / 1e3) for k in keys if v[k]))
for name, H, cins, cout, k, act, epi in CASES:
    if only and only not in name: continue
    g = torch.Generator().manual_seed(0)
    views = []
    for c in cins:
        views.append(View(planar_from_nchw(torch.randn(N, phys(c), H, H, generator=g).cuda()), phys(c), 0, c))
    w = (torch.randn(cout, sum(cins), k, k, generator=g) * 0.05).cuda(); b = torch.zeros(cout, device="cuda")
    table = PackTable("cuda"); layer = ConvLayer(table, w, b, cins, act); table.launch(s())
    out = new_act(N, H, H, cout, "cuda")
    x1 = View(planar_from_nchw(torch.randn(N, phys(cout), H, H, generator=g).cuda()), phys(cout))
    x2 = View(planar_from_nchw(torch.randn(N, phys(cout), H, H, generator=g).cuda()), phys(cout))
    seg = SegSpec(out, 0)
    if epi == "add": seg.add = x1
    if epi in ("mul", "muladd"): seg.mul, seg.mul_act = x1, 1
    if epi == "muladd": seg.add = x2
    ln = layer.forward(views, [seg], N, H, H)
    dw = torch.zeros_like(w); db = torch.zeros_like(b)
    lw = layer.wgrad(views, x1, dw, db, N, H, H)
    print(name)
    for tag, fn, base, keys in (("conv", ln, 32, [33, 34] + list(range(35, 47)) + list(range(50, 58)) + [60] + list(range(70, 76)) + list(range(80, 88)) + list(range(90, 114))),
                                ("wgrad", lw, 1, [0] + list(range(2, 18)) + [20, 21, 22])):
        lib.cg_debug_timeline(None)
        for _ in range(3): fn(s())
        torch.cuda.synchronize(); tl.zero_(); lib.cg_debug_timeline(tl.data_ptr())
        fn(s()); torch.cuda.synchronize()
        show(tag, base, keys)
lib.cg_debug_timeline(None)
