"""Per-launch timing table of one training step (eager launches, CUDA events around each launch)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "causal-gen_b200"))
import torch
from bench import conv_bytes, wgrad_bytes, synthetic_host_batches
from causalgen_b200 import HVAE
from causalgen_b200.presets import init_like_reference_main, make_args
from causalgen_b200.trainer import Trainer
cfgname = sys.argv[1] if len(sys.argv) > 1 else "ukbb192"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
margs = make_args(cfgname)
torch.manual_seed(7)
model = init_like_reference_main(HVAE(margs)).cuda()
tr = Trainer(model, B, beta=margs.beta, use_graph=False)
xs, pas = synthetic_host_batches(margs, B, 1, 1)
for _ in range(2): tr.step(xs[0], pas[0])
prog = tr.prog
s = torch.cuda.current_stream().cuda_stream
rows = []
for rep in range(2):
    evs = []
    for t in prog.zero: t.zero_()
    tr.eng.flat_grad.zero_(); tr.eng.pack_weights(s); torch.cuda.synchronize()
    for ln in prog.launches:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ln(s); e1.record(); evs.append((ln, e0, e1))
    torch.cuda.synchronize()
agg = {}
for ln, e0, e1 in evs:
    ms = e0.elapsed_time(e1); name = getattr(ln, "name", "pyop")
    if name == "cg_conv2d":
        a = ln.keep[0]; cin = sum(a.src[i].C for i in range(a.nsrc))
        key = (name, a.H, a.ksize, cin, a.cout, a.nsrc, int(bool(a.seg[0].mul))); by = conv_bytes(a)
    elif name == "cg_conv2d_wgrad":
        a = ln.keep[0]; cin = sum(a.src[i].C for i in range(a.nsrc))
        key = (name, a.H, a.ksize if a.taps > 1 else 1, cin, a.dy_c, a.nsrc, 0); by = wgrad_bytes(a)
    else:
        key = (name, 0, 0, 0, 0, 0, 0); by = 0
    t = agg.setdefault(key, [0, 0.0, 0.0]); t[0] += 1; t[1] += ms; t[2] += by
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.2f} ms over {len(evs)} launches")
print("name                H  k  cin cout nsrc bwd | count   ms     us/launch  GB/s   share")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
    print("%-18s %3d %2d %4d %4d %2d %2d | %4d %8.3f %9.1f %7.0f %6.3f" % (k + (v[0], v[1], 1e3 * v[1] / v[0], v[2] / (v[1] / 1e3) / 1e9 if v[1] > 0 else 0, v[1] / tot)))
