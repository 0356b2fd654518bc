"""Where the counterfactual pass spends its time: per-phase and per-kernel-family CUDA-event timings."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "causal-gen_b200"))
import torch
from bench import synthetic_host_batches
from causalgen_b200 import HVAE, counterfactual
from causalgen_b200.presets import init_like_reference_main, make_args
cfgname = sys.argv[1] if len(sys.argv) > 1 else "ukbb192"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
margs = make_args(cfgname)
torch.manual_seed(7)
model = init_like_reference_main(HVAE(margs)).cuda().eval()
xs, pas = synthetic_host_batches(margs, B, 2, 1)
x = (xs[0].cuda().float() - 127.5) / 127.5
pa, cf = pas[0].cuda(), pas[1].cuda()
for _ in range(2): counterfactual(model, x, pa, cf)
def ev(): return torch.cuda.Event(enable_timing=True)
def timed(fn, n=3):
    torch.cuda.synchronize(); a, b = ev(), ev(); a.record()
    for _ in range(n): out = fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n, out
t_all, _ = timed(lambda: counterfactual(model, x, pa, cf))
t_abd, zs = timed(lambda: model.abduct(x, pa, t=1.0))
zs = [z["z"] if isinstance(z, dict) else z for z in zs]
t_dec, _ = timed(lambda: model._decode(zs, [cf, pa], None, None))
print(f"counterfactual {t_all:.2f} ms = abduct {t_abd:.2f} + decode(2 parent sets) {t_dec:.2f} + combine  (batch {B})")
eng = model.engine()
s = torch.cuda.current_stream().cuda_stream
for key, prog in eng.programs.items():
    evs = []
    torch.cuda.synchronize()
    for ln in prog.launches:
        a, b = ev(), ev(); a.record(); ln(s); b.record(); evs.append((getattr(ln, "name", "pyop"), a, b))
    torch.cuda.synchronize()
    fam = {}
    for name, a, b in evs: fam[name] = fam.get(name, 0.0) + a.elapsed_time(b)
    print(key[0], "launches", len(evs), "sum %.2f ms" % sum(fam.values()), {k: round(v, 2) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])})
