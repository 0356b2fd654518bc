"""Small drivers for `ncu --set full` captures (one GPU, few launches).
usage: ncu_targets.py dmol      -- cmnist + DmolNet head, batch 1024: one eager train step (dmol_fwd / dmol_bwd) + predict
       ncu_targets.py lik       -- ukbb192 batch 128 likelihood kernels only (dgauss fwd / bwd / sample)
Run under:  ncu --set full --clock-control none --import-source on -k regex:<pattern> --profile-from-start off -o <out> python tools/ncu_targets.py <kind>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "causal-gen_b200"))
import torch
from bench import make_trainer, synthetic_host_batches
kind = sys.argv[1] if len(sys.argv) > 1 else "dmol"
if kind == "dmol":
    margs, model, tr = make_trainer("cmnist", 1024, use_graph=False, x_like="diag_dmol")
else:
    margs, model, tr = make_trainer("ukbb192", 128, use_graph=False)
xs, pas = synthetic_host_batches(margs, tr.N, 1, 1)
for _ in range(2):
    tr.step(xs[0], pas[0])
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(xs[0], pas[0])
if kind == "dmol":
    model.eval()
    with torch.no_grad():
        model.sample(pas[0].cuda())
torch.cuda.synchronize()
torch.cuda.profiler.stop()
