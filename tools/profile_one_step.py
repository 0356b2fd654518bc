"""Runs warm-up steps, then exactly one eager training step inside a cudaProfiler range (for ncu
--profile-from-start off)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "causal-gen_b200"))
import torch
from bench import synthetic_host_batches
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
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(xs[0], pas[0])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
