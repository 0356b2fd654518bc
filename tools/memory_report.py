"""Where the activation memory of one training program goes (built on the CPU with CAUSALGEN_B200_TRACE_ONLY=1 at a small
batch and scaled linearly): buffers allocated while the forward launches are recorded (kept for the backward pass) against
buffers allocated while the backward launches are recorded (gradients: each lives for a few launches only).
usage: python tools/memory_report.py [config] [batch_to_report]"""
import os
import sys

os.environ["CAUSALGEN_B200_TRACE_ONLY"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "causal-gen_b200"), os.path.join(ROOT, "oracle")]
import torch  # noqa: E402
import hvae_oracle as O  # noqa: E402
import causalgen_b200._lib as L  # noqa: E402
from causalgen_b200 import HVAE, engine as E, ops  # noqa: E402


class Fake:  # the planners are host arithmetic; everything else is a no-op without a GPU
    def __init__(self, real):
        self.real = real

    def __getattr__(self, n):
        if n in ("cg_conv_nchunk", "cg_conv_nchunk_ex", "cg_packed_weight_bytes", "cg_packed_weight_bytes_nc", "cg_version",
                 "cg_last_error", "cg_conv_fold_ok", "cg_conv2d_wgrad_launches"):
            return getattr(self.real, n)
        return lambda *a: 0


L._lib = Fake(L.load())
name = sys.argv[1] if len(sys.argv) > 1 else "ukbb192"
report_b = int(sys.argv[2]) if len(sys.argv) > 2 else 128
B = 1
log = []
real_new_act = ops.new_act


def counting_new_act(*a, **k):
    v = real_new_act(*a, **k)
    log.append((v.t.numel() * v.t.element_size(), tuple(v.t.shape)))
    return v


ops.new_act = E.new_act = counting_new_act
cfg = O.make_cfg(name)
m = HVAE(cfg)
x8, pa, _ = O.synthetic_batch(cfg, B, 1)
eng = m.engine()
# forward buffers are allocated while the forward launches are recorded: remember how many existed when the first backward
# helper ran
split = {}
for fn_name in ("_decoder_bwd", "_encoder_bwd"):
    real = getattr(E.Engine, fn_name)

    def wrapped(self, *a, _real=real, **k):
        split.setdefault("n", len(log))
        return _real(self, *a, **k)
    setattr(E.Engine, fn_name, wrapped)
prog = eng.build_elbo(B, True, False)
total = sum(b for b, _ in log)
nf = split.get("n", len(log))
fwd, bwd = sum(b for b, _ in log[:nf]), sum(b for b, _ in log[nf:])
print(f"{name}: {len(log)} planar buffers, {total / 2**20:.1f} MiB per image -> {total * report_b / 2**30:.1f} GiB at batch {report_b}")
print(f"  allocated while recording the forward pass (activations, masks, statistics: live until their backward launch): "
      f"{nf} buffers, {fwd / 2**20:.1f} MiB per image = {fwd * report_b / 2**30:.1f} GiB")
print(f"  allocated while recording the backward pass (gradients: each is dead a few launches later): {len(log) - nf} buffers, "
      f"{bwd / 2**20:.1f} MiB per image = {bwd * report_b / 2**30:.1f} GiB; largest single one {max(b for b, _ in log[nf:]) / 2**20:.1f} MiB per image")
by_res = {}
for b, shp in log:
    r = shp[2] if len(shp) == 5 else shp[1]
    by_res[r] = by_res.get(r, 0) + b
for r in sorted(by_res, reverse=True):
    print(f"  {r:4d}^2: {by_res[r] / 2**20:8.1f} MiB per image ({100 * by_res[r] / total:4.1f} %)")
