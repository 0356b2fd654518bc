"""Golden trajectory of the reference's EMA (src/utils.py:100-220) on a one-parameter model whose weight is set to a
known sequence: pins the decay warm-up schedule used by the optimiser tail (cg_optim_advance).  Build container only:
    python tests/golden/make_golden_ema.py"""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, "/root/reference/src")
for m in ("imageio", "send2trash", "matplotlib", "matplotlib.pyplot", "seaborn"):
    sys.modules.setdefault(m, types.ModuleType(m))
import utils as ref_utils  # noqa: E402

model = torch.nn.Linear(1, 1, bias=False)
ema = ref_utils.EMA(model, beta=0.999, update_after_step=100)
rng = np.random.default_rng(3)
steps = 1400
p_seq = np.cumsum(rng.standard_normal(steps)).astype(np.float32)
out = np.zeros(steps, dtype=np.float32)
for s in range(steps):
    with torch.no_grad():
        model.weight.fill_(float(p_seq[s]))
    ema.update()
    out[s] = float(ema.ema_model.weight)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ema_schedule.npz"), p=p_seq, ema=out,
                    beta=np.float32(0.999), update_after=np.int32(100))
print(out[98:106], p_seq[98:106])
