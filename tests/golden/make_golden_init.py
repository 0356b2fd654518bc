"""SHA-256 of the freshly initialised reference HVAE (src/vae.py, torch.manual_seed(7)) per config: pins that the
drop-in's parameter containers draw the same numbers in the same order under the same seed (construction order,
init scaling src/vae.py:121-122,303-308).  Build container only:  python tests/golden/make_golden_init.py"""
import hashlib
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (puts /root/reference/src on the path)
import vae as ref_vae  # noqa: E402


def digest(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


out = {}
for name, (cfg_name, hps_name, extra, _) in MG.CASES.items():
    torch.manual_seed(7)
    model = ref_vae.HVAE(MG.ref_args(hps_name, extra))
    out[name] = {"sha256": digest(model.state_dict()), "tensors": len(model.state_dict()),
                 "params": sum(p.numel() for p in model.parameters())}
    print(name, out[name])
json.dump(out, open(os.path.join(HERE, "init_digests.json"), "w"), indent=1)
