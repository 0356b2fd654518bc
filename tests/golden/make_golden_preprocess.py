"""Golden vectors for the DSCM parent preprocessing (src/pgm/dscm.py:98-132) and trainer.preprocess_batch
(src/trainer.py:16-21), produced by EXECUTING the reference's own function definitions.

src/pgm/dscm.py cannot be imported here (it pulls in Pyro through layers.py), so the three functions are cut out of the
reference files with `ast` at run time and exec'd unchanged with torch in scope -- nothing is copied into this repo.
`.cuda()` in vae_preprocess is neutralised (CPU container) by a Tensor.cuda no-op patch.

    python tests/golden/make_golden_preprocess.py      # needs /root/reference; writes tests/golden/preprocess.npz
"""
import ast
import os
from types import SimpleNamespace

import numpy as np
import torch

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))


def cut(path, names):
    src = open(path).read()
    mod = ast.parse(src)
    return "\n\n".join(ast.get_source_segment(src, n) for n in mod.body
                       if isinstance(n, ast.FunctionDef) and n.name in names)


def main():
    ns = {"torch": torch, "Tensor": torch.Tensor, "Dict": dict, "Hparams": object}
    exec(cut(os.path.join(REF, "datasets.py"), {"get_attr_max_min"}), ns)
    exec(cut(os.path.join(REF, "pgm/dscm.py"), {"ukbb_preprocess", "vae_preprocess"}), ns)
    exec(cut(os.path.join(REF, "trainer.py"), {"preprocess_batch"}), ns)
    torch.Tensor.cuda = lambda self, *a, **k: self
    g = torch.Generator().manual_seed(7)
    B = 6
    pa = {"mri_seq": torch.randint(0, 2, (B, 1), generator=g).float(),
          "brain_volume": torch.rand(B, 1, generator=g) * 2 - 1,
          "ventricle_volume": torch.rand(B, 1, generator=g) * 2 - 1,
          "sex": torch.randint(0, 2, (B, 1), generator=g).float(),
          "age": torch.rand(B, 1, generator=g) * 2 - 1}
    out = {f"in_{k}": v.numpy() for k, v in pa.items()}
    a_ukbb = SimpleNamespace(dataset="ukbb", input_res=8, parents_x=["mri_seq", "brain_volume", "ventricle_volume", "sex"])
    out["ukbb_vae_pa"] = ns["vae_preprocess"](a_ukbb, {k: v.clone() for k, v in pa.items() if k != "age"}).numpy()
    a_age = SimpleNamespace(dataset="ukbb", input_res=4, parents_x=["age", "sex"])
    out["ukbb_age_vae_pa"] = ns["vae_preprocess"](a_age, {k: pa[k].clone() for k in ("age", "sex")}).numpy()
    a_none = SimpleNamespace(dataset="morphomnist", input_res=4, parents_x=["brain_volume", "digit"])
    digit = torch.nn.functional.one_hot(torch.arange(B) % 10, 10).float()
    out["in_digit"] = digit.numpy()
    out["plain_vae_pa"] = ns["vae_preprocess"](a_none, {"brain_volume": pa["brain_volume"][:, 0].clone(), "digit": digit}).numpy()
    x8 = torch.randint(0, 256, (B, 1, 8, 8), generator=g, dtype=torch.uint8)
    batch = ns["preprocess_batch"](SimpleNamespace(device="cpu", input_res=8), {"x": x8.clone(), "pa": pa["age"].repeat(1, 3)},
                                   expand_pa=True)
    out["x8"], out["x_norm"], out["pa_expanded"] = x8.numpy(), batch["x"].numpy(), batch["pa"].numpy()
    np.savez_compressed(os.path.join(HERE, "preprocess.npz"), **out)
    print("wrote preprocess.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
