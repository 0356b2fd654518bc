"""Per-tensor gradient deviation of the CUDA path vs the fp32 oracle for one config, under the fold modes
(CAUSALGEN_B200_FOLD=0/1/2 in child processes): tells a wrong layer from bf16 rounding noise.
usage: python tools/grad_diag.py [config] [mode ...]   e.g.  ukbb192 0 1 2 "res=24;dir=f" 1@CG_NO_PDL=1
a mode is a CAUSALGEN_B200_FOLD value, optionally followed by @ENV=VALUE[,ENV=VALUE] for the child process"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "causal-gen_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch  # noqa: E402


def child(name, out, ref_path):
    import hvae_oracle as O
    from causalgen_b200 import HVAE
    cfg = O.make_cfg(name)
    sd = O.seeded_state_dict(cfg, seed=7)
    model = HVAE(cfg)
    model.load_state_dict(sd, strict=True)
    model.to("cuda").eval()
    x8, pa, _ = O.synthetic_batch(cfg, 1 if "192" in name else 2, seed=11)
    x = O.normalise_x(x8)
    pa_full = O.expand_parents(pa, cfg.input_res)
    res = {}
    if not os.path.exists(ref_path):
        sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        tape = O.NoiseTape(seed=101)
        o = O.hvae_forward(sdr, cfg, x, pa_full, tape, beta=cfg.beta, detail=True)
        o["elbo"].backward()
        torch.save({"g": {k: v.grad for k, v in sdr.items() if v.grad is not None}, "eps": tape.drawn,
                    "elbo": float(o["elbo"])}, ref_path)
    ref = torch.load(ref_path)
    eps = [e.to("cuda") for e in ref["eps"]]
    model.zero_grad()
    o = model(x.cuda(), pa_full.cuda(), beta=cfg.beta, eps=eps)
    o["elbo"].backward()
    torch.cuda.synchronize()
    res["g"] = {k: p.grad.cpu() for k, p in model.named_parameters() if p.grad is not None}
    res["elbo"] = float(o["elbo"])
    eng = model.engine()
    res["folded"] = []
    torch.save(res, out)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "ukbb192"
    modes = sys.argv[2:] if len(sys.argv) > 2 else ["0", "1", "2"]
    base = f"/tmp/grad_diag_{name}"
    for i, m in enumerate(modes):
        fold, _, extra = m.partition("@")
        env = dict(os.environ, CAUSALGEN_B200_FOLD=fold)
        env.update(kv.split("=") for kv in extra.split(",") if kv)
        subprocess.check_call([sys.executable, __file__, "--child", name, f"{base}_{i}.pt", f"{base}.ref"], env=env)
    ref = torch.load(f"{base}.ref")
    got = {m: torch.load(f"{base}_{i}.pt") for i, m in enumerate(modes)}
    gmax = max(float(g.norm()) for g in ref["g"].values())
    print(f"{name}: oracle elbo {ref['elbo']:.6f}; ours " + " ".join(f"FOLD={m}: {got[m]['elbo']:.6f}" for m in modes))
    rows = []
    for k, g in ref["g"].items():
        if float(g.norm()) < 1e-2 * gmax:
            continue
        rel = [float((got[m]["g"][k] - g).norm() / g.norm()) for m in modes]
        rows.append((max(rel), k, tuple(g.shape), float(g.norm()) / gmax, rel))
    for m in modes:
        num = sum(float((got[m]["g"][k] - g).pow(2).sum()) for k, g in ref["g"].items())
        den = sum(float(g.pow(2).sum()) for g in ref["g"].values())
        print(f"[{modes.index(m)}] FOLD={m}: global grad rel-L2 {(num / den) ** 0.5:.5f}  elbo {got[m]['elbo']:.6f}")
    rows.sort(reverse=True)
    print("%-52s %-18s %8s  " % ("tensor", "shape", "|g|/max") + " ".join(f"[{i:>6}]" for i in range(len(modes))))
    for _, k, shp, gn, rel in rows[:int(os.environ.get("DIAG_ROWS", "12"))]:
        print("%-52s %-18s %8.3f  " % (k, str(shp), gn) + " ".join("%8.4f" % r for r in rel))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        main()
