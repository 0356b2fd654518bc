"""Detailed parity report (ours vs fp32 oracle vs bf16-storage-emulating oracle) for debugging on the GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "causal-gen_b200"))
import numpy as np, torch
import hvae_oracle as O
from causalgen_b200 import HVAE, counterfactual
DEV = "cuda"
def rel(a, b): return float((a - b).norm() / (b.norm() + 1e-12))
def oracle_run(cfg, sd, x, pa_full, emu):
    O.EMULATE_BF16 = emu
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    tape = O.NoiseTape(seed=101)
    out = O.hvae_forward(sdr, cfg, x, pa_full, tape, beta=cfg.beta, detail=True)
    out["elbo"].backward()
    O.EMULATE_BF16 = False
    return out, sdr, tape
for name in sys.argv[1:] or ["tiny_ukbb", "tiny_morphomnist", "tiny_cmnist", "morphomnist", "cmnist", "ukbb192", "mimic192"]:
    B = {"ukbb192": 1, "mimic192": 1}.get(name, 2)
    cfg = O.make_cfg(name); sd = O.seeded_state_dict(cfg, 7)
    model = HVAE(cfg); model.load_state_dict(sd); model.to(DEV).eval()
    x8, pa, cf = O.synthetic_batch(cfg, B, 11); x = O.normalise_x(x8)
    pa_full, cf_full = O.expand_parents(pa, cfg.input_res), O.expand_parents(cf, cfg.input_res)
    r32, sd32, tape = oracle_run(cfg, sd, x, pa_full, False)
    r16, sd16, _ = oracle_run(cfg, sd, x, pa_full, True)
    eps = [e.to(DEV) for e in tape.drawn]
    model.zero_grad(); out = model(x.to(DEV), pa.to(DEV), beta=cfg.beta, eps=eps); out["elbo"].backward(); torch.cuda.synchronize()
    print(f"\n##### {name} B={B}")
    for k in ("elbo", "nll", "kl"):
        print(f"  {k}: ours {out[k].item():.6f} fp32 {r32[k].item():.6f} emu {r16[k].item():.6f}")
    bk = model.block_kl().cpu()
    print("  blockKL rel: vs fp32 %.4g  vs emu %.4g | emu vs fp32 %.4g" % (rel(bk, r32["block_kl"].detach()), rel(bk, r16["block_kl"].detach()), rel(r16["block_kl"].detach(), r32["block_kl"].detach())))
    named = dict(model.named_parameters())
    rows = []
    n32 = d32 = n16 = 0.0
    for k in sd:
        g32, g16 = sd32[k].grad, sd16[k].grad
        if g32 is None: continue
        g = named[k].grad.cpu()
        rows.append((rel(g, g16), rel(g, g32), rel(g16, g32), float(g32.norm()), k))
        n32 += float((g - g32).pow(2).sum()); n16 += float((g - g16).pow(2).sum()); d32 += float(g32.pow(2).sum())
    print("  global grad rel-L2: vs fp32 %.4g vs emu %.4g" % ((n32 / d32) ** .5, (n16 / d32) ** .5))
    gmax = max(r[3] for r in rows)
    rows = [r for r in rows if r[3] > 1e-3 * gmax]
    rows.sort(reverse=True)
    print("  worst params (rel vs emu, rel vs fp32, emu-vs-fp32, |g|, name):")
    for r in rows[:12]: print("    %.4f %.4f %.4f %.3e %s" % r)
    # inference
    with torch.no_grad():
        res = {}
        for emu in (False, True):
            O.EMULATE_BF16 = emu
            t2 = O.NoiseTape(seed=202)
            zs = O.hvae_abduct(sd, cfg, x, pa_full, t2, t=0.9); zs = [z["z"] for z in zs] if cfg.cond_prior else zs
            rl, rs = O.hvae_forward_latents(sd, cfg, zs, pa_full)
            cl, cs = O.hvae_forward_latents(sd, cfg, zs, cf_full)
            cfx = torch.clamp(cl + cs * (x - rl) / rs.clamp(min=1e-12), -1, 1)
            res[emu] = (zs, rl, rs, cfx, t2)
        O.EMULATE_BF16 = False
    e2 = [e.to(DEV) for e in res[False][4].drawn]
    zs = model.abduct(x.to(DEV), pa.to(DEV), t=0.9, eps=e2); zs = [z["z"] for z in zs] if cfg.cond_prior else zs
    print("  z rel (max over blocks): vs fp32 %.4g vs emu %.4g | emu vs fp32 %.4g" % (
        max(rel(a.cpu(), b) for a, b in zip(zs, res[False][0])), max(rel(a.cpu(), b) for a, b in zip(zs, res[True][0])),
        max(rel(a, b) for a, b in zip(res[True][0], res[False][0]))))
    for tag, zin in (("own z", zs), ("fp32-oracle z", [z.to(DEV) for z in res[False][0]])):
        loc, sc = model.forward_latents(zin, pa.to(DEV))
        for emu in (False, True):
            d = (loc.cpu() - res[emu][1]).abs()
            print(f"  rec loc [{tag}] vs {'emu' if emu else 'fp32'}: max {d.max():.4f} mean {d.mean():.5f} frac<=2/255 {(d <= 2/255).float().mean():.4f}  scale rel {rel(sc.cpu(), res[emu][2]):.4f}")
    d = (res[True][1] - res[False][1]).abs(); print(f"  rec loc emu vs fp32: max {d.max():.4f} mean {d.mean():.5f} frac<=2/255 {(d <= 2/255).float().mean():.4f}")
    cfx, _ = counterfactual(model, x.to(DEV), pa.to(DEV), cf.to(DEV), t_abduct=0.9, eps=[e2])
    for emu in (False, True):
        d = (cfx.cpu() - res[emu][3]).abs(); print(f"  cf_x vs {'emu' if emu else 'fp32'}: max {d.max():.4f} mean {d.mean():.5f} frac<=2/255 {(d <= 2/255).float().mean():.4f}")
    d = (res[True][3] - res[False][3]).abs(); print(f"  cf_x emu vs fp32: max {d.max():.4f} mean {d.mean():.5f} frac<=2/255 {(d <= 2/255).float().mean():.4f}")
    del model; torch.cuda.empty_cache()
