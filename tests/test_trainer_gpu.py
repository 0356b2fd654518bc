"""Parity of the path bench.py times: `Trainer.step` (uint8 normalise -> Philox-noise ELBO program -> backward ->
bucket reduce -> clip / AdamW / EMA), eager and CUDA-graph replay, against `oracle.train_step_cpu` fed the SAME
noise (every latent kernel exports the eps it drew: `Trainer(export_eps=True)`).

Reference lines: src/trainer.py:50-87 (step body), src/train_setup.py:42-53 (AdamW + LambdaLR warm-up),
src/utils.py:169-220 (EMA).  Tolerances (fixed numbers; bf16 activations, fp32 accumulation):
    elbo / nll / kl per step      rel <= 5e-3
    gradient norm per step        rel <= 2e-2
    Adam first moment  m          rel-L2 <= 3e-2      (linear in the gradients of all steps)
    Adam second moment v          rel-L2 <= 6e-2      (quadratic)
    parameter / EMA update        rel-L2 of (after - before) <= 0.2: AdamW normalises every coordinate to ~lr*sign(g)
                                  in the first steps, so coordinates whose gradient is at the bf16 noise floor flip sign;
                                  cosine similarity of the update >= 0.97 is asserted next to it
Every measured deviation is recorded in the parity report (tests/conftest.py).
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import hvae_oracle as O
from conftest import parity_report

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HP = dict(lr=2e-4, wd=0.05, betas=(0.9, 0.9), lr_warmup_steps=2, grad_clip=350.0, grad_skip=5000.0, ema_rate=0.999,
          ema_update_after=0)


def rel_l2(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def cosine(a, b):
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-30))


def make(name, B, use_graph, **over):
    from causalgen_b200 import HVAE
    from causalgen_b200.trainer import Trainer
    cfg = O.make_cfg(name, **over)
    sd = O.seeded_state_dict(cfg, seed=7)
    model = HVAE(cfg)
    model.load_state_dict(sd, strict=True)
    model.to(DEV)
    tr = Trainer(model, B, beta=cfg.beta, use_graph=use_graph, noise_seed=11, export_eps=True, **HP)
    return cfg, sd, model, tr


def oracle_run(cfg, sd, batches, eps_per_step, drops=None, betas_kl=None):
    """the reference step on the CPU, same weights / batches / noise"""
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ema = {k: v.detach().clone() for k, v in sdr.items()}
    state, hist = {}, []
    for i, ((x8, pa), eps) in enumerate(zip(batches, eps_per_step)):
        x = O.normalise_x(x8)
        lr = O.warmup_lr(HP["lr"], i + 1, HP["lr_warmup_steps"])
        out, gn = O.train_step_cpu(sdr, cfg, x, O.expand_parents(pa, cfg.input_res), O.NoiseTape(tensors=eps), state, lr=lr,
                                   wd=HP["wd"], betas=HP["betas"], grad_clip=HP["grad_clip"], grad_skip=HP["grad_skip"],
                                   step=i + 1, ema=ema, ema_update_after=HP["ema_update_after"],
                                   beta=None if betas_kl is None else betas_kl[i],
                                   drop=(1.0, 1.0) if drops is None else drops[i])
        hist.append((float(out["elbo"]), float(out["nll"]), float(out["kl"]), float(gn)))
    return sdr, ema, state, hist


def check_against_oracle(tag, cfg, sd, model, tr, batches, got_hist, eps_per_step, **okw):
    sdr, ema_ref, state, hist = oracle_run(cfg, sd, batches, eps_per_step, **okw)
    for i, (g, w) in enumerate(zip(got_hist, hist)):
        for j, k in enumerate(("elbo", "nll", "kl")):
            d = abs(g[j] - w[j]) / abs(w[j])
            parity_report(tag, f"step{i} {k} rel", d, 5e-3, f"ours {g[j]:.6f} oracle {w[j]:.6f}")
            assert d <= 5e-3, (tag, i, k, g[j], w[j])
        d = abs(g[3] - w[3]) / w[3]
        parity_report(tag, f"step{i} grad-norm rel", d, 2e-2, f"ours {g[3]:.5f} oracle {w[3]:.5f}")
        assert d <= 2e-2, (tag, i, g[3], w[3])
    names = [k for k, _ in model.named_parameters()]
    views_m, views_v = tr._views(tr.m), tr._views(tr.v)
    m_ref = torch.cat([state[k][0].flatten() for k in names if k in state])
    v_ref = torch.cat([state[k][1].flatten() for k in names if k in state])
    m_got = torch.cat([views_m[k].flatten() for k in names if k in state]).cpu()
    v_got = torch.cat([views_v[k].flatten() for k in names if k in state]).cpu()
    parity_report(tag, "Adam m rel-L2", rel_l2(m_got, m_ref), 3e-2)
    parity_report(tag, "Adam v rel-L2", rel_l2(v_got, v_ref), 6e-2)
    assert rel_l2(m_got, m_ref) <= 3e-2 and rel_l2(v_got, v_ref) <= 6e-2
    p_now = dict(model.named_parameters())
    ema_now = tr.ema_state_dict()
    dp_got = torch.cat([(p_now[k].detach().cpu() - sd[k]).flatten() for k in names])
    dp_ref = torch.cat([(sdr[k].detach() - sd[k]).flatten() for k in names])
    de_got = torch.cat([(ema_now[k].cpu() - sd[k]).flatten() for k in names])
    de_ref = torch.cat([(ema_ref[k] - sd[k]).flatten() for k in names])
    for what, a, b in (("param update", dp_got, dp_ref), ("EMA update", de_got, de_ref)):
        parity_report(tag, f"{what} rel-L2", rel_l2(a, b), 0.2)
        parity_report(tag, f"{what} cosine", 1 - cosine(a, b), 0.03, "1 - cos")
        assert rel_l2(a, b) <= 0.2 and cosine(a, b) >= 0.97, (tag, what, rel_l2(a, b), cosine(a, b))
    assert float(de_ref.norm()) > 0 and not torch.equal(de_ref, dp_ref), "EMA decay must be active in this test"


def run_steps(tr, batches):
    hist, eps = [], []
    for x8, pa in batches:
        out = tr.step(x8, pa)
        eps.append([e.cpu().clone() for e in tr.eps_out])
        hist.append((float(out[0]), float(out[1]), float(out[2]), tr.grad_norm()))
    return hist, eps


@pytest.mark.parametrize("name,B,graph", [("tiny_ukbb", 3, False), ("tiny_ukbb", 3, True), ("tiny_cmnist", 2, True),
                                          ("ukbb192", 2, True)])
def test_trainer_steps_match_oracle_train_step(name, B, graph):
    cfg, sd, model, tr = make(name, B, graph)
    nsteps = 4 if graph else 3  # graph: step 1 eager, capture before step 2, replays from then on
    batches = [O.synthetic_batch(cfg, B, seed=20 + i)[:2] for i in range(nsteps)]
    hist, eps = run_steps(tr, batches)
    assert (tr.g_fb is not None) == graph
    assert tr.skipped_updates() == 0 and int(tr.state[0]) == nsteps and int(tr.state[1]) == nsteps
    check_against_oracle(f"trainer[{name},B{B},{'graph' if graph else 'eager'}]", cfg, sd, model, tr, batches, hist, eps)


def test_trainer_graph_replay_equals_eager_and_beta_annealing():
    """same seeds, same batches: graph replay == eager launches (up to atomic-order rounding); beta warm-up
    (src/trainer.py:52-57) changes under replay because beta is a device scalar"""
    outs = {}
    for graph in (False, True):
        from causalgen_b200 import HVAE
        from causalgen_b200.trainer import Trainer
        cfg = O.make_cfg("tiny_ukbb")
        model = HVAE(cfg)
        model.load_state_dict(O.seeded_state_dict(cfg, seed=7))
        model.to(DEV)
        tr = Trainer(model, 3, beta=cfg.beta, use_graph=graph, noise_seed=5, beta_warmup_steps=4, **HP)
        hist = []
        for i in range(6):
            x8, pa, _ = O.synthetic_batch(cfg, 3, seed=40 + i)
            out = tr.step(x8, pa)
            hist.append([float(v) for v in out] + [tr.beta])
        outs[graph] = (np.array(hist), tr.flat_p.detach().cpu().clone())
    a, b = outs[False], outs[True]
    d_loss = float(np.abs(a[0][:, :3] - b[0][:, :3]).max() / np.abs(a[0][:, :3]).max())
    d_par = rel_l2(b[1], a[1])
    parity_report("graph==eager", "loss max rel", d_loss, 1e-4)
    parity_report("graph==eager", "params rel-L2 after 6 steps", d_par, 1e-4)
    assert d_loss <= 1e-4 and d_par <= 1e-4
    betas = a[0][:, 3]
    np.testing.assert_allclose(betas, [5.0 * min(1.0, (i + 1) / 4) for i in range(6)], rtol=1e-6)
    # elbo = nll + beta * kl with the annealed beta of each step, also under replay
    np.testing.assert_allclose(b[0][:, 0], b[0][:, 1] + betas * b[0][:, 2], rtol=2e-5)


def test_trainer_morphomnist_conditioning_dropout_replays_from_graph():
    """src/vae.py:234-249: the per-step dropout draw reaches the captured graph through a device scalar"""
    cfg, sd, model, tr = make("tiny_morphomnist", 3, True)
    draws = [(0, 1), (1, 0), (1, 1), (0, 1), (1, 1)]
    it = iter(draws)
    model.drop_cond = lambda: next(it)
    batches = [O.synthetic_batch(cfg, 3, seed=60 + i)[:2] for i in range(len(draws))]
    hist, eps = run_steps(tr, batches)
    assert tr.g_fb is not None, "morphomnist must replay from a graph too"
    check_against_oracle("trainer[tiny_morphomnist,graph,drop]", cfg, sd, model, tr, batches, hist, eps,
                         drops=[(float(a), float(b)) for a, b in draws])


def test_trainer_frozen_likelihood_scale_and_accumulation():
    """x_like=shared_dgauss with std_init>0 freezes x_logscale.weight (src/vae.py:340-349): no gradient, no decay, no
    Adam state, like torch AdamW over grad-None parameters.  accu_steps=2 updates at i = 0, 2, 4 (src/trainer.py:63-66)"""
    from causalgen_b200 import HVAE
    from causalgen_b200.trainer import Trainer
    cfg = O.make_cfg("tiny_ukbb", x_like="shared_dgauss", std_init=0.5)
    sd = O.seeded_state_dict(cfg, seed=7)
    model = HVAE(cfg)
    model.load_state_dict(sd, strict=True)
    model.to(DEV)
    frozen = [k for k, p in model.named_parameters() if not p.requires_grad]
    assert frozen == ["likelihood.x_logscale.weight"]
    tr = Trainer(model, 2, beta=cfg.beta, use_graph=True, noise_seed=3, accu_steps=2, export_eps=True, **HP)
    unused = [k for k, _ in model.named_parameters() if k.startswith(f"decoder.blocks.{len(model.decoder.blocks) - 1}.z_feat_proj")]
    assert tr.n_train == sum(p.numel() for k, p in model.named_parameters() if k not in frozen + unused)
    batches = [O.synthetic_batch(cfg, 2, seed=70 + i)[:2] for i in range(5)]
    hist, eps = run_steps(tr, batches)
    assert int(tr.state[0]) == 3, "updates at i = 0, 2, 4"
    after = dict(model.named_parameters())
    assert torch.equal(after[frozen[0]].detach().cpu(), sd[frozen[0]]), "frozen parameter must not move (no wd either)"
    # oracle: accumulate grads over the same micro-batches with the reference's schedule
    sdr = {k: v.clone().requires_grad_(k not in frozen) for k, v in sd.items()}
    state, upd = {}, 0
    for p in sdr.values():
        p.grad = None
    for i, ((x8, pa), e) in enumerate(zip(batches, eps)):
        out = O.hvae_forward(sdr, cfg, O.normalise_x(x8), O.expand_parents(pa, cfg.input_res), O.NoiseTape(tensors=e),
                             beta=cfg.beta)
        (out["elbo"] / 2).backward()
        if i % 2 == 0:
            upd += 1
            params = [p for p in sdr.values() if p.grad is not None]
            torch.nn.utils.clip_grad_norm_(params, HP["grad_clip"])
            lr = O.warmup_lr(HP["lr"], upd, HP["lr_warmup_steps"])
            with torch.no_grad():
                for k, p in sdr.items():
                    if p.grad is None:
                        continue
                    m, v = state.setdefault(k, (torch.zeros_like(p), torch.zeros_like(p)))
                    p.mul_(1 - lr * HP["wd"])
                    m.mul_(0.9).add_(p.grad, alpha=0.1)
                    v.mul_(0.9).addcmul_(p.grad, p.grad, value=0.1)
                    p.addcdiv_(m / (1 - 0.9 ** upd), (v / (1 - 0.9 ** upd)).sqrt().add_(1e-8), value=-lr)
            for p in sdr.values():
                p.grad = None
    # the last block's z_feat_proj is never used by the forward (src/vae.py:297-300): grad None in the reference, so AdamW
    # leaves it alone (no decay either)
    assert all(k not in state for k in unused) and len(unused) == 2
    for k in unused:
        assert torch.equal(after[k].detach().cpu(), sd[k]), k
    names = [k for k, _ in model.named_parameters() if k not in frozen + unused]
    vm = tr._views(tr.m)
    m_got = torch.cat([vm[k].flatten() for k in names]).cpu()
    m_ref = torch.cat([state[k][0].flatten() for k in names])
    parity_report("trainer[accu2,frozen]", "Adam m rel-L2", rel_l2(m_got, m_ref), 3e-2)
    assert rel_l2(m_got, m_ref) <= 3e-2
    dp_got = torch.cat([(after[k].detach().cpu() - sd[k]).flatten() for k in names])
    dp_ref = torch.cat([(sdr[k].detach() - sd[k]).flatten() for k in names])
    parity_report("trainer[accu2,frozen]", "param update cosine", 1 - cosine(dp_got, dp_ref), 0.03, "1 - cos")
    assert cosine(dp_got, dp_ref) >= 0.97


def test_trainer_state_dict_resume():
    cfg, sd, model, tr = make("tiny_ukbb", 2, False)
    batches = [O.synthetic_batch(cfg, 2, seed=80 + i)[:2] for i in range(4)]
    for x8, pa in batches[:2]:
        tr.step(x8, pa)
    snap = {k: v.clone() for k, v in tr.state_dict().items()}
    weights = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ref = [tr.step(x8, pa).clone() for x8, pa in batches[2:]]
    cfg2, _, model2, tr2 = make("tiny_ukbb", 2, False)
    model2.load_state_dict(weights)
    tr2.load_state_dict(snap)
    got = [tr2.step(x8, pa).clone() for x8, pa in batches[2:]]
    for a, b in zip(got, ref):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=2e-5)
    assert rel_l2(tr2.flat_p.cpu(), tr.flat_p.cpu()) <= 1e-5


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_nccl_two_ranks_reduced_bucket_equals_full_batch():
    """torchrun --nproc-per-node 2: the NCCL-reduced gradient bucket equals the single-process gradient of the full
    batch on the same eps, replicas stay bit-identical after graph-replayed steps (tests/_ddp_worker.py)"""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611",
                        os.path.join(ROOT, "tests", "_ddp_worker.py")], capture_output=True, text=True, env=env, timeout=900)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stderr[-3000:]
    for line in r.stdout.splitlines():
        if line.startswith("PARITY "):
            _, q, m, tol = line.split()
            parity_report("nccl world-2", q, float(m), float(tol))
    assert "DDP_OK" in r.stdout
