"""Worker of tests/test_trainer_gpu.py::test_nccl_two_ranks_reduced_bucket_equals_full_batch (run under torchrun,
one rank per GPU, NCCL).  Checks, on real GPUs and through the CUDA kernels:
  1. SUM-all-reduced gradient bucket / world == gradient of the single-process full batch on the same eps
     (src/trainer.py:62-67 under DDP);
  2. after graph-replayed steps the replicas' parameters, Adam moments and EMA are bit-identical.
Prints `PARITY <quantity> <measured> <tolerance>` lines and DDP_OK on rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "causal-gen_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import hvae_oracle as O  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from causalgen_b200 import HVAE, dp
    from causalgen_b200.trainer import Trainer
    name = os.environ.get("DDP_CFG", "tiny_ukbb")
    B = 2
    cfg = O.make_cfg(name)
    sd = O.seeded_state_dict(cfg, seed=7)
    model = HVAE(cfg)
    model.load_state_dict(sd)
    model.cuda()
    tr = Trainer(model, B, beta=cfg.beta, use_graph=True, noise_seed=11, export_eps=True, lr=1e-3, wd=0.05,
                 lr_warmup_steps=0, ema_update_after=0)
    x8_all, pa_all, _ = O.synthetic_batch(cfg, B * world, seed=90)
    lo, hi = dp.shard_batch(B * world, world, rank)
    # ---- step 1 taken apart: forward/backward, reduce, compare, then the optimiser
    tr._load_inputs(x8_all[lo:hi].cuda(), pa_all[lo:hi].cuda())
    tr._fwd_bwd()
    dp.reduce_gradients_(tr.grad)
    eps_all = []
    for e in tr.eps_out:
        parts = [torch.empty_like(e) for _ in range(world)]
        dist.all_gather(parts, e)
        eps_all.append(torch.cat(parts, 0))
    ok = True
    if rank == 0:
        full = HVAE(cfg)
        full.load_state_dict(sd)
        full.cuda().train()
        out = full(O.normalise_x(x8_all).cuda(), pa_all.cuda(), beta=cfg.beta, eps=eps_all)
        out["elbo"].backward()
        named = dict(full.named_parameters())
        name_of = {id(p): k for k, p in model.named_parameters()}
        # tr.grad covers the trainable parameters in flat-buffer order (engine.ordered_params): frozen ones and the last
        # block's never-used z_feat_proj (grad None in the reference too) sit behind it
        live = [p for p in tr.params if p.requires_grad and named[name_of[id(p)]].grad is not None]
        assert sum(p.numel() for p in live) == tr.grad.numel()
        want = torch.cat([named[name_of[id(p)]].grad.flatten() for p in live])
        got = tr.grad / world
        d = float((got - want).norm() / want.norm())
        print(f"PARITY reduced_bucket_vs_full_batch_relL2 {d:.3e} 2e-3")
        ok = ok and d <= 2e-3
    tr._optim()
    tr.steps_done += 1
    tr.iter_in_epoch += 1
    # ---- graph-replayed steps on rank-specific shards, then replica equality
    for i in range(4):
        x8, pa, _ = O.synthetic_batch(cfg, B * world, seed=91 + i)
        tr.step(x8[lo:hi], pa[lo:hi])
    assert tr.g_fb is not None
    for nm, buf in (("params", tr.flat_p), ("adam_m", tr.m), ("adam_v", tr.v), ("ema", tr.ema)):
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf)
        if rank == 0:
            d = max(float((p - parts[0]).abs().max()) for p in parts[1:])
            print(f"PARITY replica_max_abs_diff_{nm} {d:.3e} 0")
            ok = ok and d == 0.0
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.broadcast(flag, 0)
    if rank == 0 and ok:
        print("DDP_OK")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if float(flag) == 1.0 else 1)


if __name__ == "__main__":
    main()
