#!/usr/bin/env python
"""Benchmark of the hot path: HVAE ELBO training step (images/s) on synthetic UKBB-shape images.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N ...            # the reference algorithm on the host CPUs
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # one rank per GPU, weak scaling

One step = preprocess + forward + backward + (all-reduce) + clip/AdamW/EMA over one batch of
`--batch` images per GPU.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions
of `value`, `e2e`, `roofline`, `cpu_baseline`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "causal-gen_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "hvae_elbo_train_images_per_sec"
HBM_FALLBACK_GBS = 6650.0
TENSOR_FALLBACK_TFS = 1400.0
# conv FLOPs per image of one ELBO forward pass (measured on the reference with hooks);
# a training step executes 3x (forward + data-gradient + weight-gradient)
# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per conv_tc_kernel launch, averaged over the 891
# conv launches of one training step at the given per-GPU batch: profiles/r1h_ncu_launches_one_step_b128.csv
# (ncu pass of this very command).  Below the algorithmic bytes because producer->consumer tensors hit the L2.
CONV_DRAM_TRAFFIC_PER_LAUNCH = {("ukbb192", 128): 85.92e6}
FWD_GFLOP = {"ukbb192": 23.064, "mimic192": 9.127, "morphomnist": 0.0865, "cmnist": 0.0917, "mimic224": 12.460}
CF_GFLOP = {"ukbb192": 47.670, "mimic192": 19.565, "morphomnist": 0.1845, "cmnist": 0.1924, "mimic224": 26.707}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured"
    return HBM_FALLBACK_GBS, TENSOR_FALLBACK_TFS, "fallback"


def synthetic_host_batches(args_m, batch, nbatch, seed):
    """uint8 images with an MNIST/brain-like share of exact zeros + parents, in pinned host memory"""
    rng = np.random.default_rng(seed)
    C, R, ctx = args_m.input_channels, args_m.input_res, args_m.context_dim
    xs, pas = [], []
    for _ in range(nbatch):
        x = rng.integers(0, 256, (batch, C, R, R), dtype=np.uint8)
        x[rng.random((batch, C, R, R)) < 0.4] = 0
        pa = rng.standard_normal((batch, ctx)).astype(np.float32)
        if ctx >= 4:
            pa[:, 0] = rng.integers(0, 2, batch)
            pa[:, -1] = rng.integers(0, 2, batch)
        xs.append(torch.from_numpy(x))
        pas.append(torch.from_numpy(pa))
    return xs, pas


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            if not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            f = [v.strip() for v in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_setup(n_gpus):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        # stdout is ONE JSON line: NCCL prints its version banner on stdout when the communicator is created, so fd 1
        # points at stderr until the first collective has run
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            t = torch.zeros(1, device="cuda")
            dist.all_reduce(t)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    else:
        torch.cuda.set_device(0)
    return world, rank, local


def max_over_ranks(ms, world):
    if world == 1:
        return ms
    import torch.distributed as dist
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, steps, world):
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    barrier(world)
    w1 = time.time()
    return max_over_ranks(e0.elapsed_time(e1), world), w0, w1


def conv_bytes(a):
    """algorithmic HBM bytes of one cg_conv2d launch: inputs + outputs + fused addends + packed weights"""
    npix = a.N * a.H * a.W
    b = 0
    k = 0
    for i in range(a.nsrc):
        s = a.src[i]
        b += npix * s.C * 2
        k += s.C
    for i in range(a.nseg):
        sg = a.seg[i]
        b += npix * sg.cn * (4 if sg.dtype == 1 else 2)
        for extra in (sg.add, sg.add2, sg.mul):
            if extra:
                b += npix * sg.cn * 2
    b += a.ksize * a.ksize * k * a.cout * 2
    return b


def wgrad_bytes(a):
    npix = a.N * a.H * a.W
    b = npix * a.dy_c * 2
    for i in range(a.nsrc):
        b += npix * a.src[i].C * 2
    return b + a.cout_l * a.cin_l * a.ksize * a.ksize * 4


def profile_step(trainer):
    """one eager (un-graphed) step with CUDA events around every launch: per-kernel-family device time and
    the algorithmic bytes of the conv launches (for the roofline object)"""
    from causalgen_b200 import _lib as L
    prog = trainer.prog
    s = torch.cuda.current_stream().cuda_stream
    evs = []
    for t in prog.zero:
        t.zero_()
    trainer.eng.flat_grad.zero_()
    trainer.eng.pack_weights(s)
    torch.cuda.synchronize()
    for ln in prog.launches:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ln(s)
        e1.record()
        evs.append((ln, e0, e1))
    torch.cuda.synchronize()
    fam = {}
    conv_t = conv_b = wg_t = wg_b = 0.0
    nconv = nwg = 0
    for ln, e0, e1 in evs:
        ms = e0.elapsed_time(e1)
        name = getattr(ln, "name", "pyop")
        fam[name] = fam.get(name, 0.0) + ms
        if name == "cg_conv2d":
            conv_t += ms
            conv_b += conv_bytes(ln.keep[0])
            nconv += 1
        elif name == "cg_conv2d_wgrad":
            wg_t += ms
            wg_b += wgrad_bytes(ln.keep[0])
            nwg += 1
    total = sum(fam.values())
    return dict(families_ms={k: round(v, 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])},
                total_ms=total, conv_ms=conv_t, conv_bytes=conv_b, nconv=nconv, wgrad_ms=wg_t, wgrad_bytes=wg_b,
                nwgrad=nwg)


def cpu_port_step_time(name, sd_cpu, batch, budget_s, threads):
    """the reference algorithm (oracle port) on the host cores: ELBO forward + backward + clip + AdamW + EMA"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hvae_oracle as O
    torch.set_num_threads(threads)
    cfg = O.make_cfg(name)
    sd = {k: v.clone().float().requires_grad_(True) for k, v in sd_cpu.items()}
    ema = {k: v.detach().clone() for k, v in sd.items()}
    x8, pa, _ = O.synthetic_batch(cfg, batch, seed=3)
    x = O.normalise_x(x8)
    pa_full = O.expand_parents(pa, cfg.input_res)
    state = {}
    times = []
    t_start = time.time()
    step = 0
    while True:
        t0 = time.time()
        O.train_step_cpu(sd, cfg, x, pa_full, O.NoiseTape(seed=step), state, lr=1e-3, wd=0.05, step=step + 1, ema=ema)
        dt = time.time() - t0
        step += 1
        if step > 1:
            times.append(dt)
        if (time.time() - t_start > budget_s and len(times) >= 2) or len(times) >= 8:
            break
    return float(np.median(times)), len(times)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port: the reference
    package itself cannot travel to the GPU box) with all host threads, bounded sample per step"""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from causalgen_b200 import HVAE
    from causalgen_b200.presets import init_like_reference_main, make_args
    torch.manual_seed(7)
    m = init_like_reference_main(HVAE(make_args(args.config)))
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    threads = os.cpu_count() or 1
    bs = args.cpu_batch
    sec, n = cpu_port_step_time(args.config, sd, bs, budget_s=max(10.0, 4.0 * args.steps), threads=threads)
    val = bs / sec
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus,
            "steps": n, "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config} HVAE ELBO train step (fwd+bwd+clip+AdamW+EMA), CPU batch {bs}",
                       "note": "reference algorithm restated in oracle/ (torch CPU fp32); /root/reference cannot travel"},
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": "port",
                             "sample": f"{n} timed steps of batch {bs} after 1 warm-up"},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="ukbb192")
    ap.add_argument("--batch", type=int, default=128,
                    help="images per GPU per step (throughput configuration; the reference's bs=32 is also measured)")
    ap.add_argument("--no-ref-batch", action="store_true", help="skip the extra measurement at the reference bs=32")
    ap.add_argument("--cpu-batch", type=int, default=2)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-cf", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback (use --impl reference for the "
                         "CPU arm)")
    from causalgen_b200 import HVAE, counterfactual
    from causalgen_b200.presets import init_like_reference_main, make_args
    from causalgen_b200.trainer import Trainer

    world, rank, local = dist_setup(args.gpus)
    hbm_gbs, tensor_tfs, peak_src = peaks()
    margs = make_args(args.config)
    torch.manual_seed(7)
    model = init_like_reference_main(HVAE(margs)).cuda()
    sd_cpu = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()} if rank == 0 else None
    B = args.batch
    trainer = Trainer(model, B, lr=margs.lr, wd=margs.wd, betas=margs.betas, lr_warmup_steps=margs.lr_warmup_steps,
                      grad_clip=margs.grad_clip, grad_skip=margs.grad_skip, ema_rate=margs.ema_rate, beta=margs.beta,
                      use_graph=not args.no_graph, noise_seed=7)
    nb = 4
    xs_h, pas_h = synthetic_host_batches(margs, B, nb, seed=100 + rank)
    xs_h = [x.pin_memory() for x in xs_h]
    pas_h = [p.pin_memory() for p in pas_h]
    xs_d = [x.cuda() for x in xs_h]
    pas_d = [p.cuda() for p in pas_h]

    for i in range(args.warmup):
        trainer.step(xs_h[i % nb], pas_h[i % nb])
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    # device-resident arm: inputs already in HBM when the timed region starts
    ms_dev, w0, w1 = timed(lambda i: trainer.step_device(xs_d[i % nb], pas_d[i % nb]), args.steps, world)
    clocks = sampler.summary(w0, w1)
    # end-to-end arm: pinned host uint8 batch -> H2D -> step -> D2H loss, every step
    ms_e2e, w2, w3 = timed(lambda i: trainer.step(xs_h[i % nb], pas_h[i % nb]), args.steps, world)
    sampler.stop()
    loss = [float(v) for v in trainer.loss_host]
    imgs = B * world * args.steps
    value = imgs / (ms_dev / 1e3)
    e2e = imgs / (ms_e2e / 1e3)

    line = None
    if rank == 0:
        prof = profile_step(trainer)
        conv_gbs = prof["conv_bytes"] / (prof["conv_ms"] / 1e3) / 1e9
        wg_gbs = prof["wgrad_bytes"] / (prof["wgrad_ms"] / 1e3) / 1e9
        gflop_step = 3.0 * FWD_GFLOP.get(args.config, 0.0)
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.config} HVAE ELBO training step, {margs.input_channels}x{margs.input_res}x"
                                   f"{margs.input_res} uint8 images, batch {B}/GPU, fwd+bwd+"
                                   f"allreduce+clip+AdamW+EMA, reference init (seed 7)",
                       "params": int(sum(p.numel() for p in model.parameters())),
                       "l2": "per-step working set (saved activations + gradients) is several GB >> 126 MB L2; "
                             "4 distinct input batches rotate",
                       "cuda_graph": trainer.g_fb is not None, "parallelism": f"dp{world}"},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "images/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(xs_h[0].numel() + pas_h[0].numel() * 4), "d2h_bytes_per_step": 12},
            "gpu_launches": int(trainer.kernels_per_step * args.steps),
            "loss": {"elbo": loss[0], "nll": loss[1], "kl": loss[2], "skipped_updates": trainer.skipped_updates()},
            "roofline": {"bound": "hbm", "kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv: forward + data-gradient)",
                         "achieved": conv_gbs, "peak": hbm_gbs, "unit": "GB/s", "frac": conv_gbs / hbm_gbs,
                         "peak_source": peak_src, "traffic": CONV_DRAM_TRAFFIC_PER_LAUNCH.get((args.config, B)),
                         "algorithmic_bytes_per_launch": prof["conv_bytes"] / max(prof["nconv"], 1),
                         "launches_per_step": prof["nconv"], "avg_launch_us": 1e3 * prof["conv_ms"] / max(prof["nconv"], 1),
                         "algorithmic_bytes_per_step": prof["conv_bytes"],
                         "share_of_step": prof["conv_ms"] / prof["total_ms"],
                         "wgrad_kernel": {"kernel": "wgrad_mma_kernel (mma.sync, 3x3) + wgrad_tc_kernel (tcgen05, 1x1)",
                                          "achieved": wg_gbs, "frac": wg_gbs / hbm_gbs, "launches_per_step": prof["nwgrad"],
                                          "share_of_step": prof["wgrad_ms"] / prof["total_ms"]},
                         "tensor": {"conv_tflops": gflop_step * value / 1e3, "peak": tensor_tfs,
                                    "frac": gflop_step * value / 1e3 / tensor_tfs,
                                    "gflop_per_image_step": gflop_step}},
            "breakdown_ms": prof["families_ms"],
            "hbm_gb_peak_train": round(torch.cuda.max_memory_allocated() / 1e9, 2),
        }
    # counterfactual inference throughput (abduct + 2x forward_latents + combine), replicas only
    if not args.no_cf:
        model.eval()
        Bc = B
        xf = (xs_d[0].float() - 127.5) / 127.5
        pa, cfp = pas_d[0], pas_d[1]
        for _ in range(2):
            counterfactual(model, xf, pa, cfp)
        ncf = max(3, args.steps // 2)
        ms_cf, _, _ = timed(lambda i: counterfactual(model, xf, pa, cfp), ncf, world)
        from causalgen_b200 import CounterfactualGraph
        cfg_run = CounterfactualGraph(model, Bc)
        for _ in range(2):
            cfg_run(xf, pa, cfp)
        ms_cfg, _, _ = timed(lambda i: cfg_run(xf, pa, cfp), ncf, world)
        if rank == 0:
            cf_val = Bc * world * ncf / (ms_cfg / 1e3)
            line["cf_inference"] = {"metric": "counterfactual_images_per_sec", "value": cf_val, "unit": "images/s",
                                    "batch_per_gpu": Bc, "ms_per_batch": ms_cfg / ncf,
                                    "eager_value": Bc * world * ncf / (ms_cf / 1e3),
                                    "tensor_frac": CF_GFLOP.get(args.config, 0) * cf_val / 1e3 / tensor_tfs,
                                    "note": "abduct + forward_latents(cf_pa, pa) batched + combine (src/pgm/dscm.py:47-72), "
                                            "CUDA-graph replay; eager_value = same launches issued from Python"}
        del cfg_run
    # the same step at the reference's own batch size (src/hps.py ukbb192: bs=32): latency-bound regime
    if B != 32 and not args.no_ref_batch:
        del trainer
        model.train()
        model.engine().programs.clear()
        torch.cuda.empty_cache()
        tr32 = Trainer(model, 32, lr=margs.lr, wd=margs.wd, betas=margs.betas, lr_warmup_steps=margs.lr_warmup_steps,
                       grad_clip=margs.grad_clip, grad_skip=margs.grad_skip, ema_rate=margs.ema_rate, beta=margs.beta,
                       use_graph=not args.no_graph, noise_seed=7)
        x32 = [x[:32].contiguous() for x in xs_d]
        p32 = [p[:32].contiguous() for p in pas_d]
        for i in range(4):
            tr32.step_device(x32[i % nb], p32[i % nb])
        n32 = min(args.steps, 10)
        ms32, _, _ = timed(lambda i: tr32.step_device(x32[i % nb], p32[i % nb]), n32, world)
        if rank == 0:
            line["reference_batch32"] = {"value": 32 * world * n32 / (ms32 / 1e3), "unit": "images/s",
                                         "ms_per_step": ms32 / n32, "batch_per_gpu": 32}
        del tr32
    if rank == 0:
        line["hbm_gb_peak_total"] = round(torch.cuda.max_memory_allocated() / 1e9, 2)
        if world == 1 and not args.no_cpu:
            sec, n = cpu_port_step_time(args.config, sd_cpu, args.cpu_batch, args.cpu_seconds, os.cpu_count() or 1)
            line["cpu_baseline"] = {"value": args.cpu_batch / sec, "unit": "images/s", "cores": os.cpu_count() or 1,
                                    "kind": "port", "sample": f"{n} timed steps of batch {args.cpu_batch} "
                                    f"(same model/weights, fwd+bwd+clip+AdamW+EMA) after 1 warm-up"}
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
