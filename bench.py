#!/usr/bin/env python
"""Benchmark of the hot path: HVAE ELBO training step (images/s) on synthetic UKBB-shape images.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N ...            # the reference itself on the host CPUs
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # one rank per GPU, weak scaling

One step = preprocess + forward + backward + (all-reduce) + clip/AdamW/EMA over one batch of
`--batch` images per GPU.  Prints ONE JSON line (rank 0).  DESIGN.md section 6 defines `value`, `e2e`, `roofline`,
`cpu_baseline`, `configs` and `reference_gpu`.
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
# conv FLOPs per image (measured on the reference with hooks; tools/roofline.py regenerates them);
# a training step executes 3x the forward FLOPs (forward + data-gradient + weight-gradient)
FWD_GFLOP = {"ukbb192": 23.064, "mimic192": 9.127, "morphomnist": 0.0865, "cmnist": 0.0917, "mimic224": 12.460}
CF_GFLOP = {"ukbb192": 47.670, "mimic192": 19.565, "morphomnist": 0.1845, "cmnist": 0.1924, "mimic224": 26.707}
# SURVEY 8(d) "fwd layerwise bytes" per image (bf16 input + output activation bytes of every conv of one ELBO forward)
FWD_LAYERWISE_MB = {"ukbb192": 226.1, "mimic192": 282.9, "morphomnist": 5.92, "cmnist": 6.01, "mimic224": 385.4}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured"
    return HBM_FALLBACK_GBS, TENSOR_FALLBACK_TFS, "fallback"


def measured_traffic(config, batch):
    """dram__bytes_read.sum + dram__bytes_write.sum per conv_tc_kernel launch from the committed ncu pass of one step of
    this very workload (profiles/r2_traffic.json, written by tools/ncu_summary.py); None when no capture matches"""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p))
    e = d.get(f"{config}:{batch}")
    return (e["conv_dram_bytes_per_launch"], e) if e else (None, None)


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
    """ISSUED bytes of one cg_conv2d launch (padded channels, fused addends / masks, packed weights): what the
    kernel asks the memory system for -- reported next to the algorithmic figure, never as the roofline numerator"""
    npix = a.N * a.H * a.W
    b = 0
    k = 0
    for i in range(a.nsrc):
        s = a.src[i]
        b += npix * (s.c8 * 8 if s.c8 else s.C) * 2   # octets physically stored (cg_src.c8), not the K-blocks
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
    b = npix * (a.dy_c8 * 8 if a.dy_c8 else a.dy_c) * 2
    for i in range(a.nsrc):
        b += npix * (a.src[i].c8 * 8 if a.src[i].c8 else a.src[i].C) * 2
    return b + a.cout_l * a.cin_l * a.ksize * a.ksize * 4


def profile_program(prog, pre=None):
    """one eager (un-graphed) pass over a launch program with CUDA events around every launch, issued behind a spin
    kernel so the host is ahead of the device and the events bracket device time only.  Returns per-family device
    time, and for the conv launches the ALGORITHMIC bytes (bf16 input + output activation bytes, logical
    channels) next to the issued bytes."""
    s = torch.cuda.current_stream().cuda_stream
    if pre is not None:
        pre()
    torch.cuda.synchronize()
    torch.cuda._sleep(int(8e7))  # ~40 ms head start for the host
    evs = []
    for ln in prog.launches:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ln(s)
        e1.record()
        evs.append((ln, e0, e1))
    torch.cuda.synchronize()
    fam = {}
    conv_t = conv_algo = conv_issued = wg_t = wg_algo = 0.0
    nconv = nwg = 0
    for ln, e0, e1 in evs:
        ms = e0.elapsed_time(e1)
        name = getattr(ln, "name", "pyop")
        fam[name] = fam.get(name, 0.0) + ms
        if name == "cg_conv2d":
            conv_t += ms
            conv_algo += ln.algo_bytes
            conv_issued += conv_bytes(ln.keep[0])
            nconv += 1
        elif name == "cg_conv2d_wgrad":
            wg_t += ms
            wg_algo += ln.algo_bytes
            nwg += 1
    total = sum(fam.values())
    return dict(families_ms={k: round(v, 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])},
                total_ms=total, conv_ms=conv_t, conv_algo_bytes=conv_algo, conv_issued_bytes=conv_issued, nconv=nconv,
                wgrad_ms=wg_t, wgrad_algo_bytes=wg_algo, nwgrad=nwg)


def profile_step(trainer):
    def pre():
        for t in trainer.prog.zero:
            t.zero_()
        trainer.eng.flat_grad.zero_()
        trainer.eng.pack_weights(torch.cuda.current_stream().cuda_stream)
    out = profile_program(trainer.prog, pre)
    trainer.eng.flat_grad.zero_()
    return out


def roofline_object(prof, hbm_gbs, peak_src, step_ms, config, batch, tensor=None):
    conv_gbs = prof["conv_algo_bytes"] / (prof["conv_ms"] / 1e3) / 1e9
    traffic, tr_meta = measured_traffic(config, batch)
    r = {"bound": "hbm", "kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv: forward + data-gradient launches)",
         "achieved": conv_gbs, "peak": hbm_gbs, "unit": "GB/s", "frac": conv_gbs / hbm_gbs, "peak_source": peak_src,
         "traffic": traffic, "traffic_source": (tr_meta or {}).get("source"),
         "algorithmic_bytes_per_launch": prof["conv_algo_bytes"] / max(prof["nconv"], 1),
         "algorithmic_bytes_def": "SURVEY 8(d): sum over conv launches of 2 B x pixels x (logical Cin + logical Cout)",
         "issued_bytes_per_launch": prof["conv_issued_bytes"] / max(prof["nconv"], 1),
         "issued_over_algorithmic": prof["conv_issued_bytes"] / max(prof["conv_algo_bytes"], 1),
         "launches_per_step": prof["nconv"], "avg_launch_us": 1e3 * prof["conv_ms"] / max(prof["nconv"], 1),
         "algorithmic_bytes_per_step": prof["conv_algo_bytes"],
         "timing": "CUDA events around each launch of one eager pass (launching stream, host ahead of device)",
         "share_of_serialised_pass": prof["conv_ms"] / prof["total_ms"],
         "serialised_pass_ms": prof["total_ms"], "graph_step_ms": step_ms}
    if prof["nwgrad"]:
        wg = prof["wgrad_algo_bytes"] / (prof["wgrad_ms"] / 1e3) / 1e9
        r["wgrad_kernel"] = {"kernel": "wgrad_mma_kernel (mma.sync, 3x3) + wgrad_tc_kernel (tcgen05, 1x1)",
                             "achieved": wg, "frac": wg / hbm_gbs, "launches_per_step": prof["nwgrad"],
                             "share_of_serialised_pass": prof["wgrad_ms"] / prof["total_ms"]}
    if tensor is not None:
        r["tensor"] = tensor
    return r


# ---------------------------------------------------------------------------------------------- CPU / reference arms
def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hvae_oracle as O
    import ref_runner as R
    return O, R


def cpu_train_step_time(name, sd_cpu, batch, budget_s, threads, steps=None, warmup=1):
    """the reference's CPU implementation of the step on the host cores: the staged reference itself when
    baseline/_ref exists (kind "reference"), else the oracle port (kind "port").  steps / warmup: time exactly that many
    steps (the --impl reference arm); budget_s then only stops a run that would take longer (at least 2 steps)"""
    O, R = _oracle()
    tk = dict(warmup=warmup) if steps is None else dict(warmup=warmup, min_steps=min(2, steps), max_steps=steps)
    torch.set_num_threads(threads)
    cfg = O.make_cfg(name)
    x8, pa, _ = O.synthetic_batch(cfg, batch, seed=3)
    if R.available() and name in R.FLAGS:
        st = R.RefTrainStep(name, "cpu")
        sec, n = R.time_steps(lambda: st(x8, pa), lambda: None, budget_s, **tk)
        return sec, n, "reference"
    sd = {k: v.clone().float().requires_grad_(True) for k, v in sd_cpu.items()}
    ema = {k: v.detach().clone() for k, v in sd.items()}
    x, pa_full = O.normalise_x(x8), O.expand_parents(pa, cfg.input_res)
    state, step = {}, [0]

    def fn():
        step[0] += 1
        O.train_step_cpu(sd, cfg, x, pa_full, O.NoiseTape(seed=step[0]), state, lr=1e-3, wd=0.05, step=step[0], ema=ema)
    sec, n = R.time_steps(fn, lambda: None, budget_s, **tk)
    return sec, n, "port"


def cpu_cf_time(name, sd_cpu, batch, budget_s, threads):
    """abduct + 2 x forward_latents + combine (src/pgm/dscm.py:52-56) on the host cores"""
    O, R = _oracle()
    torch.set_num_threads(threads)
    cfg = O.make_cfg(name)
    x8, pa, cf = O.synthetic_batch(cfg, batch, seed=3)
    x, pa_full, cf_full = O.normalise_x(x8), O.expand_parents(pa, cfg.input_res), O.expand_parents(cf, cfg.input_res)
    if R.available() and name in R.FLAGS:
        _, model, _ = R.build(name, "cpu")
        model.eval()
        sec, n = R.time_steps(lambda: R.ref_counterfactual(model, x, pa_full, cf_full), lambda: None, budget_s)
        return sec, n, "reference"
    sd = {k: v.clone().float() for k, v in sd_cpu.items()}

    def fn():
        with torch.no_grad():
            O.counterfactual(sd, cfg, x, pa_full, cf_full, O.NoiseTape(seed=1))
    sec, n = R.time_steps(fn, lambda: None, budget_s)
    return sec, n, "port"


def reference_on_gpu(name, batch, steps):
    """the real competitor: the unmodified reference module, eager PyTorch + cuDNN on this B200, default
    TF32 and under autocast(bf16), same step (fwd + bwd + clip + AdamW + EMA), inputs resident"""
    O, R = _oracle()
    if not R.available() or name not in R.FLAGS:
        return {"unavailable": "reference not staged under baseline/_ref"}
    cfg = O.make_cfg(name)
    x8, pa, _ = O.synthetic_batch(cfg, batch, seed=3)
    x8, pa = x8.cuda(), pa.cuda()
    out = {"batch_per_gpu": batch, "what": "reference HVAE (src/vae.py) eager on cuda:0, src/trainer.py:62-87 step"}
    for mode, kw in (("tf32", dict(autocast_bf16=False, tf32=True)), ("bf16_autocast", dict(autocast_bf16=True, tf32=True))):
        try:
            st = R.RefTrainStep(name, "cuda", **kw)
            for _ in range(3):
                st(x8, pa)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                o = st(x8, pa)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[mode] = {"value": batch / (ms / 1e3), "unit": "images/s", "ms_per_step": ms,
                         "elbo": float(o["elbo"].detach())}
            del st
        except Exception as ex:  # e.g. out of memory at a large batch: report, do not hide
            out[mode] = {"error": f"{type(ex).__name__}: {str(ex)[:160]}"}
        torch.cuda.empty_cache()
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path with all host threads, bounded sample"""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    sd = None
    _, R = _oracle()
    if not (R.available() and args.config in R.FLAGS):
        from causalgen_b200 import HVAE
        from causalgen_b200.presets import init_like_reference_main, make_args
        torch.manual_seed(7)
        m = init_like_reference_main(HVAE(make_args(args.config)))
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    threads = os.cpu_count() or 1
    bs = args.cpu_batch
    # exactly --warmup untimed and --steps timed steps, each one batch of `bs` images (the bounded sample of the workload);
    # a 150 s budget only cuts a run short that would not end within a few minutes (the line reports what was timed)
    sec, n, kind = cpu_train_step_time(args.config, sd, bs, budget_s=150.0, threads=threads, steps=max(1, args.steps),
                                       warmup=max(1, args.warmup))
    val = bs / sec
    what = ("unmodified reference (baseline/_ref/src: vae.py HVAE + trainer.py step body) on the host CPUs"
            if kind == "reference" else "reference algorithm restated in oracle/ (torch CPU fp32)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus,
            "steps": n, "warmup": max(1, args.warmup), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config} HVAE ELBO train step (fwd+bwd+clip+AdamW+EMA), CPU batch {bs}",
                       "note": what},
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": kind,
                             "sample": f"{n} timed steps of batch {bs} after {max(1, args.warmup)} warm-up (median step time)"},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- GPU arms
def make_trainer(name, B, use_graph=True, **over):
    from causalgen_b200 import HVAE
    from causalgen_b200.presets import init_like_reference_main, make_args
    from causalgen_b200.trainer import Trainer
    margs = make_args(name, **over)
    torch.manual_seed(7)
    model = init_like_reference_main(HVAE(margs)).cuda()
    tr = Trainer(model, B, lr=margs.lr, wd=margs.wd, betas=margs.betas, lr_warmup_steps=margs.lr_warmup_steps,
                 grad_clip=margs.grad_clip, grad_skip=margs.grad_skip, ema_rate=margs.ema_rate, beta=margs.beta,
                 use_graph=use_graph, noise_seed=7)
    return margs, model, tr


def release(model):
    model.engine().programs.clear()
    import gc
    gc.collect()
    torch.cuda.empty_cache()


def side_config_train(name, B, steps, world, rank, hbm_gbs, tensor_tfs, **over):
    """train-step throughput of one more BASELINE config (device-resident inputs, graph replay, 4 rotating batches)"""
    margs, model, tr = make_trainer(name, B, **over)
    xs, pas = synthetic_host_batches(margs, B, 4, seed=200 + rank)
    xs, pas = [x.cuda() for x in xs], [p.cuda() for p in pas]
    for i in range(4):
        tr.step_device(xs[i % 4], pas[i % 4])
    ms, _, _ = timed(lambda i: tr.step_device(xs[i % 4], pas[i % 4]), steps, world)
    val = B * world * steps / (ms / 1e3)
    loss = [float(v) for v in tr.prog.out3]
    out = {"metric": METRIC, "value": val, "unit": "images/s", "batch_per_gpu": B, "ms_per_step": ms / steps,
           "cuda_graph": tr.g_fb is not None, "x_like": margs.x_like, "kernels_per_step": tr.kernels_per_step,
           "tensor_frac": 3 * FWD_GFLOP[name] * val / 1e3 / tensor_tfs,
           "hbm_layerwise_frac": 3 * FWD_LAYERWISE_MB[name] * 1e6 * val / 1e9 / hbm_gbs,
           "loss": {"elbo": loss[0], "nll": loss[1], "kl": loss[2]}, "skipped_updates": tr.skipped_updates()}
    del tr
    release(model)
    return out


def cf_config(name, B, steps, world, rank, hbm_gbs, tensor_tfs, model=None, profile=False, peak_src="measured"):
    """counterfactual inference (abduct + 2 x forward_latents + combine, src/pgm/dscm.py:47-72): replicas only"""
    from causalgen_b200 import CounterfactualGraph, HVAE, counterfactual
    from causalgen_b200.presets import init_like_reference_main, make_args
    margs = make_args(name)
    own = model is None
    if own:
        torch.manual_seed(7)
        model = init_like_reference_main(HVAE(margs)).cuda()
    model.eval()
    xs, pas = synthetic_host_batches(margs, B, 2, seed=300 + rank)
    x8h = xs[0].pin_memory()
    pah, cfh = pas[0].pin_memory(), pas[1].pin_memory()
    xf = (xs[0].cuda().float() - 127.5) / 127.5
    pa, cfp = pas[0].cuda(), pas[1].cuda()
    for _ in range(2):
        counterfactual(model, xf, pa, cfp)
    ms_eager, _, _ = timed(lambda i: counterfactual(model, xf, pa, cfp), max(2, steps // 2), world)
    run = CounterfactualGraph(model, B)
    for _ in range(2):
        run(xf, pa, cfp)
    ms, _, _ = timed(lambda i: run(xf, pa, cfp), steps, world)
    # end to end through the public call: pinned host uint8 image + parents in, counterfactual image back on the host
    out_h = torch.empty(B, margs.input_channels, margs.input_res, margs.input_res).pin_memory()

    def e2e_step(i):
        xd = (x8h.cuda(non_blocking=True).float() - 127.5) / 127.5
        cf_x, _ = run(xd, pah.cuda(non_blocking=True), cfh.cuda(non_blocking=True))
        out_h.copy_(cf_x, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e2e_step(0)
    ms_e2e, _, _ = timed(e2e_step, steps, world)
    val = B * world * steps / (ms / 1e3)
    out = {"metric": "counterfactual_images_per_sec", "value": val, "unit": "images/s", "batch_per_gpu": B,
           "ms_per_batch": ms / steps, "scaling": "replicas (no collective)",
           "eager_value": B * world * max(2, steps // 2) / (ms_eager / 1e3),
           "e2e": {"value": B * world * steps / (ms_e2e / 1e3), "unit": "images/s",
                   "h2d_bytes_per_step": int(x8h.numel() + 2 * pah.numel() * 4), "d2h_bytes_per_step": int(out_h.numel() * 4)},
           "tensor_frac": CF_GFLOP[name] * val / 1e3 / tensor_tfs,
           "note": "abduct + forward_latents(cf_pa) + forward_latents(pa) + combine in ONE launch program, CUDA-graph "
                   "replay; eager_value = same launches issued from Python"}
    if profile and rank == 0:
        prog = model.engine().programs[("cf", B)]
        prof = profile_program(prog, lambda: model.engine().pack_weights())
        out["roofline"] = roofline_object(prof, hbm_gbs, peak_src, ms / steps, name + ":cf", B,
                                          tensor={"conv_tflops": CF_GFLOP[name] * val / 1e3, "peak": tensor_tfs,
                                                  "frac": CF_GFLOP[name] * val / 1e3 / tensor_tfs})
        out["breakdown_ms"] = prof["families_ms"]
    del run
    if own:
        release(model)
    return out


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
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-cf", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configs (morphomnist, cmnist, ...)")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference-eager-on-B200 competitor arm")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback (use --impl reference for the "
                         "CPU arm)")
    world, rank, local = dist_setup(args.gpus)
    hbm_gbs, tensor_tfs, peak_src = peaks()
    B = args.batch
    margs, model, trainer = make_trainer(args.config, B, use_graph=not args.no_graph)
    sd_cpu = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()} if rank == 0 else None
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
    last = {}

    def e2e_step(i):
        last["loss"] = trainer.step(xs_h[i % nb], pas_h[i % nb])
    ms_e2e, w2, w3 = timed(e2e_step, args.steps, world)
    sampler.stop()
    loss = [float(v) for v in last["loss"]]
    imgs = B * world * args.steps
    value = imgs / (ms_dev / 1e3)
    e2e = imgs / (ms_e2e / 1e3)

    line = None
    if rank == 0:
        prof = profile_step(trainer)
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
            "roofline": roofline_object(
                prof, hbm_gbs, peak_src, ms_dev / args.steps, args.config, B,
                tensor={"conv_tflops": gflop_step * value / world / 1e3, "peak": tensor_tfs,
                        "frac": gflop_step * value / world / 1e3 / tensor_tfs, "gflop_per_image_step": gflop_step}),
            "hbm_layerwise_frac_step": 3 * FWD_LAYERWISE_MB.get(args.config, 0) * 1e6 * value / world / 1e9 / hbm_gbs,
            "breakdown_ms": prof["families_ms"],
            "hbm_gb_peak_train": round(torch.cuda.max_memory_allocated() / 1e9, 2),
        }
    # counterfactual inference throughput of the same config, replicas only
    del trainer
    release(model)
    if not args.no_cf:
        cf = cf_config(args.config, B, max(3, args.steps // 2), world, rank, hbm_gbs, tensor_tfs, model=model, profile=True,
                       peak_src=peak_src)
        if rank == 0:
            line["cf_inference"] = cf
        release(model)
    # the same step at the reference's own batch size (src/hps.py ukbb192: bs=32): latency-bound regime
    if B != 32 and not args.no_ref_batch:
        from causalgen_b200.trainer import Trainer
        model.train()
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
        release(model)
    del model
    torch.cuda.empty_cache()
    # the other BASELINE.json configs: driver-run numbers, same clocked run
    if not args.no_configs:
        n = min(args.steps, 10)
        cfgs = {}
        if world == 1:
            cfgs["morphomnist"] = side_config_train("morphomnist", 1024, n, world, rank, hbm_gbs, tensor_tfs)
            cfgs["morphomnist_bs32"] = side_config_train("morphomnist", 32, n, world, rank, hbm_gbs, tensor_tfs)
            cfgs["cmnist"] = side_config_train("cmnist", 1024, n, world, rank, hbm_gbs, tensor_tfs)
            cfgs["cmnist_dmol"] = side_config_train("cmnist", 1024, n, world, rank, hbm_gbs, tensor_tfs, x_like="diag_dmol")
            cfgs["mimic192"] = side_config_train("mimic192", 64, n, world, rank, hbm_gbs, tensor_tfs)
        cfgs["mimic224_cf"] = cf_config("mimic224", 32, max(3, n // 2), world, rank, hbm_gbs, tensor_tfs)
        if rank == 0:
            line["configs"] = cfgs
    if rank == 0:
        line["hbm_gb_peak_total"] = round(torch.cuda.max_memory_allocated() / 1e9, 2)
        if world == 1 and not args.no_ref_gpu:
            line["reference_gpu"] = reference_on_gpu(args.config, 32, 5)
            if B != 32:
                big = reference_on_gpu(args.config, B, 3)
                line["reference_gpu"][f"batch{B}"] = {k: v for k, v in big.items() if k in ("tf32", "bf16_autocast")}
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            sec, n, kind = cpu_train_step_time(args.config, sd_cpu, args.cpu_batch, args.cpu_seconds, cores)
            line["cpu_baseline"] = {"value": args.cpu_batch / sec, "unit": "images/s", "cores": cores,
                                    "kind": kind, "sample": f"{n} timed steps of batch {args.cpu_batch} "
                                    f"(fwd+bwd+clip+AdamW+EMA, src/trainer.py:62-87) after 1 warm-up"}
            if "cf_inference" in line:
                sec, n, kind = cpu_cf_time(args.config, sd_cpu, args.cpu_batch, args.cpu_seconds / 2, cores)
                line["cf_inference"]["cpu_baseline"] = {
                    "value": args.cpu_batch / sec, "unit": "images/s", "cores": cores, "kind": kind,
                    "sample": f"{n} timed passes of batch {args.cpu_batch} (abduct + 2 x forward_latents + combine, "
                              f"src/pgm/dscm.py:52-56, no_grad) after 1 warm-up"}
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
