"""Kernel-level trace of the REAL (CUDA-graph replayed, multi-stream) training step through torch.profiler / CUPTI:
per-kernel start / duration / stream inside the replay -> device busy fraction, per-kernel-family time under overlap,
the largest idle gaps.  Under-profiler durations are not bench numbers; the shares and the overlap picture are the point.
usage: graph_trace.py [config] [batch] [out.json]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "causal-gen_b200"))
import torch
from torch.profiler import ProfilerActivity, profile
from bench import make_trainer, synthetic_host_batches
cfg = sys.argv[1] if len(sys.argv) > 1 else "ukbb192"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "graph_trace.json")
margs, model, tr = make_trainer(cfg, B)
xs, pas = synthetic_host_batches(margs, B, 2, 1)
xs, pas = [x.cuda() for x in xs], [p.cuda() for p in pas]
for i in range(4):
    tr.step_device(xs[i % 2], pas[i % 2])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(2):
        tr.step_device(xs[i % 2], pas[i % 2])
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
rows = sorted(((e.time_range.start, e.time_range.end, e.name, getattr(e, "device_index", 0)) for e in ev), key=lambda r: r[0])
if not rows:
    print("no CUDA events captured"); sys.exit(0)
# keep the second step only (steady state): split at the largest gap between consecutive kernel starts
t0 = rows[0][0]
mid = (rows[0][0] + rows[-1][1]) / 2
second = [r for r in rows if r[0] >= mid]
first_big_gap = 0
span0, span1 = second[0][0], max(r[1] for r in second)


def fam(n):
    for k in ("conv_tc_kernel", "wgrad_mma_kernel", "wgrad1_mma_kernel", "wgrad_tc_kernel", "latent", "stem", "avgpool", "upsample",
              "dgauss", "adamw", "pack_weights", "parents_plane", "colsum", "add_kernel", "sumsq", "Memset", "Memcpy", "elementwise"):
        if k in n:
            return k
    return n[:40]


busy, cur_s, cur_e = 0.0, None, None
for s, e, n, _ in second:
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
agg = {}
for s, e, n, _ in second:
    a = agg.setdefault(fam(n), [0, 0.0])
    a[0] += 1; a[1] += e - s
# concurrency-weighted time: at each instant split the time evenly over the running kernels
points = sorted([(s, 1, fam(n)) for s, e, n, _ in second] + [(e, -1, fam(n)) for s, e, n, _ in second])
active, last, share = {}, None, {}
for t, d, f in points:
    if last is not None and active:
        tot = sum(active.values())
        for k, c in active.items():
            share[k] = share.get(k, 0.0) + (t - last) * c / tot
    active[f] = active.get(f, 0) + d
    if active[f] == 0:
        del active[f]
    last = t
gaps = []
cur_e = None
for s, e, n, _ in second:
    if cur_e is not None and s > cur_e:
        gaps.append((s - cur_e, n))
    cur_e = e if cur_e is None else max(cur_e, e)
span = span1 - span0
res = {"config": cfg, "batch": B, "kernels": len(second), "span_us": span, "busy_us": busy, "idle_us": span - busy,
       "sum_kernel_us": sum(e - s for s, e, _, _ in second),
       "families": {k: {"n": v[0], "sum_us": v[1], "attributed_us": share.get(k, 0.0)} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])},
       "idle_gaps_over_2us": len([g for g in gaps if g[0] > 2]), "idle_in_gaps_over_2us": sum(g[0] for g in gaps if g[0] > 2)}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k != "families"}))
print("%-24s %6s %10s %12s" % ("family", "n", "sum us", "attributed us"))
for k, v in res["families"].items():
    print("%-24s %6d %10.0f %12.0f" % (k, v["n"], v["sum_us"], v["attributed_us"]))
with open(out.replace(".json", "_kernels.csv"), "w") as f:
    f.write("start_us,dur_us,name\n")
    for s, e, n, _ in second:
        f.write("%.2f,%.2f,%s\n" % (s - span0, e - s, n.replace(",", ";")[:80]))
