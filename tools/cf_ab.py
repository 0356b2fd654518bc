"""counterfactual-inference throughput of one config under environment variants (child processes):
usage: python tools/cf_ab.py mimic224 32 "" CAUSALGEN_B200_PRIO=0 CAUSALGEN_B200_STEM_MMA=0 ..."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "causal-gen_b200")]

if sys.argv[1] == "--child":
    import torch
    import bench
    out = bench.cf_config(sys.argv[2], int(sys.argv[3]), 12, 1, 0, 6500.0, 1300.0)
    print("RESULT", json.dumps({k: out[k] for k in ("value", "ms_per_batch", "eager_value")}))
else:
    name, B = sys.argv[1], sys.argv[2]
    for var in sys.argv[3:]:
        env = dict(os.environ)
        env.update(kv.split("=") for kv in var.split(",") if kv)
        r = subprocess.run([sys.executable, __file__, "--child", name, B], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
        print(f"{name} B={B} [{var or 'default'}]:", line[0][7:] if line else r.stderr[-400:])
