#!/bin/bash
# round 2, session 4i: cap on the CTAs of the high-resolution weight gradients; other configs before / after the batch-dependent grid
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --no-configs --no-ref-gpu --no-cpu --no-cf > $O/r4i_bench_$name.json 2> $O/r4i_bench.err
  python -c "
import json; d=json.load(open('$O/r4i_bench_$name.json')); print('$name', round(d['value'],1), round(d['ms_per_step'],3), round(d['reference_batch32']['value'],1), d['loss']['elbo'])"
}
run default
run cap120 CG_WGRAD_MAX_CTAS=120
run cap96 CG_WGRAD_MAX_CTAS=96
run cap72 CG_WGRAD_MAX_CTAS=72
run cap48 CG_WGRAD_MAX_CTAS=48
run default2
cfg() { # name, env...
  local name=$1; shift
  env "$@" timeout 400 python bench.py --no-ref-gpu --no-cpu --no-cf --no-ref-batch > $O/r4i_bench_cfg_$name.json 2> $O/r4i_bench.err
  python -c "
import json; d=json.load(open('$O/r4i_bench_cfg_$name.json')); print('$name', round(d['value'],1), {k: round(v['value']) for k, v in d['configs'].items()})"
}
cfg wmt24 CG_WGRAD_MIN_TILES=24
cfg auto
