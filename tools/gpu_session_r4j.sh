#!/bin/bash
# round 2, session 4j: weight-gradient tiles per CTA on the non-light configs (forced through CG_WGRAD_MIN_TILES)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
run() { # name, config, batch, env...
  local name=$1 cfg=$2 b=$3; shift 3
  env "$@" timeout 300 python bench.py --config $cfg --batch $b --no-configs --no-ref-gpu --no-cpu --no-cf --no-ref-batch > $O/r4j_bench_$name.json 2> $O/r4j_bench.err
  python -c "
import json; d=json.load(open('$O/r4j_bench_$name.json')); print('$name', round(d['value'],1), round(d['ms_per_step'],3), d['loss']['elbo'])" || tail -3 $O/r4j_bench.err
}
for t in 24 12 48 96; do run mimic192_wmt$t mimic192 64 CG_WGRAD_MIN_TILES=$t; done
for t in 24 12 48 96; do run morphomnist_wmt$t morphomnist 1024 CG_WGRAD_MIN_TILES=$t; done
run ukbb_b64_auto ukbb192 64
run ukbb_b64_wmt96 ukbb192 64 CG_WGRAD_MIN_TILES=96
run ukbb_b64_wmt192 ukbb192 64 CG_WGRAD_MIN_TILES=192
