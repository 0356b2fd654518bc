#!/bin/bash
# round 2, session 4h: weight-gradient grid size, second sweep (larger minimum tile counts per CTA)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --no-configs --no-ref-gpu --no-cpu --no-cf > $O/r4h_bench_$name.json 2> $O/r4h_bench.err
  python -c "
import json; d=json.load(open('$O/r4h_bench_$name.json')); print('$name', round(d['value'],1), round(d['ms_per_step'],3), round(d['reference_batch32']['value'],1), d['loss']['elbo'])"
}
run wmt24
run wmt48 CG_WGRAD_MIN_TILES=48
run wmt72 CG_WGRAD_MIN_TILES=72
run wmt96 CG_WGRAD_MIN_TILES=96
run wmt144 CG_WGRAD_MIN_TILES=144
run wmt192 CG_WGRAD_MIN_TILES=192
run wmt384 CG_WGRAD_MIN_TILES=384
run wmt48b CG_WGRAD_MIN_TILES=48
run wmt96b CG_WGRAD_MIN_TILES=96
