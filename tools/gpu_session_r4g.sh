#!/bin/bash
# round 2, session 4g: weight-gradient grid size under the priority regime (CG_WGRAD_MIN_TILES), pool width re-check, stem kernel timing
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "=== stem bench"; for m in 1 0; do CAUSALGEN_B200_STEM_MMA=$m timeout 120 python tools/stem_bench.py 2>&1 | grep stem_mma; done | tee $O/r4g_stem_bench.txt
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --no-configs --no-ref-gpu --no-cpu --no-cf > $O/r4g_bench_$name.json 2> $O/r4g_bench.err
  python -c "
import json; d=json.load(open('$O/r4g_bench_$name.json')); print('$name', round(d['value'],1), round(d['ms_per_step'],3), round(d['reference_batch32']['value'],1), d['loss']['elbo'])"
}
run wmt24
run wmt4 CG_WGRAD_MIN_TILES=4
run wmt8 CG_WGRAD_MIN_TILES=8
run wmt12 CG_WGRAD_MIN_TILES=12
run wmt48 CG_WGRAD_MIN_TILES=48
run wmt24b
run sides10 CAUSALGEN_B200_SIDE_STREAMS=10
run sides4 CAUSALGEN_B200_SIDE_STREAMS=4
