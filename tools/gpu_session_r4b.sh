#!/bin/bash
# round 2, session 4b: per-tensor gradient deviation under the fold modes; weight-gradient pool width sweep
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "=== grad diag"; timeout 400 python tools/grad_diag.py ukbb192 0,1,2 > $O/r4b_grad_diag.txt 2>&1; tail -32 $O/r4b_grad_diag.txt
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --no-configs --no-ref-gpu --no-cpu --no-cf > $O/r4b_bench_$name.json 2> $O/r4b_bench.err
  python -c "
import json; d=json.load(open('$O/r4b_bench_$name.json')); print('$name', round(d['value'],1), round(d['ms_per_step'],3), round(d['reference_batch32']['value'],1), d['loss']['elbo'])"
}
run sides2 CAUSALGEN_B200_SIDE_STREAMS=2
run sides3 CAUSALGEN_B200_SIDE_STREAMS=3
run sides4 CAUSALGEN_B200_SIDE_STREAMS=4
run sides6 CAUSALGEN_B200_SIDE_STREAMS=6
run sides8 CAUSALGEN_B200_SIDE_STREAMS=8
run sides2b CAUSALGEN_B200_SIDE_STREAMS=2
run sides4b CAUSALGEN_B200_SIDE_STREAMS=4
run prio1_sides6 CAUSALGEN_B200_PRIO=1 CAUSALGEN_B200_SIDE_STREAMS=6
run prio1_sides8 CAUSALGEN_B200_PRIO=1 CAUSALGEN_B200_SIDE_STREAMS=8
