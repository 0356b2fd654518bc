#!/bin/bash
# round 2, session 4c: bisect the gradient deviation of the mixed fold policy; stem mma.sync kernels (tests + A/B)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "=== grad diag"
timeout 600 python tools/grad_diag.py ukbb192 0 1 1@CG_NO_PDL=1 1@CAUSALGEN_B200_SIDE_STREAMS=0 "res=24" "res=96" "res=192;kmin=64" "res=24+96+192;kmin=64;dir=f" "res=24+96+192;kmin=64;dir=b" "res=48" "res=24+48" > $O/r4c_grad_diag.txt 2>&1; grep -v Warn $O/r4c_grad_diag.txt | tail -30
echo "=== stem tests"; timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "stem" 2>&1 | tail -5
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --no-configs --no-ref-gpu --no-cpu --no-cf --no-ref-batch > $O/r4c_bench_$name.json 2> $O/r4c_bench.err
  python -c "
import json; d=json.load(open('$O/r4c_bench_$name.json')); print('$name', round(d['value'],1), round(d['ms_per_step'],3), d['loss'])"
}
run stem_direct CAUSALGEN_B200_STEM_MMA=0 CAUSALGEN_B200_FOLD=2
run stem_mma CAUSALGEN_B200_STEM_MMA=1 CAUSALGEN_B200_FOLD=2
