#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for i in 1 2; do
echo "=== old lib (r3a code), narrow"; CAUSALGEN_B200_LIB=causal-gen_b200/causalgen_b200/libcausalgen_b200_old.so MB_N=128 MB_NOWGRAD=1 timeout 200 python tools/conv_microbench.py 20 r96 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7}'
echo "=== new lib, CG_CONV_WIDE=0"; CG_CONV_WIDE=0 MB_N=128 MB_NOWGRAD=1 timeout 200 python tools/conv_microbench.py 20 r96 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7}'
echo "=== new lib, wide"; MB_N=128 MB_NOWGRAD=1 timeout 200 python tools/conv_microbench.py 20 r96 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7}'
done > $O/r3d_ab.txt 2>&1
cat $O/r3d_ab.txt
echo "=== bench old lib"; CAUSALGEN_B200_LIB=causal-gen_b200/causalgen_b200/libcausalgen_b200_old.so timeout 300 python bench.py --no-configs --no-ref-gpu --no-cpu --no-cf --no-ref-batch > $O/r3d_bench_old.json 2> $O/r3d_bench.err; python -c "
import json; d=json.load(open('$O/r3d_bench_old.json')); print(d['value'], d['ms_per_step'])"
echo "=== bench new lib wide"; timeout 300 python bench.py --no-configs --no-ref-gpu --no-cpu --no-cf --no-ref-batch > $O/r3d_bench_new.json 2> $O/r3d_bench.err; python -c "
import json; d=json.load(open('$O/r3d_bench_new.json')); print(d['value'], d['ms_per_step'])"
