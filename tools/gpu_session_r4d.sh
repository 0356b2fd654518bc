#!/bin/bash
# round 2, session 4d: GPU suite on the new defaults / kernels, layer table, in-kernel timelines of the small launches, CG_CONV_MIN_TILES
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
echo "=== hang finder"; timeout 120 python tools/find_hang.py ukbb192 1 > $O/r4d_find_hang.txt 2>&1; tail -1 $O/r4d_find_hang.txt
echo "=== tests"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -v "^trainer\[\|^graph==\|^test  \|^elbo\[\|^pixels\|^nccl\|^cf-grad\|^fold\[\|^freebits\|^predictor\|^submodules" | tail -25 > $O/r4d_pytest_gpu.txt; tail -6 $O/r4d_pytest_gpu.txt
echo "=== layer table"; timeout 300 python tools/layer_table.py ukbb192 128 > $O/r4d_layer_table_b128.txt 2>&1; grep "total\|stem\|pool\|upsample" $O/r4d_layer_table_b128.txt
echo "=== timelines (small launches, batch 128)"
for c in "48->192 r6" "dgrad 1x1 192->192 r6" "48->160 r12" "128->32 r24" "32->128 r24"; do
  CAUSALGEN_B200_LIB=causal-gen_b200/causalgen_b200/libcausalgen_b200_tl.so timeout 120 python tools/timeline.py "$c" 128 2>&1 | grep -v Warn
done > $O/r4d_timeline_small.txt 2>&1; cat $O/r4d_timeline_small.txt | cut -c1-400
echo "=== small-launch microbench"; MB_N=128 MB_NOWGRAD=1 timeout 200 python tools/conv_microbench.py 20 r6 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7}'
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --no-configs --no-ref-gpu --no-cpu --no-cf > $O/r4d_bench_$name.json 2> $O/r4d_bench.err
  python -c "
import json; d=json.load(open('$O/r4d_bench_$name.json')); print('$name', round(d['value'],1), round(d['ms_per_step'],3), round(d['reference_batch32']['value'],1), d['loss']['elbo'])"
}
run default
run mintiles1 CG_CONV_MIN_TILES=1
run mintiles3 CG_CONV_MIN_TILES=3
run default2
