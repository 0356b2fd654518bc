#!/bin/bash
# round 2, session 4f: fp16-split stem accuracy, GPU suite, early tensor-map prefetch A/B (previous commit's library vs this one)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
echo "=== hang finder"; timeout 120 python tools/find_hang.py ukbb192 1 > $O/r4f_find_hang.txt 2>&1; tail -1 $O/r4f_find_hang.txt
echo "=== stem accuracy"; timeout 200 python tools/stem_accuracy.py 2>&1 | grep -v Warn | tee $O/r4f_stem_accuracy.txt | head -5
echo "=== tests"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -v "^trainer\[\|^graph==\|^test  \|^elbo\[\|^pixels\|^nccl\|^cf-grad\|^fold\[\|^freebits\|^predictor\|^submodules" | tail -25 > $O/r4f_pytest_gpu.txt; tail -6 $O/r4f_pytest_gpu.txt
cp $O/parity_report.txt $O/r4f_parity_report.txt
echo "=== timelines (small launches, batch 128)"
for c in "48->192 r6" "dgrad 1x1 192->192 r6" "128->32 r24"; do
  CAUSALGEN_B200_LIB=causal-gen_b200/causalgen_b200/libcausalgen_b200_tl.so timeout 120 python tools/timeline.py "$c" 128 2>&1 | grep -v Warn
done > $O/r4f_timeline_small.txt 2>&1; grep -A1 "^fwd\|^dgrad" $O/r4f_timeline_small.txt | cut -c1-200
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --no-configs --no-ref-gpu --no-cpu --no-cf > $O/r4f_bench_$name.json 2> $O/r4f_bench.err
  python -c "
import json; d=json.load(open('$O/r4f_bench_$name.json')); print('$name', round(d['value'],1), round(d['ms_per_step'],3), round(d['reference_batch32']['value'],1), d['loss']['elbo'])"
}
P=causal-gen_b200/causalgen_b200/libcausalgen_b200_prev.so
run prev1 CAUSALGEN_B200_LIB=$P
run new1
run prev2 CAUSALGEN_B200_LIB=$P
run new2
