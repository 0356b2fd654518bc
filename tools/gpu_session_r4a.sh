#!/bin/bash
# round 2, session 4a: lane stream priority A/B, fold policy A/B (policy / everywhere / off), GPU test suite
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
echo "=== hang finder"; timeout 120 python tools/find_hang.py ukbb192 1 > $O/r4a_find_hang.txt 2>&1; tail -1 $O/r4a_find_hang.txt
if ! grep -q "ALL LAUNCHES COMPLETED" $O/r4a_find_hang.txt; then tail -5 $O/r4a_find_hang.txt; echo "HANG/ERROR"; exit 1; fi
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --no-configs --no-ref-gpu --no-cpu --no-cf > $O/r4a_bench_$name.json 2> $O/r4a_bench.err
  python -c "
import json; d=json.load(open('$O/r4a_bench_$name.json')); print('$name', round(d['value'],1), round(d['ms_per_step'],3), round(d['reference_batch32']['value'],1), d['loss']['elbo'])"
}
for i in 1 2; do
  run prio0_$i CAUSALGEN_B200_PRIO=0
  run prio1_$i CAUSALGEN_B200_PRIO=1
done
run fold0 CAUSALGEN_B200_FOLD=0
run fold2 CAUSALGEN_B200_FOLD=2
run fold1 CAUSALGEN_B200_FOLD=1
run prio1_sides1 CAUSALGEN_B200_PRIO=1 CAUSALGEN_B200_SIDE_STREAMS=1
run prio1_sides4 CAUSALGEN_B200_PRIO=1 CAUSALGEN_B200_SIDE_STREAMS=4
echo "=== tests"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -v "^trainer\[\|^graph==\|^test  \|^elbo\[\|^pixels\|^nccl\|^cf-grad\|^fold\[\|^freebits\|^predictor\|^submodules" | tail -12 > $O/r4a_pytest_gpu.txt; tail -4 $O/r4a_pytest_gpu.txt
