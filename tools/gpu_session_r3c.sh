#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
echo "=== microbench wide (>= 6 steps in flight)"; MB_N=128 MB_NOWGRAD=1 timeout 200 python tools/conv_microbench.py 20 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7,$8}' > $O/r3c_microbench_wide.txt 2>&1; cat $O/r3c_microbench_wide.txt
echo "=== microbench 16x8 tiles"; CG_CONV_WIDE=0 MB_N=128 MB_NOWGRAD=1 timeout 200 python tools/conv_microbench.py 20 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7,$8}' > $O/r3c_microbench_narrow.txt 2>&1; cat $O/r3c_microbench_narrow.txt
for w in 1 0 1 0; do echo "=== bench quick CG_CONV_WIDE=$w"; CG_CONV_WIDE=$w timeout 300 python bench.py --no-configs --no-ref-gpu --no-cpu --no-cf > $O/r3c_bench_w$w.json 2> $O/r3c_bench.err; python -c "
import json; d=json.load(open('$O/r3c_bench_w$w.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['reference_batch32'])"; done
echo "=== tests"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^trainer\[\|^graph==\|^test  \|^elbo\[\|^pixels\|^nccl\|^cf-grad\|^fold\[\|^freebits\|^predictor\|^submodules" | tail -30 > $O/r3c_pytest_gpu.txt; tail -4 $O/r3c_pytest_gpu.txt
