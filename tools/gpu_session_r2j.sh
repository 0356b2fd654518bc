#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
echo "=== tests"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^trainer\[\|^graph==\|^test  \|^elbo\[\|^pixels\|^nccl\|^cf-grad" | tail -40 > $O/r2j_pytest_gpu.txt; tail -6 $O/r2j_pytest_gpu.txt
echo "=== microbench B=128"; MB_N=128 timeout 600 python tools/conv_microbench.py 20 > $O/r2j_microbench_b128.txt 2>&1; cat $O/r2j_microbench_b128.txt
echo "=== bench quick"; timeout 900 python bench.py --no-configs --no-ref-gpu --no-cpu > $O/r2j_bench.json 2> $O/r2j_bench.err; python -c "
import json; d=json.load(open('$O/r2j_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['cf_inference']['value'], d['reference_batch32'])"; tail -3 $O/r2j_bench.err
