#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
echo "=== umma_rate2"; timeout 120 tools/micro/umma_rate2 > $O/r2h_umma_rate2.txt 2>&1; cat $O/r2h_umma_rate2.txt
echo "=== tests"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v "^trainer\[\|^graph==\|^test  \|^elbo\[\|^pixels\|^nccl\|^cf-grad" | tail -40 > $O/r2h_pytest_gpu.txt; tail -6 $O/r2h_pytest_gpu.txt
echo "=== bench quick"; timeout 900 python bench.py --no-configs --no-ref-gpu --no-cpu > $O/r2h_bench.json 2> $O/r2h_bench.err; python -c "
import json; d=json.load(open('$O/r2h_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['cf_inference']['value'], d['reference_batch32'])"; tail -3 $O/r2h_bench.err
