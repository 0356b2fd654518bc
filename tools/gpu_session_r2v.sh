#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
echo "=== hang finder"; timeout 120 python tools/find_hang.py ukbb192 1 > $O/r2v_find_hang.txt 2>&1; tail -1 $O/r2v_find_hang.txt
if ! grep -q "ALL LAUNCHES COMPLETED" $O/r2v_find_hang.txt; then echo "HANG persists"; exit 1; fi
echo "=== microbench"; MB_N=128 timeout 200 python tools/conv_microbench.py 20 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7,$8,$9,$10,$11}' > $O/r2v_microbench.txt 2>&1; cat $O/r2v_microbench.txt
echo "=== tests"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^trainer\[\|^graph==\|^test  \|^elbo\[\|^pixels\|^nccl\|^cf-grad\|^fold\[\|^freebits\|^predictor" | tail -30 > $O/r2v_pytest_gpu.txt; tail -6 $O/r2v_pytest_gpu.txt
echo "=== bench quick"; timeout 300 python bench.py --no-configs --no-ref-gpu --no-cpu > $O/r2v_bench.json 2> $O/r2v_bench.err; python -c "
import json; d=json.load(open('$O/r2v_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['cf_inference']['value'], d['reference_batch32'])"; tail -3 $O/r2v_bench.err
