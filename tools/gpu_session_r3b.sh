#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
echo "=== hang finder"; timeout 120 python tools/find_hang.py ukbb192 1 > $O/r3b_find_hang.txt 2>&1; tail -1 $O/r3b_find_hang.txt
if ! grep -q "ALL LAUNCHES COMPLETED" $O/r3b_find_hang.txt; then tail -5 $O/r3b_find_hang.txt; echo "HANG/ERROR"; exit 1; fi
echo "=== kernel tests"; timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "conv" 2>&1 | tail -12
echo "=== microbench wide"; MB_N=128 MB_NOWGRAD=1 timeout 200 python tools/conv_microbench.py 20 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7,$8}' > $O/r3b_microbench_wide.txt 2>&1; cat $O/r3b_microbench_wide.txt
echo "=== microbench 16x8 tiles"; CG_CONV_WIDE=0 MB_N=128 MB_NOWGRAD=1 timeout 200 python tools/conv_microbench.py 20 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7,$8}' > $O/r3b_microbench_narrow.txt 2>&1; cat $O/r3b_microbench_narrow.txt
echo "=== tests"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^trainer\[\|^graph==\|^test  \|^elbo\[\|^pixels\|^nccl\|^cf-grad\|^fold\[\|^freebits\|^predictor\|^submodules" | tail -30 > $O/r3b_pytest_gpu.txt; tail -8 $O/r3b_pytest_gpu.txt
echo "=== bench quick"; timeout 300 python bench.py --no-configs --no-ref-gpu --no-cpu > $O/r3b_bench.json 2> $O/r3b_bench.err; python -c "
import json; d=json.load(open('$O/r3b_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['cf_inference']['value'], d['reference_batch32'])"; tail -3 $O/r3b_bench.err
