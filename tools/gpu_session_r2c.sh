#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
echo "=== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^trainer\[\|^graph==\|^test  \|^elbo\[\|^pixels\|^nccl" | tail -80 > $O/r2c_pytest_gpu.txt; tail -8 $O/r2c_pytest_gpu.txt
echo "=== microbench B=128"; MB_N=128 timeout 600 python tools/conv_microbench.py 20 > $O/r2c_microbench_b128.txt 2>&1; cat $O/r2c_microbench_b128.txt
for sk in 16 18; do echo "=== microbench SKIP=$sk"; CG_DEBUG_SKIP=$sk MB_N=128 timeout 600 python tools/conv_microbench.py 20 > $O/r2c_microbench_b128_skip$sk.txt 2>&1; cat $O/r2c_microbench_b128_skip$sk.txt | awk '{print $1,$2,$3,$4,$5,$6,$7}' ; done
echo "=== bench quick"; timeout 900 python bench.py --no-configs --no-ref-gpu --no-cpu > $O/r2c_bench.json 2> $O/r2c_bench.err; python -c "
import json; d=json.load(open('$O/r2c_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['cf_inference']['value'], d['reference_batch32'])"; tail -3 $O/r2c_bench.err
