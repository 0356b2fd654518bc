#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
echo "=== umma_rate2"; timeout 120 tools/micro/umma_rate2 > $O/r2g_umma_rate2.txt 2>&1; cat $O/r2g_umma_rate2.txt
echo "=== kernel + model tests"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^trainer\[\|^graph==\|^test  \|^elbo\[\|^pixels\|^nccl" | tail -40 > $O/r2g_pytest_gpu.txt; tail -6 $O/r2g_pytest_gpu.txt; grep "cf-grad" $O/parity_report.txt
echo "=== microbench B=128 (tile commit)"; MB_N=128 timeout 600 python tools/conv_microbench.py 20 > $O/r2g_microbench_b128.txt 2>&1; cat $O/r2g_microbench_b128.txt
echo "=== microbench B=128 (stage commit)"; CG_STAGE_COMMIT=1 MB_N=128 timeout 600 python tools/conv_microbench.py 20 > $O/r2g_microbench_b128_stagecommit.txt 2>&1; awk '{print $1,$2,$3,$4,$5,$6,$7}' $O/r2g_microbench_b128_stagecommit.txt
echo "=== bench quick"; timeout 900 python bench.py --no-configs --no-ref-gpu --no-cpu > $O/r2g_bench.json 2> $O/r2g_bench.err; python -c "
import json; d=json.load(open('$O/r2g_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['cf_inference']['value'], d['reference_batch32'])"; tail -3 $O/r2g_bench.err
