#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
echo "=== hang finder"; timeout 120 python tools/find_hang.py ukbb192 1 > $O/r3a_find_hang.txt 2>&1; tail -1 $O/r3a_find_hang.txt
if ! grep -q "ALL LAUNCHES COMPLETED" $O/r3a_find_hang.txt; then tail -5 $O/r3a_find_hang.txt; echo "HANG/ERROR"; exit 1; fi
echo "=== tests"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^trainer\[\|^graph==\|^test  \|^elbo\[\|^pixels\|^nccl\|^cf-grad\|^fold\[\|^freebits\|^predictor\|^submodules" | tail -30 > $O/r3a_pytest_gpu.txt; tail -8 $O/r3a_pytest_gpu.txt
echo "=== microbench (8-granular storage)"; MB_N=128 timeout 200 python tools/conv_microbench.py 20 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7,$8,$9,$10,$11}' > $O/r3a_microbench.txt 2>&1; cat $O/r3a_microbench.txt
echo "=== microbench PAD16"; CAUSALGEN_B200_PAD16=1 MB_N=128 timeout 200 python tools/conv_microbench.py 20 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7,$8,$9,$10,$11}' > $O/r3a_microbench_pad16.txt 2>&1; cat $O/r3a_microbench_pad16.txt
echo "=== bench quick"; timeout 300 python bench.py --no-configs --no-ref-gpu --no-cpu > $O/r3a_bench.json 2> $O/r3a_bench.err; python -c "
import json; d=json.load(open('$O/r3a_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['issued_over_algorithmic'], d['cf_inference']['value'], d['reference_batch32'], d['hbm_gb_peak_train'])"; tail -3 $O/r3a_bench.err
echo "=== bench quick PAD16"; CAUSALGEN_B200_PAD16=1 timeout 300 python bench.py --no-configs --no-ref-gpu --no-cpu --no-cf > $O/r3a_bench_pad16.json 2> $O/r3a_bench_pad16.err; python -c "
import json; d=json.load(open('$O/r3a_bench_pad16.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['issued_over_algorithmic'])"
