#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
echo "=== tests"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -v "^trainer\[\|^graph==\|^test  \|^elbo\[\|^pixels\|^nccl\|^cf-grad\|^fold\[\|^freebits\|^predictor" | tail -30 > $O/r2u_pytest_gpu.txt; tail -6 $O/r2u_pytest_gpu.txt
echo "=== microbench folded (two producers)"; CAUSALGEN_B200_FOLD=1 MB_N=128 MB_NOWGRAD=1 timeout 200 python tools/conv_microbench.py 20 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7,$8}' > $O/r2u_microbench_fold.txt 2>&1; cat $O/r2u_microbench_fold.txt
