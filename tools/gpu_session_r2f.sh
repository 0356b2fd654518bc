#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
echo "=== umma_rate2"; timeout 120 tools/micro/umma_rate2 > $O/r2f_umma_rate2.txt 2>&1; cat $O/r2f_umma_rate2.txt
echo "=== tests"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v "^trainer\[\|^graph==\|^test  \|^elbo\[\|^pixels\|^nccl" | tail -60 > $O/r2f_pytest_gpu.txt; tail -12 $O/r2f_pytest_gpu.txt; grep "cf-grad" $O/parity_report.txt
