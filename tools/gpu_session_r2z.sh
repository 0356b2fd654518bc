#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
echo "=== submodule tests"; timeout 600 python -m pytest tests/test_hvae_gpu.py -q -x -k "submodule" 2>&1 | tail -25
echo "=== tests"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -v "^trainer\[\|^graph==\|^test  \|^elbo\[\|^pixels\|^nccl\|^cf-grad\|^fold\[\|^freebits\|^predictor\|^submodules" | tail -12 > $O/r2z_pytest_gpu.txt; tail -5 $O/r2z_pytest_gpu.txt
