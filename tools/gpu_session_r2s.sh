#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
echo "=== predictor tests"; timeout 600 python -m pytest tests/test_predictors_gpu.py -q -x 2>&1 | tail -25
