#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "=== failing tests"; timeout 900 python -m pytest tests/test_trainer_gpu.py -q -m gpu 2>&1 | grep -v "^trainer\[\|^graph==\|^test  " | tail -120 > $O/r2b_trainer_tests.txt; tail -5 $O/r2b_trainer_tests.txt
echo "=== tma rate"; timeout 300 tools/micro/tma_rate > $O/r2b_tma_rate.txt 2>&1; cat $O/r2b_tma_rate.txt
