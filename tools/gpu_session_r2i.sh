#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "=== umma_rate2 variants"; timeout 120 tools/micro/umma_rate2 > $O/r2i_umma_rate2.txt 2>&1; cat $O/r2i_umma_rate2.txt
