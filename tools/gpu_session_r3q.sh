#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for sk in 0 2 4 8 6 10 12 14; do
echo "=== skip=$sk (2 no act | 4 no mma | 8 no epilogue stores)"
CG_DEBUG_SKIP=$sk MB_N=128 MB_NOWGRAD=1 timeout 120 python tools/conv_microbench.py 20 "r96" 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7,$8}' | grep -v "1x1"
CG_DEBUG_SKIP=$sk MB_N=128 MB_NOWGRAD=1 timeout 120 python tools/conv_microbench.py 20 "r192" 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7,$8}' | grep -v "^case"
done > $O/r3q_ablation.txt 2>&1
cat $O/r3q_ablation.txt
