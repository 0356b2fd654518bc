#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for fold in 1 0; do for sk in 0 2 4 6 8 14; do
echo "=== fold=$fold skip=$sk (2 no act | 4 no mma | 8 no epilogue stores)"
CAUSALGEN_B200_FOLD=$fold CG_DEBUG_SKIP=$sk MB_N=128 MB_NOWGRAD=1 timeout 120 python tools/conv_microbench.py 20 "r96" 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7,$8}' | grep -v "1x1"
CAUSALGEN_B200_FOLD=$fold CG_DEBUG_SKIP=$sk MB_N=128 MB_NOWGRAD=1 timeout 120 python tools/conv_microbench.py 20 "r192" 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7,$8}' | grep -v "^case"
done; done > $O/r2o_ablation.txt 2>&1
cat $O/r2o_ablation.txt
