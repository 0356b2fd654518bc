#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
echo "=== kernel tests"; timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "conv or latent" 2>&1 | tail -5
echo "=== free bits"; timeout 600 python -m pytest tests/test_hvae_gpu.py -q -x -k "free_bits" 2>&1 | tail -12
export CAUSALGEN_B200_FOLD=0
echo "=== timeline (nine-tap)"; for sk in 0 14; do for c in "fwd 64->16 r96" "fwd 32->8 r192" "dgrad 16->64 r96 mul+add"; do echo "skip=$sk"; CG_DEBUG_SKIP=$sk CAUSALGEN_B200_LIB=causal-gen_b200/causalgen_b200/libcausalgen_b200_tl.so timeout 120 python tools/timeline.py "$c" 128 2>&1 | grep -v "^  wgrad"; done; done > $O/r2p_timeline.txt 2>&1; cat $O/r2p_timeline.txt
echo "=== microbench nine-tap, tile walk"; for sk in 0 14; do echo "skip=$sk"; CG_DEBUG_SKIP=$sk MB_N=128 timeout 200 python tools/conv_microbench.py 20 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7,$8}'; done > $O/r2p_microbench.txt 2>&1; cat $O/r2p_microbench.txt
