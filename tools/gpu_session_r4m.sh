#!/bin/bash
# round 2, session 4m: which change moved the MIMIC-224 counterfactual throughput (1627 -> 1503 images/s)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 500 python tools/cf_ab.py mimic224 32 "" CAUSALGEN_B200_PRIO=0 CAUSALGEN_B200_STEM_MMA=0 CAUSALGEN_B200_FOLD=0 "CAUSALGEN_B200_PRIO=0,CAUSALGEN_B200_STEM_MMA=0,CAUSALGEN_B200_FOLD=0" "" 2>&1 | tee $O/r4m_cf_ab.txt
