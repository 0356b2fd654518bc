#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
echo "=== tests"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^trainer\[\|^graph==\|^test  \|^elbo\[\|^pixels\|^nccl" | tail -60 > $O/r2d_pytest_gpu.txt; tail -8 $O/r2d_pytest_gpu.txt; grep "cf-grad" $O/parity_report.txt
echo "=== timeline steady state"; CG_TL_FIRST=30 CAUSALGEN_B200_LIB=causal-gen_b200/causalgen_b200/libcausalgen_b200_tl.so timeout 600 python tools/timeline.py "r96" 128 2>&1 | grep -v "^  wgrad" > $O/r2d_timeline_ss.txt
CG_TL_FIRST=30 CAUSALGEN_B200_LIB=causal-gen_b200/causalgen_b200/libcausalgen_b200_tl.so timeout 600 python tools/timeline.py "r192" 128 2>&1 | grep -v "^  wgrad" >> $O/r2d_timeline_ss.txt
CG_TL_FIRST=8 CAUSALGEN_B200_LIB=causal-gen_b200/causalgen_b200/libcausalgen_b200_tl.so timeout 600 python tools/timeline.py "r48" 128 2>&1 | grep -v "^  wgrad" >> $O/r2d_timeline_ss.txt
cat $O/r2d_timeline_ss.txt
echo "=== graph trace"; timeout 600 python tools/graph_trace.py ukbb192 128 $O/r2d_graph_trace.json 2>&1 | tail -40
