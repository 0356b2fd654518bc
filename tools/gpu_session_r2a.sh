#!/bin/bash
# Round-2 session A: parity of the benchmarked path, the new bench line, baseline measurements for the kernel work.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r2a_gpu.txt 2>&1
echo "=== new tests"; timeout 900 python -m pytest tests/test_trainer_gpu.py tests/test_kernels_gpu.py -q -m gpu -x -k "trainer or normalise or mix_cf or optimizer or ema_decay" 2>&1 | tail -60 > $O/r2a_new_tests.txt; tail -5 $O/r2a_new_tests.txt
echo "=== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -150 > $O/r2a_pytest_gpu.txt; tail -5 $O/r2a_pytest_gpu.txt
echo "=== bench"; timeout 900 python bench.py > $O/r2a_bench.json 2> $O/r2a_bench.err; tail -c 600 $O/r2a_bench.json; tail -5 $O/r2a_bench.err
echo "=== microbench B=128"; MB_N=128 timeout 600 python tools/conv_microbench.py 20 > $O/r2a_microbench_b128.txt 2>&1; tail -25 $O/r2a_microbench_b128.txt
echo "=== layer table"; timeout 600 python tools/layer_table.py ukbb192 128 > $O/r2a_layer_table_b128.txt 2>&1; head -3 $O/r2a_layer_table_b128.txt
echo "=== timeline B=128"; CAUSALGEN_B200_LIB=causal-gen_b200/causalgen_b200/libcausalgen_b200_tl.so timeout 600 python tools/timeline.py "" 128 > $O/r2a_timeline_b128.txt 2>&1; tail -4 $O/r2a_timeline_b128.txt
echo "=== ncu launch list"; timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file $O/r2a_ncu_raw.csv python tools/profile_one_step.py ukbb192 128 > $O/r2a_ncu.log 2>&1; tail -2 $O/r2a_ncu.log
python tools/ncu_summary.py $O/r2a_ncu_raw.csv $O/r2a_ncu ukbb192 128 | head -30
