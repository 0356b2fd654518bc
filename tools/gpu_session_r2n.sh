#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
echo "=== kernel tests (fold, likelihood)"; timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -s -k "conv or dgauss or dmol or cf_combine" 2>&1 | tail -40 > $O/r2n_kernel_tests.txt; grep "^fold\[" $O/r2n_kernel_tests.txt; tail -5 $O/r2n_kernel_tests.txt
if ! tail -3 $O/r2n_kernel_tests.txt | grep -q "passed" || tail -3 $O/r2n_kernel_tests.txt | grep -q "failed"; then echo "KERNEL TESTS FAILED"; exit 1; fi
echo "=== microbench B=128 folded"; MB_N=128 timeout 300 python tools/conv_microbench.py 20 > $O/r2n_microbench_b128_fold.txt 2>&1; cat $O/r2n_microbench_b128_fold.txt
echo "=== microbench B=128 nine-tap"; CAUSALGEN_B200_FOLD=0 MB_N=128 timeout 300 python tools/conv_microbench.py 20 > $O/r2n_microbench_b128_nofold.txt 2>&1; awk '{print $1,$2,$3,$4,$5,$6,$7}' $O/r2n_microbench_b128_nofold.txt
echo "=== tests"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^trainer\[\|^graph==\|^test  \|^elbo\[\|^pixels\|^nccl\|^cf-grad\|^fold\[" | tail -40 > $O/r2n_pytest_gpu.txt; tail -6 $O/r2n_pytest_gpu.txt
echo "=== bench quick"; timeout 300 python bench.py --no-configs --no-ref-gpu --no-cpu > $O/r2n_bench.json 2> $O/r2n_bench.err; python -c "
import json; d=json.load(open('$O/r2n_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['cf_inference']['value'], d['reference_batch32'])"; tail -3 $O/r2n_bench.err
