#!/bin/bash
# round 2, session 4n: FOLD as a template parameter of the conv kernel -- GPU suite, bench line, counterfactual configs
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
echo "=== hang finder"; timeout 120 python tools/find_hang.py ukbb192 1 > $O/r4n_find_hang.txt 2>&1; tail -1 $O/r4n_find_hang.txt
if ! grep -q "ALL LAUNCHES COMPLETED" $O/r4n_find_hang.txt; then tail -5 $O/r4n_find_hang.txt; echo "HANG/ERROR"; exit 1; fi
echo "=== tests"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -v "^trainer\[\|^graph==\|^test  \|^elbo\[\|^pixels\|^nccl\|^cf-grad\|^fold\[\|^freebits\|^predictor\|^submodules" | tail -12 > $O/r4n_pytest_gpu.txt; tail -4 $O/r4n_pytest_gpu.txt
cp $O/parity_report.txt $O/r4n_parity_report.txt
echo "=== bench"; timeout 900 python bench.py --no-cpu --no-ref-gpu > $O/r4n_bench.json 2> $O/r4n_bench.err; python -c "
import json; d=json.load(open('$O/r4n_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cf_inference']['value'], d['reference_batch32']['value'], {k: round(v['value']) for k, v in d['configs'].items()})"; tail -2 $O/r4n_bench.err
echo "=== bench FOLD=0"; CAUSALGEN_B200_FOLD=0 timeout 300 python bench.py --no-configs --no-ref-gpu --no-cpu --no-cf > $O/r4n_bench_fold0.json 2>> $O/r4n_bench.err; python -c "
import json; d=json.load(open('$O/r4n_bench_fold0.json')); print('fold0', d['value'], d['ms_per_step'], d['reference_batch32']['value'])"
