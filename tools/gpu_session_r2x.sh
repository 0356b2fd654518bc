#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
nvidia-smi -L
echo "=== tests (2 GPUs visible: the NCCL test runs)"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -v "^trainer\[\|^graph==\|^test  \|^elbo\[\|^pixels\|^cf-grad\|^fold\[\|^freebits\|^predictor" | tail -30 > $O/r2x_pytest_gpu.txt; tail -5 $O/r2x_pytest_gpu.txt; grep "^nccl" $O/parity_report.txt
echo "=== ncu dgauss"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:dgauss --profile-from-start off -c 2 -o $O/r2x_dgauss -f python tools/ncu_targets.py lik > $O/r2x_ncu_dgauss.log 2>&1; tail -1 $O/r2x_ncu_dgauss.log
echo "=== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 4 --no-configs --no-ref-gpu --no-cpu > $O/r2x_bench_n2.json 2> $O/r2x_bench_n2.err; python -c "
import json; d=json.load(open('$O/r2x_bench_n2.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"; tail -2 $O/r2x_bench_n2.err
echo "=== bench N=2 batch 32"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --batch 32 --no-configs --no-ref-gpu --no-cpu > $O/r2x_bench_n2_b32.json 2> $O/r2x_bench_n2_b32.err; python -c "
import json; d=json.load(open('$O/r2x_bench_n2_b32.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"; tail -2 $O/r2x_bench_n2_b32.err
echo "=== bench N=1 batch 32"; timeout 300 python bench.py --steps 20 --warmup 5 --batch 32 --no-configs --no-ref-gpu --no-cpu > $O/r2x_bench_n1_b32.json 2> $O/r2x_bench_n1_b32.err; python -c "
import json; d=json.load(open('$O/r2x_bench_n1_b32.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"
