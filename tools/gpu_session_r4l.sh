#!/bin/bash
# round 2, final 2-GPU session: NCCL bucket parity test, weak scaling at 128 and 32 images per GPU (N=1 on the same box beside it)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
nvidia-smi -L
echo "=== NCCL world-2 parity"; timeout 600 python -m pytest tests/test_trainer_gpu.py -q -m gpu -k "nccl" 2>&1 | tail -3; grep "^nccl" $O/parity_report.txt | tee $O/r4l_nccl_world2_parity.txt
B="--no-configs --no-ref-gpu --no-cpu --no-cf --no-ref-batch"
echo "=== bench N=1"; timeout 300 python bench.py $B > $O/r4l_bench_n1_b128.json 2> $O/r4l_bench.err; python -c "
import json; d=json.load(open('$O/r4l_bench_n1_b128.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"
echo "=== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 4 $B > $O/r4l_bench_n2_b128.json 2> $O/r4l_bench_n2.err; python -c "
import json; d=json.load(open('$O/r4l_bench_n2_b128.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"; tail -2 $O/r4l_bench_n2.err
echo "=== bench N=1 batch 32"; timeout 300 python bench.py --steps 20 --warmup 5 --batch 32 $B > $O/r4l_bench_n1_b32.json 2> $O/r4l_bench.err; python -c "
import json; d=json.load(open('$O/r4l_bench_n1_b32.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"
echo "=== bench N=2 batch 32"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --batch 32 $B > $O/r4l_bench_n2_b32.json 2> $O/r4l_bench_n2_b32.err; python -c "
import json; d=json.load(open('$O/r4l_bench_n2_b32.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"; tail -2 $O/r4l_bench_n2_b32.err
