#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
MASTER_ADDR=127.0.0.1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 tests/_ddp_worker.py > $O/r2y_ddp.out 2> $O/r2y_ddp.err
echo rc=$?; cat $O/r2y_ddp.out; grep -v "^\*\*\*\|OMP_NUM" $O/r2y_ddp.err | head -60
