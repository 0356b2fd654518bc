#!/bin/bash
# Runs each kernel test in its own process under a timeout so one hung kernel cannot hide the rest.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for t in test_mix_cf_and_layout_glue test_optimizer_tail test_pool_and_upsample test_latent \
         test_dgauss test_dmol test_stem test_conv_forward test_conv_segments test_conv_centre test_conv_dgrad test_conv_many test_wgrad_mma test_colsum test_ema_decay; do
  echo "=== $t"
  timeout 240 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "$t" -x 2>&1 | tail -25
done
