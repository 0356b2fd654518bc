#!/bin/bash
# Model-level parity tests, one pytest process per test group under a timeout.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for t in test_cpu_model_fails test_state_dict test_elbo_kl_and_gradients test_abduct_forward test_counterfactual_particles test_counterfactual_fused test_counterfactual_mimic224 test_counterfactual_graph test_mediator test_dmol_likelihood test_conditioning; do
  echo "=== $t"
  timeout 600 python -m pytest tests/test_hvae_gpu.py -q -m gpu -k "$t" 2>&1 | tail -${TAILN:-40}
done
