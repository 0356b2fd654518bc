#!/bin/bash
cd "$(dirname "$0")/.."
for i in 1 2 3; do timeout 300 python -m pytest tests/test_hvae_gpu.py -q -x -k "counterfactual_gradients or free_bits" 2>&1 | grep -E "passed|failed|Error|assert |rel|cos" | head -12; done
