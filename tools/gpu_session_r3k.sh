#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python tools/layer_table.py ukbb192 128 > $O/r3k_layer_table_b128.txt 2>&1; head -45 $O/r3k_layer_table_b128.txt
timeout 300 python tools/graph_trace.py ukbb192 128 $O/r3k_graph_trace.json 2>&1 | tail -12
