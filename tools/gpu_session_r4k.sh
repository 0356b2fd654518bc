#!/bin/bash
# round 2, final evidence session: ncu launch list (-> traffic), full bench line, graph trace, ncu --set full of the folded conv
# and the stem kernels, layer table, GPU suite with the parity report
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "=== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file $O/r4k_ncu_raw.csv python tools/profile_one_step.py ukbb192 128 > $O/r4k_ncu.log 2>&1; tail -2 $O/r4k_ncu.log
python tools/ncu_summary.py $O/r4k_ncu_raw.csv $O/r4k_ncu ukbb192 128 | head -16
cp $O/r2_traffic.json profiles/r2_traffic.json
echo "=== bench (full default line)"; timeout 900 python bench.py > $O/r4k_bench.json 2> $O/r4k_bench.err; python -c "
import json; d=json.load(open('$O/r4k_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('traffic'), d['cf_inference']['value'], d['reference_batch32']['value'], {k: round(v['value']) for k, v in d['configs'].items()}, d['reference_gpu'], d['cpu_baseline'], d['clocks'])"; tail -3 $O/r4k_bench.err
echo "=== graph trace"; timeout 300 python tools/graph_trace.py ukbb192 128 $O/r4k_graph_trace.json 2>&1 | tail -25 > $O/r4k_graph_trace.txt; head -12 $O/r4k_graph_trace.txt
echo "=== ncu full"; MB_N=128 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 4 -c 1 -o $O/r4k_conv_fwd64_16_folded -f python tools/conv_microbench.py 2 "fwd 64->16" > $O/r4k_ncu_conv1.log 2>&1; tail -1 $O/r4k_ncu_conv1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stem_fwd_mma -s 2 -c 1 -o $O/r4k_stem_fwd_mma -f python tools/stem_bench.py 128 192 2 > $O/r4k_ncu_stem1.log 2>&1; tail -1 $O/r4k_ncu_stem1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stem_wgrad_mma -s 2 -c 1 -o $O/r4k_stem_wgrad_mma -f python tools/stem_bench.py 128 192 2 > $O/r4k_ncu_stem2.log 2>&1; tail -1 $O/r4k_ncu_stem2.log
ls -la $O/r4k*.ncu-rep
echo "=== layer table"; timeout 300 python tools/layer_table.py ukbb192 128 > $O/r4k_layer_table_b128.txt 2>&1; head -3 $O/r4k_layer_table_b128.txt
echo "=== tests"; rm -f $O/parity_report.txt; timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -v "^trainer\[\|^graph==\|^test  \|^elbo\[\|^pixels\|^nccl\|^cf-grad\|^fold\[\|^freebits\|^predictor\|^submodules" | tail -12 > $O/r4k_pytest_gpu.txt; tail -4 $O/r4k_pytest_gpu.txt
cp $O/parity_report.txt $O/r4k_parity_report.txt
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
