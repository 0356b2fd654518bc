#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "=== bench (full default line)"; timeout 600 python bench.py > $O/r3m_bench.json 2> $O/r3m_bench.err; python -c "
import json; d=json.load(open('$O/r3m_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['cf_inference']['value'], d['reference_batch32'], {k: round(v['value']) for k, v in d['configs'].items()}, d['reference_gpu'].get('bf16_autocast'), d['cpu_baseline'])"; tail -3 $O/r3m_bench.err
echo "=== graph trace"; timeout 300 python tools/graph_trace.py ukbb192 128 $O/r3m_graph_trace.json 2>&1 | tail -25 > $O/r3m_graph_trace.txt; head -12 $O/r3m_graph_trace.txt
echo "=== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file $O/r3m_ncu_raw.csv python tools/profile_one_step.py ukbb192 128 > $O/r3m_ncu.log 2>&1; tail -2 $O/r3m_ncu.log
python tools/ncu_summary.py $O/r3m_ncu_raw.csv $O/r3m_ncu ukbb192 128 | head -14
echo "=== ncu full: conv"; MB_N=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 4 -c 1 -o $O/r3m_conv_fwd64_16 -f python tools/conv_microbench.py 2 "fwd 64->16" > $O/r3m_ncu_conv1.log 2>&1; tail -1 $O/r3m_ncu_conv1.log
MB_N=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 4 -c 1 -o $O/r3m_conv_fwd16_64 -f python tools/conv_microbench.py 2 "fwd 16->64" > $O/r3m_ncu_conv2.log 2>&1; tail -1 $O/r3m_ncu_conv2.log
ls -la $O/*.ncu-rep
echo "=== tests"; rm -f $O/parity_report.txt; timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -v "^trainer\[\|^graph==\|^test  \|^elbo\[\|^pixels\|^nccl\|^cf-grad\|^fold\[\|^freebits\|^predictor\|^submodules" | tail -12 > $O/r3m_pytest_gpu.txt; tail -4 $O/r3m_pytest_gpu.txt
