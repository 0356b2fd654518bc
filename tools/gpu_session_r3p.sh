#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
for i in 1 2; do
echo "=== fp32 adds"; CAUSALGEN_B200_LIB=causal-gen_b200/causalgen_b200/libcausalgen_b200_prev.so MB_N=128 MB_NOWGRAD=1 timeout 200 python tools/conv_microbench.py 20 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7}'
echo "=== packed adds"; MB_N=128 MB_NOWGRAD=1 timeout 200 python tools/conv_microbench.py 20 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7}'
done > $O/r3p_ab.txt 2>&1
python - <<'P'
import re
txt=open('gpurun_out/r3p_ab.txt').read()
blocks=re.split(r'=== ',txt)[1:]
res={}
for b in blocks:
    lines=b.splitlines(); name=lines[0]
    for l in lines[1:]:
        m=re.match(r'(.*?)\s+([\d.]+)\s+(\d+)\s*\|',l)
        if m: res.setdefault(m.group(1).strip(),{}).setdefault(name,[]).append(float(m.group(2)))
names=["fp32 adds","packed adds"]
print('%-28s'%'case',' | '.join('%-20s'%n for n in names))
for k,v in res.items():
    print('%-28s'%k,' | '.join('%-20s'%(' '.join('%.1f'%x for x in v.get(n,[]))) for n in names))
P
for l in prev "" prev ""; do echo "=== bench lib=$l"; if [ -n "$l" ]; then export CAUSALGEN_B200_LIB=causal-gen_b200/causalgen_b200/libcausalgen_b200_prev.so; else unset CAUSALGEN_B200_LIB; fi; timeout 300 python bench.py --no-configs --no-ref-gpu --no-cpu --no-cf --no-ref-batch > $O/r3p_bench_$l.json 2> $O/r3p_bench.err; python -c "
import json; d=json.load(open('$O/r3p_bench_$l.json')); print(d['value'], d['ms_per_step'], d['loss'])"; done
unset CAUSALGEN_B200_LIB
echo "=== tests"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -v "^trainer\[\|^graph==\|^test  \|^pixels\|^nccl\|^fold\[\|^predictor\|^submodules" | tail -40 > $O/r3p_pytest_gpu.txt; tail -4 $O/r3p_pytest_gpu.txt; grep "^elbo\[\|^cf-grad\|^freebits" $O/parity_report.txt | grep "grad\|block KL max" | head -24
