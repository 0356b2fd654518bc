#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
echo "=== hang finder"; timeout 120 python tools/find_hang.py ukbb192 1 > $O/r3n_find_hang.txt 2>&1; tail -1 $O/r3n_find_hang.txt
if ! grep -q "ALL LAUNCHES COMPLETED" $O/r3n_find_hang.txt; then tail -5 $O/r3n_find_hang.txt; echo "HANG/ERROR"; exit 1; fi
for i in 1 2; do
echo "=== plain waits"; MB_N=128 MB_NOWGRAD=1 timeout 200 python tools/conv_microbench.py 20 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7}'
echo "=== hint 2us"; CAUSALGEN_B200_LIB=causal-gen_b200/causalgen_b200/libcausalgen_b200_hint.so MB_N=128 MB_NOWGRAD=1 timeout 200 python tools/conv_microbench.py 20 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7}'
done > $O/r3n_ab.txt 2>&1
python - <<'P'
import re
txt=open('gpurun_out/r3n_ab.txt').read()
blocks=re.split(r'=== ',txt)[1:]
res={}
for b in blocks:
    lines=b.splitlines(); name=lines[0]
    for l in lines[1:]:
        m=re.match(r'(.*?)\s+([\d.]+)\s+(\d+)\s*\|',l)
        if m: res.setdefault(m.group(1).strip(),{}).setdefault(name,[]).append(float(m.group(2)))
names=["plain waits","hint 2us"]
print('%-28s'%'case',' | '.join('%-20s'%n for n in names))
for k,v in res.items():
    print('%-28s'%k,' | '.join('%-20s'%(' '.join('%.1f'%x for x in v.get(n,[]))) for n in names))
P
for l in "" hint "" hint; do echo "=== bench lib=$l"; if [ -n "$l" ]; then export CAUSALGEN_B200_LIB=causal-gen_b200/causalgen_b200/libcausalgen_b200_hint.so; else unset CAUSALGEN_B200_LIB; fi; timeout 300 python bench.py --no-configs --no-ref-gpu --no-cpu --no-cf --no-ref-batch > $O/r3n_bench_$l.json 2> $O/r3n_bench.err; python -c "
import json; d=json.load(open('$O/r3n_bench_$l.json')); print(d['value'], d['ms_per_step'])"; done
unset CAUSALGEN_B200_LIB
echo "=== tests"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
