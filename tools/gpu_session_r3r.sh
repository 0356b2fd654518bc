#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.txt
echo "=== hang finder"; timeout 120 python tools/find_hang.py ukbb192 1 > $O/r3r_find_hang.txt 2>&1; tail -1 $O/r3r_find_hang.txt
if ! grep -q "ALL LAUNCHES COMPLETED" $O/r3r_find_hang.txt; then tail -5 $O/r3r_find_hang.txt; echo "HANG/ERROR"; exit 1; fi
echo "=== kernel tests"; timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -s -k "conv" 2>&1 | grep "fold\[\|passed\|failed\|Error" | head -12
for i in 1 2; do
echo "=== nine taps"; CAUSALGEN_B200_FOLD=0 MB_N=128 MB_NOWGRAD=1 timeout 200 python tools/conv_microbench.py 20 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7}'
echo "=== folded"; MB_N=128 MB_NOWGRAD=1 timeout 200 python tools/conv_microbench.py 20 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7}'
done > $O/r3r_ab.txt 2>&1
python - <<'P'
import re
txt=open('gpurun_out/r3r_ab.txt').read()
blocks=re.split(r'=== ',txt)[1:]
res={}
for b in blocks:
    lines=b.splitlines(); name=lines[0]
    for l in lines[1:]:
        m=re.match(r'(.*?)\s+([\d.]+)\s+(\d+)\s*\|',l)
        if m: res.setdefault(m.group(1).strip(),{}).setdefault(name,[]).append(float(m.group(2)))
names=["nine taps","folded"]
print('%-28s'%'case',' | '.join('%-20s'%n for n in names))
for k,v in res.items():
    print('%-28s'%k,' | '.join('%-20s'%(' '.join('%.1f'%x for x in v.get(n,[]))) for n in names))
P
for f in 0 1 0 1; do echo "=== bench FOLD=$f"; CAUSALGEN_B200_FOLD=$f timeout 300 python bench.py --no-configs --no-ref-gpu --no-cpu --no-cf > $O/r3r_bench_f$f.json 2> $O/r3r_bench.err; python -c "
import json; d=json.load(open('$O/r3r_bench_f$f.json')); print(d['value'], d['ms_per_step'], d['reference_batch32']['value'], d['loss'])"; done
echo "=== tests"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -v "^trainer\[\|^graph==\|^test  \|^elbo\[\|^pixels\|^nccl\|^cf-grad\|^fold\[\|^freebits\|^predictor\|^submodules" | tail -12 > $O/r3r_pytest_gpu.txt; tail -4 $O/r3r_pytest_gpu.txt
