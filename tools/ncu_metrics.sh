#!/bin/bash
# usage: ncu_metrics.sh file.ncu-rep  -> key metrics per captured launch (run where ncu is installed; no GPU needed)
ncu -i "$1" --page raw --csv 2>/dev/null | python3 -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__block_size']
idx = [(w, hdr.index(w)) for w in want if w in hdr]
units = rows[1]
for r in rows[2:]:
    print('---')
    for w, i in idx:
        print('%-72s %s %s' % (w, r[i], units[i]))
"
