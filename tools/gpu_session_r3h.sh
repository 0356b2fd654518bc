#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
MB_N=128 MB_NOWGRAD=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 4 -c 1 -o $O/r3h_conv_fwd16_64 -f python tools/conv_microbench.py 2 "fwd 16->64" > $O/r3h_ncu1.log 2>&1; tail -1 $O/r3h_ncu1.log
MB_N=128 MB_NOWGRAD=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 4 -c 1 -o $O/r3h_conv_fwd64_16 -f python tools/conv_microbench.py 2 "fwd 64->16" > $O/r3h_ncu2.log 2>&1; tail -1 $O/r3h_ncu2.log
