#!/usr/bin/env bash
# ncu capture of the order-2 per-element kernels on the C4 configuration (3-D p=2 adaptive ball), dense and sum-factorised operator
O=gpurun_out/r02_c4; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mv3 -s 8 -c 2 -f -o $O/prof_c4 python tools/bench_configs.py c4 > $O/ncu_c4.log 2>&1
tail -3 $O/ncu_c4.log
python tools/ncu_summary.py $O/prof_c4.ncu-rep k_mv3 12 > $O/ncu_c4_summary.txt 2>&1; head -60 $O/ncu_c4_summary.txt
