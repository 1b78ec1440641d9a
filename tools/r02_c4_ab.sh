#!/usr/bin/env bash
# order-2 per-element kernels: register-lean (default) against node records in registers (libdkt_nolean.so), C4 config + parity tests
O=gpurun_out/r02_c4ab; mkdir -p $O
for v in "" nolean ""; do
  L=$PWD/dendro-kt_b200/lib/libdkt${v:+_$v}.so
  DKT_LIB=$L timeout 300 python tools/bench_configs.py c4 > $O/c4_${v:-lean}.json 2>&1
  echo "variant '${v:-lean}': $(grep -o '"ms_per_matvec": [0-9.]*' $O/c4_${v:-lean}.json) $(grep -o '"frac_of_hbm_peak": [0-9.]*' $O/c4_${v:-lean}.json)"
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_d4p2.py -m gpu -q -x -k "p2 or order or d3 or factor or dof or d4p2" 2>&1 | tail -2
