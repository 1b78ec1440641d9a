#!/usr/bin/env bash
# run the bench once per library variant (one gpurun call)
for v in "" _h5 _h3 _r2 _r4 _pd; do
  L=$PWD/dendro-kt_b200/lib/libdkt$v.so
  [ -f "$L" ] || continue
  echo -n "variant '$v': "
  DKT_LIB=$L timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1
done
