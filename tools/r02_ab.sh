#!/usr/bin/env bash
# A/B timing of library variants on the bench tree (one B200): tools/r02_ab.sh OUTDIR [variant names...]; "" = libdkt.so
O=gpurun_out/$1; shift
mkdir -p $O
for v in "$@"; do
  L=$PWD/dendro-kt_b200/lib/libdkt${v:+_$v}.so
  [ -f "$L" ] || { echo "no $L"; continue; }
  DKT_LIB=$L timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_${v:-base}.json 2> $O/bench_${v:-base}.err
  echo "variant '${v:-base}': $(grep -o '"ms_per_step": [0-9.]*' $O/bench_${v:-base}.json | head -1) $(grep -o '"frac": [0-9.]*' $O/bench_${v:-base}.json)"
done
if [ -n "$NCU_VARIANT" ] || [ -n "$NCU" ]; then
  L=$PWD/dendro-kt_b200/lib/libdkt${NCU_VARIANT:+_$NCU_VARIANT}.so
  DKT_LIB=$L timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mvf -s 6 -c 1 -f -o $O/prof_mvf \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline > $O/ncu_full.log 2>&1
fi
