#!/usr/bin/env bash
# tools/build_variant.sh NAME "-DMACRO=V ..." : builds dendro-kt_b200/lib/libdkt_NAME.so with extra macros
# (A/B experiments in ONE gpurun call: DKT_LIB=.../libdkt_NAME.so python bench.py).  Only the translation unit named by
# DKT_VARIANT_UNIT (default dkt_family) is recompiled with the macros.
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
D=dendro-kt_b200
mkdir -p $D/lib/var_$NAME
for u in dkt_api dkt_build dkt_matvec dkt_chunks dkt_family dkt_dist dkt_solve dkt_tree; do
  if [ "$u" = "${DKT_VARIANT_UNIT:-dkt_family}" ] || [ ! -f $D/lib/$u.o ]; then
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -x cu -c $D/csrc/$u.cu -o $D/lib/var_$NAME/$u.o &
  else cp $D/lib/$u.o $D/lib/var_$NAME/$u.o; fi
done
wait
/usr/local/cuda/bin/nvcc -shared -o $D/lib/libdkt_$NAME.so $D/lib/var_$NAME/*.o $D/lib/dkt_sfc.o -gencode arch=compute_100a,code=sm_100a -ldl
rm -rf $D/lib/var_$NAME
echo built $D/lib/libdkt_$NAME.so
