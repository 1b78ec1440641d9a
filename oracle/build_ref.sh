#!/usr/bin/env bash
# Builds the reference oracle: the UNMODIFIED reference sources under /root/reference compiled
# for a single rank, wrapped by oracle/ref_driver.cpp.  TEST INFRASTRUCTURE ONLY.
#
#   outputs: oracle/_ref/libdktref_morton.so, oracle/_ref/libdktref_hilbert.so   (git-ignored)
#
# The reference is read-only and has six functions that fall off the end of a non-void
# function (UB; g++ >= 8 at -O1+ turns that into a crash: SURVEY.md §8c item 3).  The sources
# are therefore copied to a scratch directory OUTSIDE the repository, the missing `return`
# statements are appended there (nothing else is touched), and only the shared objects come back.
set -euo pipefail
REF=${DKT_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
if [ ! -d "$REF/include" ]; then
  echo "[build_ref] $REF not present: keeping prebuilt oracle/_ref (if any)"; exit 0
fi
mkdir -p "$OUT"
W=$(mktemp -d /tmp/dktref.XXXXXX)
trap 'rm -rf "$W"' EXIT
for d in include src FEM test array; do cp -r "$REF/$d" "$W/$d"; done

# insert_after <file> <line> <expected text on that line> <text>
insert_after() {
  local f=$W/$1 ln=$2 expect=$3 text=$4
  if ! sed -n "${ln}p" "$f" | grep -qF -- "$expect"; then
    echo "[build_ref] patch site moved: $1:$ln does not contain '$expect'" >&2; exit 1
  fi
  sed -i "${ln}a\\$text" "$f"
}
# (file, after-line, sanity text, inserted statement) -- SURVEY.md Appendix A
insert_after include/nsort.tcc 119 "m_owner = other.m_owner;" "    return *this;"
insert_after include/octUtils.h 401 "delete[] varVal;" "  return 0;"
# two sites in each of heatMat.cpp / heatVec.cpp; patch the later one first so line numbers hold
insert_after FEM/examples/src/heatMat.cpp 138 "out[bdyIndex[i]]=0.0;" "    return true;"
insert_after FEM/examples/src/heatMat.cpp 127 "out[bdyIndex[i]]=0.0;" "    return true;"
insert_after FEM/examples/src/heatVec.cpp 103 "out[bdyIndex[i]]=0.0;" "    return true;"
insert_after FEM/examples/src/heatVec.cpp 92 "out[bdyIndex[i]]=0.0;" "    return true;"

CXX=${CXX:-g++}
SHIM=$HERE/shim   # build_variant switches it to shim_mp for the multi-process oracle
FLAGS_TAIL="-std=c++11 -O3 -DNDEBUG -w -fpermissive -fPIC -DWITH_BLAS_LAPACK -DUSE_64BIT_INDICES \
 -DSPLITTER_SELECTION_FIX -DNUM_NPES_THRESHOLD=2 -DDENDRO_VTU_ASCII \
 -I$W/include -I$W/FEM/include -I$W/array/include -I$W/FEM/examples/include -I$W/test"
SRCS="src/binUtils.cpp src/parUtils.cpp src/point.cpp src/profiler.cpp src/KDhcurvedata.cpp src/KDhcurvedata_DATA.cpp \
 src/tsort.cpp src/nsort.cpp src/treeNode.cpp src/oda.cpp FEM/src/tensor.cpp FEM/src/refel.cpp FEM/src/basis.cpp \
 FEM/examples/src/heatMat.cpp FEM/examples/src/heatVec.cpp"

build_variant() {
  local name=$1; shift
  local extra="$*"
  local FLAGS="-I$SHIM $FLAGS_TAIL"
  mkdir -p "$W/obj_$name"
  local objs=""
  if [ "$SHIM" = "$HERE/shim_mp" ]; then
    $CXX -std=c++11 -O2 -fPIC -I$SHIM -c "$HERE/shim_mp/mpi_mp.cpp" -o "$W/obj_$name/mpi_mp.o" &
    objs="$W/obj_$name/mpi_mp.o"
  fi
  for s in $SRCS; do
    o=$W/obj_$name/$(echo "$s" | tr '/' '_').o
    $CXX $FLAGS $extra -c "$W/$s" -o "$o" &
    objs="$objs $o"
  done
  $CXX $FLAGS $extra -c "$HERE/ref_driver.cpp" -o "$W/obj_$name/ref_driver.o" &
  $CXX $FLAGS -c "$HERE/shim/lapack_shim.cpp" -o "$W/obj_$name/lapack_shim.o" &
  wait
  $CXX -shared -Wl,--no-undefined -o "$OUT/libdktref_$name.so" $objs "$W/obj_$name/ref_driver.o" "$W/obj_$name/lapack_shim.o"
  echo "[build_ref] built $OUT/libdktref_$name.so"
}
build_variant morton
build_variant hilbert -DHILBERT_ORDERING
# the same library over the multi-process MPI stand-in (oracle/shim_mp): the reference on several ranks, SURVEY 8f N3
SHIM=$HERE/shim_mp
build_variant mp_morton
