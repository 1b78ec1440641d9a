#!/usr/bin/env bash
# Drop-in check (TEST INFRASTRUCTURE): the reference's own example operators FEM/examples/src/heatMat.cpp and heatVec.cpp are
# compiled UNMODIFIED, where they lie under /root/reference, against THIS repository's host headers (dendro-kt_b200/include)
# and linked with tests/cpp/heat_dropin_main.cpp + libdkt.so.   output: oracle/_ref/heat_dropin (git-ignored; travels to the
# GPU box like the other built files).  -O0 -fno-unreachable-traps: both sources have bool functions that fall off their end
# (SURVEY 8c item 3); gcc 13 plants a trap there at -O0 unless told not to, and optimises the epilogue away above -O0.
set -euo pipefail
REF=${DKT_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(dirname "$HERE")
OUT=$HERE/_ref
if [ ! -d "$REF/FEM/examples/src" ]; then
  echo "[build_dropin] $REF not present: keeping prebuilt oracle/_ref/heat_dropin (if any)"; exit 0
fi
mkdir -p "$OUT"
INC="-I$ROOT/dendro-kt_b200/include -I$REF/FEM/examples/include"
CXX=${CXX:-g++}
$CXX -std=c++14 -O0 -fno-unreachable-traps -w $INC -c "$REF/FEM/examples/src/heatMat.cpp" -o "$OUT/dropin_heatMat.o"
$CXX -std=c++14 -O0 -fno-unreachable-traps -w $INC -c "$REF/FEM/examples/src/heatVec.cpp" -o "$OUT/dropin_heatVec.o"
$CXX -std=c++14 -O1 -w $INC "$ROOT/tests/cpp/heat_dropin_main.cpp" "$OUT/dropin_heatMat.o" "$OUT/dropin_heatVec.o" -o "$OUT/heat_dropin" \
  -L"$ROOT/dendro-kt_b200/lib" -ldkt -Wl,-rpath,'$ORIGIN/../../dendro-kt_b200/lib'
rm -f "$OUT/dropin_heatMat.o" "$OUT/dropin_heatVec.o"
echo "[build_dropin] built $OUT/heat_dropin"
