#!/bin/bash
# First GPU session of the sibling-group tables (DKT_GROUPS, dkt_chunks.cu k_mvg): parity, then A/B timing, then ncu.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/r02_groups_ab.sh'
# Everything lands in gpurun_out/r02_groups/.
set -u
out=gpurun_out/r02_groups
mkdir -p $out
python __graft_entry__.py > $out/build.log 2>&1
DKT_TEST_GROUPS=1 timeout 900 python -m pytest tests/test_zz_gpu_groups.py -x -q -m gpu > $out/pytest_groups.log 2>&1
echo "pytest groups rc=$?" | tee -a $out/summary.txt
DKT_TEST_D4P2=1 timeout 300 python -m pytest tests/test_zz_gpu_d4p2.py -x -q -m gpu > $out/pytest_d4p2.log 2>&1
echo "pytest d4p2 rc=$?" | tee -a $out/summary.txt
# the peer-memory exchange protocol with all ranks of a partition in one process on this one GPU
DKT_TEST_P2P=1 timeout 600 python -m pytest tests/test_zz_gpu_p2p_local.py -x -q -m gpu > $out/pytest_p2p_local.log 2>&1
echo "pytest p2p local rc=$?" | tee -a $out/summary.txt
# memory and shared-memory race checks of the group kernels on one small 4-D tree (the CPU emulation cannot see
# real races or misaligned accesses)
DKT_TEST_GROUPS=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_gpu_groups.py -x -q -m gpu \
  -k "ex3-d4-p1-morton-3" > $out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a $out/summary.txt
DKT_TEST_GROUPS=1 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_zz_gpu_groups.py -x -q -m gpu \
  -k "ex3-d4-p1-morton-3" > $out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a $out/summary.txt
for g in 0 2 3 2,1 3,2; do
  timeout 600 python bench.py --groups $g --steps 20 --warmup 5 --no-experimental --no-cpu-baseline > $out/bench_g$g.json 2> $out/bench_g$g.err
  echo "bench groups=$g rc=$?" | tee -a $out/summary.txt
done
# the sets of one matvec on several streams (DKT_MV_STREAMS): small sets fill the tails of the big ones
for st in 2 4; do
  DKT_MV_STREAMS=$st timeout 600 python bench.py --groups 2 --steps 20 --warmup 5 --no-experimental --no-cpu-baseline > $out/bench_g2_s$st.json 2> $out/bench_g2_s$st.err
  echo "bench groups=2 streams=$st rc=$?" | tee -a $out/summary.txt
done
DKT_MV_STREAMS=2 timeout 600 python bench.py --groups 0 --steps 20 --warmup 5 --no-experimental --no-cpu-baseline > $out/bench_g0_s2.json 2> $out/bench_g0_s2.err
# compile-time variants, built BEFORE the gpurun call in the container (they travel with the snapshot), e.g.
#   tools/build_variant.sh m3  -DDKT_GRP_MINB_REG=3 -DDKT_GRP_MINB_HANG=3                 # 168-register budget
#   tools/build_variant.sh t96 -DDKT_GRP_TPB=96 -DDKT_GRP_MINB_REG=3 -DDKT_GRP_MINB_HANG=3 # three CTAs of regular quads per SM
for L in dendro-kt_b200/lib/libdkt_*.so; do
  [ -f "$L" ] || continue
  v=$(basename $L .so)
  for g in 2 2,1; do
    DKT_LIB=$PWD/$L timeout 600 python bench.py --groups $g --steps 20 --warmup 5 --no-experimental --no-cpu-baseline > $out/bench_${v}_g$g.json 2> $out/bench_${v}_g$g.err
    echo "bench $v groups=$g rc=$?" | tee -a $out/summary.txt
    python - <<PY
import json
try:
    d = json.loads(open("$out/bench_${v}_g$g.json").read().strip().splitlines()[-1])
    print("$v groups=$g", "ms", round(d["ms_per_step"], 4), "DOF/s %.3e" % d["value"], "frac", round(d["roofline"]["frac"], 4))
except Exception as e:
    print("$v groups=$g", "no result", e)
PY
  done
done
# launch list (per-kernel device time) and one full capture of the group kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches_g2.csv \
  python bench.py --groups 2 --steps 3 --warmup 2 --no-experimental --no-cpu-baseline > $out/ncu_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_mvg -c 4 -o $out/mvg_g2 \
  python bench.py --groups 2 --steps 2 --warmup 1 --no-experimental --no-cpu-baseline > $out/ncu_full.log 2>&1
python tools/ncu_summary.py $out/mvg_g2.ncu-rep k_mvg > $out/ncu_mvg_g2.txt 2>&1
tail -n 3 $out/pytest_groups.log
for f in $out/bench_g*_s*.json; do python - <<PY
import json
try:
    d = json.loads(open("$f").read().strip().splitlines()[-1])
    print("$f", "ms", round(d["ms_per_step"], 4), "DOF/s %.3e" % d["value"], "frac", round(d["roofline"]["frac"], 4))
except Exception as e:
    print("$f", "no result", e)
PY
done
for g in 0 2 3 2,1 3,2; do python - <<PY
import json
try:
    d = json.loads(open("$out/bench_g$g.json").read().strip().splitlines()[-1])
    print("groups=$g", "ms", round(d["ms_per_step"], 4), "DOF/s %.3e" % d["value"], "frac", round(d["roofline"]["frac"], 4))
except Exception as e:
    print("groups=$g", "no result", e)
PY
done
