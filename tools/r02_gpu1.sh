#!/usr/bin/env bash
# Round-2 GPU session 1 (one B200): parity of the sibling-family path (the order-1 default), the opt-in paths that
# never ran under the driver (4-D order 2, peer-memory exchange in one process), A/B timing families vs per-element
# tables, compute-sanitizer, ncu launch list + full capture of k_mvf.  Everything lands in gpurun_out/r02a/.
O=gpurun_out/r02a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
DKT_TEST_P2P=1 DKT_TEST_D4P2=1 timeout 600 python -m pytest tests/test_zz_gpu_p2p_local.py tests/test_zz_gpu_d4p2.py -m gpu -q > $O/pytest_optin.log 2>&1
echo "optin rc=$?" >> $O/pytest_optin.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_fam.json 2> $O/bench_fam.err
timeout 300 python bench.py --families 0 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_elem.json 2> $O/bench_elem.err
# other configs (3-D uniform, 2-D)
timeout 300 python tools/bench_configs.py > $O/configs.txt 2>&1
# memcheck + racecheck of the family kernel on small trees
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ex3-d4-p1-morton-3 or ball-d3-p1-morton-6 or ball-d2" > $O/memcheck.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ex3-d4-p1-morton-3 or ball-d2" > $O/racecheck.log 2>&1
# launch list of the bench step, then the full capture of the family kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_fam.csv \
  python bench.py --steps 3 --warmup 1 --no-cpu-baseline > $O/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mvf -s 2 -c 2 -f -o $O/prof_mvf \
  python bench.py --steps 3 --warmup 1 --no-cpu-baseline > $O/ncu_full.log 2>&1
ls -la $O
tail -3 $O/pytest_gpu.log $O/pytest_optin.log
cat $O/bench_fam.json | head -c 3000
echo
cat $O/bench_elem.json | head -c 1500
