#!/usr/bin/env bash
# Final single-GPU evidence of the round: GPU test-suite, smoke, bench, launch list with DRAM bytes, full ncu capture of k_mvf.
O=gpurun_out/r02_final; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -n 3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; head -c 600 $O/bench.json; echo
timeout 900 python bench.py --impl reference --steps 10 --warmup 10 > $O/bench_ref.json 2> $O/bench_ref.err; head -c 400 $O/bench_ref.json; echo
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_mvf|k_mv3' --csv \
  --log-file $O/launches_step.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline > $O/bench_under_ncu.json 2> $O/ncu1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_all.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2> $O/ncu2.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mvf -s 6 -c 1 -f -o $O/prof_mvf \
  python bench.py --steps 3 --warmup 1 --no-cpu-baseline > $O/ncu_full.log 2>&1
timeout 400 python tools/bench_configs.py > $O/configs.txt 2>&1; cat $O/configs.txt
# one GPU at the per-GPU size of an 8x larger job (the N=8 bench tree on ONE device): 9.7e7 elements
timeout 600 python bench.py --per-gpu-elems 9.6e7 --no-cpu-baseline > $O/bench_large.json 2> $O/bench_large.err; head -c 700 $O/bench_large.json; echo
ls -la $O
