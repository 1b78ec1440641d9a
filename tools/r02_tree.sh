#!/usr/bin/env bash
# GPU session for the tree pipeline: parity tests + timing
O=gpurun_out/r02_tree; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tree.py tests/test_cpp_api.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest.log
timeout 600 python tools/tree_bench.py > $O/tree_bench.json 2> $O/tree_bench.err; echo "bench rc=$?"; cat $O/tree_bench.json; tail -3 $O/tree_bench.err
timeout 600 python tools/tree_bench.py --dim 3 --depth 16 --points 5000000 --ref-points 50000 > $O/tree_bench_d3.json 2>> $O/tree_bench.err; cat $O/tree_bench_d3.json
