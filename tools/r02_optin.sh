O=gpurun_out/r02_optin; mkdir -p $O
DKT_TEST_D4P2=1 timeout 600 python -m pytest tests/test_zz_gpu_d4p2.py -m gpu -q > $O/d4p2.log 2>&1; tail -n 3 $O/d4p2.log
DKT_TEST_P2P=1 timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_zz_gpu_p2p_local.py -m gpu -q -x -k "0-2" > $O/p2p_memcheck.log 2>&1; grep -m 20 "Invalid\|at \|ERROR SUMMARY\|passed\|failed" $O/p2p_memcheck.log
