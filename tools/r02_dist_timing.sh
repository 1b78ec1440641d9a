O=gpurun_out/r02_dist; mkdir -p $O
DKT_DIST_P2P=1 timeout 600 python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 bash tools/rank0_ncu.sh bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_ncu0.json 2> $O/bench_ncu0.err
tail -n 3 $O/bench_ncu0.err
ls -la $O
