#!/usr/bin/env bash
# Multi-GPU session (gpurun --gpus N): tools/r02_dist_ab.sh N [notest]
#   1. tests/test_gpu_dist.py (tests/dist_check.py under torchrun: partitioned matvec gathered to the single-rank order against the
#      reference's golden vectors; NCCL and peer-memory exchange)
#   2. weak-scaling bench with both exchanges.  Results in gpurun_out/r02_dist/.
N=${1:-2}
O=gpurun_out/r02_dist
mkdir -p $O
if [ -z "$2" ]; then
  timeout 1500 python -m pytest tests/test_gpu_dist.py -m gpu -q -x > $O/pytest_dist_n$N.log 2>&1; echo "rc=$?" >> $O/pytest_dist_n$N.log
  tail -n 3 $O/pytest_dist_n$N.log
fi
for p2p in 0 1; do
  tag=n${N}_p2p${p2p}
  DKT_DIST_P2P=$p2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$p2p \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_$tag.json 2> $O/bench_$tag.err
  echo "$tag: $(grep -o '"ms_per_step": [0-9.]*' $O/bench_$tag.json | head -1) $(grep -o '"value": [0-9.e+]*' $O/bench_$tag.json | head -1)"
done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
echo "n1: $(grep -o '"ms_per_step": [0-9.]*' $O/bench_n1.json | head -1)"
