#!/bin/bash
# Multi-GPU A/B of the exchange variants and the sibling-group tables (run with N GPUs of one box):
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1500 -- 'bash tools/r02_dist_ab.sh 2'
# DKT_DIST_P2P=1: peer-memory exchange (dkt_dist.cu, k_p2p_*); 0: ncclSend/Recv groups.  Results in gpurun_out/r02_dist/.
set -u
N=${1:-2}
out=gpurun_out/r02_dist
mkdir -p $out
python __graft_entry__.py > $out/build.log 2>&1
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) "$@"; }
for p2p in 0 1; do
  for g in 0 2; do
    tag="n${N}_p2p${p2p}_g${g}"
    DKT_P2P_CHECK=1 DKT_DIST_P2P=$p2p DKT_GROUPS=$g run tests/dist_check.py > $out/check_$tag.log 2>&1
    echo "dist_check $tag rc=$?" | tee -a $out/summary.txt
    DKT_DIST_P2P=$p2p run bench.py --gpus $N --steps 20 --warmup 5 --groups $g --no-cpu-baseline > $out/bench_$tag.json 2> $out/bench_$tag.err
    echo "bench $tag rc=$?" | tee -a $out/summary.txt
    python - <<PY
import json
try:
    d = json.loads(open("$out/bench_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "ms", round(d["ms_per_step"], 4), "DOF/s %.3e" % d["value"])
except Exception as e:
    print("$tag", "no result", e)
PY
  done
done
