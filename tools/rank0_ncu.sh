#!/usr/bin/env bash
# torchrun helper: rank 0 runs under a single-pass ncu duration listing, the other ranks plainly (diagnostics; peer-memory exchange only)
if [ "$LOCAL_RANK" = "0" ]; then
  exec ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_mvf|k_mv3|k_p2p" -c 400 --csv --log-file gpurun_out/r02_dist/launches_rank0.csv python "$@"
else
  exec python "$@"
fi
