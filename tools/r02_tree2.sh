#!/usr/bin/env bash
O=gpurun_out/r02_tree2; mkdir -p $O
DKT_TREE_TIMING=1 timeout 600 python tools/tree_bench.py --points 1000000 --ref-points 2000 > $O/bench_1m.json 2> $O/bench_1m.err; cat $O/bench_1m.json; grep "dkt tree" $O/bench_1m.err | tail -12
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_tree.csv python tools/tree_bench.py --points 300000 --ref-points 1000 > $O/ncu.log 2>&1
python - <<'PY'
import csv, collections, re
rows = list(csv.reader(open("gpurun_out/r02_tree2/launches_tree.csv", errors="ignore")))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if "Kernel Name" in r:
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except Exception:
            continue
        u = d.get("Metric Unit", "")
        us = v / 1000.0 if u in ("ns", "nsecond") else v * 1000.0 if u in ("ms", "msecond") else v
        k = re.sub(r"<.*", "", d["Kernel Name"])[:60]
        agg[k][0] += 1; agg[k][1] += us
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print("%-60s %6d launches %12.1f us" % (k, n, us))
PY
