#!/usr/bin/env python
"""Writes profiles/traffic.json from an ncu CSV of one bench run taken with
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:'k_mvf|k_mv3' --csv --log-file <csv> python bench.py --steps 3 --warmup 1 --no-cpu-baseline
The kernels of the LAST matvec step are summed (one k_mvf + the per-element kernels of the singles); the output vector's
cudaMemsetAsync is not a kernel and is added as 8 bytes per node.  bench.py reports the sum as roofline.traffic when the kernel
sources and the workload are the ones of this capture.
usage: tools/ncu_traffic.py <csv> <bench json line file>"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
kn, mn, mv, mu, idc = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
launches = {}
for r in rows[h + 1:]:
    if len(r) <= mv:
        continue
    d = launches.setdefault(int(r[idc]), {"name": r[kn]})
    v = float(r[mv].replace(",", ""))
    unit = r[mu]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1.0)
    d[r[mn]] = v * scale
ids = sorted(launches)
# the last step: walk back from the end to the last k_mvf, take it and the k_mv3 launches around it up to the previous k_mvf
last = max(i for i in ids if "k_mvf" in launches[i]["name"])
prev = max([i for i in ids if "k_mvf" in launches[i]["name"] and i < last] or [-1])
step = [i for i in ids if prev < i <= last or (i > last and "k_mv3" in launches[i]["name"])]
step = [i for i in step if i == last or "k_mv3" in launches[i]["name"]][-3:] if len(step) > 3 else step
line = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
n_nodes, n_elem = line["config"]["n_nodes"], line["config"]["n_elem"]
kern = [{"kernel": launches[i]["name"][:60], "dram_read": launches[i].get("dram__bytes_read.sum", 0.0),
         "dram_write": launches[i].get("dram__bytes_write.sum", 0.0), "us": launches[i].get("gpu__time_duration.sum", 0.0)} for i in step]
total = sum(k["dram_read"] + k["dram_write"] for k in kern) + 8.0 * n_nodes
out = {"kernel_source_sha": bench.kernel_source_sha(), "n_elem": n_elem, "n_nodes": n_nodes, "dram_bytes_per_step": total,
       "memset_bytes_added": 8.0 * n_nodes, "kernels": kern, "alg_bytes_per_step": line["roofline"]["alg_bytes_per_step_per_gpu"],
       "how": "ncu dram__bytes_read.sum + dram__bytes_write.sum of the kernels of one matvec step + 8 B per node for the cudaMemsetAsync"}
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
