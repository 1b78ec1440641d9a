#!/usr/bin/env python
"""Table sizes of the chunk sets for one tree and several DKT_GROUPS specs (analysis tool, CPU only): runs the table
construction of dkt_chunks.cu under the CUDA-on-CPU emulation (tests/emu) on the oracle's tables and prints slots,
chunk nodes and the DRAM bytes per element the tables imply (index stream + node records + node values + output).
    python tools/group_stats.py [level=6]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "dendro-kt_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import flat  # noqa: E402
import dkt.trees as T  # noqa: E402
import emu_chunks  # noqa: E402


def main():
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    dim, md = 4, 12
    xyz, lev = T.moving_ball_tree(dim, level, md)
    t = flat.build_tables(xyz, lev, dim, 1, md)
    ne, nn = len(t.mv_lev), len(t.node_lev)
    print("4-D moving-ball tree, max level %d: %d elements, %d nodes, %d hanging elements" % (level, ne, nn, len(t.hang_idx)))
    u = np.ones(nn)
    print("%-6s %8s %8s %8s %8s | %s" % ("spec", "slots/el", "nodes/el", "B/el", "maxnode", "sets (kind rows g units chunks upc)"))
    for spec in ("0", "2", "2,1", "3,2", "3"):
        v, sets = emu_chunks.matvec(t, u, md, groups=spec)
        slots = nodes = bytes_ = 0.0
        maxn = 0
        for kind, rows, g, units, chunks, upc, maxnloc, total, phase in sets:
            if kind == 0:
                spu, rec = rows * 16, 6
            else:
                L3 = 3 ** g
                spu = (L3 * 2 ** (dim - g) + (16 if rows == 2 else 0) + 1) & ~1
                rec = 4
            slots += units * spu
            nodes += total
            bytes_ += units * spu * 4 + total * (rec + 8)  # slot words + node records + gathered node values
            maxn = max(maxn, maxnloc)
        bytes_ += nn * 16  # zero + write (or accumulate) the output once
        print("%-6s %8.2f %8.2f %8.1f %8d | %s" % (spec, slots / ne, nodes / ne, bytes_ / ne, maxn,
                                                   " ".join("(%d %d %d %d %d %d)" % s[:6] for s in sets)))


if __name__ == "__main__":
    main()
