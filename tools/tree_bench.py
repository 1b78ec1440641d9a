"""Times dkt_tree_from_points (GPU: Morton sort, split nodes, bottom-up balancing, leaves, tree order) on the SURVEY 8d C3
recipe (shell points + guard) and, on a bounded sample, the reference's distTreeBalancing on one host core (oracle/_ref).
    python tools/tree_bench.py [--dim 4] [--depth 12] [--points 2000000] [--guard 3]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dendro-kt_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=4)
    ap.add_argument("--depth", type=int, default=12)
    ap.add_argument("--points", type=int, default=2000000)
    ap.add_argument("--guard", type=int, default=3)
    ap.add_argument("--ref-points", type=int, default=20000)
    a = ap.parse_args()
    import torch
    import dkt
    pts = dkt.trees.shell_points(a.dim, a.points, a.depth, guard_level=a.guard)
    d = torch.from_numpy(pts.astype(np.int64)).to(torch.int32).cuda()
    out = dict(dim=a.dim, max_depth=a.depth, n_points=len(pts))
    for bal in (False, True):
        best = 1e30
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            t = dkt.Tree(d, a.dim, a.depth, 1, balance=bal)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
            n = t.n_elem
            fl = t.finest_level
            t.close()
        out["balance" if bal else "construct"] = dict(leaves=n, finest_level=fl, seconds=best, leaves_per_s=n / best)
    t = dkt.Tree(d, a.dim, a.depth, 1, balance=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    da = t.da(order=1)
    torch.cuda.synchronize()
    out["da"] = dict(seconds=time.perf_counter() - t0, n_nodes=da.n_nodes, tree_class=da.tree_class, n_hanging_elem=da.n_hanging)
    da.close()
    t.close()
    try:
        import dktref
        rp = dkt.trees.shell_points(a.dim, a.ref_points, a.depth, guard_level=a.guard)
        R = dktref.Reference(a.dim, a.depth)
        t0 = time.perf_counter()
        rt = R.tree_from_points(rp, max_pts=1, balance=True)
        dt = time.perf_counter() - t0
        out["reference_1core"] = dict(n_points=len(rp), leaves=len(rt), seconds=dt, leaves_per_s=len(rt) / dt)
        g = dkt.Tree(rp, a.dim, a.depth, 1, balance=True)
        gx, gl = g.export()
        rx, rl = rt.export()
        out["reference_1core"]["identical_to_gpu"] = bool(np.array_equal(gx, rx) and np.array_equal(gl, rl))
    except Exception as e:
        out["reference_1core"] = "unavailable: %s" % e
    print(json.dumps(out))


if __name__ == "__main__":
    main()
