"""Generates tests/golden/tree-*.npz from the REFERENCE itself (oracle/_ref): points in, the leaves of
SFC_Tree::distTreeConstruction and SFC_Tree::distTreeBalancing (src/tsort.cpp:647-716, 862-877) out, in the reference's
tree order.  Run in the build container only:    python tests/golden/make_tree_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "dendro-kt_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import dktref  # noqa: E402
import dkt  # noqa: E402

# name: (points, dim, max_depth, max_pts, sfc)
def cases():
    T = dkt.trees
    out = {}
    for sfc in ("morton", "hilbert"):
        out["tree-gauss-d2-%s" % sfc] = (T.gaussian_points(2, 400, 12, seed=11), 2, 12, 1, sfc)
        out["tree-gauss-d3-%s" % sfc] = (T.gaussian_points(3, 300, 9, seed=12), 3, 9, 1, sfc)
        out["tree-gauss-d4-%s" % sfc] = (T.gaussian_points(4, 120, 7, seed=13), 4, 7, 1, sfc)
    p = T.gaussian_points(3, 600, 8, seed=14)
    p[:40] = p[40:80]  # coincident points: regions that cannot be resolved stop at max_depth (src/tsort.cpp:630-639)
    out["tree-gauss-d3-maxpts4-dups"] = (p, 3, 8, 4, "morton")
    out["tree-gauss-d2-maxpts7"] = (T.gaussian_points(2, 500, 10, seed=15, sigma=0.1), 2, 10, 7, "morton")
    out["tree-shell-d4-guard"] = (T.shell_points(4, 300, 7, guard_level=2), 4, 7, 1, "morton")
    out["tree-shell-d3-guard"] = (T.shell_points(3, 500, 9, guard_level=3), 3, 9, 1, "hilbert")
    out["tree-corners-d3"] = (np.array([[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 255], [128, 128, 128], [127, 127, 127]], dtype=np.uint32),
                              3, 8, 1, "morton")
    out["tree-single-d4"] = (np.array([[5, 9, 3, 60]], dtype=np.uint32), 4, 6, 1, "morton")
    return out


def main():
    for name, (pts, dim, md, mp, sfc) in cases().items():
        R = dktref.Reference(dim, md, sfc)
        cx, cl = R.tree_from_points(pts, max_pts=mp, balance=False).export()
        bx, bl = R.tree_from_points(pts, max_pts=mp, balance=True).export()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), pts=pts, dim=dim, max_depth=md, max_pts=mp, sfc=sfc, construct_xyz=cx,
                            construct_lev=cl, balance_xyz=bx, balance_lev=bl)
        print("%-30s %6d points -> %7d leaves constructed, %7d balanced" % (name, len(pts), len(cl), len(bl)))


if __name__ == "__main__":
    main()
