"""Shared by the tree tests: fixtures of tests/golden/tree-*.npz (made by tests/golden/make_tree_golden.py from the reference)
and size-independent checks of a linear tree."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TREE_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "tree-*.npz")))
# fixtures of the matvec suite whose input tree is the reference's distTreeBalancing of dkt.trees.gaussian_points(.., seed=7)
POINT_CLOUD_CASES = ["gauss-d2-p1-morton", "gauss-d3-p1-morton", "gauss-d3-p1-hilbert", "gauss-d4-p1-morton", "gauss-d4-p1-hilbert",
                     "gaussguard-d3-p1-morton", "gaussguard-d4-p1-morton"]


def load_tree_case(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    return dict(name=name, pts=g["pts"], dim=int(g["dim"]), max_depth=int(g["max_depth"]), max_pts=int(g["max_pts"]), sfc=str(g["sfc"]),
                construct=(g["construct_xyz"], g["construct_lev"]), balance=(g["balance_xyz"], g["balance_lev"]))


def point_cloud_inputs(name):
    """(points, dim, max_depth, sfc, reference elem_xyz, elem_lev) of a gauss*/gaussguard* matvec fixture
    (tests/golden/make_golden.py::point_cloud_case)."""
    import dkt
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    parts = name.split("-")
    kind, dim, sfc = parts[0], int(parts[1][1:]), parts[3]
    n = {2: 300, 3: 300, 4: 150}[dim]
    pts = dkt.trees.gaussian_points(dim, n, 14, seed=7, guard_level=3 if kind == "gaussguard" else None)
    return pts, dim, 14, sfc, g["elem_xyz"], g["elem_lev"]


def morton(xyz, dim, depth):
    """Full-depth Morton keys (python ints would overflow numpy only above 64 bits: dim * depth <= 56 here)."""
    x = np.asarray(xyz, dtype=np.uint64)
    k = np.zeros(len(x), dtype=np.uint64)
    for l in range(depth):
        for d in range(dim):
            k |= ((x[:, d] >> np.uint64(l)) & np.uint64(1)) << np.uint64(l * dim + d)
    return k


def check_complete(xyz, lev, dim, depth):
    """The leaves tile the domain: disjoint key ranges that add up to the whole cube."""
    k = morton(xyz, dim, depth)
    size = np.uint64(1) << (np.uint64(dim) * (np.uint64(depth) - lev.astype(np.uint64)))
    o = np.argsort(k)
    k, size = k[o], size[o]
    assert k[0] == 0
    assert np.array_equal(k[1:], (k + size)[:-1]), "gaps or overlaps between consecutive leaves"
    assert int(k[-1]) + int(size[-1]) == 1 << (dim * depth)
    # anchors are aligned to the leaf size
    cell = np.uint64(1) << (np.uint64(depth) - lev.astype(np.uint64))
    assert not np.any(np.asarray(xyz, dtype=np.uint64) % cell[:, None])


def check_balanced(xyz, lev, dim, depth, sample=None, seed=0):
    """2:1 across faces, edges and corners: the leaf holding the centre of any same-size neighbour cell is at most one level coarser."""
    import itertools
    k = morton(xyz, dim, depth)
    o = np.argsort(k)
    ks, ls = k[o], lev[o].astype(np.int64)
    idx = np.arange(len(lev))
    if sample is not None and sample < len(lev):
        idx = np.random.default_rng(seed).choice(len(lev), sample, replace=False)
    x = np.asarray(xyz, dtype=np.int64)[idx]
    l = lev[idx].astype(np.int64)
    cell = 1 << (depth - l)
    for off in itertools.product((-1, 0, 1), repeat=dim):
        if not any(off):
            continue
        q = x + np.array(off)[None, :] * cell[:, None]
        ok = np.all((q >= 0) & (q < (1 << depth)), axis=1)
        kq = morton(q[ok].astype(np.uint64), dim, depth)
        j = np.searchsorted(ks, kq, side="right") - 1
        assert np.all(ls[j] >= l[ok] - 1), "a neighbour is more than one level coarser"


def check_construction(pts, xyz, lev, dim, depth, max_pts):
    """Every leaf coarser than max_depth holds at most max_pts points and every leaf's parent holds more (or is the root)."""
    kp = np.sort(morton(pts, dim, depth))
    k = morton(xyz, dim, depth)
    bits = (np.uint64(dim) * (np.uint64(depth) - lev.astype(np.uint64)))
    cnt = np.searchsorted(kp, k + (np.uint64(1) << bits)) - np.searchsorted(kp, k)
    assert np.all((cnt <= max_pts) | (lev == depth))
    pb = bits + np.uint64(dim)
    pk = (k >> pb) << pb
    pcnt = np.searchsorted(kp, pk + (np.uint64(1) << pb)) - np.searchsorted(kp, pk)
    assert np.all((pcnt > max_pts) | (lev == 1))
