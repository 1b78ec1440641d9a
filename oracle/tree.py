"""TEST INFRASTRUCTURE ONLY (see oracle/README.md): a literal NumPy/Python restatement of the reference's tree construction
and 2:1 balancing from points, single rank.  Never imported by the product.

Follows, step by step (paths relative to the reference repository):
  * SFC_Tree::locTreeConstruction   src/tsort.cpp:566-644   top-down: the children of a region are examined, a child holding
                                                          more than maxPtsPerRegion seeds *of level >= its own* (coarser seeds
                                                          are the "ancestor" bucket of SFC_bucketing, src/tsort.cpp:25-41,
                                                          include/tsort.tcc:185-265, and drop out) is split while it is
                                                          coarser than m_uiMaxDepth, else appended as a leaf
  * SFC_Tree::distTreeConstruction  src/tsort.cpp:647-716   (one rank: construction, sort, duplicates removed)
  * SFC_Tree::propagateNeighbours   src/tsort.cpp:775-823   bottom-up over the levels: parent and the parent's neighbours
                                                          (TreeNode::appendAllNeighbours, include/treeNode.tcc:525-585, clipped
                                                          at the domain: getNeighbour1d :453-482) join the next coarser level
  * SFC_Tree::distTreeBalancing     src/tsort.cpp:862-877   construction, propagateNeighbours, construction with 1 seed per region
The output is the SET of leaves; the reference returns them in tree (SFC) order, which oracle/flat.py / the library's
element sort reproduce separately.  Pinned against oracle/_ref (the reference itself) by tests/test_tree_oracle.py and the
committed fixtures tests/golden/tree-*.npz."""
import itertools

import numpy as np


def _construct(seed_xyz, seed_lev, dim, max_depth, max_pts):
    """Leaves (xyz, lev) of locTreeConstruction over seeds given as anchors + levels."""
    seed_xyz = np.asarray(seed_xyz, dtype=np.int64).reshape(-1, dim)
    seed_lev = np.asarray(seed_lev, dtype=np.int64)
    out_xyz, out_lev = [], []
    if len(seed_lev) == 0:
        return np.zeros((0, dim), dtype=np.uint32), np.zeros(0, dtype=np.uint8)
    # explicit stack of (indices of the seeds inside the region with level >= sLev - 1, region anchor, sLev)
    stack = [(np.arange(len(seed_lev)), np.zeros(dim, dtype=np.int64), 1)]
    while stack:
        idx, anchor, slev = stack.pop()
        idx = idx[seed_lev[idx] >= slev]  # the others are the ancestor bucket
        sh = max_depth - slev
        child = np.zeros(len(idx), dtype=np.int64)
        for d in range(dim):
            child |= ((seed_xyz[idx, d] >> sh) & 1) << d
        for c in range(1 << dim):
            ca = anchor + np.array([((c >> d) & 1) << sh for d in range(dim)], dtype=np.int64)
            if slev < max_depth:
                sub = idx[child == c]
                if len(sub) > max_pts:
                    stack.append((sub, ca, slev + 1))
                    continue
            out_xyz.append(ca)
            out_lev.append(slev)
    return np.array(out_xyz, dtype=np.uint32).reshape(-1, dim), np.array(out_lev, dtype=np.uint8)


def construct(points, dim, max_depth, max_pts=1):
    """distTreeConstruction of points (integer coordinates in [0, 2^max_depth), level max_depth each)."""
    points = np.asarray(points, dtype=np.int64).reshape(-1, dim)
    return _construct(points, np.full(len(points), max_depth, dtype=np.int64), dim, max_depth, max_pts)


def propagate_neighbours(xyz, lev, dim, max_depth):
    """The seed set of propagateNeighbours as (xyz, lev), every level without duplicates."""
    levels = [set() for _ in range(max_depth + 1)]
    for x, l in zip(np.asarray(xyz, dtype=np.int64).reshape(-1, dim), lev):
        levels[int(l)].add(tuple(int(v) for v in x))
    offs = [o for o in itertools.product((0, 1, -1), repeat=dim) if any(o)]
    size = 1 << max_depth
    for l in range(max_depth, 0, -1):
        lp = l - 1
        plen = 1 << (max_depth - lp)
        parents = {tuple(v & ~(plen - 1) for v in x) for x in levels[l]}
        for p in parents:
            levels[lp].add(p)
            for o in offs:
                q = tuple(p[d] + o[d] * plen for d in range(dim))
                if all(0 <= v < size for v in q):  # getNeighbour1d without includeDomBdry
                    levels[lp].add(q)
    sx = [x for l in range(max_depth + 1) for x in sorted(levels[l])]
    sl = [l for l in range(max_depth + 1) for _ in levels[l]]
    return np.array(sx, dtype=np.int64).reshape(-1, dim), np.array(sl, dtype=np.int64)


def balance(points, dim, max_depth, max_pts=1):
    """distTreeBalancing: leaves of the 2:1-balanced tree of the points."""
    x0, l0 = construct(points, dim, max_depth, max_pts)
    sx, sl = propagate_neighbours(x0, l0, dim, max_depth)
    return _construct(sx, sl, dim, max_depth, 1)


def canonical(xyz, lev):
    """Rows (lev, x...) sorted lexicographically: compares leaf SETS."""
    a = np.concatenate([np.asarray(lev, dtype=np.int64)[:, None], np.asarray(xyz, dtype=np.int64)], axis=1)
    return a[np.lexsort(a.T[::-1])]
