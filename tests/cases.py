"""Seeded tree cases shared by the golden generator, the oracle tests and the GPU parity tests.
Every case is a pure function of its name: (dim, order, max_depth, sfc, xyz, lev)."""
import numpy as np

import flat


def example_tree(dim, which, depth, max_depth):
    """Analytic trees of test/testAdaptiveExamples.h (Example1: centre-refined :36-78, Example2:
    uniform :84-112, Example3: boundary fringe :119-171), restated with plain arrays."""
    nch = 1 << dim
    out = []

    def child(cell, c):
        xyz, lev = cell
        h = 1 << (max_depth - lev - 1)
        return (tuple(x + (h if (c >> d) & 1 else 0) for d, x in enumerate(xyz)), lev + 1)

    def touches(cell):
        xyz, lev = cell
        h = 1 << (max_depth - lev)
        full = 1 << max_depth
        return any(x == 0 or x + h == full for x in xyz)

    root = (tuple([0] * dim), 0)
    if which == 1:
        def corner(e, ch):
            if e[1] >= depth:
                out.append(e)
            else:
                for oc in range(nch):
                    if oc != ch:
                        out.append(child(e, oc))
                corner(child(e, ch), ch)
        for ch in range(nch):
            corner(child(root, ch), nch - 1 - ch)
    elif which == 2:
        def fill(e):
            if e[1] >= depth:
                out.append(e)
            else:
                for c in range(nch):
                    fill(child(e, c))
        fill(root)
    else:
        def sub(e):
            if e[1] >= depth:
                out.append(e)
            else:
                for c in range(nch):
                    f = child(e, c)
                    if touches(f):
                        sub(f)
                    else:
                        out.append(f)
        for c in range(nch):
            sub(child(root, c))
    xyz = np.array([e[0] for e in out], dtype=np.uint32)
    lev = np.array([e[1] for e in out], dtype=np.uint8)
    return xyz, lev


def make_case(name):
    """name: '<kind>-d<dim>-p<order>-<sfc>[-args]'."""
    import dkt
    parts = name.split("-")
    kind, dim, order, sfc = parts[0], int(parts[1][1:]), int(parts[2][1:]), parts[3]
    arg = int(parts[4]) if len(parts) > 4 else 0
    md = 12
    if kind in ("ex1", "ex2", "ex3"):
        xyz, lev = example_tree(dim, int(kind[2]), arg, md)
    elif kind == "ball":
        xyz, lev = dkt.trees.moving_ball_tree(dim, arg, md)
    elif kind == "uniform":
        xyz, lev = dkt.trees.uniform_tree(dim, arg, md)
    else:
        raise ValueError(name)
    # shuffle so that the SFC sort is exercised
    rng = np.random.default_rng(len(lev))
    perm = rng.permutation(len(lev))
    return dict(name=name, dim=dim, order=order, max_depth=md, sfc=sfc, xyz=xyz[perm], lev=lev[perm])


# trees whose elements come from the REFERENCE's own pipeline (points -> distTreeBalancing);
# they exist only as committed fixtures (tests/golden/*.npz), see make_golden.py
POINT_CLOUD_CASES = [
    "gauss-d2-p1-morton", "gauss-d2-p2-morton", "gauss-d3-p1-morton", "gauss-d3-p2-morton", "gauss-d4-p1-morton",
    "gaussguard-d3-p1-morton", "gaussguard-d4-p1-morton", "gauss-d3-p1-hilbert", "gauss-d4-p1-hilbert",
]
GENERATED_CASES = [
    "ex1-d2-p1-morton-3", "ex1-d2-p1-morton-6", "ex1-d3-p1-morton-4", "ex1-d4-p1-morton-3", "ex1-d2-p2-morton-4",
    "ex1-d3-p2-morton-3", "ex2-d2-p1-morton-3", "ex2-d3-p1-morton-3", "ex2-d4-p1-morton-2", "ex2-d3-p2-morton-2",
    "ex3-d2-p1-morton-4", "ex3-d3-p1-morton-3", "ex3-d4-p1-morton-3", "ex3-d2-p2-morton-3", "ex3-d3-p2-morton-3",
    "ball-d2-p1-morton-7", "ball-d3-p1-morton-6", "ball-d4-p1-morton-5", "ball-d3-p2-morton-5",
    "ex1-d2-p1-hilbert-5", "ex3-d3-p1-hilbert-3", "ex3-d4-p1-hilbert-3", "ex3-d3-p2-hilbert-3",
]
ALL_CASES = GENERATED_CASES + POINT_CLOUD_CASES
# 4-D order 2 (81 nodes per element): table construction is the generic one; the matvec runs on the loop-based flat
# kernels (dkt_matvec.cu k_mv_big).  Checked on the CPU (oracle + emulated pipeline); GPU tests opt-in until confirmed.
D4P2_CASES = ["ex1-d4-p2-morton-3", "ex1-d4-p2-hilbert-3"]


def dense_operator(dim, order, seed=5):
    """Random dense (non-symmetric) reference matrix: exercises every entry and the orientation."""
    N = (order + 1) ** dim
    return np.random.default_rng(seed).uniform(-1.0, 1.0, (N, N))


def input_vector(n, seed=99):
    return np.random.default_rng(seed).uniform(-1.0, 1.0, n)


def oracle_tables_for(case, tables=None):
    if tables is None:
        tables = sfc_tables_for(case)
    return flat.build_tables(case["xyz"], case["lev"], case["dim"], case["order"], case["max_depth"], tables)


def sfc_tables_for(case):
    """SFC tables for the oracle: Morton is trivial; Hilbert comes from the golden fixture (the
    reference's own tables) so that the oracle never depends on the product."""
    if case["sfc"] == "morton":
        return flat.morton_tables(case["dim"])
    g = case.get("golden")
    if g is not None and "rot_perm" in g:
        return flat.SfcTables(case["dim"], g["rot_perm"], g["rot_inv"], g["htab"])
    raise RuntimeError("hilbert case without golden tables")
