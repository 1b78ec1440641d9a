"""CPU tests: the oracle restatement (oracle/flat.py) against the committed golden vectors taken
from the reference, and - when oracle/_ref is present - against the reference live."""
import os

import numpy as np
import pytest

import cases
import flat

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    parts = name.split("-")
    case = dict(name=name, dim=int(g["dim"]), order=int(g["order"]), max_depth=int(g["max_depth"]), sfc=parts[3],
                xyz=g["in_xyz"], lev=g["in_lev"], golden=g)
    return case


@pytest.mark.parametrize("name", cases.ALL_CASES + cases.D4P2_CASES)
def test_oracle_matches_golden(name):
    case = load_case(name)
    g = case["golden"]
    t = cases.oracle_tables_for(case)
    # tree order, node set, node ORDER and levels: bit-exact
    assert np.array_equal(t.elem_xyz, g["elem_xyz"]) and np.array_equal(t.elem_lev, g["elem_lev"])
    assert np.array_equal(t.node_xyz, g["node_xyz"])
    assert np.array_equal(t.node_lev, g["node_lev"])
    assert np.array_equal(t.bdy_ids, g["bdy"])
    # the matvec visits exactly as many elements as the reference calls eleOp on (phantoms included)
    assert len(t.mv_lev) == int(g["ncalls"])
    assert t.tree_class in "ABP"
    n = len(g["node_lev"])
    K = cases.dense_operator(case["dim"], case["order"])
    u = cases.input_vector(n)
    v = flat.matvec(t, u, K, alpha=float(g["alpha"]), scale=float(g["scale"]), ip0=g["ip0"], ip1=g["ip1"])
    assert np.abs(v - g["v_dense"]).max() <= 1e-13 * np.abs(g["v_dense"]).max()
    vd = flat.matvec(t, u, K, alpha=float(g["alpha"]), scale=float(g["scale"]), ip0=g["ip0"], ip1=g["ip1"], dirichlet=True)
    assert np.abs(vd - g["v_dense_diri"]).max() <= 1e-13 * np.abs(g["v_dense_diri"]).max()
    vi = flat.matvec(t, np.ones(n), ip0=g["ip0"], ip1=g["ip1"])
    assert np.abs(vi - g["v_id"]).max() <= 1e-13 * np.abs(g["v_id"]).max()


def test_reference_known_answer_example1_2d():
    """The vector the reference's own test pins (test/testMatvec.cpp:324-399, Example1<2> depth 3,
    identity operator, u = 1): interior nodes first, the 16 boundary nodes last, and (2,2) gets 5
    where a conservative operator would give 6 (quirk Q1); sum 104, not 112."""
    g = load_case("ex1-d2-p1-morton-3")["golden"]
    unit = 1 << (int(g["max_depth"]) - 3)
    got = [(int(x) // unit, int(y) // unit, int(round(v))) for (x, y), v in zip(g["node_xyz"], g["v_id"])]
    want = [(3, 3, 4), (2, 2, 5), (5, 3, 4), (6, 2, 5), (3, 5, 4), (2, 6, 5), (5, 5, 4), (6, 6, 5), (4, 2, 5), (4, 3, 4), (4, 4, 4),
            (4, 5, 4), (4, 6, 5), (2, 4, 5), (3, 4, 4), (5, 4, 4), (6, 4, 5), (0, 0, 1), (2, 0, 2), (0, 2, 2), (4, 0, 2), (6, 0, 2),
            (0, 4, 2), (0, 6, 2), (8, 0, 1), (8, 2, 2), (8, 4, 2), (8, 6, 2), (0, 8, 1), (2, 8, 2), (4, 8, 2), (6, 8, 2), (8, 8, 1)]
    assert got == want
    assert g["v_id"].sum() == 104.0


@pytest.mark.parametrize("dim,which,depth,order", [(2, 1, 3, 1), (3, 1, 3, 1), (4, 1, 3, 1), (2, 2, 2, 1), (3, 2, 2, 1), (4, 2, 2, 1),
                                                   (2, 3, 3, 1), (3, 3, 3, 1), (4, 3, 3, 1), (2, 1, 4, 2), (3, 3, 3, 2)])
def test_node_count_formulas(dim, which, depth, order):
    """Closed-form CG node counts of test/testAdaptiveExamples.h:42-45, 88-91, 128-141
    (the known-answer check of test/testCountCGNodes.cpp:142-175)."""
    xyz, lev = cases.example_tree(dim, which, depth, 10)
    t = flat.build_tables(xyz, lev, dim, order, 10)
    p = order
    if which == 1:
        want = (depth - 2) * ((4 * p - 1) ** dim - (2 * p - 1) ** dim) + (4 * p + 1) ** dim
    elif which == 2:
        want = (2 ** depth * p + 1) ** dim
    else:
        want = (2 ** depth * p + 1) ** dim
        for l in range(2, depth):
            want += ((2 ** l - 2) * p + 1) ** dim - ((2 ** (l + 1) - 4) * p + 1) ** dim
    assert len(t.node_lev) == want


def test_q1_sums():
    """Sum of A*1 with the identity operator: the Q1 deficits of SURVEY.md §8a."""
    for dim, which, want in [(2, 1, 104), (3, 1, 864), (4, 1, 7136), (2, 3, 200), (3, 3, 3504), (4, 3, 59936)]:
        xyz, lev = cases.example_tree(dim, which, 3, 8)
        t = flat.build_tables(xyz, lev, dim, 1, 8)
        v = flat.matvec(t, np.ones(len(t.node_lev)))
        assert abs(v.sum() - want) < 1e-9
        v2 = flat.matvec(t, np.ones(len(t.node_lev)), q1_mask=False)
        assert abs(v2.sum() - (1 << dim) * len(lev)) < 1e-9  # conservative without the mask


def test_oracle_against_live_reference():
    """When the reference library is present (build container and the GPU box alike), compare on
    a tree that is NOT among the fixtures."""
    import dktref
    if not dktref.available("morton"):
        pytest.skip("oracle/_ref not built")
    import dkt
    dim, md = 3, 11
    xyz, lev = dkt.trees.moving_ball_tree(dim, 5, md, radius=0.11)
    R = dktref.Reference(dim, md)
    tree = R.tree_from_elements(xyz, lev, sort=True)
    da = R.da(tree, 1)
    t = flat.build_tables(xyz, lev, dim, 1, md)
    nx, nl = da.nodes()
    assert np.array_equal(nx, t.node_xyz) and np.array_equal(nl, t.node_lev)
    re = R.refel(1)
    K = flat.laplace_kref(dim, 1)
    u = cases.input_vector(len(nl), seed=3)
    vr, _, ncalls = da.matvec(u, dktref.OP_DENSE, K, alpha=dim - 2.0)
    vo = flat.matvec(t, u, K, alpha=dim - 2.0, ip0=re["ip0"], ip1=re["ip1"])
    assert ncalls == len(t.mv_lev)
    assert np.abs(vr - vo).max() <= 1e-13 * np.abs(vr).max()


@pytest.mark.parametrize("fixture", ["heatmat-d3-p1-ball", "heatmat-d3-p1-ex3"])
def test_reference_heatmat_operator(fixture):
    """The reference's own example operator HeatEq::HeatMat<3> (FEM/examples/src/heatMat.cpp:46-139):
    its elemental operator is h^1 * K_ref on axis-aligned cells; the flat matvec with the probed K_ref and
    Dirichlet rows reproduces feMatrix::matVec of the reference, and K_ref is the unit-cube Laplacian."""
    import dkt
    g = dict(np.load(os.path.join(GOLDEN, fixture + ".npz")))
    t = flat.build_tables(g["in_xyz"], g["in_lev"], 3, 1, int(g["max_depth"]))
    assert np.array_equal(t.node_xyz, g["node_xyz"])
    assert abs(float(g["heat_alpha"]) - 1.0) < 1e-12
    u = cases.input_vector(len(t.node_lev))
    v = flat.matvec(t, u, g["heat_kref"], alpha=float(g["heat_alpha"]), ip0=g["ip0"], ip1=g["ip1"], dirichlet=True)
    assert np.abs(v - g["v_heat"]).max() <= 1e-13 * np.abs(g["v_heat"]).max()
    K = dkt.operators.laplace_kref(3, 1)
    assert np.abs(K - g["heat_kref"]).max() <= 1e-13 * np.abs(K).max()
