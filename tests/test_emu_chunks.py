"""CPU tests of the chunked matvec's LOGIC: dendro-kt_b200/csrc/dkt_chunks.cu is compiled with g++ against
tests/emu/cuda_emu.h (threads of a block = fibers, see there) and its table construction + kernels are run on
the oracle's element->node tables, then compared with the golden vectors taken from the reference and with the
oracle.  This covers the sibling-family sets (the order-1 default) and the per-element sets (DKT_FAMILIES=0, order 2,
general operators) without a GPU; it says nothing about performance and is not a CPU path of the product."""
import numpy as np
import pytest

import cases
import emu_chunks
import flat
from test_oracle import load_case

ORDER1 = [c for c in cases.ALL_CASES if "-p1-" in c]
ORDER2 = ["ex1-d2-p2-morton-4", "ex3-d3-p2-morton-3", "gauss-d3-p2-morton"]
TOL = 1e-12  # BASELINE.json: output vectors within 1e-12 relative


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def _tables(name):
    case = load_case(name)
    return case, case["golden"], cases.oracle_tables_for(case)


@pytest.mark.parametrize("name", ORDER1 + ORDER2)
def test_emulated_element_sets_match_reference(name):
    """default tables (one unit per element), dense operator / Dirichlet / identity against the golden vectors"""
    case, g, t = _tables(name)
    n = len(g["node_lev"])
    K = cases.dense_operator(case["dim"], case["order"])
    u = cases.input_vector(n)
    kw = dict(alpha=float(g["alpha"]), scale=float(g["scale"]), ip0=g["ip0"], ip1=g["ip1"])
    big = len(t.mv_lev) > 50000
    v, sets = emu_chunks.matvec(t, u, case["max_depth"], kref=K, families=0, **kw)
    assert all(s[0] == 0 for s in sets)
    assert rel(v, g["v_dense"]) <= TOL
    if not big:
        v, _ = emu_chunks.matvec(t, u, case["max_depth"], kref=K, dirichlet=True, order=2, families=0, **kw)
        assert rel(v, g["v_dense_diri"]) <= TOL
        v, _ = emu_chunks.matvec(t, np.ones(n), case["max_depth"], ip0=g["ip0"], ip1=g["ip1"], order=1, families=0)
        assert rel(v, g["v_id"]) <= TOL


@pytest.mark.parametrize("name", ORDER1)
def test_emulated_family_sets_match_reference(name):
    """sibling-family tables + family kernel (the order-1 default): identity against the reference's golden vector,
    Walsh-Hadamard Laplacian (+ Dirichlet) against the oracle; every fiber order must give the same vector; a general
    dense operator on the same DA takes the per-element tables built on first use"""
    case, g, t = _tables(name)
    dim, md = case["dim"], case["max_depth"]
    n = len(g["node_lev"])
    big = len(t.mv_lev) > 50000
    v, sets = emu_chunks.matvec(t, np.ones(n), md, ip0=g["ip0"], ip1=g["ip1"])
    assert rel(v, g["v_id"]) <= TOL
    # every visited element is in exactly one set
    assert units_of(sets, dim) == len(t.mv_lev)
    K = flat.laplace_kref(dim, 1)
    u = cases.input_vector(n)
    for diri in ((False,) if big else (False, True)):
        ref = flat.matvec(t, u, Kref=K, alpha=dim - 2.0, scale=0.7, ip0=g["ip0"], ip1=g["ip1"], dirichlet=diri)
        for order in ((2,) if big else (0, 1, 2)):
            v, _ = emu_chunks.matvec(t, u, md, kref=K, alpha=dim - 2.0, scale=0.7, ip0=g["ip0"], ip1=g["ip1"], dirichlet=diri, order=order)
            assert rel(v, ref) <= TOL
    if not big:
        Kd = cases.dense_operator(dim, 1)
        v, _ = emu_chunks.matvec(t, u, md, kref=Kd, alpha=float(g["alpha"]), scale=float(g["scale"]), ip0=g["ip0"], ip1=g["ip1"])
        assert rel(v, g["v_dense"]) <= TOL


def units_of(sets, dim):
    return sum(s[3] * ((1 << dim) if s[0] == 2 else 1) for s in sets)


def test_family_sets_cover_most_of_an_adaptive_tree():
    """on the benchmark's tree family nearly all leaves sit in complete sibling families, hanging ones included"""
    import dkt
    for dim, level in ((3, 6), (4, 5)):
        xyz, lev = dkt.trees.moving_ball_tree(dim, level, 10)
        t = flat.build_tables(xyz, lev, dim, 1, 10)
        u = cases.input_vector(len(t.node_lev))
        K = flat.laplace_kref(dim, 1)
        v, sets = emu_chunks.matvec(t, u, 10, kref=K, alpha=dim - 2.0)
        assert rel(v, flat.matvec(t, u, Kref=K, alpha=dim - 2.0)) <= TOL
        in_families = sum(s[3] << dim for s in sets if s[0] == 2)
        assert in_families >= 0.9 * len(t.mv_lev)
        assert len(t.hang_idx) > 0.2 * len(t.mv_lev)


@pytest.mark.parametrize("name,families", [("ball-d3-p1-morton-6", 0), ("ball-d3-p1-morton-6", 1), ("ex3-d4-p1-hilbert-3", 1),
                                           ("gauss-d4-p1-morton", 1), ("ball-d2-p1-morton-7", 1), ("ball-d4-p1-morton-5", 1)])
def test_emulated_phased_sets(name, families):
    """partitioned-DA layout ([interior | boundary] lists, three phases, visit positions with an offset):
    the phases together must give the single-pass vector"""
    case, g, t = _tables(name)
    dim, md = case["dim"], case["max_depth"]
    n = len(g["node_lev"])
    K = flat.laplace_kref(dim, 1)
    u = cases.input_vector(n)
    ref = flat.matvec(t, u, Kref=K, alpha=dim - 2.0, scale=0.7, ip0=g["ip0"], ip1=g["ip1"], dirichlet=True)
    v, sets = emu_chunks.matvec(t, u, md, kref=K, alpha=dim - 2.0, scale=0.7, ip0=g["ip0"], ip1=g["ip1"], dirichlet=True,
                                families=families, order=2, phased_seed=11)
    assert rel(v, ref) <= TOL
    assert {s[8] for s in sets} == {0, 1}
    assert units_of(sets, dim) == len(t.mv_lev)


@pytest.mark.parametrize("seed", [3, 8, 15, 21])
def test_emulated_random_trees(seed):
    """random 2:1-balanced trees (refinement towards random spheres, some at a domain corner -> phantom elements),
    family / per-element tables, random phase layout / fiber order: the emulated chunk path against the oracle"""
    import dkt
    rng = np.random.default_rng(seed)
    dim = int(rng.choice([2, 3, 3, 4]))
    md = 8
    maxl = {2: 7, 3: 5, 4: 4}[dim]
    pts = rng.uniform(0.0, 1.0, (int(rng.integers(1, 4)), dim))
    if rng.random() < 0.5:
        pts[0] = rng.choice([0.0, 1.0], dim)
    rad = rng.uniform(0.0, 0.3, len(pts))

    def g(ctr):
        d = np.full(len(ctr), 1e9)
        for p, r in zip(pts, rad):
            d = np.minimum(d, np.abs(np.abs(ctr - p).max(axis=1) - r))
        return d

    xyz, lev = dkt.trees._refine(np, dim, md, int(rng.integers(1, 3)), maxl, g)
    t = flat.build_tables(xyz, lev, dim, 1, md)
    if t.tree_class == "U" or len(t.mv_lev) > 40000:
        pytest.skip("class-U or too large for the CPU suite")
    u = rng.uniform(-1, 1, len(t.node_lev))
    K = flat.laplace_kref(dim, 1) + 0.3 * flat.mass_kref(dim, 1)
    ref = flat.matvec(t, u, Kref=K, alpha=dim - 2.0, scale=0.9, dirichlet=True)
    for spec in (0, 1):
        v, _ = emu_chunks.matvec(t, u, md, kref=K, alpha=dim - 2.0, scale=0.9, dirichlet=True, families=spec, order=int(rng.integers(0, 3)),
                                 phased_seed=(None if rng.random() < 0.5 else int(rng.integers(0, 100))))
        assert rel(v, ref) <= TOL, spec


@pytest.mark.parametrize("name", ["ex1-d2-p2-morton-4", "ex3-d3-p2-morton-3", "gauss-d3-p2-morton", "ball-d3-p1-morton-6", "ex3-d4-p1-morton-3"])
def test_emulated_sum_factorised_operator(name):
    """DKT_OP_KRON (sum of Kronecker products of 1-D matrices, the form of HeatMat/HeatVec): axis passes in registers at order
    2, expanded to the dense matrix (and from there to the Walsh-Hadamard form / family kernel) at order 1 - against the oracle
    with the dense matrix the terms stand for"""
    import dkt
    case, g, t = _tables(name)
    dim, order, md = case["dim"], case["order"], case["max_depth"]
    n = len(g["node_lev"])
    u = cases.input_vector(n)
    terms = np.concatenate([dkt.operators.laplace_terms(dim, order), 0.3 * dkt.operators.mass_terms(dim, order)])
    K = np.zeros(((order + 1) ** dim,) * 2)
    for term in terms:
        T = np.ones((1, 1))
        for d in range(dim):
            T = np.kron(term[d].T, T)  # A[k, j]: in k -> out j  => matrix entry [j, k]
        K += T
    for diri in (False, True):
        ref = flat.matvec(t, u, Kref=K, alpha=1.25, scale=0.7, ip0=g["ip0"], ip1=g["ip1"], dirichlet=diri)
        v, _ = emu_chunks.matvec(t, u, md, kron=terms, alpha=1.25, scale=0.7, ip0=g["ip0"], ip1=g["ip1"], dirichlet=diri, order=2)
        assert rel(v, ref) <= TOL
