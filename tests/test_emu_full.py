"""CPU tests of the PRODUCT's single-rank pipeline under the CUDA-on-CPU emulation (tests/emu): dkt_build.cu (tree order,
CG node set and order, element->node and hanging tables, phantom elements), dkt_chunks.cu (chunk tables, kernels) and
dkt_sfc.cpp compiled with g++ and driven like dkt_da_create / dkt_matvec.  Same assertions as the GPU parity test
(tests/test_gpu_parity.py): construction-time tables bit-exact against the reference's golden fixtures, vectors within
1e-12.  This checks the library's LOGIC without a GPU; it is not a CPU path of the product (libdkt.so has none)."""
import numpy as np
import pytest

import cases
import emu_full
import flat
from test_oracle import load_case

TOL = 1e-12
INVALID = 0xFFFFFFFF


def _sorted_rows(xyz, lev, *cols):
    key = np.lexsort(tuple(xyz[:, d] for d in range(xyz.shape[1])) + (lev,))
    return [c[key] for c in (xyz, lev) + cols]


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("name", cases.ALL_CASES + cases.D4P2_CASES)
def test_emulated_pipeline_matches_reference(name):
    case = load_case(name)
    g = case["golden"]
    dim, order, md = case["dim"], case["order"], case["max_depth"]
    sfc = 1 if case["sfc"] == "hilbert" else 0
    da = emu_full.EmuDA(case["xyz"], case["lev"], dim, order, md, sfc=sfc, ip0=g["ip0"], ip1=g["ip1"])
    e = da.export()
    # --- construction-time tables: bit-exact against the reference -------------------------------
    assert np.array_equal(e["elem_xyz"], g["elem_xyz"]) and np.array_equal(e["elem_lev"], g["elem_lev"])
    assert np.array_equal(e["node_xyz"], g["node_xyz"]), "CG node order differs from DA::getTNCoords()"
    assert np.array_equal(e["node_lev"], g["node_lev"])
    assert np.array_equal(e["bdy"], g["bdy"])
    assert da.n_mv_elem == int(g["ncalls"])
    # --- flat tables against the oracle ------------------------------------------------------------
    t = cases.oracle_tables_for(case)
    assert da.tree_class == t.tree_class
    oe2n = np.where(t.e2n < 0, INVALID, t.e2n).astype(np.uint32)
    for x, y in zip(_sorted_rows(e["mv_xyz"], e["mv_lev"], e["e2n"]), _sorted_rows(t.mv_xyz, t.mv_lev, oe2n)):
        assert np.array_equal(x, y)
    nh = da.n_hanging
    assert nh == len(t.hang_idx)
    if nh:
        op = np.where(t.pnode < 0, INVALID, t.pnode).astype(np.uint32)
        a = _sorted_rows(e["mv_xyz"][-nh:], e["mv_lev"][-nh:], e["pnode"], e["child"])
        b = _sorted_rows(t.mv_xyz[t.hang_idx], t.mv_lev[t.hang_idx], op, t.child.astype(np.uint8))
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    # --- matvec against the reference's own output ---------------------------------------------------
    n = da.n_nodes
    K = cases.dense_operator(dim, order)
    u = cases.input_vector(n)
    assert rel(da.matvec(u, kref=K, alpha=float(g["alpha"]), scale=float(g["scale"])), g["v_dense"]) <= TOL
    if len(t.mv_lev) < 50000:
        assert rel(da.matvec(u, kref=K, alpha=float(g["alpha"]), scale=float(g["scale"]), dirichlet=True), g["v_dense_diri"]) <= TOL
        assert rel(da.matvec(np.ones(n)), g["v_id"]) <= TOL
        # the flat kernels (independent second implementation) and the mathematically consistent variant without Q1
        assert rel(da.matvec(u, kref=K, alpha=float(g["alpha"]), scale=float(g["scale"]), flat=True), g["v_dense"]) <= TOL
        assert rel(da.matvec(u, kref=K, alpha=float(g["alpha"]), scale=float(g["scale"]), dirichlet=True, flat=True), g["v_dense_diri"]) <= TOL
        vo2 = flat.matvec(t, u, K, alpha=float(g["alpha"]), scale=float(g["scale"]), ip0=g["ip0"], ip1=g["ip1"], q1_mask=False)
        assert rel(da.matvec(u, kref=K, alpha=float(g["alpha"]), scale=float(g["scale"]), q1_mask=False), vo2) <= TOL
        if order == 1:  # Walsh-Hadamard fast path against the dense product of the same operator
            Kl = flat.laplace_kref(dim, 1)
            v_fast = da.matvec(u, kref=Kl, alpha=dim - 2.0)
            v_dense = da.matvec(u, kref=Kl, alpha=dim - 2.0, fastpath=False)
            assert rel(v_fast, v_dense) <= TOL
    da_fam_v = da.matvec(u, kref=flat.laplace_kref(dim, 1), alpha=dim - 2.0, scale=0.7, dirichlet=True) if order == 1 else None
    da.close()
    # --- the same tree with per-element chunk tables only (DKT_FAMILIES=0; order 1 runs on family tables by default) ----
    if order == 1 and len(t.mv_lev) < 50000:
        dg = emu_full.EmuDA(case["xyz"], case["lev"], dim, order, md, sfc=sfc, ip0=g["ip0"], ip1=g["ip1"], families=0)
        assert rel(dg.matvec(np.ones(n)), g["v_id"]) <= TOL
        Kl = flat.laplace_kref(dim, 1)
        vo = flat.matvec(t, u, Kl, alpha=dim - 2.0, scale=0.7, ip0=g["ip0"], ip1=g["ip1"], dirichlet=True)
        assert rel(dg.matvec(u, kref=Kl, alpha=dim - 2.0, scale=0.7, dirichlet=True), vo) <= TOL
        assert rel(da_fam_v, vo) <= TOL
        dg.close()


def test_emulated_pipeline_refuses_class_u():
    """the stock moving-ball sphere touches the domain boundary with level jumps: class U"""
    import dkt
    dim, md = 4, 8
    c = np.full(dim, 0.25)

    def g(ctr):
        return np.abs(np.sqrt(((ctr - c) ** 2).sum(axis=1)) - 0.25) * 1.1

    xyz, lev = dkt.trees._refine(np, dim, md, 1, 4, g)
    t = flat.build_tables(xyz, lev, dim, 1, md)
    if t.tree_class != "U":
        pytest.skip("generator did not produce a class-U tree")
    with pytest.raises(RuntimeError, match="class-U"):
        emu_full.EmuDA(xyz, lev, dim, 1, md)


def test_emulated_cg_solver():
    """dkt_cg_solve (HeatMat::cgSolve with resident vectors) under the emulation - the block reductions with warp shuffles
    included: converges to the manufactured solution on a uniform grid and follows the oracle's iterates on an adaptive tree"""
    import dkt
    from test_gpu_parity import _oracle_cg
    dim, md = 3, 10
    xyz, lev = dkt.trees.uniform_tree(dim, 3, md)
    da = emu_full.EmuDA(xyz, lev, dim, 1, md)
    K = flat.laplace_kref(dim, 1)
    t = flat.build_tables(xyz, lev, dim, 1, md)
    xt = cases.input_vector(da.n_nodes)
    xt[t.bdy_ids] = 0.0
    b = da.matvec(xt, kref=K, alpha=dim - 2.0, dirichlet=True)
    x, it, resid, status = da.cg_solve(b, K, dim - 2.0, max_iter=300, tol=1e-11)
    assert status == 0 and resid <= 1e-11 and it < 300
    assert np.abs(x - xt).max() <= 1e-8 * np.abs(xt).max()
    da.close()
    case = load_case("ex3-d3-p1-morton-3")
    t = cases.oracle_tables_for(case)
    da = emu_full.EmuDA(case["xyz"], case["lev"], 3, 1, case["max_depth"])
    bb = cases.input_vector(da.n_nodes, seed=4)
    bb[t.bdy_ids] = 0.0
    xo, ito, ro = _oracle_cg(t, K, 1.0, bb, 12, 0.0)
    xg, itg, rg, _ = da.cg_solve(bb, K, 1.0, max_iter=12, tol=0.0)
    assert itg == ito == 12
    assert np.abs(xg - xo).max() <= 1e-9 * np.abs(xo).max() and abs(rg - ro) <= 1e-9 * ro
    da.close()


@pytest.mark.parametrize("seed", [46, 47, 49, 52, 55, 59])
def test_emulated_pipeline_random_trees(seed):
    """random 2:1-balanced trees (dims 2-4, orders 1-2, Morton / Hilbert, depths 8-14, phantom elements, a class-U one): the
    emulated construction must equal the oracle's tables bit for bit, the matvec its vector"""
    import dkt
    tabs = {d: cases.sfc_tables_for(load_case(n)) for d, n in ((2, "ex1-d2-p1-hilbert-5"), (3, "ex3-d3-p1-hilbert-3"), (4, "ex3-d4-p1-hilbert-3"))}
    rng = np.random.default_rng(9000 + seed)
    dim = int(rng.choice([2, 3, 4]))
    order = int(rng.choice([1, 2]))
    hil = int(rng.integers(0, 2))
    md = int(rng.choice([8, 10, 14]))
    maxl = {2: 7, 3: 5, 4: 4}[dim] - (order == 2)
    pts = rng.uniform(0, 1, (int(rng.integers(1, 4)), dim))
    if rng.random() < 0.5:
        pts[0] = rng.choice([0.0, 1.0], dim)
    rad = rng.uniform(0, 0.35, len(pts))

    def g(ctr):
        d = np.full(len(ctr), 1e9)
        for q, r in zip(pts, rad):
            d = np.minimum(d, np.abs(np.sqrt(((ctr - q) ** 2).sum(axis=1)) - r))
        return d

    xyz, lev = dkt.trees._refine(np, dim, md, int(rng.integers(0, 3)), maxl, g)
    perm = rng.permutation(len(lev))
    xyz, lev = xyz[perm], lev[perm]
    t = flat.build_tables(xyz, lev, dim, order, md, tabs[dim] if hil else None)
    if t.tree_class == "U":
        with pytest.raises(RuntimeError, match="class-U"):
            emu_full.EmuDA(xyz, lev, dim, order, md, sfc=hil)
        return
    da = emu_full.EmuDA(xyz, lev, dim, order, md, sfc=hil)
    e = da.export()
    assert da.tree_class == t.tree_class and da.n_mv_elem == len(t.mv_lev) and da.n_hanging == len(t.hang_idx)
    assert np.array_equal(e["elem_xyz"], t.elem_xyz) and np.array_equal(e["elem_lev"], t.elem_lev)
    assert np.array_equal(e["node_xyz"], t.node_xyz) and np.array_equal(e["node_lev"], t.node_lev) and np.array_equal(e["bdy"], t.bdy_ids)
    oe2n = np.where(t.e2n < 0, INVALID, t.e2n).astype(np.uint32)
    for x, y in zip(_sorted_rows(e["mv_xyz"], e["mv_lev"], e["e2n"]), _sorted_rows(t.mv_xyz, t.mv_lev, oe2n)):
        assert np.array_equal(x, y)
    n = da.n_nodes
    u = rng.uniform(-1, 1, n)
    N = (order + 1) ** dim
    K = rng.uniform(-1, 1, (N, N))
    assert rel(da.matvec(u, kref=K, alpha=1.3, scale=0.6, dirichlet=True), flat.matvec(t, u, K, alpha=1.3, scale=0.6, dirichlet=True)) <= TOL
    da.close()
