"""GPU parity tests: the CUDA path (through the C ABI) against the oracle restatement and the
golden vectors taken from the reference.  Integer tables bit-exact; fp64 vectors within 1e-12
relative (the tolerance BASELINE.json's north_star states)."""
import numpy as np
import pytest

import cases
import flat
from test_oracle import load_case

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _sfc(dkt, case):
    return dkt.SFC_HILBERT if case["sfc"] == "hilbert" else dkt.SFC_MORTON


def _sorted_rows(xyz, lev, *cols):
    """Canonical row order (level, coordinates) to compare element tables stored in different orders."""
    key = np.lexsort(tuple(xyz[:, d] for d in range(xyz.shape[1])) + (lev,))
    return [c[key] for c in (xyz, lev) + cols]


@pytest.mark.parametrize("name", cases.ALL_CASES)
def test_tables_and_matvec_match_reference(dkt, name):
    case = load_case(name)
    g = case["golden"]
    da = dkt.DA(case["xyz"], case["lev"], case["dim"], case["order"], case["max_depth"], sfc=_sfc(dkt, case), ip0=g["ip0"],
                ip1=g["ip1"])
    # --- construction-time tables: bit-exact against the reference -------------------------------
    exyz, elev = da.elements()
    assert np.array_equal(exyz, g["elem_xyz"]) and np.array_equal(elev, g["elem_lev"])
    nxyz, nlev = da.nodes()
    assert np.array_equal(nxyz, g["node_xyz"]), "CG node order differs from DA::getTNCoords()"
    assert np.array_equal(nlev, g["node_lev"])
    assert np.array_equal(da.boundary_ids(), g["bdy"])
    assert da.n_mv_elem == int(g["ncalls"])
    # --- flat tables against the oracle ------------------------------------------------------------
    t = cases.oracle_tables_for(case)
    assert da.tree_class == t.tree_class
    tb = da.tables()
    oe2n = np.where(t.e2n < 0, dkt.INVALID, t.e2n).astype(np.uint32)
    a = _sorted_rows(tb["mv_xyz"], tb["mv_lev"], tb["e2n"])
    b = _sorted_rows(t.mv_xyz, t.mv_lev, oe2n)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    nh = da.n_hanging
    assert nh == len(t.hang_idx)
    if nh:
        hx, hl = tb["mv_xyz"][-nh:], tb["mv_lev"][-nh:]
        a = _sorted_rows(hx, hl, tb["pnode"], tb["child"])
        op = np.where(t.pnode < 0, dkt.INVALID, t.pnode).astype(np.uint32)
        b = _sorted_rows(t.mv_xyz[t.hang_idx], t.mv_lev[t.hang_idx], op, t.child.astype(np.uint8))
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    # --- matvec against the reference's own output ---------------------------------------------------
    n = da.n_nodes
    K = cases.dense_operator(case["dim"], case["order"])
    u = cases.input_vector(n)
    for flat_path in (False, True):  # the chunked (production) kernels and the flat ones
        v = da.matvec(dkt.Operator.dense(K, float(g["alpha"])), u, scale=float(g["scale"]), flat=flat_path)
        assert np.abs(v - g["v_dense"]).max() <= TOL * np.abs(g["v_dense"]).max()
        vd = da.matvec(dkt.Operator.dense(K, float(g["alpha"]), dirichlet=True), u, scale=float(g["scale"]), flat=flat_path)
        assert np.abs(vd - g["v_dense_diri"]).max() <= TOL * np.abs(g["v_dense_diri"]).max()
        vi = da.matvec(dkt.Operator.identity(), np.ones(n), flat=flat_path)
        assert np.abs(vi - g["v_id"]).max() <= TOL * np.abs(g["v_id"]).max()
    # the mathematically consistent variant (no Q1 mask) agrees with the oracle's
    v2 = da.matvec(dkt.Operator.dense(K, float(g["alpha"])), u, scale=float(g["scale"]), q1_mask=False)
    vo2 = flat.matvec(t, u, K, alpha=float(g["alpha"]), scale=float(g["scale"]), ip0=g["ip0"], ip1=g["ip1"], q1_mask=False)
    assert np.abs(v2 - vo2).max() <= TOL * np.abs(vo2).max()
    da.close()


@pytest.mark.parametrize("name", ["ball-d2-p1-morton-7", "ball-d3-p1-morton-6", "ball-d4-p1-morton-5", "gauss-d4-p1-morton"])
def test_structured_operator_fast_path(dkt, name):
    """Laplacian / mass / Helmholtz reference matrices have Walsh-Hadamard diagonal form at order 1;
    the library applies them in that form.  Same answer as the dense product and as the oracle."""
    case = load_case(name)
    g = case["golden"]
    dim = case["dim"]
    da = dkt.DA(case["xyz"], case["lev"], dim, 1, case["max_depth"], ip0=g["ip0"], ip1=g["ip1"])
    t = cases.oracle_tables_for(case)
    u = cases.input_vector(da.n_nodes)
    for K, alpha in ((dkt.operators.laplace_kref(dim, 1), dim - 2.0), (dkt.operators.mass_kref(dim, 1), float(dim))):
        vo = flat.matvec(t, u, K, alpha=alpha, ip0=g["ip0"], ip1=g["ip1"])
        op = dkt.Operator.dense(K, alpha)
        v_fast = da.matvec(op, u)
        v_dense = da.matvec(op, u, fastpath=False)
        v_flat = da.matvec(op, u, flat=True)
        for v in (v_fast, v_dense, v_flat):
            assert np.abs(v - vo).max() <= TOL * np.abs(vo).max()
    da.close()


@pytest.mark.parametrize("name", ["ex3-d3-p2-morton-3", "gauss-d3-p2-morton", "ex1-d2-p2-morton-4", "ball-d4-p1-morton-5"])
def test_sum_factorised_operator(dkt, name):
    """DKT_OP_KRON - the sum-factorised form of HeatMat/HeatVec (FEM/examples/src/heatMat.cpp:46-117): Laplacian + mass as
    Kronecker terms of 1-D matrices, axis passes in registers at order 2 - against the oracle with the dense matrix the terms
    stand for, and against the dense device operator"""
    case = load_case(name)
    g = case["golden"]
    dim, order = case["dim"], case["order"]
    da = dkt.DA(case["xyz"], case["lev"], dim, order, case["max_depth"], ip0=g["ip0"], ip1=g["ip1"])
    t = cases.oracle_tables_for(case)
    u = cases.input_vector(da.n_nodes)
    mt = dkt.operators.mass_terms(dim, order)
    mt[0, 0] *= 0.3  # the factor of a term goes into ONE of its 1-D matrices
    terms = np.concatenate([dkt.operators.laplace_terms(dim, order), mt])
    K = dkt.operators.laplace_kref(dim, order) + 0.3 * dkt.operators.mass_kref(dim, order)
    for diri in (False, True):
        vo = flat.matvec(t, u, K, alpha=1.25, scale=0.7, ip0=g["ip0"], ip1=g["ip1"], dirichlet=diri)
        v = da.matvec(dkt.Operator.kron(terms, 1.25, dirichlet=diri), u, scale=0.7)
        vd = da.matvec(dkt.Operator.dense(K, 1.25, dirichlet=diri), u, scale=0.7)
        vf = da.matvec(dkt.Operator.kron(terms, 1.25, dirichlet=diri), u, scale=0.7, flat=True)
        for x in (v, vd, vf):
            assert np.abs(x - vo).max() <= TOL * np.abs(vo).max()
    da.close()


@pytest.mark.parametrize("fixture", ["heatmat-d3-p1-ball", "heatmat-d3-p1-ex3"])
def test_reference_heatmat_operator(dkt, fixture):
    """v = A u of the reference's own HeatEq::HeatMat<3> (stiffness operator + Dirichlet pre/postMatVec,
    FEM/examples/src/heatMat.cpp:46-139) against the device operator built from its probed K_ref, and
    against the library's own unit-cube Laplacian."""
    import os
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", fixture + ".npz")))
    da = dkt.DA(g["in_xyz"], g["in_lev"], 3, 1, int(g["max_depth"]), ip0=g["ip0"], ip1=g["ip1"])
    u = cases.input_vector(da.n_nodes)
    for K in (g["heat_kref"], dkt.operators.laplace_kref(3, 1)):
        for kw in (dict(), dict(fastpath=False), dict(flat=True)):
            v = da.matvec(dkt.Operator.dense(K, float(g["heat_alpha"]), dirichlet=True), u, **kw)
            assert np.abs(v - g["v_heat"]).max() <= TOL * np.abs(g["v_heat"]).max()
    da.close()


def test_device_pointer_path_and_repeatability(dkt):
    import torch
    dim, md = 3, 12
    xyz, lev = dkt.trees.moving_ball_tree(dim, 7, md)
    da = dkt.DA(xyz, lev, dim, 1, md)
    t = flat.build_tables(xyz, lev, dim, 1, md)
    K = flat.laplace_kref(dim, 1)
    u = cases.input_vector(da.n_nodes)
    vo = flat.matvec(t, u, K, alpha=dim - 2.0)
    op = dkt.Operator.dense(K, dim - 2.0)
    vh = da.matvec(op, u)
    ud = torch.from_numpy(u).cuda()
    vd = torch.empty_like(ud)
    da.matvec(op, ud, vd)
    torch.cuda.synchronize()
    assert np.abs(vh - vo).max() <= TOL * np.abs(vo).max()
    assert np.abs(vd.cpu().numpy() - vo).max() <= TOL * np.abs(vo).max()
    assert da.last_kernel_ms() > 0
    da.close()


def test_elements_on_device_and_presorted(dkt):
    import torch
    dim, md = 4, 10
    xyz, lev = dkt.trees.moving_ball_tree(dim, 5, md, use_torch=True)
    da = dkt.DA(xyz, lev, dim, 1, md)
    t = flat.build_tables(xyz.cpu().numpy().astype(np.uint32), lev.cpu().numpy(), dim, 1, md)
    nx, nl = da.nodes()
    assert np.array_equal(nx, t.node_xyz) and np.array_equal(nl, t.node_lev)
    ex, el = da.elements()
    da2 = dkt.DA(ex, el, dim, 1, md, presorted=True)
    nx2, _ = da2.nodes()
    assert np.array_equal(nx2, nx)
    da.close()
    da2.close()


def test_properties_at_scale(dkt):
    """Size-independent properties on a tree the oracle cannot reach in seconds (3-D uniform L7 =
    2.1e6 elements and a 4-D ball with ~2e6 elements):
    * Laplacian annihilates constants (also across hanging faces, interpolation is exact on them)
    * linearity
    * identity operator on a uniform tree counts incident elements: sum = 2^dim * nElem."""
    import torch
    for dim, xyz, lev, md in [(3,) + dkt.trees.uniform_tree_torch(3, 7, 12) + (12,),
                              (4,) + dkt.trees.moving_ball_tree(4, 6, 10, use_torch=True) + (10,)]:
        da = dkt.DA(xyz, lev, dim, 1, md)
        K = flat.laplace_kref(dim, 1)
        op = dkt.Operator.dense(K, dim - 2.0)
        n = da.n_nodes
        ones = torch.ones(n, dtype=torch.float64, device="cuda")
        v = da.matvec(op, ones)
        torch.cuda.synchronize()
        assert float(v.abs().max()) < 1e-11
        g = torch.Generator(device="cuda").manual_seed(1)
        a = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
        b = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
        va, vb, vab = da.matvec(op, a), da.matvec(op, b), da.matvec(op, 2.0 * a - 3.0 * b)
        torch.cuda.synchronize()
        ref = 2.0 * va - 3.0 * vb
        assert float((vab - ref).abs().max()) <= 1e-11 * float(ref.abs().max())
        vi = da.matvec(dkt.Operator.identity(), ones)
        torch.cuda.synchronize()
        if da.tree_class == "A":
            assert abs(float(vi.sum()) - (1 << dim) * da.n_elem) < 1e-6 * da.n_elem
        da.close()


@pytest.mark.parametrize("nranks", [2, 3, 8])
def test_partition_tables_are_consistent(dkt, nranks):
    """All ranks' partitions built on ONE GPU (dry run, no communicator): every node has exactly one
    owner, what r sends to p is what p expects from r, element work is split by weight."""
    dim, md = 4, 10
    xyz, lev = dkt.trees.moving_ball_tree(dim, 5, md)
    da1 = dkt.DA(xyz, lev, dim, 1, md)
    n_global, n_mv = da1.n_nodes, da1.n_mv_elem
    da1.close()
    owned_all, send, recv, work = [], [], [], []
    for r in range(nranks):
        da = dkt.DA(xyz, lev, dim, 1, md, rank=r, nranks=nranks, dryrun=True)
        assert da.n_global_nodes == n_global
        owned_all.append(da.owned_ids())
        sc, rc = da.exchange_counts()
        send.append(sc)
        recv.append(rc)
        work.append(3 * (da.n_mv_elem - da.n_hanging) + 5 * da.n_hanging)
        assert da.n_nodes > 0 and da.n_mv_elem > 0
        with pytest.raises(dkt.DktError):
            da.matvec(dkt.Operator.identity(), np.zeros(da.n_nodes))
        da.close()
    allo = np.concatenate(owned_all)
    assert len(allo) == n_global and len(np.unique(allo)) == n_global
    for r in range(nranks):
        for p in range(nranks):
            assert send[r][p] == recv[p][r]
        assert send[r][r] == 0 and recv[r][r] == 0
    assert max(work) <= 1.05 * (sum(work) / nranks) + 5 * 256


def _oracle_cg(t, K, alpha, b, max_iter, tol):
    """HeatMat::cgSolve (FEM/examples/src/heatMat.cpp:165-325) on top of the oracle matvec."""
    mv = lambda x: flat.matvec(t, x, K, alpha=alpha, dirichlet=True)
    x = np.zeros_like(b)
    normb = np.abs(b).max() or 1.0
    r0 = b - mv(x)
    p = r0.copy()
    resid = np.abs(r0).max() / normb
    it = 0
    for it in range(1, max_iter + 1):
        Ap = mv(p)
        a = (r0 @ r0) / (p @ Ap)
        x += a * p
        r1 = r0 - a * Ap
        resid = np.abs(r1).max() / normb
        if resid <= tol:
            break
        beta = (r1 @ r1) / (r0 @ r0)
        p = r1 + beta * p
        r0 = r1
    return x, it, resid


def test_resident_cg_solver(dkt):
    """dkt_cg_solve = HeatMat::cgSolve with resident vectors: (a) converges on a uniform grid (SPD
    operator) to the manufactured solution; (b) on an adaptive tree follows the oracle's iterates."""
    import torch
    dim, md = 3, 10
    xyz, lev = dkt.trees.uniform_tree(dim, 4, md)
    da = dkt.DA(xyz, lev, dim, 1, md)
    K = dkt.operators.laplace_kref(dim, 1)
    op = dkt.Operator.dense(K, dim - 2.0, dirichlet=True)
    xt = cases.input_vector(da.n_nodes)
    xt[da.boundary_ids()] = 0.0
    b = da.matvec(op, xt)
    x, it, resid, ok = da.cg_solve(op, b, max_iter=500, tol=1e-11)
    assert ok and resid <= 1e-11 and it < 500
    assert np.abs(x - xt).max() <= 1e-8 * np.abs(xt).max()
    xd, it2, _, ok2 = da.cg_solve(op, torch.from_numpy(b).cuda(), max_iter=500, tol=1e-11)  # device vectors
    assert ok2 and it2 == it and np.abs(xd.cpu().numpy() - x).max() <= 1e-12 * np.abs(x).max()
    da.close()
    case = load_case("ball-d3-p1-morton-6")
    t = cases.oracle_tables_for(case)
    da = dkt.DA(case["xyz"], case["lev"], 3, 1, case["max_depth"])
    bb = cases.input_vector(da.n_nodes, seed=4)
    bb[da.boundary_ids()] = 0.0
    xo, ito, ro = _oracle_cg(t, K, 1.0, bb, 12, 0.0)
    xg, itg, rg, _ = da.cg_solve(op, bb, max_iter=12, tol=0.0)
    assert itg == ito == 12
    assert np.abs(xg - xo).max() <= 1e-9 * np.abs(xo).max() and abs(rg - ro) <= 1e-9 * ro
    da.close()


def class_u_tree():
    """A ball of radius 0.25 centred at 0.25 (the stock testMovingBall parameters,
    test/testMovingBall.cpp:122-175) touches the domain boundary with level jumps."""
    import dkt as m
    g = m.trees._ball_g(np, 4, 0.25, (0.25, 0.25, 0.25), 0.25, 0.25, 0.75)
    return m.trees._refine(np, 4, 10, 2, 4, g)


def test_class_u_tree_is_refused(dkt):
    """On such a tree the reference reads undefined parent values (FEM/include/matvec.h:439-447);
    the library detects it at construction and refuses unless told otherwise."""
    xyz, lev = class_u_tree()
    t = flat.build_tables(xyz, lev, 4, 1, 10)
    assert t.tree_class == "U"
    with pytest.raises(dkt.DktError, match="class-U"):
        dkt.DA(xyz, lev, 4, 1, 10)
    da = dkt.DA(xyz, lev, 4, 1, 10, allow_undefined=True)
    assert da.tree_class == "U" and da.n_mv_elem == len(t.mv_lev)
    nx, nl = da.nodes()
    assert np.array_equal(nx, t.node_xyz) and np.array_equal(nl, t.node_lev)
    # with absent parents read as zero (the intended semantics) both sides still agree
    u = cases.input_vector(da.n_nodes)
    K = dkt.operators.laplace_kref(4, 1)
    v = da.matvec(dkt.Operator.dense(K, 2.0), u)
    vo = flat.matvec(t, u, K, alpha=2.0)
    assert np.abs(v - vo).max() <= TOL * np.abs(vo).max()
    da.close()


def test_smoke_case(dkt):
    import smoke_case
    smoke_case.run(dkt)


@pytest.mark.parametrize("name", ["ball-d4-p1-morton-5", "gauss-d3-p1-hilbert", "ball-d3-p2-morton-5"])
def test_interleaved_dof_vectors(dkt, name):
    """dof > 1 (include/oda.h:296-322 layout [abc][abc]..): every component gets the scalar operator.  The reference's traversal is
    dof == 1 only (FEM/include/matvec.h:27), so each component is compared with the reference's scalar vector."""
    import torch
    case = load_case(name)
    g = case["golden"]
    da = dkt.DA(case["xyz"], case["lev"], case["dim"], case["order"], case["max_depth"], sfc=_sfc(dkt, case), ip0=g["ip0"], ip1=g["ip1"])
    n = da.n_nodes
    K = cases.dense_operator(case["dim"], case["order"])
    u = cases.input_vector(n)
    dof = 3
    U = np.stack([u, -2.0 * u, u[::-1]], axis=1)  # (n, dof), C order = interleaved
    op = dkt.Operator.dense(K, float(g["alpha"]))
    want0 = g["v_dense"]
    want2 = da.matvec(op, np.ascontiguousarray(u[::-1]), scale=float(g["scale"]))
    for dev in (False, True):
        if dev:
            V = da.matvec(op, torch.from_numpy(U).cuda().reshape(-1), scale=float(g["scale"]), dof=dof)
            torch.cuda.synchronize()
            V = V.cpu().numpy().reshape(n, dof)
        else:
            V = da.matvec(op, U.reshape(-1), scale=float(g["scale"]), dof=dof).reshape(n, dof)
        s = np.abs(want0).max()
        assert np.abs(V[:, 0] - want0).max() <= TOL * s
        assert np.abs(V[:, 1] + 2.0 * want0).max() <= 2 * TOL * s
        assert np.abs(V[:, 2] - want2).max() <= TOL * np.abs(want2).max()
    da.close()
