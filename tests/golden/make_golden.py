"""Generates tests/golden/*.npz from the REFERENCE itself (oracle/_ref, built from
/root/reference by oracle/build_ref.sh).  Run in the build container only:

    python tests/golden/make_golden.py

Each fixture holds, for one seeded case: the input elements, the reference's tree order, its CG
nodes in DA order (getTNCoords), boundary ids, RefElement's 1-D interpolation matrices, an input
vector u and v = A u from feMatrix::matVec for (a) the identity elemental operator with u = 1
(test/testMatvec.cpp:324-399), (b) a random dense K_ref with level scaling, (c) order 1: the unit-cell Laplacian
(--add-laplacian extends existing fixtures in place), plus the number of eleOp calls.  Hilbert cases also store the reference's SFC tables.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "dendro-kt_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import cases  # noqa: E402
import dktref  # noqa: E402
import dkt  # noqa: E402

ALPHA, SCALE = 1.5, 0.7


def point_cloud_case(name):
    parts = name.split("-")
    kind, dim, order, sfc = parts[0], int(parts[1][1:]), int(parts[2][1:]), parts[3]
    md = 14
    n = {2: 300, 3: 300, 4: 150}[dim]
    pts = dkt.trees.gaussian_points(dim, n, md, seed=7, guard_level=3 if kind == "gaussguard" else None)
    R = dktref.Reference(dim, md, sfc)
    tree = R.tree_from_points(pts, max_pts=1, flex=0.3, balance=True)
    xyz, lev = tree.export()
    rng = np.random.default_rng(len(lev))
    perm = rng.permutation(len(lev))
    return dict(name=name, dim=dim, order=order, max_depth=md, sfc=sfc, xyz=xyz[perm], lev=lev[perm])


def generate(case):
    R = dktref.Reference(case["dim"], case["max_depth"], case["sfc"])
    tree = R.tree_from_elements(case["xyz"], case["lev"], sort=True)
    exyz, elev = tree.export()
    da = R.da(tree, case["order"])
    nxyz, nlev = da.nodes()
    re = R.refel(case["order"])
    u = cases.input_vector(da.num_nodes)
    K = cases.dense_operator(case["dim"], case["order"])
    v_dense, _, ncalls = da.matvec(u, dktref.OP_DENSE, K, alpha=ALPHA, scale=SCALE)
    v_dense_diri, _, _ = da.matvec(u, dktref.OP_DENSE, K, alpha=ALPHA, scale=SCALE, dirichlet=True)
    v_id, _, _ = da.matvec(np.ones(da.num_nodes))
    out = dict(dim=case["dim"], order=case["order"], max_depth=case["max_depth"], in_xyz=case["xyz"], in_lev=case["lev"],
               elem_xyz=exyz, elem_lev=elev, node_xyz=nxyz, node_lev=nlev, bdy=da.boundary_ids(), ip0=re["ip0"], ip1=re["ip1"],
               v_dense=v_dense, v_dense_diri=v_dense_diri, v_id=v_id, ncalls=np.int64(ncalls), alpha=ALPHA, scale=SCALE)
    if case["sfc"] == "hilbert":
        rot, h = R.tables()
        nch = 1 << case["dim"]
        out.update(rot_perm=rot[:, :nch], rot_inv=rot[:, nch:], htab=h)
    return out


def generate_heatmat(name):
    """The reference's own example operator, HeatEq::HeatMat<3> (FEM/examples/src/heatMat.cpp:46-139):
    v = A u through feMatrix::matVec with its Dirichlet pre/postMatVec, plus the reference-cell matrix
    recovered by probing HeatMat<3>::elementalMatVec with unit vectors on cells of two levels."""
    case = cases.make_case(name)
    R = dktref.Reference(3, case["max_depth"], "morton")
    tree = R.tree_from_elements(case["xyz"], case["lev"], sort=True)
    da = R.da(tree, 1)
    n = da.num_nodes
    u = cases.input_vector(n)
    v_heat, _, _ = da.matvec(u, dktref.OP_HEATMAT)
    mi = np.array([[(r >> d) & 1 for d in range(3)] for r in range(8)], dtype=np.float64)

    def probe(level):
        h = 2.0 ** -level
        K = np.zeros((8, 8))
        for j in range(8):
            e = np.zeros(8)
            e[j] = 1.0
            K[:, j] = da.heat_elemental(e, (mi * h).ravel())
        return K
    K2, K3 = probe(2), probe(3)
    alpha = float(np.log2(K2[0, 0] / K3[0, 0]))
    assert np.abs(K3 - K2 * 2.0 ** -alpha).max() <= 1e-13 * np.abs(K3).max()
    re = R.refel(1)
    return dict(dim=3, order=1, max_depth=case["max_depth"], in_xyz=case["xyz"], in_lev=case["lev"], ip0=re["ip0"], ip1=re["ip1"],
                v_heat=v_heat, heat_kref=K2 * 2.0 ** (alpha * 2), heat_alpha=alpha, node_xyz=da.nodes()[0])


def add_laplacian(name):
    """Order-1 fixtures also carry v = A u of the reference for the unit-cell Laplacian K_ref (a Walsh-Hadamard-form
    operator: what the sibling-family kernel serves), K_e = 0.7 h^(dim-2) K_ref, with and without Dirichlet rows."""
    path = os.path.join(HERE, name + ".npz")
    g = dict(np.load(path))
    dim, md = int(g["dim"]), int(g["max_depth"])
    sfc = "hilbert" if "hilbert" in name else "morton"
    R = dktref.Reference(dim, md, sfc)
    tree = R.tree_from_elements(g["in_xyz"], g["in_lev"], sort=True)
    da = R.da(tree, 1)
    u = cases.input_vector(da.num_nodes)
    K = dkt.operators.laplace_kref(dim, 1)
    g["v_lap"], _, _ = da.matvec(u, dktref.OP_DENSE, K, alpha=dim - 2.0, scale=SCALE)
    g["v_lap_diri"], _, _ = da.matvec(u, dktref.OP_DENSE, K, alpha=dim - 2.0, scale=SCALE, dirichlet=True)
    np.savez_compressed(path, **g)
    print("%-28s + v_lap, v_lap_diri  %6.1f KB" % (name, os.path.getsize(path) / 1024))


HEATMAT_CASES = {"heatmat-d3-p1-ball": "ball-d3-p1-morton-6", "heatmat-d3-p1-ex3": "ex3-d3-p1-morton-3"}


def main():
    if "--add-laplacian" in sys.argv:  # extends the committed order-1 fixtures in place
        for name in cases.ALL_CASES:
            if "-p1-" in name:
                add_laplacian(name)
        return
    for fixture, name in HEATMAT_CASES.items():
        g = generate_heatmat(name)
        path = os.path.join(HERE, fixture + ".npz")
        np.savez_compressed(path, **g)
        print("%-28s nN=%6d alpha=%.3f  %6.1f KB" % (fixture, len(g["v_heat"]), g["heat_alpha"], os.path.getsize(path) / 1024))
    if "--heatmat-only" in sys.argv:
        return
    for name in cases.ALL_CASES + cases.D4P2_CASES:
        case = point_cloud_case(name) if name in cases.POINT_CLOUD_CASES else cases.make_case(name)
        g = generate(case)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **g)
        if "-p1-" in name and name in cases.ALL_CASES:
            add_laplacian(name)
        print("%-28s nE=%6d nN=%6d calls=%6d  %6.1f KB" % (name, len(g["elem_lev"]), len(g["node_lev"]), int(g["ncalls"]),
                                                            os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
