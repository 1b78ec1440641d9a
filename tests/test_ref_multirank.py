"""The reference on SEVERAL ranks inside this container: oracle/_ref/libdktref_mp_morton.so is the reference built
over the multi-process MPI stand-in oracle/shim_mp.  Checked here:
  * the stand-in itself (collectives, point-to-point, communicators) on 4 ranks,
  * the reference's distributed DA - node ownership, scatter maps (include/nsort.tcc:535-872), readFromGhost / writeToGhosts
    (include/oda.tcc:212-435) - and its distributed feMatrix::matVec: on 2, 3 and 4 ranks the owned nodes are a partition of the
    single-rank node set and the gathered vectors equal the single-rank ones (and the committed golden vectors),
  * distTreeConstruction on several ranks returns the single-rank tree.
This is what makes "N-GPU result gathered to the single-rank order == reference" (tests/dist_check.py) a statement about the
reference's own multi-rank path and not only about its single-rank one."""
import os
import subprocess

import numpy as np
import pytest

import cases
import dktref
import dktref_mp
import tree as otree
from test_oracle import load_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not dktref_mp.available(), reason="oracle/_ref/libdktref_mp_morton.so is not built (needs the reference sources)")


def test_mpi_stand_in_unit(tmp_path):
    exe = str(tmp_path / "test_mpi_mp")
    shim = os.path.join(ROOT, "oracle", "shim_mp")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-I", shim, os.path.join(ROOT, "tests", "cpp", "test_mpi_mp.cpp"), os.path.join(shim, "mpi_mp.cpp"),
                           "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.count("OK") == 4, (r.stdout, r.stderr)


def _ufun(xyz, md):
    x = xyz.astype(np.float64) / (1 << md)
    return np.sin(3.0 * x[:, 0] + 1.0) + x[:, 1] * x[:, -1] - 0.3 * x.sum(1) ** 2


def _da_job(rank, nranks, exyz, elev, dim, order, md, K, alpha, scale, dirichlet):
    R = dktref_mp.session(dim, md)
    n = len(elev)
    lo, hi = rank * n // nranks, (rank + 1) * n // nranks  # a contiguous piece of the sorted tree, like DistTree hands it over
    tree = R.tree_from_elements(exyz[lo:hi], elev[lo:hi], sort=False)
    da = R.da(tree, order)
    nloc, beg, ntot, npes, rk, nglob = dktref_mp.local_info(da)
    nx, nl = da.nodes()
    lx, ll = nx[beg:beg + nloc], nl[beg:beg + nloc]
    v, _, nc = da.matvec(_ufun(lx, md), dktref.OP_DENSE, K, alpha=alpha, scale=scale, dirichlet=dirichlet)
    v1, _, _ = da.matvec(np.ones(nloc))
    return dict(xyz=lx, lev=ll, v=v, v1=v1, info=(nloc, beg, ntot, npes, rk, nglob), calls=nc)


@pytest.mark.parametrize("name", ["ex1-d2-p1-morton-6", "ball-d3-p1-morton-6", "gauss-d3-p1-morton", "gauss-d4-p1-morton", "ball-d3-p2-morton-5",
                                  "ex3-d3-p2-morton-3"])
def test_reference_multirank_da_and_matvec(name):
    case = load_case(name)
    g = case["golden"]
    dim, order, md = case["dim"], case["order"], case["max_depth"]
    K = cases.dense_operator(dim, order)
    n = len(g["node_lev"])
    order_g = np.lexsort(g["node_xyz"].T[::-1])
    want = {}
    for diri in (False, True):
        want[diri] = dktref.Reference(dim, md).da(dktref.Reference(dim, md).tree_from_elements(g["elem_xyz"], g["elem_lev"], sort=False), order)
    for nranks in (2, 3, 4):
        for diri in (False, True):
            res = dktref_mp.run(nranks, _da_job, g["elem_xyz"], g["elem_lev"], dim, order, md, K, float(g["alpha"]), float(g["scale"]), diri)
            assert [r["info"][3] for r in res] == [nranks] * nranks and [r["info"][4] for r in res] == list(range(nranks))
            assert all(r["info"][5] == n for r in res), "global node count"
            xyz = np.concatenate([r["xyz"] for r in res])
            lev = np.concatenate([r["lev"] for r in res])
            assert len(xyz) == n, "the owned nodes of the ranks partition the node set"
            key = np.lexsort(xyz.T[::-1])
            assert np.array_equal(xyz[key], g["node_xyz"][order_g]) and np.array_equal(lev[key], g["node_lev"][order_g])
            v1 = np.concatenate([r["v1"] for r in res])[key]
            assert np.array_equal(v1, g["v_id"][order_g]), "identity operator, u = 1: the reference's own known answer, bit for bit"
            # dense operator on u = f(coordinates): against the single-rank reference on the same input
            da1 = want[diri]
            nx1, _ = da1.nodes()
            v_single, _, calls = da1.matvec(_ufun(nx1, md), dktref.OP_DENSE, K, alpha=float(g["alpha"]), scale=float(g["scale"]), dirichlet=diri)
            o1 = np.lexsort(nx1.T[::-1])
            v = np.concatenate([r["v"] for r in res])[key]
            assert np.abs(v - v_single[o1]).max() <= 1e-13 * np.abs(v_single).max()
            assert sum(r["calls"] for r in res) == calls == int(g["ncalls"]), "eleOp calls (phantom elements included) add up"


def _tree_job(rank, nranks, pts, dim, md, max_pts):
    R = dktref_mp.session(dim, md)
    return R.tree_from_points(pts[rank::nranks], max_pts=max_pts, balance=False).export()


@pytest.mark.parametrize("dim,md,n,max_pts", [(2, 10, 500, 1), (3, 8, 400, 1), (4, 6, 200, 3)])
def test_reference_multirank_tree_construction(dim, md, n, max_pts):
    import dkt
    pts = dkt.trees.gaussian_points(dim, n, md, seed=21 + dim)
    ref = otree.canonical(*otree.construct(pts, dim, md, max_pts))
    for nranks in (2, 3):
        res = dktref_mp.run(nranks, _tree_job, pts, dim, md, max_pts)
        x = np.concatenate([r[0] for r in res])
        l = np.concatenate([r[1] for r in res])
        assert np.array_equal(otree.canonical(x, l), ref)
