"""GPU parity of 4-D order 2 (81 nodes per element: generic table construction + the loop-based flat kernels k_mv_big
of dkt_matvec.cu).  The same assertions run on the CPU under the emulation (tests/test_emu_full.py).  Confirmed on a B200 in
round 2 (gpurun_out/r02_optin/d4p2.log: 2 passed)."""
import os

import numpy as np
import pytest

import cases
from test_oracle import load_case

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.mark.parametrize("name", cases.D4P2_CASES)
def test_d4p2_matches_reference(dkt, name):
    case = load_case(name)
    g = case["golden"]
    sfc = dkt.SFC_HILBERT if case["sfc"] == "hilbert" else dkt.SFC_MORTON
    da = dkt.DA(case["xyz"], case["lev"], 4, 2, case["max_depth"], sfc=sfc, ip0=g["ip0"], ip1=g["ip1"])
    exyz, elev = da.elements()
    assert np.array_equal(exyz, g["elem_xyz"]) and np.array_equal(elev, g["elem_lev"])
    nxyz, nlev = da.nodes()
    assert np.array_equal(nxyz, g["node_xyz"]) and np.array_equal(nlev, g["node_lev"])
    assert np.array_equal(da.boundary_ids(), g["bdy"])
    n = da.n_nodes
    K = cases.dense_operator(4, 2)
    u = cases.input_vector(n)
    v = da.matvec(dkt.Operator.dense(K, float(g["alpha"])), u, scale=float(g["scale"]))
    assert np.abs(v - g["v_dense"]).max() <= TOL * np.abs(g["v_dense"]).max()
    vd = da.matvec(dkt.Operator.dense(K, float(g["alpha"]), dirichlet=True), u, scale=float(g["scale"]))
    assert np.abs(vd - g["v_dense_diri"]).max() <= TOL * np.abs(g["v_dense_diri"]).max()
    vi = da.matvec(dkt.Operator.identity(), np.ones(n))
    assert np.abs(vi - g["v_id"]).max() <= TOL * np.abs(g["v_id"]).max()
    da.close()
