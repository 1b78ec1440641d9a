"""GPU parity of the OPT-IN sibling-group tables (DKT_GROUPS=g, dendro-kt_b200/csrc/dkt_chunks.cu: k_mvg).
Their logic is covered on the CPU by tests/test_emu_chunks.py; these tests run the real kernels and are enabled
with DKT_TEST_GROUPS=1 (tools/r02_groups_ab.sh) until the path has been confirmed on a B200 and made the default."""
import os

import numpy as np
import pytest

import cases
import flat
from test_oracle import load_case

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("DKT_TEST_GROUPS") != "1", reason="opt-in: DKT_TEST_GROUPS=1")]
TOL = 1e-12
GROUP_G = {2: ("2",), 3: ("3", "3,2"), 4: ("2", "3", "2,1", "3,2")}


@pytest.fixture
def groups_env():
    old = os.environ.get("DKT_GROUPS")
    yield
    if old is None:
        os.environ.pop("DKT_GROUPS", None)
    else:
        os.environ["DKT_GROUPS"] = old


@pytest.mark.parametrize("name", [c for c in cases.ALL_CASES if "-p1-" in c])
def test_group_tables_match_reference(dkt, groups_env, name):
    case = load_case(name)
    g = case["golden"]
    dim = case["dim"]
    t = cases.oracle_tables_for(case)
    sfc = dkt.SFC_HILBERT if case["sfc"] == "hilbert" else dkt.SFC_MORTON
    for gg in GROUP_G[dim]:
        os.environ["DKT_GROUPS"] = str(gg)
        da = dkt.DA(case["xyz"], case["lev"], dim, 1, case["max_depth"], sfc=sfc, ip0=g["ip0"], ip1=g["ip1"])
        n = da.n_nodes
        vi = da.matvec(dkt.Operator.identity(), np.ones(n))
        assert np.abs(vi - g["v_id"]).max() <= TOL * np.abs(g["v_id"]).max()
        u = cases.input_vector(n)
        for K, alpha in ((dkt.operators.laplace_kref(dim, 1), dim - 2.0), (dkt.operators.mass_kref(dim, 1), float(dim))):
            for diri in (False, True):
                vo = flat.matvec(t, u, K, alpha=alpha, scale=0.7, ip0=g["ip0"], ip1=g["ip1"], dirichlet=diri)
                v = da.matvec(dkt.Operator.dense(K, alpha, dirichlet=diri), u, scale=0.7)
                assert np.abs(v - vo).max() <= TOL * np.abs(vo).max()
        # a dense operator without Walsh-Hadamard form falls back to the flat kernels: still the reference's answer
        Kd = cases.dense_operator(dim, 1)
        v = da.matvec(dkt.Operator.dense(Kd, float(g["alpha"])), u, scale=float(g["scale"]))
        assert np.abs(v - g["v_dense"]).max() <= TOL * np.abs(g["v_dense"]).max()
        da.close()


def test_group_tables_at_scale(dkt, groups_env):
    """2e6-element 4-D tree: group tables against the default tables on device vectors; repeatable"""
    import torch
    dim, md = 4, 12
    xyz, lev = dkt.trees.moving_ball_tree(dim, 7, md, use_torch=True)
    op = dkt.Operator.dense(dkt.operators.laplace_kref(dim, 1), dim - 2.0)
    outs = []
    for gg in ("0", "2", "3", "2,1"):
        os.environ["DKT_GROUPS"] = gg
        da = dkt.DA(xyz, lev, dim, 1, md)
        u = torch.rand(da.n_nodes, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5)) * 2 - 1
        v = torch.zeros_like(u)
        da.matvec(op, u, v)
        torch.cuda.synchronize()
        v2 = torch.zeros_like(u)
        da.matvec(op, u, v2)
        torch.cuda.synchronize()
        assert torch.equal(v, v2) or (v - v2).abs().max().item() <= 1e-13 * v.abs().max().item()
        outs.append(v.cpu().numpy())
        da.close()
    for v in outs[1:]:
        assert np.abs(v - outs[0]).max() <= TOL * np.abs(outs[0]).max()
