"""The C++ host API (dendro-kt_b200/include: ot::DA, feMatrix<LeafT,dim>::matVec with the
elementalMatVec callback, feVector::computeVec) driven the way a user of the reference drives it."""
import os
import subprocess

import numpy as np
import pytest

import cases
import flat
from test_oracle import load_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "dendro-kt_b200", "include")
LIBDIR = os.path.join(ROOT, "dendro-kt_b200", "lib")
SRC = os.path.join(ROOT, "tests", "cpp", "test_host_api.cpp")


@pytest.fixture(scope="session")
def host_api_binary(dkt, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cpp") / "test_host_api")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-Werror", "-I", INC, SRC, "-o", out, "-L", LIBDIR, "-ldkt",
                           "-Wl,-rpath," + LIBDIR])
    return out


def test_host_api_compiles(host_api_binary):
    """CPU: the header-only host layer compiles warning-free against the C ABI and links to libdkt.so."""
    assert os.path.exists(host_api_binary)


def _write_inputs(d, case, K, u, alpha, scale):
    case["xyz"].astype(np.uint32).tofile(os.path.join(d, "elem_xyz.bin"))
    case["lev"].astype(np.uint8).tofile(os.path.join(d, "elem_lev.bin"))
    K.astype(np.float64).tofile(os.path.join(d, "K.bin"))
    u.astype(np.float64).tofile(os.path.join(d, "u.bin"))
    np.array([alpha, scale], dtype=np.float64).tofile(os.path.join(d, "params.bin"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ex1-d2-p1-morton-6", "ball-d3-p1-morton-6", "gauss-d4-p1-morton", "ex3-d3-p2-morton-3"])
def test_host_api_matvec(host_api_binary, tmp_path, name):
    case = load_case(name)
    g = case["golden"]
    dim, order, md = case["dim"], case["order"], case["max_depth"]
    t = cases.oracle_tables_for(case)
    n = len(t.node_lev)
    K = cases.dense_operator(dim, order)
    u = cases.input_vector(n)
    alpha, scale = 1.5, 0.7
    d = str(tmp_path)
    _write_inputs(d, case, K, u, alpha, scale)
    for mode, diri in ((0, False), (1, True), (3, False)):
        rc = subprocess.run([host_api_binary, str(dim), str(order), str(md), str(mode), d], capture_output=True, text=True)
        assert rc.returncode == 0, rc.stderr
        nodes = np.fromfile(os.path.join(d, "nodes.bin"), dtype=np.uint32).reshape(-1, dim + 1)
        assert np.array_equal(nodes[:, :dim], g["node_xyz"]) and np.array_equal(nodes[:, dim], g["node_lev"])
        v = np.fromfile(os.path.join(d, "v.bin"), dtype=np.float64)
        vo = flat.matvec(t, u, K, alpha=alpha, scale=scale, dirichlet=diri)  # exact interpolation on both sides
        assert np.abs(v - vo).max() <= 1e-12 * np.abs(vo).max()
    # a callback that depends on the element position cannot be represented: refused loudly
    rc = subprocess.run([host_api_binary, str(dim), str(order), str(md), "2", d], capture_output=True, text=True)
    assert rc.returncode == 3 and "refused" in rc.stderr
