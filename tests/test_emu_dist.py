"""CPU test of the multi-rank path under the CUDA-on-CPU emulation (tests/emu/emu_dist.cpp): every rank runs the library's
own build_da + partition_da (dry run) + build_chunks in one process; the harness moves the ghost values with the
library's send / receive lists around the library's (phased) chunk matvec - or (p2p=1) the library's own peer-memory flow run_matvec_dist_p2p runs stage by
stage over all ranks, twice; the gathered result must equal the single-rank vector of the oracle.  Covers ownership, local numbering, exchange lists, the [interior | boundary] element
order with comm/compute phases, and the sibling-family tables on real partitions - without GPUs or NCCL."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cases
import flat
from test_oracle import load_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "dendro-kt_b200", "csrc")
LIB = os.path.join(EMU, "_build", "libdkt_emu_dist.so")
TOL = 1e-12


def _lib():
    import emu_build
    L = C.CDLL(emu_build.build())
    L.emu_last_error.restype = C.c_char_p
    L.emu_dist_matvec.restype = C.c_int
    L.emu_dist_matvec.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                  C.c_double, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int]
    return L


@pytest.mark.parametrize("name,R,families,overlap,p2p", [
    ("ball-d2-p1-morton-7", 2, "0", "0", 0), ("ball-d2-p1-morton-7", 3, "1", "1", 1),
    ("ball-d3-p1-morton-6", 3, "0", "1", 1), ("ball-d3-p1-morton-6", 8, "1", "1", 0), ("ball-d3-p1-morton-6", 2, "1", "0", 1),
    ("gauss-d4-p1-morton", 4, "0", "1", 0), ("gauss-d4-p1-morton", 3, "1", "1", 1), ("ex3-d4-p1-hilbert-3", 5, "1", "1", 0),
    ("ex3-d4-p1-hilbert-3", 8, "1", "0", 1), ("gauss-d3-p2-morton", 3, "0", "1", 0), ("gauss-d3-p2-morton", 4, "0", "1", 1),
])
def test_emulated_partitioned_matvec(name, R, families, overlap, p2p):
    case = load_case(name)
    g = case["golden"]
    dim, order, md = case["dim"], case["order"], case["max_depth"]
    t = cases.oracle_tables_for(case)
    n = len(g["node_lev"])
    u = cases.input_vector(n)
    if order == 1:
        K, alpha = flat.laplace_kref(dim, 1), dim - 2.0
    else:
        K, alpha = cases.dense_operator(dim, order), 1.5
    ref = flat.matvec(t, u, K, alpha=alpha, scale=0.7, ip0=g["ip0"], ip1=g["ip1"], dirichlet=True)
    xyz = np.ascontiguousarray(case["xyz"], dtype=np.uint32)
    lev = np.ascontiguousarray(case["lev"], dtype=np.uint8)
    ip0 = np.ascontiguousarray(np.asarray(g["ip0"], dtype=np.float64).ravel())
    ip1 = np.ascontiguousarray(np.asarray(g["ip1"], dtype=np.float64).ravel())
    Kc = np.ascontiguousarray(K.ravel())
    v = np.full(n, np.nan)
    info = np.zeros(8 * R, dtype=np.uint64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    old = {k: os.environ.get(k) for k in ("DKT_FAMILIES", "DKT_DIST_OVERLAP")}
    os.environ["DKT_FAMILIES"] = families
    os.environ["DKT_DIST_OVERLAP"] = overlap
    try:
        L = _lib()
        rc = L.emu_dist_matvec(dim, order, md, 1 if case["sfc"] == "hilbert" else 0, p(xyz), p(lev), len(lev), p(ip0), p(ip1), R, 1, p(Kc),
                               alpha, 1, 0.7, p(u), p(v), n, p(info), p2p, 1 if (p2p and R % 2) else 0)
    finally:
        for k, val in old.items():
            if val is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = val
    assert rc == 0, L.emu_last_error().decode()
    info = info.reshape(R, 8)
    assert int(info[:, 0].sum()) == n, "every node has exactly one owner"
    assert int(info[:, 2].sum()) == len(t.mv_lev), "every visited element belongs to exactly one rank"
    assert all(int(x) == int(overlap) for x in info[:, 6])
    assert np.abs(v - ref).max() <= TOL * np.abs(ref).max()
