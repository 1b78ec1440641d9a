"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, generates SFC tables whose ORDER equals the reference's, and refuses to compute
without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

import cases
import flat
from test_oracle import load_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(dkt):
    hdr = open(os.path.join(ROOT, "include", "dkt.h")).read()
    declared = set(re.findall(r"\b(dkt_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"dkt_da"}
    assert declared, "no declarations found"
    L = dkt.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(dkt.DECLARED_SYMBOLS) == declared
    assert b"sm_100a" in L.dkt_version()


def test_no_cpu_fallback(dkt):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    xyz, lev = dkt.trees.uniform_tree(2, 2, 8)
    with pytest.raises(dkt.DktError, match="no CUDA device"):
        dkt.DA(xyz, lev, 2, 1, 8)


def test_invalid_arguments(dkt):
    xyz, lev = dkt.trees.uniform_tree(2, 2, 8)
    for kw in (dict(dim=5), dict(order=3), dict(max_depth=31), dict(order=0)):
        args = dict(dim=2, order=1, max_depth=8)
        args.update(kw)
        with pytest.raises(dkt.DktError):
            dkt.DA(xyz, lev, args["dim"], args["order"], args["max_depth"])


@pytest.mark.parametrize("dim", [2, 3, 4])
def test_morton_tables(dkt, dim):
    perm, inv, h = dkt.sfc_tables(dim, dkt.SFC_MORTON)
    assert perm.shape == (1, 1 << dim)
    assert np.array_equal(perm[0], np.arange(1 << dim)) and np.array_equal(inv[0], np.arange(1 << dim)) and not h.any()


@pytest.mark.parametrize("name", ["ex1-d2-p1-hilbert-5", "ex3-d3-p1-hilbert-3", "ex3-d4-p1-hilbert-3"])
def test_hilbert_tables_give_reference_order(dkt, name):
    """The library's generated Hilbert tables (own state numbering) must order cells exactly like
    the reference's KDhcurvedata tables (stored in the golden fixture)."""
    case = load_case(name)
    g = case["golden"]
    dim = case["dim"]
    ref = flat.SfcTables(dim, g["rot_perm"], g["rot_inv"], g["htab"])
    perm, inv, h = dkt.sfc_tables(dim, dkt.SFC_HILBERT)
    mine = flat.SfcTables(dim, perm, inv, h)
    # every state is a consistent permutation pair
    for r in range(len(perm)):
        assert np.array_equal(inv[r][perm[r]], np.arange(1 << dim))
    depth = {2: 5, 3: 4, 4: 3}[dim]
    xyz, lev = cases.example_tree(dim, 2, depth, 10)
    assert np.array_equal(flat.sort_elements(xyz, lev, 10, mine), flat.sort_elements(xyz, lev, 10, ref))
    # and on the adaptive fixture tree itself
    o = flat.sort_elements(case["xyz"], case["lev"], case["max_depth"], mine)
    assert np.array_equal(case["xyz"][o], g["elem_xyz"])


def test_procedural_trees_are_balanced(dkt):
    """Generators of dkt.trees: complete (volumes sum to 1) and 2:1 balanced incl. corners
    (checked through the node rule: on a balanced tree no lattice location mixes 3 levels)."""
    for dim, ml in ((2, 8), (3, 6), (4, 5)):
        md = 10
        xyz, lev = dkt.trees.moving_ball_tree(dim, ml, md)
        vol = np.sum(np.power(2.0, -dim * lev.astype(np.float64)))
        assert abs(vol - 1.0) < 1e-12
        t = flat.build_tables(xyz, lev, dim, 1, md)
        assert t.tree_class == "B"
        # neighbours across every lattice node differ by <= 1 level
        N = t.N
        lv = t.mv_lev.astype(np.int64)
        mx = np.zeros(len(t.node_lev), dtype=np.int64)
        mn = np.full(len(t.node_lev), 99, dtype=np.int64)
        f = t.e2n >= 0
        np.maximum.at(mx, t.e2n[f], np.repeat(lv, N).reshape(-1, N)[f])
        np.minimum.at(mn, t.e2n[f], np.repeat(lv, N).reshape(-1, N)[f])
        assert (mx - mn).max() <= 1


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/dkt.h must compile as C99 with no C++ or torch types."""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "dkt.h"\nint main(void){ dkt_op op; dkt_sizes s; (void)op; (void)s; return DKT_OK; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)])


def test_reference_cell_operators(dkt):
    """Host-side operator setup: symmetric, constants in the Laplacian's null space, mass sums to the cell
    volume, and at order 1 both are diagonal in the Walsh-Hadamard basis (the form the kernels exploit)."""
    for dim in (2, 3, 4):
        for order in (1, 2):
            K, M = dkt.operators.laplace_kref(dim, order), dkt.operators.mass_kref(dim, order)
            N = (order + 1) ** dim
            assert K.shape == (N, N) and np.abs(K - K.T).max() < 1e-13 and np.abs(M - M.T).max() < 1e-13
            assert np.abs(K @ np.ones(N)).max() < 1e-12 and abs(M.sum() - 1.0) < 1e-12
            if order == 1:
                H = np.array([[(-1.0) ** bin(i & j).count("1") for j in range(N)] for i in range(N)])
                for A in (K, M):
                    D = H @ A @ H / N
                    assert np.abs(D - np.diag(np.diag(D))).max() <= 1e-13 * np.abs(D).max()


def test_reference_arm_json_contract():
    """`bench.py --impl reference` (the reference's CPU matvec through oracle/_ref) prints the contract's keys."""
    import json
    import subprocess
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dktref
    if not dktref.available("morton"):
        pytest.skip("oracle/_ref not built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-procs", "2",
                          "--ref-level", "4"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "DOF/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] == 2
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and line["config"]["mode"] in ("replicas", "distributed")
    assert line["reference_replicas"]["replicas"] == 2 and line["reference_replicas"]["value"] > 0
    mpi = line["reference_mpi"]  # ONE distributed job over oracle/shim_mp (or the reason it could not run)
    assert ("value" in mpi and mpi["ranks"] == 2 and mpi["n_nodes"] == line["config"]["n_nodes"]) or "unavailable" in mpi
