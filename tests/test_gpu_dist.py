"""Multi-GPU parity under pytest: launches tests/dist_check.py with torchrun on every power-of-two rank count the box offers
(2, 4, 8), once with the NCCL ghost exchange and once with the peer-memory exchange (DKT_DIST_P2P=1).  The script gathers
the partitioned results to the single-rank node order and compares them with the golden vectors taken from the reference.
Skipped on a box with one GPU."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("p2p", ["0", "1"])
@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_partitioned_matvec_matches_reference(dkt, nranks, p2p):
    if _ngpu() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    env = dict(os.environ)
    env["DKT_DIST_P2P"] = p2p
    if p2p == "1":
        env["DKT_P2P_CHECK"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "DIST_CHECK PASS" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])


@pytest.mark.parametrize("name,nranks", [("ball-d4-p1-morton-5", 2), ("gauss-d4-p1-morton", 2), ("ball-d3-p2-morton-5", 2), ("ball-d4-p1-morton-5", 4)])
def test_cpp_distributed_api(dkt, tmp_path, name, nranks):
    """The reference's distributed C++ API through the drop-in headers, one process per GPU (tests/cpp/test_dist_api.cpp):
    ot::DA built by every rank, feMatrix::matVec through the elementalMatVec callback with and without Dirichlet hooks,
    readFromGhost / writeToGhosts on host vectors.  Gathered to the single-rank order and compared with the reference's
    golden vectors."""
    if _ngpu() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    import numpy as np
    import cases
    from test_oracle import load_case
    inc, libdir = os.path.join(ROOT, "dendro-kt_b200", "include"), os.path.join(ROOT, "dendro-kt_b200", "lib")
    exe = str(tmp_path / "test_dist_api")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-Werror", "-I", inc, os.path.join(ROOT, "tests", "cpp", "test_dist_api.cpp"), "-o", exe,
                           "-L", libdir, "-ldkt", "-Wl,-rpath," + libdir])
    case = load_case(name)
    g = case["golden"]
    dim, order, md = case["dim"], case["order"], case["max_depth"]
    n = len(g["node_lev"])
    K, u = cases.dense_operator(dim, order), cases.input_vector(n)
    d = str(tmp_path)
    case["xyz"].astype(np.uint32).tofile(os.path.join(d, "elem_xyz.bin"))
    case["lev"].astype(np.uint8).tofile(os.path.join(d, "elem_lev.bin"))
    K.astype(np.float64).tofile(os.path.join(d, "K.bin"))
    u.astype(np.float64).tofile(os.path.join(d, "u.bin"))
    np.array([float(g["alpha"]), float(g["scale"])], dtype=np.float64).tofile(os.path.join(d, "params.bin"))
    import flat
    t = cases.oracle_tables_for(case)
    for diri in (0, 1):
        procs = []
        for r in range(nranks):
            env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(nranks), CUDA_VISIBLE_DEVICES=str(r), DKT_NCCL_ID_FILE=os.path.join(d, "ncclid%d" % diri))
            procs.append(subprocess.Popen([exe, str(dim), str(order), str(md), str(diri), d], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env))
        outs = [p.communicate(timeout=600) for p in procs]
        assert all(p.returncode == 0 for p in procs), outs
        v = np.full(n, np.nan)
        owners = np.zeros(n, dtype=np.int64)
        per_rank = []
        for r in range(nranks):
            ids = np.fromfile(os.path.join(d, "rank%d_ids.bin" % r), dtype=np.uint32)
            v[ids] = np.fromfile(os.path.join(d, "rank%d_v.bin" % r), dtype=np.float64)
            owners[ids] += 1
            per_rank.append(ids)
        assert np.all(owners == 1), "every node has exactly one owner"
        # exact interpolation on both sides (the host RefElement's matrices differ from the reference's LAPACK ones in the last bits)
        want = flat.matvec(t, u, K, alpha=float(g["alpha"]), scale=float(g["scale"]), dirichlet=bool(diri))
        assert np.abs(v - want).max() <= 1e-12 * np.abs(want).max()
        # node coordinates of the local vectors and the two ghost exchanges
        ghosted_by = np.zeros(n, dtype=np.int64)
        ghost_ids = []
        for r in range(nranks):
            ids = per_rank[r]
            gw = np.fromfile(os.path.join(d, "rank%d_ghost.bin" % r), dtype=np.float64)
            ntot = len(gw) // 2
            gids = gw[len(ids):ntot].astype(np.int64)  # what readFromGhost delivered: the owners' single-rank positions
            assert np.all((gids >= 0) & (gids < n)) and not np.isin(gids, ids).any()
            nodes = np.fromfile(os.path.join(d, "rank%d_nodes.bin" % r), dtype=np.uint32).reshape(-1, dim + 1)
            allids = np.concatenate([ids.astype(np.int64), gids])
            assert np.array_equal(nodes[:, :dim], g["node_xyz"][allids]) and np.array_equal(nodes[:, dim], g["node_lev"][allids])
            np.add.at(ghosted_by, gids, 1)
            ghost_ids.append((ids, gw[ntot:ntot + len(ids)]))
        for ids, w in ghost_ids:  # writeToGhosts of ones: 1 + number of ranks that ghost the node
            assert np.array_equal(w, 1.0 + ghosted_by[ids])
