"""N>1 logic on CPU: two gloo processes run the partitioned matvec protocol (ghost read, local
matvec on the rank's elements, ghost write-back + accumulate) with the oracle as the local
operator, and the gathered result must equal the single-rank oracle result.  This is the host-side
model of dendro-kt_b200/csrc/dkt_dist.cu; the GPU version is checked by tests/dist_check.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
import flat
import partition_model
from test_oracle import load_case


def _local_matvec(t, part, u_local, K, alpha, scale, ip0, ip1):
    """The oracle restricted to one rank's elements, on its local (owned + ghost) vector."""
    import copy
    sub = copy.copy(t)
    e = part["elems"]
    hang_pos = {int(x): i for i, x in enumerate(t.hang_idx)}
    sub.e2n = np.where(t.e2n[e] >= 0, part["g2l"][np.clip(t.e2n[e], 0, None)], -1)
    sub.mv_lev = t.mv_lev[e]
    is_h = np.array([int(x) in hang_pos for x in e])
    sub.hang_idx = np.nonzero(is_h)[0]
    rows = np.array([hang_pos[int(x)] for x in e[is_h]], dtype=np.int64)
    pn = t.pnode[rows] if len(rows) else np.zeros((0, t.N), dtype=np.int64)
    sub.pnode = np.where(pn >= 0, part["g2l"][np.clip(pn, 0, None)], -1)
    sub.child = t.child[rows] if len(rows) else np.zeros(0, dtype=np.int64)
    sub.bdy_ids = np.zeros(0, dtype=np.int64)
    return flat.matvec(sub, u_local, K, alpha=alpha, scale=scale, ip0=ip0, ip1=ip1)


def _worker(rank, world, port, name, result_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = load_case(name)
    g = case["golden"]
    t = cases.oracle_tables_for(case)
    parts = partition_model.partition(t, world)
    me = parts[rank]
    n = len(t.node_lev)
    K = cases.dense_operator(case["dim"], case["order"])
    u = cases.input_vector(n)
    n_owned = len(me["owned"])
    u_local = np.zeros(len(me["local"]))
    u_local[:n_owned] = u[me["owned"]]
    # ghost read: owners -> ghosts (include/oda.tcc:212-315)
    reqs, recv_bufs = [], {}
    for p in range(world):
        if p == rank:
            continue
        if len(me["sends"][p]):
            reqs.append(dist.isend(torch.from_numpy(u_local[me["g2l"][me["sends"][p]]].copy()), p))
        if len(me["ghosts"][p]):
            recv_bufs[p] = torch.zeros(len(me["ghosts"][p]), dtype=torch.float64)
            reqs.append(dist.irecv(recv_bufs[p], p))
    for r in reqs:
        r.wait()
    for p, b in recv_bufs.items():
        u_local[me["g2l"][me["ghosts"][p]]] = b.numpy()
    v_local = _local_matvec(t, me, u_local, K, float(g["alpha"]), float(g["scale"]), g["ip0"], g["ip1"])
    # ghost write: partial sums -> owners, accumulated (include/oda.tcc:319-435)
    reqs, recv_bufs = [], {}
    for p in range(world):
        if p == rank:
            continue
        if len(me["ghosts"][p]):
            reqs.append(dist.isend(torch.from_numpy(v_local[me["g2l"][me["ghosts"][p]]].copy()), p))
        if len(me["sends"][p]):
            recv_bufs[p] = torch.zeros(len(me["sends"][p]), dtype=torch.float64)
            reqs.append(dist.irecv(recv_bufs[p], p))
    for r in reqs:
        r.wait()
    for p, b in recv_bufs.items():
        np.add.at(v_local, me["g2l"][me["sends"][p]], b.numpy())
    full = torch.zeros(n, dtype=torch.float64)
    full[torch.from_numpy(me["owned"])] = torch.from_numpy(v_local[:n_owned])
    dist.all_reduce(full)
    if rank == 0:
        np.save(result_path, full.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["ex1-d3-p1-morton-4", "gauss-d3-p1-morton", "ex3-d2-p2-morton-3"])
def test_two_rank_protocol_matches_single_rank(tmp_path, name):
    world = 2
    port = 29600 + (hash(name) % 200)
    result = str(tmp_path / "v.npy")
    mp.spawn(_worker, args=(world, port, name, result), nprocs=world, join=True)
    v = np.load(result)
    g = load_case(name)["golden"]
    assert np.abs(v - g["v_dense"]).max() <= 1e-12 * np.abs(g["v_dense"]).max()


def test_partition_model_invariants():
    case = load_case("gauss-d4-p1-morton")
    t = cases.oracle_tables_for(case)
    for world in (2, 3, 8):
        parts = partition_model.partition(t, world)
        owned = np.concatenate([p["owned"] for p in parts])
        assert len(owned) == len(t.node_lev) and len(np.unique(owned)) == len(owned)  # exactly one owner each
        assert sum(len(p["elems"]) for p in parts) == len(t.mv_lev)
        for r, p in enumerate(parts):
            for q in range(world):  # what r sends to q is exactly what q ghosts from r, same order
                assert np.array_equal(p["sends"][q], parts[q]["ghosts"][r])
