"""Multi-GPU check, launched as:  torchrun --nproc-per-node N tests/dist_check.py
Every rank builds the partitioned DA of the same tree; the owned pieces of v = A u are gathered on
rank 0 (in single-rank DA order) and compared with (a) the oracle on a small tree and (b) the
single-GPU result on a larger one."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "dendro-kt_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import dkt  # noqa: E402


def bcast_id(rank):
    t = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(dkt.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, 0)
    return bytes(t.cpu().numpy().tobytes())


def gathered_matvec(da, op, u_global, n_global, rank, world, **kw):
    ids = torch.from_numpy(da.owned_ids().astype(np.int64)).cuda()
    u_loc = torch.from_numpy(u_global).cuda()[ids].contiguous()
    v_loc = da.matvec(op, u_loc, **kw)
    torch.cuda.synchronize()
    # the in-place ghosted-vector path must give the same owned values
    ug = torch.zeros(da.n_nodes + da.n_ghost_nodes, dtype=torch.float64, device="cuda")
    ug[:da.n_nodes] = u_loc
    vg = torch.empty_like(ug)
    da.matvec(op, ug, vg, ghosted=True, **kw)
    torch.cuda.synchronize()
    assert float((vg[:da.n_nodes] - v_loc).abs().max()) <= 1e-12 * max(float(v_loc.abs().max()), 1e-300)
    full = torch.zeros(n_global, dtype=torch.float64, device="cuda")
    full[ids] = v_loc
    owned = torch.zeros(n_global, dtype=torch.float64, device="cuda")
    owned[ids] = 1.0
    dist.all_reduce(full)
    dist.all_reduce(owned)
    assert float(owned.min()) == 1.0 and float(owned.max()) == 1.0, "every node must have exactly one owner"
    return full.cpu().numpy()


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    import cases
    import flat
    ok = True
    # (a) small trees against the oracle, all element kinds (regular, hanging, phantom, p=2, Dirichlet)
    for name in ("ball-d4-p1-morton-5", "gauss-d4-p1-morton", "ball-d3-p2-morton-5", "gauss-d3-p1-hilbert", "ex1-d2-p1-morton-6"):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from test_oracle import load_case
        case = load_case(name)
        g = case["golden"]
        sfc = dkt.SFC_HILBERT if case["sfc"] == "hilbert" else dkt.SFC_MORTON
        da = dkt.DA(case["xyz"], case["lev"], case["dim"], case["order"], case["max_depth"], sfc=sfc, ip0=g["ip0"], ip1=g["ip1"],
                    rank=rank, nranks=world, nccl_id=bcast_id(rank))
        n = len(g["node_lev"])
        assert da.n_global_nodes == n
        K = cases.dense_operator(case["dim"], case["order"])
        u = cases.input_vector(n)
        for diri, want in ((False, g["v_dense"]), (True, g["v_dense_diri"])):
            v = gathered_matvec(da, dkt.Operator.dense(K, float(g["alpha"]), dirichlet=diri), u, n, rank, world, scale=float(g["scale"]))
            err = np.abs(v - want).max() / np.abs(want).max()
            if rank == 0:
                print("%-24s ranks=%d owned=%d ghost=%d dirichlet=%d rel err vs reference %.2e" % (name, world, da.n_nodes,
                                                                                                  da.n_ghost_nodes, diri, err))
            ok &= err < 1e-12
        if case["order"] == 1:  # the sibling-family kernel under a partition: the reference's Laplacian vectors
            Kl = dkt.operators.laplace_kref(case["dim"], 1)
            for diri, want in ((False, g["v_lap"]), (True, g["v_lap_diri"])):
                v = gathered_matvec(da, dkt.Operator.dense(Kl, case["dim"] - 2.0, dirichlet=diri), u, n, rank, world, scale=0.7)
                err = np.abs(v - want).max() / np.abs(want).max()
                if rank == 0:
                    print("%-24s ranks=%d laplacian dirichlet=%d rel err vs reference %.2e" % (name, world, diri, err))
                ok &= err < 1e-12
        da.close()
    # (b) a larger tree against the single-GPU path
    dim, md = 4, 10
    xyz, lev = dkt.trees.moving_ball_tree(dim, 6, md, use_torch=True)
    K = dkt.operators.laplace_kref(dim, 1)
    op = dkt.Operator.dense(K, dim - 2.0)
    da1 = dkt.DA(xyz, lev, dim, 1, md)
    n = da1.n_nodes
    u = cases.input_vector(n)
    v1 = da1.matvec(op, u)
    nx1, nl1 = da1.nodes()
    b1 = da1.boundary_ids()
    da1.close()
    daN = dkt.DA(xyz, lev, dim, 1, md, rank=rank, nranks=world, nccl_id=bcast_id(rank))
    vN = gathered_matvec(daN, op, u, n, rank, world)
    err = np.abs(vN - v1).max() / np.abs(v1).max()
    if rank == 0:
        print("4-D ball level 6: %d elements, %d nodes, ranks=%d: rel diff vs single GPU %.2e" % (da1.n_elem, n, world, err))
    ok &= err < 1e-12
    # DA getters of a partitioned DA: node coordinates of the owned segment and the owned boundary nodes are the single-rank ones
    nxN, nlN = daN.nodes()
    oid = daN.owned_ids().astype(np.int64)
    okg = len(nlN) == daN.n_nodes + daN.n_ghost_nodes and np.array_equal(nxN[:daN.n_nodes], nx1[oid]) and np.array_equal(nlN[:daN.n_nodes], nl1[oid])
    okg &= np.array_equal(np.sort(oid[daN.boundary_ids().astype(np.int64)]), np.sort(np.intersect1d(oid, b1.astype(np.int64))))
    if rank == 0:
        print("partitioned DA getters (nodes, boundary ids):", "ok" if okg else "MISMATCH")
    ok &= bool(okg)
    # (c) the ghost exchanges on their own (dkt_ghost_read/write): ghost copies equal the owners' values; a write-back of ones
    # adds, to every owned node, the number of ranks that ghost it
    ids = torch.from_numpy(daN.owned_ids().astype(np.int64)).cuda()
    w = torch.zeros(daN.n_nodes + daN.n_ghost_nodes, dtype=torch.float64, device="cuda")
    w[:daN.n_nodes] = ids.to(torch.float64)
    torch.cuda.synchronize()  # the DA works on its own stream
    daN.ghost_read(w)
    torch.cuda.synchronize()
    full = torch.zeros(n, dtype=torch.float64, device="cuda")
    full[ids] = 1.0
    dist.all_reduce(full)
    gids = w[daN.n_nodes:].to(torch.int64)
    okc = bool(((gids >= 0) & (gids < n)).all()) and bool((full[gids] == 1.0).all()) and not bool(torch.isin(gids, ids).any())
    w2 = torch.ones_like(w)
    torch.cuda.synchronize()
    daN.ghost_write(w2)
    torch.cuda.synchronize()
    cnt = torch.zeros(n, dtype=torch.float64, device="cuda")
    cnt[gids] += 0.0
    cnt.index_add_(0, gids, torch.ones_like(gids, dtype=torch.float64))
    dist.all_reduce(cnt)
    okc &= bool((w2[:daN.n_nodes] == 1.0 + cnt[ids]).all())
    if rank == 0:
        print("ghost read/write entry points:", "ok" if okc else "MISMATCH")
    ok &= okc
    daN.close()
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_CHECK", "PASS" if float(t) == 1.0 else "FAIL")
    dist.destroy_process_group()
    sys.exit(0 if float(t) == 1.0 else 1)


if __name__ == "__main__":
    main()
