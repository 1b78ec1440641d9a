"""The peer-memory ghost exchange (DKT_DIST_P2P, dendro-kt_b200/csrc/dkt_dist.cu: k_p2p_*) on ONE GPU: the dry-run
DAs of all ranks of a partition live in this process, dkt_p2p_attach_local wires their exchange buffers directly, and
the ranks' matvecs run concurrently on their own streams - same kernels, flags and protocol as between processes,
without IPC and NCCL.  Opt-in (DKT_TEST_P2P=1) and for small partitions only: on ONE GPU the spinning wait kernels of one rank
can fill the device before the other rank's put kernels are resident (round 2: the level-5 tree deadlocks and the 30 s
time-out traps), which cannot happen between GPUs.  The protocol between processes is covered by tests/test_gpu_dist.py
(torchrun, DKT_DIST_P2P=1), green on 2 B200s."""
import os

import numpy as np
import pytest

import cases

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("DKT_TEST_P2P") != "1", reason="opt-in: DKT_TEST_P2P=1")]
TOL = 1e-12


@pytest.mark.parametrize("nranks", [2, 3, 8])
@pytest.mark.parametrize("families", ["0", "1"])
def test_p2p_exchange_in_one_process(dkt, nranks, families):
    import torch
    dim, md = 4, 10
    old = os.environ.get("DKT_FAMILIES")
    os.environ["DKT_FAMILIES"] = families
    os.environ.pop("DKT_P2P_CHECK", None)  # it synchronises inside dkt_matvec: the ranks must be enqueued back to back here
    try:
        xyz, lev = dkt.trees.moving_ball_tree(dim, 5, md)
        op = dkt.Operator.dense(dkt.operators.laplace_kref(dim, 1), dim - 2.0, dirichlet=True)
        da1 = dkt.DA(xyz, lev, dim, 1, md)
        n = da1.n_nodes
        u = cases.input_vector(n)
        v_ref = da1.matvec(op, u)
        da1.close()
        das = [dkt.DA(xyz, lev, dim, 1, md, rank=r, nranks=nranks, dryrun=True) for r in range(nranks)]
        dkt.p2p_attach_local(das)
        ids = [torch.from_numpy(d.owned_ids().astype(np.int64)).cuda() for d in das]
        ug = torch.from_numpy(u).cuda()
        for rep in range(3):  # epochs advance together on all ranks
            ins = [ug[i].contiguous() for i in ids]
            outs = [torch.empty_like(x) for x in ins]
            torch.cuda.synchronize()
            for d, x, y in zip(das, ins, outs):
                d.matvec(op, x, y)  # enqueued on the DA's own stream; returns without waiting
            torch.cuda.synchronize()
            full = torch.zeros(n, dtype=torch.float64, device="cuda")
            for i, y in zip(ids, outs):
                full[i] = y
            v = full.cpu().numpy()
            assert np.abs(v - v_ref).max() <= TOL * np.abs(v_ref).max()
        for d in das:
            d.close()
    finally:
        if old is None:
            os.environ.pop("DKT_FAMILIES", None)
        else:
            os.environ["DKT_FAMILIES"] = old
