#!/usr/bin/env python
"""Throughput of the other BASELINE.json configs on one GPU (parity cases, not the headline bench):
C1 2-D p=1 adaptive, C2 3-D uniform p=1 (L=8 and L=9 bracket ~5e7 elements), C4 3-D p=2 adaptive,
plus one CG solve.  Prints one JSON line per config."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dendro-kt_b200"))
import dkt  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def run(name, xyz, lev, dim, order, md, steps=10, warm=3):
    t0 = time.time()
    da = dkt.DA(xyz, lev, dim, order, md)
    tb = time.time() - t0
    K = dkt.operators.laplace_kref(dim, order)
    op = dkt.Operator.dense(K, dim - 2.0)
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    da.set_stream(st.cuda_stream)
    u = torch.rand(da.n_nodes, dtype=torch.float64, device="cuda")
    v = torch.empty_like(u)
    for _ in range(warm):
        da.matvec(op, u, v)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        da.matvec(op, u, v)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ach = da.alg_bytes / (ms * 1e-3) / 1e9
    print(json.dumps({"config": name, "dim": dim, "order": order, "n_elem": da.n_elem, "n_nodes": da.n_nodes,
                      "n_hanging_elem": da.n_hanging, "tree_class": da.tree_class, "ms_per_matvec": ms,
                      "dof_per_s": da.n_nodes / (ms * 1e-3), "elem_per_s": da.n_elem / (ms * 1e-3), "alg_GBs": ach,
                      "frac_of_hbm_peak": ach / PEAK, "da_build_s": round(tb, 3)}), flush=True)
    return da, op


def main():
    which = sys.argv[1:] or ["c1", "c2a", "c2b", "c3pts", "c4", "cg"]
    if "c3pts" in which:
        # SURVEY 8d C3 as the reference builds it: shell points + level-3 guard -> distTreeBalancing -> DA -> matvec, all on the GPU
        pts = torch.from_numpy(dkt.trees.shell_points(4, 500000, 12, guard_level=3).astype(np.int64)).to(torch.int32).cuda()
        torch.cuda.synchronize()
        t0 = time.time()
        tr = dkt.Tree(pts, 4, 12, 1, balance=True)
        torch.cuda.synchronize()
        tt = time.time() - t0
        x, l = tr.export_torch()
        tr.close()
        print(json.dumps({"config": "C3 tree pipeline", "n_points": int(pts.shape[0]), "leaves": int(l.numel()), "tree_build_s": round(tt, 4)}), flush=True)
        run("C3 4-D p=1, shell points + guard through dkt_tree_from_points (distTreeBalancing)", x, l, 4, 1, 12)[0].close()
    if "c1" in which:
        x, l = dkt.trees.moving_ball_tree(2, 14, 16, use_torch=True)
        run("C1 2-D p=1 adaptive ball level 14", x, l, 2, 1, 16)[0].close()
    if "c2a" in which:
        x, l = dkt.trees.uniform_tree_torch(3, 8, 12)
        run("C2 3-D uniform p=1 L=8", x, l, 3, 1, 12)[0].close()
    if "c2b" in which:
        x, l = dkt.trees.uniform_tree_torch(3, 9, 12)
        run("C2 3-D uniform p=1 L=9", x, l, 3, 1, 12)[0].close()
    if "c4" in which:
        x, l = dkt.trees.moving_ball_tree(3, 10, 14, use_torch=True)
        run("C4 3-D p=2 adaptive ball level 10", x, l, 3, 2, 14)[0].close()
    if "cg" in which:
        x, l = dkt.trees.uniform_tree_torch(3, 7, 12)
        da = dkt.DA(x, l, 3, 1, 12)
        op = dkt.Operator.dense(dkt.operators.laplace_kref(3, 1), 1.0, dirichlet=True)
        xt = torch.rand(da.n_nodes, dtype=torch.float64, device="cuda")
        xt[torch.from_numpy(da.boundary_ids().astype(np.int64)).cuda()] = 0.0
        b = da.matvec(op, xt)
        torch.cuda.synchronize()
        t0 = time.time()
        xs, it, resid, ok = da.cg_solve(op, b, max_iter=2000, tol=1e-10)
        torch.cuda.synchronize()
        dt = time.time() - t0
        print(json.dumps({"config": "CG 3-D uniform L=7 Dirichlet Laplacian", "n_nodes": da.n_nodes, "iterations": it, "residual": resid,
                          "converged": ok, "seconds": dt, "ms_per_iteration": 1e3 * dt / max(it, 1),
                          "max_err_vs_manufactured": float((xs - xt).abs().max())}), flush=True)
        da.close()


if __name__ == "__main__":
    main()
