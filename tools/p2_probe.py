#!/usr/bin/env python
"""3-D order-2 matvec (BASELINE config C4) with the identity, the dense and (if the library has it) the Kronecker-sum
operator: where does the time go?  One JSON line per operator."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dendro-kt_b200"))
import dkt  # noqa: E402

level = int(sys.argv[1]) if len(sys.argv) > 1 else 10
x, l = dkt.trees.moving_ball_tree(3, level, 14, use_torch=True)
da = dkt.DA(x, l, 3, 2, 14)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
da.set_stream(st.cuda_stream)
u = torch.rand(da.n_nodes, dtype=torch.float64, device="cuda")
v = torch.empty_like(u)
ops = [("identity", dkt.Operator.identity()), ("dense laplace", dkt.Operator.dense(dkt.operators.laplace_kref(3, 2), 1.0))]
if hasattr(dkt.Operator, "kron"):
    ops.append(("kronecker laplace", dkt.Operator.kron(dkt.operators.laplace_terms(3, 2), 1.0)))
res = {}
for name, op in ops:
    for _ in range(3):
        da.matvec(op, u, v)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10):
        da.matvec(op, u, v)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    res[name] = v.clone()
    print(json.dumps({"operator": name, "n_elem": da.n_elem, "n_nodes": da.n_nodes, "n_hanging_elem": da.n_hanging, "ms_per_matvec": ms,
                      "dof_per_s": da.n_nodes / (ms * 1e-3), "alg_GBs": da.alg_bytes / (ms * 1e-3) / 1e9}), flush=True)
if "kronecker laplace" in res:
    d = (res["kronecker laplace"] - res["dense laplace"]).abs().max().item() / res["dense laplace"].abs().max().item()
    print(json.dumps({"kronecker_vs_dense_max_rel_diff": d}))
