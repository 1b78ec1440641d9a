"""CPU test of the peer-memory ghost exchange (dendro-kt_b200/csrc/dkt_p2p.cuh, opt-in DKT_DIST_P2P): its kernels and
pointer tables run under the CUDA-on-CPU emulation (tests/emu) with all ranks of a partition in one process, on the
send/receive lists of the NumPy partition model.  Ghost values must equal the owners' values, owners must receive the
sum of the ghosting ranks' partial sums - for several epochs, since the buffers are re-used without double buffering."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cases
import partition_model
from test_oracle import load_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "dendro-kt_b200", "csrc")
LIB = os.path.join(EMU, "_build", "libdkt_emu_p2p.so")


def _lib():
    import emu_build
    L = C.CDLL(emu_build.build())
    L.emu_p2p_error.restype = C.c_char_p
    return L


@pytest.mark.parametrize("name,R", [("ball-d2-p1-morton-7", 2), ("ball-d3-p1-morton-6", 3), ("ex3-d4-p1-morton-3", 5), ("gauss-d3-p1-morton", 8)])
def test_emulated_p2p_exchange(name, R):
    case = load_case(name)
    t = cases.oracle_tables_for(case)
    parts = partition_model.partition(t, R)
    n = len(t.node_lev)
    rng = np.random.default_rng(R)
    u = rng.uniform(-1, 1, n)
    send_off = np.zeros((R, R + 1), dtype=np.uint64)
    recv_off = np.zeros((R, R + 1), dtype=np.uint64)
    n_owned = np.zeros(R, dtype=np.uint64)
    sidx, sidx_base, vec_base = [], np.zeros(R, dtype=np.uint64), np.zeros(R, dtype=np.uint64)
    ins, outs, partial = [], [], []
    off_s = off_v = 0
    for r, me in enumerate(parts):
        n_owned[r] = len(me["owned"])
        send_off[r, 1:] = np.cumsum([len(me["sends"][p]) for p in range(R)])
        recv_off[r, 1:] = np.cumsum([len(me["ghosts"][p]) for p in range(R)])
        s = np.concatenate([me["g2l"][me["sends"][p]] for p in range(R)]).astype(np.uint32)
        assert (s < n_owned[r]).all()
        sidx.append(s)
        sidx_base[r], vec_base[r] = off_s, off_v
        off_s += len(s)
        nl = len(me["local"])
        off_v += nl
        x = np.full(nl, np.nan)
        x[:len(me["owned"])] = u[me["owned"]]
        ins.append(x)
        pr = rng.uniform(-1, 1, nl)  # what the local matvec left: owned entries and ghost partial sums
        partial.append(pr)
        outs.append(pr.copy())
    sidx = np.ascontiguousarray(np.concatenate(sidx)) if off_s else np.zeros(1, dtype=np.uint32)
    in_all = np.ascontiguousarray(np.concatenate(ins))
    out_all = np.ascontiguousarray(np.concatenate(outs))
    epochs = 3
    L = _lib()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = L.emu_p2p_exchange(C.c_int(R), p(send_off), p(recv_off), p(n_owned), p(sidx), p(sidx_base), p(in_all), p(out_all), p(vec_base),
                            C.c_int(epochs))
    assert rc == 0, L.emu_p2p_error().decode()
    # expected: ghosts hold the owners' values; owners accumulated the ghosting ranks' partial sums once per epoch
    for r, me in enumerate(parts):
        a, no, nl = int(vec_base[r]), int(n_owned[r]), len(me["local"])
        assert np.array_equal(in_all[a:a + nl], u[me["local"]])
        add = np.zeros(n)
        for q, other in enumerate(parts):
            if q == r:
                continue
            g = other["ghosts"][r]  # nodes rank q ghosts that r owns
            if len(g):
                add[g] += partial[q][other["g2l"][g]]
        exp = partial[r][:no] + epochs * add[me["owned"]]
        assert np.abs(out_all[a:a + no] - exp).max() <= 1e-13
        # the ghost partial sums themselves stay where they were
        assert np.array_equal(out_all[a + no:a + nl], partial[r][no:])
