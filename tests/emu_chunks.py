"""Python side of the CUDA-on-CPU emulation of dkt_chunks.cu (tests/emu/): builds libdkt_emu.so with g++ and
feeds it the ORACLE's flat tables.  Test infrastructure only - see tests/emu/cuda_emu.h."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "dendro-kt_b200", "csrc")
LIB = os.path.join(EMU, "_build", "libdkt_emu.so")
INVALID = 0xFFFFFFFF
_lib = None


def build():
    import emu_build
    return emu_build.build()


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.emu_last_error.restype = C.c_char_p
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def matvec(t, u, max_depth, kref=None, alpha=0.0, scale=1.0, ip0=None, ip1=None, dirichlet=False, families=1, flags=0, order=0,
           phased_seed=None, kron=None):
    """Emulated dkt chunk path on the oracle's FlatTables `t`.  Returns (v, sets) where sets lists
    (kind, rows, slots per unit, units, chunks, units per chunk, max nodes per chunk, total chunk nodes, phase); kind 2 = sibling
    families (families=0 builds per-element sets only, DKT_FAMILIES=0).
    phased_seed: emulate a partitioned DA with comm/compute overlap - a random third of the elements counts as
    "boundary", the lists are ordered [interior | boundary] like dkt_dist.cu does, visit positions get an offset,
    and the phases (interior, boundary) run one after the other."""
    import flat
    dim, N = t.dim, t.N
    nMv = len(t.mv_lev)
    hang = np.zeros(nMv, dtype=bool)
    hang[t.hang_idx] = True
    nReg = int((~hang).sum())
    phased, nRegInt, nHangInt, src0 = 0, nReg, nMv - nReg, 0
    if phased_seed is None:
        perm = np.concatenate([np.nonzero(~hang)[0], np.nonzero(hang)[0]])  # regular first, hanging after (dkt_build.cu)
    else:
        bd = np.random.default_rng(phased_seed).random(nMv) < 0.33
        parts = [np.nonzero(~hang & ~bd)[0], np.nonzero(~hang & bd)[0], np.nonzero(hang & ~bd)[0], np.nonzero(hang & bd)[0]]
        perm = np.concatenate(parts)
        phased, nRegInt, nHangInt, src0 = 1, len(parts[0]), len(parts[2]), 1000
    e2n = np.ascontiguousarray(np.where(t.e2n < 0, INVALID, t.e2n)[perm].astype(np.uint32))
    hpos = np.full(nMv, -1, dtype=np.int64)
    hpos[t.hang_idx] = np.arange(len(t.hang_idx))
    pn = t.pnode[hpos[perm[nReg:]]] if len(t.hang_idx) else np.zeros((0, N), dtype=np.int64)
    pnode = np.ascontiguousarray(np.where(pn < 0, INVALID, pn).astype(np.uint32))
    xyz = np.ascontiguousarray(t.mv_xyz[perm].astype(np.uint32))
    lev = np.ascontiguousarray(t.mv_lev[perm].astype(np.uint8))
    src = np.ascontiguousarray((perm + src0).astype(np.uint32))
    nNodes = len(t.node_lev)
    isbdy = np.zeros(nNodes, dtype=np.uint8)
    isbdy[t.bdy_ids] = 1
    if ip0 is None:
        ip0, ip1 = flat.default_interp(t.order)
    ip0 = np.ascontiguousarray(np.asarray(ip0, dtype=np.float64).ravel())
    ip1 = np.ascontiguousarray(np.asarray(ip1, dtype=np.float64).ravel())
    kr = None if kref is None else np.ascontiguousarray(np.asarray(kref, dtype=np.float64).ravel())
    op_kind = 0 if kr is None else 1
    if kron is not None:  # sum-factorised operator: (terms, dim, M, M)
        kr = np.ascontiguousarray(np.asarray(kron, dtype=np.float64).ravel())
        op_kind = 2 | (len(kron) << 4)
    u = np.ascontiguousarray(np.asarray(u, dtype=np.float64))
    out = np.full(nNodes, np.nan)
    info = np.zeros(128, dtype=np.uint64)
    old = {k: os.environ.get(k) for k in ("DKT_FAMILIES", "DKT_EMU_ORDER")}
    os.environ["DKT_FAMILIES"] = str(families)
    os.environ["DKT_EMU_ORDER"] = str(order)
    try:
        L = lib()
        # the fiber order is read once per process: set it through the exported state instead
        L.emu_set_order(C.c_int(order))
        rc = L.emu_matvec(C.c_int(dim), C.c_int(t.order), C.c_int(max_depth), C.c_uint64(nMv), C.c_uint64(nReg), C.c_uint64(nNodes),
                          _p(e2n), _p(pnode), _p(xyz), _p(lev), _p(src), _p(isbdy), _p(ip0), _p(ip1),
                          C.c_int(op_kind), None if kr is None else _p(kr), C.c_double(alpha), C.c_int(int(dirichlet)),
                          _p(u), _p(out), C.c_double(scale), C.c_uint(flags), _p(info),
                          C.c_int(phased), C.c_uint64(nRegInt), C.c_uint64(nHangInt), C.c_uint64(src0))
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    if rc != 0:
        raise RuntimeError("emu_matvec rc=%d: %s" % (rc, L.emu_last_error().decode()))
    sets = [tuple(int(x) for x in info[8 * i:8 * i + 7]) + (int(info[8 * i + 7]) & ((1 << 56) - 1), int(info[8 * i + 7]) >> 56)
            for i in range(16) if info[8 * i + 4]]
    return out, sets
