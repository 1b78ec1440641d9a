"""ctypes binding of oracle/_ref/libdktref_{morton,hilbert}.so - the reference itself, built by
oracle/build_ref.sh.  TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}

OP_IDENTITY, OP_DENSE, OP_HEATMAT = 0, 1, 2


def available(sfc="morton"):
    return os.path.exists(os.path.join(_HERE, "_ref", "libdktref_%s.so" % sfc))


def _lib(sfc):
    if sfc in _libs:
        return _libs[sfc]
    L = C.CDLL(os.path.join(_HERE, "_ref", "libdktref_%s.so" % sfc))
    vp, i32, i64, f64 = C.c_void_p, C.c_int, C.c_long, C.c_double
    L.dktref_init.argtypes = [i32, i32]
    L.dktref_tables.argtypes = [i32, vp, vp]
    L.dktref_tree_from_points.restype = vp
    L.dktref_tree_from_points.argtypes = [i32, vp, i64, i32, f64, i32]
    L.dktref_tree_from_elements.restype = vp
    L.dktref_tree_from_elements.argtypes = [i32, vp, vp, i64, i32]
    L.dktref_tree_example.restype = vp
    L.dktref_tree_example.argtypes = [i32, i32, i32, i32]
    L.dktref_tree_size.restype = i64
    L.dktref_tree_size.argtypes = [vp]
    L.dktref_tree_export.argtypes = [vp, vp, vp]
    L.dktref_tree_destroy.argtypes = [vp]
    L.dktref_da_create.restype = vp
    L.dktref_da_create.argtypes = [vp, i32]
    L.dktref_da_num_nodes.restype = i64
    L.dktref_da_num_nodes.argtypes = [vp]
    L.dktref_da_export_nodes.argtypes = [vp, vp, vp]
    L.dktref_da_boundary.restype = i64
    L.dktref_da_boundary.argtypes = [vp, vp]
    L.dktref_da_destroy.argtypes = [vp]
    L.dktref_refel.restype = f64
    L.dktref_refel.argtypes = [i32, i32, vp, vp, vp, vp, vp]
    L.dktref_matvec.restype = f64
    L.dktref_matvec.argtypes = [vp, i32, vp, f64, i32, vp, vp, f64, i32, i32, vp]
    L.dktref_heat_elemental.argtypes = [vp, vp, vp, vp]
    _libs[sfc] = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefTree:
    def __init__(self, ref, handle):
        self.ref, self.h = ref, handle

    def __len__(self):
        return int(self.ref.L.dktref_tree_size(self.h))

    def export(self):
        n = len(self)
        xyz = np.zeros((n, self.ref.dim), dtype=np.uint32)
        lev = np.zeros(n, dtype=np.uint8)
        self.ref.L.dktref_tree_export(self.h, _p(xyz), _p(lev))
        return xyz, lev

    def __del__(self):
        if self.h:
            self.ref.L.dktref_tree_destroy(self.h)
            self.h = None


class RefDA:
    def __init__(self, ref, handle, order):
        self.ref, self.h, self.order = ref, handle, order

    @property
    def num_nodes(self):
        return int(self.ref.L.dktref_da_num_nodes(self.h))

    def nodes(self):
        n = self.num_nodes
        xyz = np.zeros((n, self.ref.dim), dtype=np.uint32)
        lev = np.zeros(n, dtype=np.uint8)
        self.ref.L.dktref_da_export_nodes(self.h, _p(xyz), _p(lev))
        return xyz, lev

    def boundary_ids(self):
        n = int(self.ref.L.dktref_da_boundary(self.h, None))
        ids = np.zeros(n, dtype=np.uint32)
        self.ref.L.dktref_da_boundary(self.h, _p(ids))
        return ids

    def matvec(self, u, kind=OP_IDENTITY, Kref=None, alpha=0.0, dirichlet=False, scale=1.0, nwarm=0, niter=1):
        """Returns (v, seconds_per_call, eleOp_calls_per_matvec)."""
        u = np.ascontiguousarray(u, dtype=np.float64)
        v = np.zeros_like(u)
        K = None if Kref is None else np.ascontiguousarray(Kref, dtype=np.float64)
        nc = C.c_long(0)
        secs = self.ref.L.dktref_matvec(self.h, kind, None if K is None else _p(K), float(alpha), int(dirichlet), _p(u), _p(v),
                                        float(scale), nwarm, niter, C.byref(nc))
        return v, secs, nc.value

    def heat_elemental(self, ein, coords):
        ein = np.ascontiguousarray(ein, dtype=np.float64)
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        out = np.zeros_like(ein)
        self.ref.L.dktref_heat_elemental(self.h, _p(ein), _p(out), _p(coords))
        return out

    def __del__(self):
        if self.h:
            self.ref.L.dktref_da_destroy(self.h)
            self.h = None


class Reference:
    """One (dim, maxDepth, sfc) session with the reference library."""

    def __init__(self, dim, max_depth, sfc="morton"):
        self.L = _lib(sfc)
        self.dim, self.max_depth, self.sfc = dim, max_depth, sfc
        if self.L.dktref_init(dim, max_depth):
            raise ValueError("bad dim")

    def reinit(self):
        self.L.dktref_init(self.dim, self.max_depth)

    def tables(self):
        nch = 1 << self.dim
        if not self.L.dktref_is_hilbert():  # Morton build: one identity rotation (src/KDhcurvedata.cpp:47-57)
            ident = np.arange(nch, dtype=np.int8)[None, :]
            return np.concatenate([ident, ident], axis=1), np.zeros((1, nch), dtype=np.int32)
        nrot = self.L.dktref_tables(self.dim, None, None)
        rot = np.zeros(nrot, dtype=np.int8)
        htab = np.zeros(nrot // 2, dtype=np.int32)
        self.L.dktref_tables(self.dim, _p(rot), _p(htab))
        return rot.reshape(-1, 2 * nch), htab.reshape(-1, nch)

    def tree_from_points(self, pts, max_pts=1, flex=0.3, balance=True):
        self.reinit()
        pts = np.ascontiguousarray(pts, dtype=np.uint32)
        return RefTree(self, self.L.dktref_tree_from_points(self.dim, _p(pts), len(pts), max_pts, flex, int(balance)))

    def tree_from_elements(self, xyz, lev, sort=True):
        self.reinit()
        xyz = np.ascontiguousarray(xyz, dtype=np.uint32)
        lev = np.ascontiguousarray(lev, dtype=np.uint8)
        return RefTree(self, self.L.dktref_tree_from_elements(self.dim, _p(xyz), _p(lev), len(lev), int(sort)))

    def tree_example(self, which, depth, sort=True):
        self.reinit()
        return RefTree(self, self.L.dktref_tree_example(self.dim, which, depth, int(sort)))

    def da(self, tree, order=1):
        self.reinit()
        return RefDA(self, self.L.dktref_da_create(tree.h, order), order)

    def refel(self, order):
        M = order + 1
        ip0, ip1, Q, Dg = (np.zeros((M, M)) for _ in range(4))
        w = np.zeros(M)
        sz = self.L.dktref_refel(self.dim, order, _p(ip0), _p(ip1), _p(Q), _p(Dg), _p(w))
        return dict(ip0=ip0, ip1=ip1, Q=Q, Dg=Dg, w=w, elem_sz=sz)
