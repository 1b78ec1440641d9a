"""Python side of the FULL emulated pipeline (tests/emu/emu_full.cpp): dkt_build.cu + dkt_chunks.cu + dkt_sfc.cpp compiled
with g++ against tests/emu/cuda_emu.h and driven like dkt_da_create / dkt_matvec.  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "dendro-kt_b200", "csrc")
LIB = os.path.join(EMU, "_build", "libdkt_emu_full.so")
_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    import emu_build
    L = C.CDLL(emu_build.build())
    L.emu_last_error.restype = C.c_char_p
    L.emu_da_create.restype = C.c_void_p
    L.emu_da_create.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint]
    L.emu_da_sizes.argtypes = [C.c_void_p, C.c_void_p]
    L.emu_da_export.argtypes = [C.c_void_p] * 11
    L.emu_da_matvec.restype = C.c_int
    L.emu_da_matvec.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_uint]
    L.emu_da_destroy.argtypes = [C.c_void_p]
    L.emu_da_cg.restype = C.c_int
    L.emu_da_cg.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_double,
                            C.c_uint, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.emu_tree_from_points.restype = C.c_void_p
    L.emu_tree_from_points.argtypes = [C.c_int] * 3 + [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int]
    L.emu_tree_size.restype = C.c_uint64
    L.emu_tree_size.argtypes = [C.c_void_p]
    L.emu_tree_export.argtypes = [C.c_void_p] * 3
    L.emu_tree_destroy.argtypes = [C.c_void_p]
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class EmuDA:
    """dkt.DA's construction and matvec, emulated.  families: DKT_FAMILIES (0: per-element chunk tables only)."""

    def __init__(self, xyz, lev, dim, order, max_depth, sfc=0, ip0=None, ip1=None, families=1, flags=0):
        import flat
        L = lib()
        xyz = np.ascontiguousarray(xyz, dtype=np.uint32)
        lev = np.ascontiguousarray(lev, dtype=np.uint8)
        if ip0 is None:
            ip0, ip1 = flat.default_interp(order)
        ip0 = np.ascontiguousarray(np.asarray(ip0, dtype=np.float64).ravel())
        ip1 = np.ascontiguousarray(np.asarray(ip1, dtype=np.float64).ravel())
        old = os.environ.get("DKT_FAMILIES")
        os.environ["DKT_FAMILIES"] = str(families)
        try:
            self._h = L.emu_da_create(dim, order, max_depth, sfc, _p(xyz), _p(lev), len(lev), _p(ip0), _p(ip1), flags)
        finally:
            if old is None:
                os.environ.pop("DKT_FAMILIES", None)
            else:
                os.environ["DKT_FAMILIES"] = old
        if not self._h:
            raise RuntimeError("emu_da_create: " + L.emu_last_error().decode())
        s = np.zeros(16, dtype=np.uint64)
        L.emu_da_sizes(self._h, _p(s))
        (self.n_elem, self.n_mv_elem, self.n_reg, self.n_hanging, self.n_nodes, self.n_boundary, self.n_split, tc, self.N, self.finest_level,
         self.n_sets) = [int(x) for x in s[:11]]
        self.tree_class = "ABPU"[tc]
        self.dim = dim

    def export(self):
        d, N = self.dim, self.N
        out = dict(elem_xyz=np.zeros((self.n_elem, d), np.uint32), elem_lev=np.zeros(self.n_elem, np.uint8),
                   node_xyz=np.zeros((self.n_nodes, d), np.uint32), node_lev=np.zeros(self.n_nodes, np.uint8),
                   bdy=np.zeros(self.n_boundary, np.uint32), mv_xyz=np.zeros((self.n_mv_elem, d), np.uint32),
                   mv_lev=np.zeros(self.n_mv_elem, np.uint8), e2n=np.zeros((self.n_mv_elem, N), np.uint32),
                   pnode=np.zeros((self.n_hanging, N), np.uint32), child=np.zeros(self.n_hanging, np.uint8))
        lib().emu_da_export(self._h, *[_p(out[k]) for k in ("elem_xyz", "elem_lev", "node_xyz", "node_lev", "bdy", "mv_xyz", "mv_lev",
                                                            "e2n", "pnode", "child")])
        return out

    def matvec(self, u, kref=None, alpha=0.0, scale=1.0, dirichlet=False, flags=0, flat=False, q1_mask=True, fastpath=True):
        flags |= (4 if flat else 0) | (0 if q1_mask else 2) | (0 if fastpath else 8)
        u = np.ascontiguousarray(u, dtype=np.float64)
        kr = None if kref is None else np.ascontiguousarray(np.asarray(kref, dtype=np.float64).ravel())
        out = np.full(self.n_nodes, np.nan)
        rc = lib().emu_da_matvec(self._h, 0 if kr is None else 1, _p(kr), alpha, int(dirichlet), _p(u), _p(out), scale, flags)
        if rc:
            raise RuntimeError("emu_da_matvec rc=%d: %s" % (rc, lib().emu_last_error().decode()))
        return out

    def cg_solve(self, b, kref, alpha, max_iter, tol, x0=None, dirichlet=True, scale=1.0, flags=0):
        """dkt_cg_solve: returns (x, iterations, residual, status)"""
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros_like(b) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).copy()
        kr = np.ascontiguousarray(np.asarray(kref, dtype=np.float64).ravel())
        t, it, st = C.c_double(tol), C.c_int(0), C.c_int(0)
        rc = lib().emu_da_cg(self._h, 1, _p(kr), alpha, int(dirichlet), _p(x), _p(b), max_iter, C.byref(t), scale, flags, C.byref(it), C.byref(st))
        if rc:
            raise RuntimeError("emu_da_cg rc=%d: %s" % (rc, lib().emu_last_error().decode()))
        return x, it.value, t.value, st.value

    def close(self):
        if self._h:
            lib().emu_da_destroy(self._h)
            self._h = None


def tree_from_points(pts, dim, max_depth, max_pts=1, balance=True, sfc=0):
    """dkt_tree.cu's build_tree, emulated: (xyz, lev) of the leaves in tree order."""
    L = lib()
    pts = np.ascontiguousarray(pts, dtype=np.uint32).reshape(-1, dim)
    h = L.emu_tree_from_points(dim, max_depth, sfc, _p(pts), len(pts), max_pts, int(balance))
    if not h:
        raise RuntimeError("emu_tree_from_points: " + L.emu_last_error().decode())
    n = L.emu_tree_size(h)
    xyz = np.zeros((n, dim), dtype=np.uint32)
    lev = np.zeros(n, dtype=np.uint8)
    L.emu_tree_export(h, _p(xyz), _p(lev))
    L.emu_tree_destroy(h)
    return xyz, lev
