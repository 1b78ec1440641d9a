"""Python host binding of libdkt.so (ctypes over the C ABI in include/dkt.h).

The product path: every call goes to the CUDA library; there is no Python/NumPy compute
fallback.  Importing works without a GPU (so that symbol checks can run on CPU boxes); any
compute entry point fails loudly without one.
"""
import ctypes as C
import os

import numpy as np

from . import operators, trees  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.environ.get("DKT_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libdkt.so")

OK = 0
SFC_MORTON, SFC_HILBERT = 0, 1
ELEMS_ON_DEVICE, ELEMS_PRESORTED, ALLOW_UNDEFINED, DIST_DRYRUN = 1, 2, 4, 8
VEC_HOST, VEC_DEVICE, NO_Q1_MASK, MV_FLAT, MV_NO_FASTPATH, VEC_GHOSTED = 0, 1, 2, 4, 8, 16
OP_IDENTITY, OP_DENSE, OP_KRON = 0, 1, 2
CLASS_NAMES = "ABPU"
INVALID = 0xFFFFFFFF

DECLARED_SYMBOLS = [
    "dkt_last_error", "dkt_version", "dkt_sfc_tables", "dkt_da_create", "dkt_da_destroy", "dkt_da_sizes",
    "dkt_nccl_unique_id", "dkt_da_create_dist", "dkt_p2p_attach_local", "dkt_da_export_owned_ids", "dkt_da_export_exchange",
    "dkt_da_export_elements", "dkt_da_export_nodes", "dkt_da_export_boundary", "dkt_da_export_tables", "dkt_matvec",
    "dkt_cg_solve", "dkt_ghost_read_begin", "dkt_ghost_read_end", "dkt_ghost_write_begin", "dkt_ghost_write_end", "dkt_last_kernel_ms", "dkt_da_chunk_info", "dkt_da_stream", "dkt_da_set_stream", "dkt_kernel_launch_count",
    "dkt_ghost_read_host", "dkt_ghost_write_host", "dkt_matvec_dof", "dkt_tree_from_points", "dkt_tree_size", "dkt_tree_export", "dkt_tree_device_ptrs", "dkt_tree_destroy",
]


class DktError(RuntimeError):
    pass


class _Op(C.Structure):
    _fields_ = [("kind", C.c_int), ("kref", C.c_void_p), ("alpha", C.c_double), ("dirichlet", C.c_int), ("terms", C.c_int)]


class _Sizes(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("n_elem", "n_mv_elem", "n_nodes", "n_boundary", "n_hanging", "n_split", "alg_bytes")] + \
               [(n, C.c_int) for n in ("nodes_per_elem", "tree_class", "finest_level", "n_ranks")] + \
               [(n, C.c_uint64) for n in ("n_global_nodes", "n_ghost_nodes", "n_global_elem")]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIBPATH):
        raise DktError("libdkt.so is not built (%s): run `python __graft_entry__.py`; there is no fallback path" % _LIBPATH)
    L = C.CDLL(_LIBPATH)
    vp, i32, u32, u64, f64 = C.c_void_p, C.c_int, C.c_uint, C.c_uint64, C.c_double
    L.dkt_last_error.restype = C.c_char_p
    L.dkt_version.restype = C.c_char_p
    L.dkt_sfc_tables.argtypes = [i32, i32, vp, vp]
    L.dkt_da_create.argtypes = [i32, i32, i32, i32, vp, vp, u64, vp, vp, u32, C.POINTER(vp)]
    L.dkt_da_create_dist.argtypes = [i32, i32, i32, i32, vp, vp, u64, vp, vp, u32, i32, i32, vp, C.POINTER(vp)]
    L.dkt_nccl_unique_id.argtypes = [vp]
    L.dkt_da_export_owned_ids.argtypes = [vp, vp]
    L.dkt_p2p_attach_local.argtypes = [vp, i32]
    L.dkt_da_export_exchange.argtypes = [vp, vp, vp]
    L.dkt_da_destroy.argtypes = [vp]
    L.dkt_da_sizes.argtypes = [vp, C.POINTER(_Sizes)]
    L.dkt_da_export_elements.argtypes = [vp, vp, vp]
    L.dkt_da_export_nodes.argtypes = [vp, vp, vp]
    L.dkt_da_export_boundary.argtypes = [vp, vp]
    L.dkt_da_export_tables.argtypes = [vp, vp, vp, vp, vp, vp]
    L.dkt_matvec.argtypes = [vp, C.POINTER(_Op), vp, vp, f64, u32]
    L.dkt_matvec_dof.argtypes = [vp, C.POINTER(_Op), vp, vp, f64, u32, i32]
    L.dkt_cg_solve.argtypes = [vp, C.POINTER(_Op), vp, vp, i32, C.POINTER(f64), f64, u32, C.POINTER(i32), C.POINTER(i32)]
    L.dkt_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.dkt_da_stream.restype = vp
    L.dkt_da_stream.argtypes = [vp]
    L.dkt_da_set_stream.argtypes = [vp, vp]
    L.dkt_da_chunk_info.argtypes = [vp, vp]
    L.dkt_tree_from_points.argtypes = [i32, i32, i32, vp, u64, u64, i32, u32, C.POINTER(vp)]
    L.dkt_tree_size.argtypes = [vp, C.POINTER(u64), C.POINTER(i32)]
    L.dkt_tree_export.argtypes = [vp, vp, vp, u32]
    L.dkt_tree_device_ptrs.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.dkt_tree_destroy.argtypes = [vp]
    for f in ("dkt_ghost_read_begin", "dkt_ghost_read_end", "dkt_ghost_write_begin", "dkt_ghost_write_end"):
        getattr(L, f).argtypes = [vp, vp]
    L.dkt_kernel_launch_count.restype = u64
    _lib = L
    return L


def check_symbols():
    L = lib()
    missing = [s for s in DECLARED_SYMBOLS if not hasattr(L, s)]
    if missing:
        raise DktError("libdkt.so lacks symbols declared in include/dkt.h: %s" % missing)
    return True


def _check(rc):
    if rc != OK:
        raise DktError("libdkt error %d: %s" % (rc, lib().dkt_last_error().decode()))


def _ptr(a):
    """Host numpy array or CUDA torch tensor -> raw pointer."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(a.data_ptr())  # torch tensor


def _is_cuda(a):
    return (not isinstance(a, np.ndarray)) and hasattr(a, "is_cuda") and a.is_cuda


def sfc_tables(dim, sfc=SFC_MORTON):
    """(rot_perm, rot_inv, htab) generated by the library (replaces _InitializeHcurve)."""
    L = lib()
    nrot = L.dkt_sfc_tables(dim, sfc, None, None)
    if nrot <= 0:
        raise DktError("bad dim/sfc")
    nch = 1 << dim
    rot = np.zeros(nrot * 2 * nch, dtype=np.int8)
    h = np.zeros(nrot * nch, dtype=np.int32)
    L.dkt_sfc_tables(dim, sfc, _ptr(rot), _ptr(h))
    rot = rot.reshape(nrot, 2 * nch)
    return rot[:, :nch].copy(), rot[:, nch:].copy(), h.reshape(nrot, nch)


class Operator:
    """Device form of a leaf class's elementalMatVec (FEM/include/feMatrix.h:56)."""

    def __init__(self, kind=OP_IDENTITY, kref=None, alpha=0.0, dirichlet=False, terms=0):
        self.kind, self.alpha, self.dirichlet, self.terms = kind, float(alpha), bool(dirichlet), int(terms)
        self.kref = None if kref is None else np.ascontiguousarray(kref, dtype=np.float64)

    @staticmethod
    def identity(dirichlet=False):
        return Operator(OP_IDENTITY, dirichlet=dirichlet)

    @staticmethod
    def dense(kref, alpha, dirichlet=False):
        return Operator(OP_DENSE, kref, alpha, dirichlet)

    @staticmethod
    def kron(terms, alpha, dirichlet=False):
        """Sum-factorised operator (DKT_OP_KRON): terms[t][d] is the M x M matrix of term t along axis d, A[k, j] taking
        input index k to output index j; K_e = scale * h^alpha * sum_t kron(terms[t][dim-1], .., terms[t][0])."""
        a = np.ascontiguousarray(terms, dtype=np.float64)
        assert a.ndim == 4 and a.shape[2] == a.shape[3]
        return Operator(OP_KRON, a, alpha, dirichlet, terms=a.shape[0])

    def _c(self):
        return _Op(self.kind, None if self.kref is None else self.kref.ctypes.data, self.alpha, int(self.dirichlet), self.terms)


class Tree:
    """A linear tree built on the GPU from points: SFC_Tree<T,dim>::distTreeBalancing (balance=True) or
    distTreeConstruction (balance=False) of the reference on one rank (src/tsort.cpp:647-716, 862-877).

    pts: (n, dim) integer coordinates in [0, 2^max_depth) - numpy array or CUDA torch tensor."""

    def __init__(self, pts, dim, max_depth, max_pts=1, balance=True, sfc=SFC_MORTON):
        L = lib()
        self.dim, self.max_depth, self.sfc = dim, max_depth, sfc
        flags = 0
        if _is_cuda(pts):
            import torch
            pts = pts.to(torch.int32).contiguous() if pts.dtype not in (torch.int32, torch.uint32) else pts.contiguous()
            n = int(pts.numel()) // dim
            flags |= ELEMS_ON_DEVICE
        else:
            pts = np.ascontiguousarray(pts, dtype=np.uint32).reshape(-1, dim)
            n = len(pts)
        h = C.c_void_p()
        _check(L.dkt_tree_from_points(dim, max_depth, sfc, _ptr(pts), n, int(max_pts), int(bool(balance)), flags, C.byref(h)))
        self._h = h
        ne, fl = C.c_uint64(), C.c_int()
        _check(L.dkt_tree_size(self._h, C.byref(ne), C.byref(fl)))
        self.n_elem, self.finest_level = int(ne.value), int(fl.value)

    def __len__(self):
        return self.n_elem

    def export(self):
        """(xyz, lev) numpy arrays of the leaves in tree order."""
        xyz = np.zeros((self.n_elem, self.dim), dtype=np.uint32)
        lev = np.zeros(self.n_elem, dtype=np.uint8)
        _check(lib().dkt_tree_export(self._h, _ptr(xyz), _ptr(lev), 0))
        return xyz, lev

    def export_torch(self):
        """(xyz int32, lev uint8) CUDA tensors (copies; usable after the tree is closed)."""
        import torch
        xyz = torch.empty((self.n_elem, self.dim), dtype=torch.int32, device="cuda")
        lev = torch.empty(self.n_elem, dtype=torch.uint8, device="cuda")
        _check(lib().dkt_tree_export(self._h, _ptr(xyz), _ptr(lev), ELEMS_ON_DEVICE))
        return xyz, lev

    def da(self, order=1, **kw):
        """ot::DA over this tree, built from the device arrays without a host round trip."""
        xyz, lev = self.export_torch()
        return DA(xyz, lev, self.dim, order, self.max_depth, sfc=self.sfc, presorted=True, **kw)

    def close(self):
        if self._h:
            lib().dkt_tree_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DA:
    """Mirror of ot::DA<dim> (include/oda.h:41-525) for the matvec path, single rank.

    xyz: (n, dim) uint32 anchors, lev: (n,) uint8 levels - numpy arrays or CUDA torch tensors.
    """

    def __init__(self, xyz, lev, dim, order=1, max_depth=30, sfc=SFC_MORTON, ip0=None, ip1=None, presorted=False,
                 allow_undefined=False, rank=0, nranks=1, nccl_id=None, dryrun=False):
        L = lib()
        self.dim, self.order, self.max_depth, self.sfc = dim, order, max_depth, sfc
        flags = (ELEMS_PRESORTED if presorted else 0) | (ALLOW_UNDEFINED if allow_undefined else 0)
        if _is_cuda(xyz):
            import torch
            xyz = xyz.to(torch.int32).contiguous() if xyz.dtype not in (torch.int32, torch.uint32) else xyz.contiguous()
            lev = lev.to(torch.uint8).contiguous()
            n = int(lev.numel())
            flags |= ELEMS_ON_DEVICE
        else:
            xyz = np.ascontiguousarray(xyz, dtype=np.uint32)
            lev = np.ascontiguousarray(lev, dtype=np.uint8)
            n = len(lev)
        self._keep = (xyz, lev)
        ip0 = None if ip0 is None else np.ascontiguousarray(ip0, dtype=np.float64)
        ip1 = None if ip1 is None else np.ascontiguousarray(ip1, dtype=np.float64)
        h = C.c_void_p()
        if nranks > 1 and dryrun:
            flags |= DIST_DRYRUN
            _check(L.dkt_da_create_dist(dim, order, max_depth, sfc, _ptr(xyz), _ptr(lev), n, _ptr(ip0), _ptr(ip1), flags, rank,
                                        nranks, None, C.byref(h)))
        elif nranks > 1:
            if nccl_id is None or len(nccl_id) != 128:
                raise DktError("nccl_id must be the 128 bytes of dkt.nccl_unique_id() from rank 0")
            idbuf = (C.c_char * 128).from_buffer_copy(bytes(nccl_id))
            _check(L.dkt_da_create_dist(dim, order, max_depth, sfc, _ptr(xyz), _ptr(lev), n, _ptr(ip0), _ptr(ip1), flags, rank,
                                        nranks, C.cast(idbuf, C.c_void_p), C.byref(h)))
        else:
            _check(L.dkt_da_create(dim, order, max_depth, sfc, _ptr(xyz), _ptr(lev), n, _ptr(ip0), _ptr(ip1), flags, C.byref(h)))
        self.rank, self.nranks = rank, nranks
        self._h = h
        self._keep = None
        s = _Sizes()
        _check(L.dkt_da_sizes(self._h, C.byref(s)))
        self.sizes = s
        self.n_elem, self.n_mv_elem, self.n_nodes = int(s.n_elem), int(s.n_mv_elem), int(s.n_nodes)
        self.n_boundary, self.n_hanging, self.n_split = int(s.n_boundary), int(s.n_hanging), int(s.n_split)
        self.alg_bytes, self.N = int(s.alg_bytes), int(s.nodes_per_elem)
        self.n_global_nodes, self.n_ghost_nodes = int(s.n_global_nodes), int(s.n_ghost_nodes)
        self.tree_class = CLASS_NAMES[s.tree_class]
        self.finest_level = int(s.finest_level)

    # --- DA getters ---------------------------------------------------------------------------
    def elements(self):
        xyz = np.zeros((self.n_elem, self.dim), dtype=np.uint32)
        lev = np.zeros(self.n_elem, dtype=np.uint8)
        _check(lib().dkt_da_export_elements(self._h, _ptr(xyz), _ptr(lev)))
        return xyz, lev

    def nodes(self):
        """getTNCoords(): (xyz, level) of every CG node in DA order; partitioned DA: of the local vector [owned | ghosts]."""
        n = self.n_nodes + self.n_ghost_nodes
        xyz = np.zeros((n, self.dim), dtype=np.uint32)
        lev = np.zeros(n, dtype=np.uint8)
        _check(lib().dkt_da_export_nodes(self._h, _ptr(xyz), _ptr(lev)))
        return xyz, lev

    def boundary_ids(self):
        ids = np.zeros(self.n_boundary, dtype=np.uint32)
        _check(lib().dkt_da_export_boundary(self._h, _ptr(ids)))
        return ids

    def owned_ids(self):
        """Partitioned DA: index in the single-rank DA order of each node this rank owns."""
        ids = np.zeros(self.n_nodes, dtype=np.uint32)
        _check(lib().dkt_da_export_owned_ids(self._h, _ptr(ids)))
        return ids

    def exchange_counts(self):
        """Partitioned DA: (send_counts, recv_counts) per peer rank."""
        sc = np.zeros(self.nranks, dtype=np.uint64)
        rcv = np.zeros(self.nranks, dtype=np.uint64)
        _check(lib().dkt_da_export_exchange(self._h, _ptr(sc), _ptr(rcv)))
        return sc, rcv

    def tables(self):
        nm, nh, N = self.n_mv_elem, self.n_hanging, self.N
        out = dict(mv_xyz=np.zeros((nm, self.dim), dtype=np.uint32), mv_lev=np.zeros(nm, dtype=np.uint8),
                   e2n=np.zeros((nm, N), dtype=np.uint32), pnode=np.zeros((nh, N), dtype=np.uint32),
                   child=np.zeros(nh, dtype=np.uint8))
        _check(lib().dkt_da_export_tables(self._h, _ptr(out["mv_xyz"]), _ptr(out["mv_lev"]), _ptr(out["e2n"]),
                                          _ptr(out["pnode"]), _ptr(out["child"])))
        return out

    # --- the hot path -----------------------------------------------------------------------------
    def matvec(self, op, u, out=None, scale=1.0, q1_mask=True, flat=False, fastpath=True, ghosted=False, dof=1):
        """feMatrix::matVec(in, out, scale).  numpy arrays (host path: H2D + kernels + D2H inside the
        call) or CUDA torch tensors (device path: kernels only, asynchronous on self.stream).
        dof > 1: u holds dof values per node, interleaved ([abc][abc]..), the operator acts on every component."""
        flags = (0 if q1_mask else NO_Q1_MASK) | (MV_FLAT if flat else 0) | (0 if fastpath else MV_NO_FASTPATH)
        if _is_cuda(u):
            import torch
            if out is None:
                out = torch.empty_like(u)
            if not getattr(self, "_user_stream", False):
                # the DA launches on its own stream: order it after the work that produced `u`
                torch.cuda.current_stream().synchronize()
            want = (self.n_nodes + (self.n_ghost_nodes if ghosted else 0)) * dof
            assert u.dtype == torch.float64 and u.is_contiguous() and out.is_contiguous() and u.numel() == want == out.numel()
            flags |= VEC_DEVICE | (VEC_GHOSTED if ghosted else 0)
        else:
            u = np.ascontiguousarray(u, dtype=np.float64)
            assert u.size == self.n_nodes * dof
            if out is None:
                out = np.empty_like(u)
        c = op._c()
        if dof != 1:
            _check(lib().dkt_matvec_dof(self._h, C.byref(c), _ptr(u), _ptr(out), float(scale), flags, int(dof)))
        else:
            _check(lib().dkt_matvec(self._h, C.byref(c), _ptr(u), _ptr(out), float(scale), flags))
        return out

    def ghost_read(self, vec):
        """readFromGhostBegin + End on a CUDA tensor in the ghosted layout [owned | ghosts] (asynchronous on self.stream)."""
        assert _is_cuda(vec) and vec.numel() == self.n_nodes + self.n_ghost_nodes
        _check(lib().dkt_ghost_read_begin(self._h, _ptr(vec)))
        _check(lib().dkt_ghost_read_end(self._h, _ptr(vec)))

    def ghost_write(self, vec):
        """writeToGhostsBegin + End: the ghost entries are added to their owners' entries."""
        assert _is_cuda(vec) and vec.numel() == self.n_nodes + self.n_ghost_nodes
        _check(lib().dkt_ghost_write_begin(self._h, _ptr(vec)))
        _check(lib().dkt_ghost_write_end(self._h, _ptr(vec)))

    def chunk_info(self):
        a = np.zeros(16, dtype=np.uint64)
        _check(lib().dkt_da_chunk_info(self._h, _ptr(a)))
        keys = ("chunks", "units_per_chunk", "max_nodes", "max_run", "total_nodes")
        return {"regular": dict(zip(keys, map(int, a[:5]))), "hanging": dict(zip(keys, map(int, a[5:10]))),
                "families": dict(zip(keys, map(int, a[10:15]))), "elements_in_families": int(a[15])}

    def cg_solve(self, op, b, x0=None, max_iter=100, tol=1e-10, scale=1.0):
        """HeatMat::cgSolve with resident vectors.  Returns (x, iterations, residual, converged)."""
        flags = 0
        if _is_cuda(b):
            import torch
            x = torch.zeros_like(b) if x0 is None else x0.clone()
            flags |= VEC_DEVICE
            if not getattr(self, "_user_stream", False):
                # the DA launches on its own stream: order it after the work that produced b and x (as matvec does)
                torch.cuda.current_stream().synchronize()
        else:
            b = np.ascontiguousarray(b, dtype=np.float64)
            x = np.zeros_like(b) if x0 is None else np.array(x0, dtype=np.float64)
        c = op._c()
        t, it, st = C.c_double(tol), C.c_int(0), C.c_int(1)
        _check(lib().dkt_cg_solve(self._h, C.byref(c), _ptr(x), _ptr(b), int(max_iter), C.byref(t), float(scale), flags,
                                  C.byref(it), C.byref(st)))
        return x, it.value, t.value, st.value == 0

    def last_kernel_ms(self):
        ms = C.c_float(0)
        _check(lib().dkt_last_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)

    @property
    def stream(self):
        return lib().dkt_da_stream(self._h)

    def set_stream(self, cuda_stream):
        """Launch on the caller's stream (an int cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream)."""
        self._user_stream = cuda_stream is not None
        if cuda_stream is None:
            h = None  # back to the DA's own stream
        else:
            h = C.c_void_p(int(cuda_stream) or 1)  # 0 is the legacy default stream == cudaStreamLegacy (0x1)
        _check(lib().dkt_da_set_stream(self._h, h))

    def close(self):
        if getattr(self, "_h", None):
            lib().dkt_da_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def p2p_attach_local(das):
    """Test hook: wire the peer-memory exchange of the dry-run DAs of ranks 0..n-1 of one partition, all living in
    this process (dkt_p2p_attach_local)."""
    arr = (C.c_void_p * len(das))(*[d._h for d in das])
    _check(lib().dkt_p2p_attach_local(C.cast(arr, C.c_void_p), len(das)))


def nccl_unique_id():
    """128-byte NCCL unique id (call on rank 0, broadcast to the other ranks)."""
    buf = (C.c_char * 128)()
    _check(lib().dkt_nccl_unique_id(C.cast(buf, C.c_void_p)))
    return bytes(buf)


def kernel_launch_count():
    return int(lib().dkt_kernel_launch_count())
