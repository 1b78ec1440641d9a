// C ABI of libdkt.so (declared in include/dkt.h).
#include "dkt_internal.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>

namespace dkt
{
static thread_local std::string g_err;
uint64_t g_launches = 0;
void set_error(const std::string &msg) { g_err = msg; }

// exact parent->child 1-D interpolation on uniform nodes in [-1,1] (what RefElement's ip_1D_0/1
// approximate, FEM/src/refel.cpp:112-163): A[k*M+j] = l_k(x_j^child), Lagrange basis l_k.
static void exact_interp(int order, double ip[2][MAX_M * MAX_M])
{
  const int M = order + 1;
  for (int b = 0; b < 2; b++)
    for (int k = 0; k < M; k++)
      for (int j = 0; j < M; j++)
      {
        const double xj = -1.0 + 2.0 * j / order;
        const double xc = 0.5 * (xj + (b ? 1.0 : -1.0));
        double l = 1.0;
        for (int m = 0; m < M; m++)
          if (m != k)
          {
            const double xm = -1.0 + 2.0 * m / order, xk = -1.0 + 2.0 * k / order;
            l *= (xc - xm) / (xk - xm);
          }
        ip[b][k * M + j] = l;
      }
}
} // namespace dkt

using namespace dkt;

struct dkt_da
{
  DA d;
  Dist dist;
};
struct dkt_tree
{
  Tree t;
};

#define CKA(call)                                                                                 \
  do                                                                                              \
  {                                                                                               \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
    {                                                                                             \
      set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                              \
      return DKT_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)

// interleaved [node][component] <-> component-major [component][node]
__global__ void k_dof_split(const double *in, size_t n, int dof, double *out)
{
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n * dof) return;
  out[(i % dof) * n + i / dof] = in[i];
}
__global__ void k_dof_merge(const double *in, size_t n, int dof, double *out)
{
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n * dof) return;
  out[i] = in[(i % dof) * n + i / dof];
}

extern "C"
{
  const char *dkt_last_error(void) { return g_err.c_str(); }
  const char *dkt_version(void) { return "dkt-b200 0.1 (sm_100a)"; }
  uint64_t dkt_kernel_launch_count(void) { return g_launches; }

  int dkt_sfc_tables(int dim, int sfc_mode, char *rotations, int *hilbert_table)
  {
    if (dim < 2 || dim > 4 || (sfc_mode != DKT_SFC_MORTON && sfc_mode != DKT_SFC_HILBERT)) return -1;
    SfcTables t;
    make_sfc_tables(dim, sfc_mode, t);
    const int nch = t.nch;
    for (int r = 0; r < t.nrot; r++)
      for (int i = 0; i < nch; i++)
      {
        if (rotations)
        {
          rotations[r * 2 * nch + i] = (char)t.rot_perm[r * nch + i];
          rotations[r * 2 * nch + nch + i] = (char)t.rot_inv[r * nch + i];
        }
        if (hilbert_table) hilbert_table[r * nch + i] = t.htab[r * nch + i];
      }
    return t.nrot;
  }

  int dkt_nccl_unique_id(void *out128)
  {
    if (!out128) { set_error("NULL argument"); return DKT_ERR_INVALID; }
    return nccl_unique_id(out128);
  }
  int dkt_tree_from_points(int dim, int max_depth, int sfc_mode, const uint32_t *pts_xyz, uint64_t n_pts, uint64_t max_pts_per_region,
                           int balance, unsigned flags, dkt_tree **out)
  {
    if (!out) { set_error("out is NULL"); return DKT_ERR_INVALID; }
    *out = nullptr;
    if (!pts_xyz) { set_error("pts_xyz is NULL"); return DKT_ERR_INVALID; }
    if (sfc_mode != DKT_SFC_MORTON && sfc_mode != DKT_SFC_HILBERT) { set_error("bad sfc_mode"); return DKT_ERR_INVALID; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
      set_error("no CUDA device: libdkt has no CPU path");
      return DKT_ERR_CUDA;
    }
    dkt_tree *t = new (std::nothrow) dkt_tree;
    if (!t) { set_error("out of host memory"); return DKT_ERR_INVALID; }
    t->t.dim = dim; t->t.max_depth = max_depth; t->t.sfc_mode = sfc_mode;
    cudaGetDevice(&t->t.device);
    const int rc = build_tree(t->t, pts_xyz, n_pts, max_pts_per_region, balance != 0, flags);
    if (rc != DKT_OK) { free_tree(t->t); delete t; return rc; }
    *out = t;
    return DKT_OK;
  }
  int dkt_tree_size(const dkt_tree *t, uint64_t *n_elem, int *finest_level)
  {
    if (!t || !n_elem) { set_error("NULL argument"); return DKT_ERR_INVALID; }
    *n_elem = t->t.n;
    if (finest_level) *finest_level = t->t.finest_level;
    return DKT_OK;
  }
  int dkt_tree_export(const dkt_tree *t, uint32_t *elem_xyz, uint8_t *elem_lev, unsigned flags)
  {
    if (!t || !elem_xyz || !elem_lev) { set_error("NULL argument"); return DKT_ERR_INVALID; }
    CKA(cudaSetDevice(t->t.device));
    const cudaMemcpyKind kind = (flags & DKT_ELEMS_ON_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (t->t.n)
    {
      CKA(cudaMemcpy(elem_xyz, t->t.d_xyz, t->t.n * t->t.dim * sizeof(uint32_t), kind));
      CKA(cudaMemcpy(elem_lev, t->t.d_lev, t->t.n, kind));
    }
    return DKT_OK;
  }
  int dkt_tree_device_ptrs(const dkt_tree *t, const uint32_t **elem_xyz, const uint8_t **elem_lev)
  {
    if (!t || !elem_xyz || !elem_lev) { set_error("NULL argument"); return DKT_ERR_INVALID; }
    *elem_xyz = t->t.d_xyz;
    *elem_lev = t->t.d_lev;
    return DKT_OK;
  }
  int dkt_tree_destroy(dkt_tree *t)
  {
    if (!t) return DKT_OK;
    cudaSetDevice(t->t.device);
    free_tree(t->t);
    delete t;
    return DKT_OK;
  }
  int dkt_da_create(int dim, int order, int max_depth, int sfc_mode, const uint32_t *elem_xyz, const uint8_t *elem_lev,
                    uint64_t n_elem, const double *ip0, const double *ip1, unsigned flags, dkt_da **out)
  {
    return dkt_da_create_dist(dim, order, max_depth, sfc_mode, elem_xyz, elem_lev, n_elem, ip0, ip1, flags, 0, 1, nullptr, out);
  }
  int dkt_da_export_exchange(const dkt_da *da, uint64_t *send_counts, uint64_t *recv_counts)
  {
    if (!da || !send_counts || !recv_counts) { set_error("NULL argument"); return DKT_ERR_INVALID; }
    if (!da->dist.active) { set_error("not a partitioned DA"); return DKT_ERR_INVALID; }
    for (int p = 0; p < da->dist.nranks; p++)
    {
      send_counts[p] = da->dist.send_off[p + 1] - da->dist.send_off[p];
      recv_counts[p] = da->dist.recv_off[p + 1] - da->dist.recv_off[p];
    }
    return DKT_OK;
  }
  int dkt_da_export_owned_ids(const dkt_da *da, uint32_t *ids)
  {
    if (!da || !ids) { set_error("NULL argument"); return DKT_ERR_INVALID; }
    if (!da->dist.active) { set_error("not a partitioned DA"); return DKT_ERR_INVALID; }
    CKA(cudaSetDevice(da->d.device));
    if (da->dist.nOwned) CKA(cudaMemcpy(ids, da->dist.d_owned_gid, da->dist.nOwned * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return DKT_OK;
  }
  int dkt_da_create_dist(int dim, int order, int max_depth, int sfc_mode, const uint32_t *elem_xyz, const uint8_t *elem_lev,
                         uint64_t n_elem, const double *ip0, const double *ip1, unsigned flags, int rank, int nranks,
                         const void *nccl_id, dkt_da **out)
  {
    if (nranks > 1 && !nccl_id && !(flags & DKT_DIST_DRYRUN)) { set_error("nccl_id is NULL"); return DKT_ERR_INVALID; }
    if (!out) { set_error("out is NULL"); return DKT_ERR_INVALID; }
    *out = nullptr;
    if (dim < 2 || dim > 4) { set_error("dim must be 2, 3 or 4"); return DKT_ERR_INVALID; }
    if (order < 1 || order > 2) { set_error("order must be 1 or 2 (low-order node resolver, include/nsort.tcc:1267)"); return DKT_ERR_UNSUPPORTED; }
    if (max_depth < 1 || max_depth > 30) { set_error("max_depth must be in 1..30"); return DKT_ERR_INVALID; }
    if (sfc_mode != DKT_SFC_MORTON && sfc_mode != DKT_SFC_HILBERT) { set_error("bad sfc_mode"); return DKT_ERR_INVALID; }
    if (!elem_xyz || !elem_lev) { set_error("element arrays are NULL"); return DKT_ERR_INVALID; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
      set_error("no CUDA device: libdkt has no CPU path");
      return DKT_ERR_CUDA;
    }
    dkt_da *h = new (std::nothrow) dkt_da;
    if (!h) { set_error("out of host memory"); return DKT_ERR_INVALID; }
    DA &d = h->d;
    d.dim = dim; d.order = order; d.max_depth = max_depth; d.sfc_mode = sfc_mode;
    d.M = order + 1;
    d.N = 1;
    for (int i = 0; i < dim; i++) d.N *= d.M;
    if ((ip0 == nullptr) != (ip1 == nullptr)) { delete h; set_error("ip0 and ip1 must both be given or both NULL"); return DKT_ERR_INVALID; }
    if (ip0)
    {
      std::memcpy(d.ip[0], ip0, sizeof(double) * d.M * d.M);
      std::memcpy(d.ip[1], ip1, sizeof(double) * d.M * d.M);
    }
    else
      exact_interp(order, d.ip);
    int rc = build_da(d, elem_xyz, elem_lev, n_elem, flags);
    if (rc == DKT_OK && nranks > 1) rc = partition_da(d, h->dist, rank, nranks, (flags & DKT_DIST_DRYRUN) ? nullptr : nccl_id);
    if (rc == DKT_OK) rc = build_chunks(d);
    if (rc != DKT_OK)
    {
      const std::string keep = g_err;
      free_dist(h->dist);
      free_da(d);
      delete h;
      g_err = keep;
      return rc;
    }
    *out = h;
    return DKT_OK;
  }

  int dkt_da_destroy(dkt_da *da)
  {
    if (!da) return DKT_OK;
    cudaSetDevice(da->d.device);
    free_dist(da->dist);
    free_da(da->d);
    delete da;
    return DKT_OK;
  }

  int dkt_da_sizes(const dkt_da *da, dkt_sizes *s)
  {
    if (!da || !s) { set_error("NULL argument"); return DKT_ERR_INVALID; }
    const DA &d = da->d;
    std::memset(s, 0, sizeof(*s));
    const Dist &ds = da->dist;
    s->n_elem = d.nElem; s->n_mv_elem = d.nMv; s->n_nodes = ds.active ? ds.nOwned : d.nNodes; s->n_boundary = d.nBdy;
    s->n_hanging = d.nHang;
    s->n_split = d.nSplit; s->nodes_per_elem = d.N; s->tree_class = d.tree_class; s->finest_level = d.finest_level;
    s->n_ranks = ds.active ? ds.nranks : 1;
    s->n_global_nodes = ds.active ? ds.nGlobalNodes : d.nNodes;
    s->n_ghost_nodes = ds.active ? ds.nGhost : 0;
    s->n_global_elem = d.nElem;
    // read u + write v, the uint32 element->node table, and per hanging element
    // the parent-cell node ids + child number/masks
    // (+ 16 B per ghost node when partitioned: 8 B each way, send + receive)
    s->alg_bytes = s->n_nodes * 16ull + d.nMv * (uint64_t)d.N * 4ull + d.nHang * ((uint64_t)d.N * 4ull + 8ull) +
                   s->n_ghost_nodes * 16ull;
    return DKT_OK;
  }

#define D2H(dst, src, bytes)                                                              \
  do                                                                                      \
  {                                                                                       \
    if ((dst) && (bytes) > 0) CKA(cudaMemcpy((dst), (src), (bytes), cudaMemcpyDeviceToHost)); \
  } while (0)

  int dkt_da_export_elements(const dkt_da *da, uint32_t *xyz, uint8_t *lev)
  {
    if (!da) { set_error("NULL da"); return DKT_ERR_INVALID; }
    const DA &d = da->d;  // partitioned DA: the WHOLE tree in tree order (every rank holds it)
    CKA(cudaSetDevice(d.device));
    D2H(xyz, d.d_elem_xyz, d.nElem * d.dim * sizeof(uint32_t));
    D2H(lev, d.d_elem_lev, d.nElem);
    return DKT_OK;
  }
  int dkt_da_export_nodes(const dkt_da *da, uint32_t *xyz, uint8_t *lev)
  {
    if (!da) { set_error("NULL da"); return DKT_ERR_INVALID; }
    const DA &d = da->d;  // partitioned DA: the local vector [owned | ghosts], n_nodes + n_ghost_nodes entries
    CKA(cudaSetDevice(d.device));
    D2H(xyz, d.d_node_xyz, d.nNodes * d.dim * sizeof(uint32_t));
    D2H(lev, d.d_node_lev, d.nNodes);
    return DKT_OK;
  }
  int dkt_da_export_boundary(const dkt_da *da, uint32_t *ids)
  {
    if (!da) { set_error("NULL da"); return DKT_ERR_INVALID; }
    const DA &d = da->d;  // partitioned DA: the OWNED nodes on the domain boundary (local indices)
    CKA(cudaSetDevice(d.device));
    D2H(ids, d.d_bdy, d.nBdy * sizeof(uint32_t));
    return DKT_OK;
  }
  int dkt_da_export_tables(const dkt_da *da, uint32_t *mv_xyz, uint8_t *mv_lev, uint32_t *e2n, uint32_t *pnode, uint8_t *child)
  {
    if (!da) { set_error("NULL da"); return DKT_ERR_INVALID; }
    if (da->dist.active) { set_error("table exports are available on single-rank DAs only"); return DKT_ERR_UNSUPPORTED; }
    const DA &d = da->d;
    CKA(cudaSetDevice(d.device));
    D2H(mv_xyz, d.d_mv_xyz, d.nMv * d.dim * sizeof(uint32_t));
    D2H(mv_lev, d.d_mv_lev, d.nMv);
    D2H(e2n, d.d_e2n, d.nMv * d.N * sizeof(uint32_t));
    D2H(pnode, d.d_pnode, d.nHang * d.N * sizeof(uint32_t));
    D2H(child, d.d_child, d.nHang);
    return DKT_OK;
  }

  int dkt_p2p_attach_local(dkt_da **das, int n)
  {
    if (!das || n < 2 || n > 64) { set_error("bad arguments"); return DKT_ERR_INVALID; }
    std::vector<Dist *> r(n, nullptr);
    for (int i = 0; i < n; i++)
    {
      if (!das[i]) { set_error("NULL da"); return DKT_ERR_INVALID; }
      r[i] = &das[i]->dist;
    }
    return p2p_attach_local(r.data(), n);
  }

  int dkt_matvec(dkt_da *da, const dkt_op *op, const double *in, double *out, double scale, unsigned flags)
  {
    if (!da || !op || !in || !out) { set_error("NULL argument"); return DKT_ERR_INVALID; }
    DA &d = da->d;
    CKA(cudaSetDevice(d.device));
    if (da->dist.active && da->dist.nranks > 1 && !da->dist.comm && !da->dist.p2p) { set_error("dry-run partition: no communicator"); return DKT_ERR_INVALID; }
    const size_t bytes = (da->dist.active ? da->dist.nOwned : d.nNodes) * sizeof(double);
    const double *din = in;
    double *dout = out;
    if ((flags & DKT_VEC_GHOSTED) && (!(flags & DKT_VEC_DEVICE) || !da->dist.active))
    {
      set_error("DKT_VEC_GHOSTED needs device vectors on a partitioned DA");
      return DKT_ERR_INVALID;
    }
    if (!(flags & DKT_VEC_DEVICE))
    {
      if (!d.d_in)
      {
        CKA(cudaMalloc((void **)&d.d_in, bytes));
        CKA(cudaMalloc((void **)&d.d_out, bytes));
      }
      CKA(cudaMemcpyAsync(d.d_in, in, bytes, cudaMemcpyHostToDevice, d.stream));
      din = d.d_in;
      dout = d.d_out;
    }
    CKA(cudaEventRecord(d.ev0, d.stream));
    // the chunk tables carry quirk Q1 statically; the conservative variant runs on the flat kernels
    if (flags & DKT_NO_Q1_MASK) flags |= DKT_MV_FLAT;
    const int rc = da->dist.active ? run_matvec_dist(d, da->dist, op, din, dout, scale, flags)
                   : (flags & DKT_MV_FLAT) ? run_matvec(d, op, din, dout, scale, flags)
                                           : run_matvec_chunked(d, op, din, dout, scale, flags);
    if (rc != DKT_OK) return rc;
    CKA(cudaEventRecord(d.ev1, d.stream));
    if (!(flags & DKT_VEC_DEVICE))
    {
      CKA(cudaMemcpyAsync(out, d.d_out, bytes, cudaMemcpyDeviceToHost, d.stream));
      CKA(cudaStreamSynchronize(d.stream));
    }
    return DKT_OK;
  }

  // dof > 1: vectors interleaved per node, [abc][abc].. (include/oda.h:296-322), the elemental operator applied to
  // every component.  The components are split into contiguous vectors, each takes the scalar path, and the results are merged:
  // the cost per DOF is the scalar one plus two streaming passes.
  int dkt_matvec_dof(dkt_da *da, const dkt_op *op, const double *in, double *out, double scale, unsigned flags, int dof)
  {
    if (dof == 1) return dkt_matvec(da, op, in, out, scale, flags);
    if (!da || !op || !in || !out) { set_error("NULL argument"); return DKT_ERR_INVALID; }
    if (dof < 1 || dof > 64) { set_error("dof must be in 1..64"); return DKT_ERR_INVALID; }
    if (flags & DKT_VEC_GHOSTED) { set_error("dkt_matvec_dof takes owned-length vectors"); return DKT_ERR_INVALID; }
    DA &d = da->d;
    CKA(cudaSetDevice(d.device));
    const size_t n = da->dist.active ? da->dist.nOwned : d.nNodes, tot = n * (size_t)dof;
    const bool host = !(flags & DKT_VEC_DEVICE);
    const size_t need = (host ? 3 : 2) * tot;
    if (d.dof_cap < need)
    {
      cudaFree(d.d_dof);
      d.d_dof = nullptr;
      d.dof_cap = 0;
      CKA(cudaMalloc((void **)&d.d_dof, std::max<size_t>(need, 1) * sizeof(double)));
      d.dof_cap = need;
    }
    double *cin = d.d_dof, *cout = d.d_dof + tot, *stage = d.d_dof + 2 * tot;
    const double *src = in;
    if (host)
    {
      CKA(cudaMemcpyAsync(stage, in, tot * sizeof(double), cudaMemcpyHostToDevice, d.stream));
      src = stage;
    }
    if (tot)
    {
      k_dof_split<<<(unsigned)((tot + 255) / 256), 256, 0, d.stream>>>(src, n, dof, cin);
      g_launches++;
    }
    for (int k = 0; k < dof; k++)
    {
      const int rc = dkt_matvec(da, op, cin + (size_t)k * n, cout + (size_t)k * n, scale, flags | DKT_VEC_DEVICE);
      if (rc != DKT_OK) return rc;
    }
    double *dst = host ? stage : out;
    if (tot)
    {
      k_dof_merge<<<(unsigned)((tot + 255) / 256), 256, 0, d.stream>>>(cout, n, dof, dst);
      g_launches++;
    }
    CKA(cudaGetLastError());
    if (host)
    {
      CKA(cudaMemcpyAsync(out, stage, tot * sizeof(double), cudaMemcpyDeviceToHost, d.stream));
      CKA(cudaStreamSynchronize(d.stream));
    }
    return DKT_OK;
  }

  // ---- ghost exchanges of ot::DA (include/oda.h:300-322) on a device vector in the ghosted layout ---------------------------
  static int ghost_begin(dkt_da *da, double *vec, int which)
  {
    if (!da || !vec) { set_error("NULL argument"); return DKT_ERR_INVALID; }
    CKA(cudaSetDevice(da->d.device));
    if (!da->dist.active) return DKT_OK;  // single rank: nothing to exchange (include/oda.tcc:216,324)
    return ghost_exchange_begin(da->d, da->dist, vec, which);
  }
  int dkt_ghost_read_begin(dkt_da *da, double *vec) { return ghost_begin(da, vec, 0); }
  int dkt_ghost_write_begin(dkt_da *da, double *vec) { return ghost_begin(da, vec, 1); }
  int dkt_ghost_read_end(dkt_da *da, double *vec)
  {
    if (!da || !vec) { set_error("NULL argument"); return DKT_ERR_INVALID; }
    return da->dist.active ? ghost_exchange_end(da->d, da->dist) : DKT_OK;
  }
  int dkt_ghost_write_end(dkt_da *da, double *vec) { return dkt_ghost_read_end(da, vec); }
  // the same exchanges on a HOST vector [owned | ghosts] (what DA::readFromGhostBegin/End get from an application)
  static int ghost_host(dkt_da *da, double *vec, int which)
  {
    if (!da || !vec) { set_error("NULL argument"); return DKT_ERR_INVALID; }
    if (!da->dist.active || da->dist.nranks <= 1) return DKT_OK;
    DA &d = da->d;
    CKA(cudaSetDevice(d.device));
    const size_t nO = da->dist.nOwned, nG = da->dist.nGhost;
    double *tmp = nullptr;
    CKA(cudaMalloc((void **)&tmp, std::max<size_t>(nO + nG, 1) * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(tmp, vec, (nO + nG) * sizeof(double), cudaMemcpyHostToDevice, d.stream);
    int rc = e == cudaSuccess ? ghost_exchange_begin(d, da->dist, tmp, which) : DKT_ERR_CUDA;
    if (rc == DKT_OK) rc = ghost_exchange_end(d, da->dist);
    if (rc == DKT_OK)
    {
      // read: the ghost segment changed; write: the owned segment accumulated what came back
      e = which == 0 ? cudaMemcpyAsync(vec + nO, tmp + nO, nG * sizeof(double), cudaMemcpyDeviceToHost, d.stream)
                     : cudaMemcpyAsync(vec, tmp, nO * sizeof(double), cudaMemcpyDeviceToHost, d.stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
      if (e != cudaSuccess) { set_error(std::string("ghost exchange: ") + cudaGetErrorString(e)); rc = DKT_ERR_CUDA; }
    }
    cudaFree(tmp);
    return rc;
  }
  int dkt_ghost_read_host(dkt_da *da, double *vec) { return ghost_host(da, vec, 0); }
  int dkt_ghost_write_host(dkt_da *da, double *vec) { return ghost_host(da, vec, 1); }

  int dkt_cg_solve(dkt_da *da, const dkt_op *op, double *x, const double *b, int max_iter, double *tol, double scale,
                   unsigned flags, int *iters, int *status)
  {
    if (!da || !op || !x || !b || !tol || !iters || !status) { set_error("NULL argument"); return DKT_ERR_INVALID; }
    if (flags & DKT_VEC_GHOSTED) { set_error("dkt_cg_solve takes owned-length vectors"); return DKT_ERR_INVALID; }
    DA &d = da->d;
    CKA(cudaSetDevice(d.device));
    if (da->dist.active && da->dist.nranks > 1 && !da->dist.comm) { set_error("dry-run partition: no communicator"); return DKT_ERR_INVALID; }
    const size_t bytes = (da->dist.active ? da->dist.nOwned : d.nNodes) * sizeof(double);
    double *dx = x;
    const double *db = b;
    double *tmp = nullptr;
    if (!(flags & DKT_VEC_DEVICE))
    {
      CKA(cudaMalloc((void **)&tmp, 2 * std::max<size_t>(bytes, 8)));
      dx = tmp;
      db = tmp + bytes / sizeof(double);
      CKA(cudaMemcpyAsync(dx, x, bytes, cudaMemcpyHostToDevice, d.stream));
      CKA(cudaMemcpyAsync((void *)db, b, bytes, cudaMemcpyHostToDevice, d.stream));
    }
    const int rc = cg_solve(d, &da->dist, op, dx, db, max_iter, tol, scale, flags, iters, status);
    if (rc == DKT_OK && !(flags & DKT_VEC_DEVICE))
    {
      CKA(cudaMemcpyAsync(x, dx, bytes, cudaMemcpyDeviceToHost, d.stream));
      CKA(cudaStreamSynchronize(d.stream));
    }
    cudaFree(tmp);
    return rc;
  }

  int dkt_last_kernel_ms(dkt_da *da, float *ms)
  {
    if (!da || !ms) { set_error("NULL argument"); return DKT_ERR_INVALID; }
    CKA(cudaEventSynchronize(da->d.ev1));
    CKA(cudaEventElapsedTime(ms, da->d.ev0, da->d.ev1));
    return DKT_OK;
  }
  int dkt_da_chunk_info(const dkt_da *da, uint64_t out[16])
  {
    if (!da || !out) { set_error("NULL argument"); return DKT_ERR_INVALID; }
    // aggregated over the regular per-element sets (out[0..4]), the hanging ones (out[5..9]) and the sibling-family sets
    // (out[10..14]; out[15] = elements in families)
    for (int i = 0; i < 16; i++) out[i] = 0;
    for (const ChunkSet &cs : da->d.sets)
    {
      uint64_t *o = out + (cs.kind == 2 ? 10 : cs.rows == 1 ? 0 : 5);
      if (cs.kind == 2) out[15] += cs.nElem << da->d.dim;
      o[0] += cs.nChunks; o[1] = cs.elemsPerChunk; o[2] = std::max<uint64_t>(o[2], cs.maxNloc);
      o[3] = std::max<uint64_t>(o[3], cs.maxLen); o[4] += cs.totalNodes;
    }
    return DKT_OK;
  }
  void *dkt_da_stream(dkt_da *da) { return da ? (void *)da->d.stream : nullptr; }
  int dkt_da_set_stream(dkt_da *da, void *cuda_stream)
  {
    if (!da) { set_error("NULL da"); return DKT_ERR_INVALID; }
    da->d.stream = cuda_stream ? (cudaStream_t)cuda_stream : da->d.own_stream;
    return DKT_OK;
  }
}
