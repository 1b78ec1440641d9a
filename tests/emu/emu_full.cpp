// The whole single-rank pipeline of libdkt under the CUDA-on-CPU emulation (tests only): dkt_build.cu's table
// construction + dkt_chunks.cu's chunk tables and kernels, driven like dkt_da_create / dkt_matvec drive them.
// Lets the CPU test-suite check the PRODUCT's construction code bit-exactly against the reference's golden fixtures.
#include "dkt_internal.h"
#include "cuda_emu.h"

#include <string>

namespace dkt { extern std::string g_emu_err; }
#define g_err g_emu_err

using namespace dkt;


extern "C" void *emu_da_create(int dim, int order, int max_depth, int sfc, const uint32_t *xyz, const uint8_t *lev, uint64_t n,
                               const double *ip0, const double *ip1, unsigned flags)
{
  DA *da = new DA;
  da->dim = dim; da->order = order; da->max_depth = max_depth; da->sfc_mode = sfc;
  da->M = order + 1;
  da->N = 1;
  for (int i = 0; i < dim; i++) da->N *= da->M;
  for (int i = 0; i < da->M * da->M; i++) { da->ip[0][i] = ip0[i]; da->ip[1][i] = ip1[i]; }
  int rc = build_da(*da, xyz, lev, n, flags);
  if (rc == DKT_OK) rc = build_chunks(*da);
  if (rc != DKT_OK)
  {
    const std::string keep = g_err;
    free_da(*da);
    delete da;
    g_err = keep;
    return nullptr;
  }
  return da;
}
extern "C" void emu_da_sizes(void *h, uint64_t *out)
{
  const DA &d = *(DA *)h;
  out[0] = d.nElem; out[1] = d.nMv; out[2] = d.nReg; out[3] = d.nHang; out[4] = d.nNodes; out[5] = d.nBdy; out[6] = d.nSplit;
  out[7] = (uint64_t)d.tree_class; out[8] = (uint64_t)d.N; out[9] = (uint64_t)d.finest_level; out[10] = d.sets.size();
}
template <typename T>
static void cp(T *dst, const T *src, size_t n)
{
  if (dst && src && n) memcpy(dst, src, n * sizeof(T));
}
extern "C" void emu_da_export(void *h, uint32_t *elem_xyz, uint8_t *elem_lev, uint32_t *node_xyz, uint8_t *node_lev, uint32_t *bdy,
                              uint32_t *mv_xyz, uint8_t *mv_lev, uint32_t *e2n, uint32_t *pnode, uint8_t *child)
{
  const DA &d = *(DA *)h;
  cp(elem_xyz, d.d_elem_xyz, d.nElem * d.dim); cp(elem_lev, d.d_elem_lev, d.nElem);
  cp(node_xyz, d.d_node_xyz, d.nNodes * d.dim); cp(node_lev, d.d_node_lev, d.nNodes);
  cp(bdy, d.d_bdy, d.nBdy);
  cp(mv_xyz, d.d_mv_xyz, d.nMv * d.dim); cp(mv_lev, d.d_mv_lev, d.nMv);
  cp(e2n, d.d_e2n, d.nMv * d.N); cp(pnode, d.d_pnode, d.nHang * d.N); cp(child, d.d_child, d.nHang);
}
extern "C" int emu_da_matvec(void *h, int kind, const double *kref, double alpha, int dirichlet, const double *in, double *out, double scale,
                             unsigned flags)
{
  DA &d = *(DA *)h;
  dkt_op op;
  op.kind = kind; op.kref = kref; op.alpha = alpha; op.dirichlet = dirichlet; op.terms = 0;
  double *din = nullptr, *dout = nullptr;
  cudaMalloc(&din, d.nNodes * sizeof(double));
  cudaMalloc(&dout, d.nNodes * sizeof(double));
  memcpy(din, in, d.nNodes * sizeof(double));
  // like dkt_matvec: the Q1-free variant and DKT_MV_FLAT run on the flat kernels
  if (flags & DKT_NO_Q1_MASK) flags |= DKT_MV_FLAT;
  const int rc = (flags & DKT_MV_FLAT) ? run_matvec(d, &op, din, dout, scale, flags) : run_matvec_chunked(d, &op, din, dout, scale, flags);
  memcpy(out, dout, d.nNodes * sizeof(double));
  cudaFree(din);
  cudaFree(dout);
  return rc;
}
// dkt_cg_solve on device vectors (HeatMat::cgSolve, FEM/examples/src/heatMat.cpp:165-325)
extern "C" int emu_da_cg(void *h, int kind, const double *kref, double alpha, int dirichlet, double *x, const double *b, int max_iter,
                         double *tol, double scale, unsigned flags, int *iters, int *status)
{
  DA &d = *(DA *)h;
  dkt_op op;
  op.kind = kind; op.kref = kref; op.alpha = alpha; op.dirichlet = dirichlet; op.terms = 0;
  double *dx = nullptr, *db = nullptr;
  cudaMalloc(&dx, d.nNodes * sizeof(double));
  cudaMalloc(&db, d.nNodes * sizeof(double));
  memcpy(dx, x, d.nNodes * sizeof(double));
  memcpy(db, b, d.nNodes * sizeof(double));
  const int rc = cg_solve(d, nullptr, &op, dx, db, max_iter, tol, scale, flags, iters, status);
  memcpy(x, dx, d.nNodes * sizeof(double));
  cudaFree(dx);
  cudaFree(db);
  return rc;
}
extern "C" void emu_da_destroy(void *h)
{
  DA *da = (DA *)h;
  free_da(*da);
  delete da;
}

// ---- dkt_tree.cu: trees from points -------------------------------------------------------------------------------------
extern "C" void *emu_tree_from_points(int dim, int max_depth, int sfc, const uint32_t *pts, uint64_t n, uint64_t max_pts, int balance)
{
  Tree *t = new Tree;
  t->dim = dim; t->max_depth = max_depth; t->sfc_mode = sfc;
  if (build_tree(*t, pts, n, max_pts, balance != 0, 0) != DKT_OK)
  {
    const std::string keep = g_err;
    free_tree(*t);
    delete t;
    g_err = keep;
    return nullptr;
  }
  return t;
}
extern "C" uint64_t emu_tree_size(void *h) { return ((Tree *)h)->n; }
extern "C" void emu_tree_export(void *h, uint32_t *xyz, uint8_t *lev)
{
  const Tree &t = *(Tree *)h;
  cp(xyz, t.d_xyz, t.n * t.dim);
  cp(lev, t.d_lev, t.n);
}
extern "C" void emu_tree_destroy(void *h)
{
  free_tree(*(Tree *)h);
  delete (Tree *)h;
}
