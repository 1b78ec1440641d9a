/* C-ABI driver around the UNMODIFIED reference (paralab/Dendro-KT) for use as the parity oracle.
 * TEST INFRASTRUCTURE - never linked or loaded by the product path (dendro-kt_b200/).
 *
 * Built by oracle/build_ref.sh against the sources where they lie under /root/reference
 * (plus the five `return` statements SURVEY.md App. A lists, applied to a scratch copy),
 * with oracle/shim/mpi.h (single rank) and oracle/shim/lapack_shim.cpp.  Output:
 * oracle/_ref/libdktref_{morton,hilbert}.so.
 *
 * Everything here calls the reference's own public API:
 *   ot::SFC_Tree<T,dim>::distTreeBalancing / locTreeSort   (include/tsort.h:228-304)
 *   ot::DA<dim>(tree, nEle, comm, order)                     (include/oda.h:150, src/oda.cpp:46)
 *   feMatrix<LeafT,dim>::matVec -> fem::matvec               (FEM/include/feMatrix.h:190, matvec.h:232)
 *   HeatEq::HeatMat<dim>                                     (FEM/examples/src/heatMat.cpp:46)
 */
#include "mpi.h"
#include "treeNode.h"
#include "tsort.h"
#include "nsort.h"
#include "hcurvedata.h"
#include "oda.h"
#include "feMatrix.h"
#include "heatMat.h"
#include "refel.h"
#include "testAdaptiveExamples.h"

#include <chrono>
#include <cmath>
#include <cstdint>
#include <vector>

namespace
{
using T = unsigned int;

int g_dim = 0;

struct TreeBase { int m_dim; virtual ~TreeBase() {} };
template <unsigned dim> struct TreeH : TreeBase { std::vector<ot::TreeNode<T, dim>> v; };

struct DABase { int m_dim; int m_order; virtual ~DABase() {} };
template <unsigned dim> struct DAH : DABase { ot::DA<dim> *da = nullptr; ~DAH() { delete da; } };

/* Operator kinds understood by dktref_matvec. */
enum { OP_IDENTITY = 0, OP_DENSE = 1, OP_HEATMAT = 2 };

/* A feMatrix leaf whose elemental operator is either the identity (test/testMatvec.cpp:177-198)
 * or out = scale * 2^(-alpha*L) * Kref^T-convention-free dense product:
 *   out[i] = scale * s(L) * sum_j Kref[i*N+j] * in[j],   L = level recovered from coords. */
template <unsigned dim>
class ProbeMat : public feMatrix<ProbeMat<dim>, dim>
{
public:
  int kind = OP_IDENTITY;
  const double *Kref = nullptr;
  double alpha = 0.0;
  int dirichlet = 0;
  long ncalls = 0;
  unsigned N = 0;
  ot::DA<dim> *m_da;

  ProbeMat(ot::DA<dim> *da) : feMatrix<ProbeMat<dim>, dim>(da, 1), m_da(da) { N = da->getNumNodesPerElement(); }

  virtual void elementalMatVec(const VECType *in, VECType *out, double *coords, double scale)
  {
    ncalls++;
    if (kind == OP_IDENTITY)
    {
      for (unsigned i = 0; i < N; i++) out[i] = in[i];
      return;
    }
    const double h = coords[(N - 1) * dim + 0] - coords[0];
    const double s = scale * std::pow(h, alpha);
    for (unsigned i = 0; i < N; i++)
    {
      double acc = 0.0;
      for (unsigned j = 0; j < N; j++) acc += Kref[i * N + j] * in[j];
      out[i] = s * acc;
    }
  }

  bool preMatVec(const VECType *, VECType *out, double)
  {
    if (dirichlet) zeroBdy(out);
    return true;
  }
  bool postMatVec(const VECType *, VECType *out, double)
  {
    if (dirichlet) zeroBdy(out);
    return true;
  }

private:
  void zeroBdy(VECType *v)
  {
    std::vector<unsigned int> b;
    m_da->getBoundaryNodeIndices(b);
    for (unsigned i : b) v[i] = 0.0;
  }
};

template <unsigned dim>
TreeBase *tree_from_points(const uint32_t *pts, long n, int maxPts, double flex, int balance)
{
  std::vector<ot::TreeNode<T, dim>> points;
  points.reserve(n);
  for (long i = 0; i < n; i++)
  {
    std::array<T, dim> c;
    for (unsigned d = 0; d < dim; d++) c[d] = pts[i * dim + d];
    points.push_back(ot::TreeNode<T, dim>(c, m_uiMaxDepth));
  }
  TreeH<dim> *h = new TreeH<dim>;
  h->m_dim = dim;
  if (balance)
    ot::SFC_Tree<T, dim>::distTreeBalancing(points, h->v, maxPts, flex, MPI_COMM_WORLD);
  else
    ot::SFC_Tree<T, dim>::distTreeConstruction(points, h->v, maxPts, flex, MPI_COMM_WORLD);
  return h;
}

template <unsigned dim>
TreeBase *tree_from_elements(const uint32_t *xyz, const uint8_t *lev, long n, int sort)
{
  TreeH<dim> *h = new TreeH<dim>;
  h->m_dim = dim;
  h->v.reserve(n);
  for (long i = 0; i < n; i++)
  {
    std::array<T, dim> c;
    for (unsigned d = 0; d < dim; d++) c[d] = xyz[i * dim + d];
    h->v.push_back(ot::TreeNode<T, dim>(1, c, lev[i]));
  }
  if (sort)
    ot::SFC_Tree<T, dim>::template locTreeSort<ot::TreeNode<T, dim>>(&(*h->v.begin()), 0, (ot::RankI)h->v.size(), 1, m_uiMaxDepth, 0);
  return h;
}

template <unsigned dim>
TreeBase *tree_example(int which, int depth, int sort)
{
  TreeH<dim> *h = new TreeH<dim>;
  h->m_dim = dim;
  if (which == 1) Example1<dim>::fill_tree(depth, h->v);
  else if (which == 2) Example2<dim>::fill_tree(depth, h->v);
  else Example3<dim>::fill_tree(depth, h->v);
  if (sort)
    ot::SFC_Tree<T, dim>::template locTreeSort<ot::TreeNode<T, dim>>(&(*h->v.begin()), 0, (ot::RankI)h->v.size(), 1, m_uiMaxDepth, 0);
  return h;
}

template <unsigned dim>
void tree_export(TreeBase *b, uint32_t *xyz, uint8_t *lev)
{
  TreeH<dim> *h = static_cast<TreeH<dim> *>(b);
  for (size_t i = 0; i < h->v.size(); i++)
  {
    for (unsigned d = 0; d < dim; d++) xyz[i * dim + d] = h->v[i].getX(d);
    lev[i] = (uint8_t)h->v[i].getLevel();
  }
}

template <unsigned dim>
DABase *da_create(TreeBase *b, int order)
{
  TreeH<dim> *h = static_cast<TreeH<dim> *>(b);
  DAH<dim> *d = new DAH<dim>;
  d->m_dim = dim;
  d->m_order = order;
  d->da = new ot::DA<dim>(&(*h->v.cbegin()), (unsigned)h->v.size(), MPI_COMM_WORLD, (unsigned)order, 100, 0.3);
  return d;
}

template <unsigned dim>
void da_export_nodes(DABase *b, uint32_t *xyz, uint8_t *lev)
{
  ot::DA<dim> *da = static_cast<DAH<dim> *>(b)->da;
  const ot::TreeNode<T, dim> *tn = da->getTNCoords();
  const unsigned n = da->getTotalNodalSz();
  for (unsigned i = 0; i < n; i++)
  {
    for (unsigned d = 0; d < dim; d++) xyz[i * dim + d] = tn[i].getX(d);
    lev[i] = (uint8_t)tn[i].getLevel();
  }
}

template <unsigned dim>
long da_boundary(DABase *b, uint32_t *ids)
{
  ot::DA<dim> *da = static_cast<DAH<dim> *>(b)->da;
  std::vector<unsigned int> v;
  da->getBoundaryNodeIndices(v);
  if (ids) for (size_t i = 0; i < v.size(); i++) ids[i] = v[i];
  return (long)v.size();
}

template <unsigned dim>
double da_matvec(DABase *b, int kind, const double *Kref, double alpha, int dirichlet, const double *in, double *out,
                 double scale, int nwarm, int niter, long *ncalls)
{
  ot::DA<dim> *da = static_cast<DAH<dim> *>(b)->da;
  double secs = 0.0;
  if (kind == OP_HEATMAT)
  {
    HeatEq::HeatMat<dim> mat(da, 1);
    /* domain [-0.5,0.5]^dim as in bench/src/matvec_bench_adaptive.cpp:160-166 */
    mat.setProblemDimensions(Point<dim>(-0.5, -0.5, -0.5), Point<dim>(0.5, 0.5, 0.5));
    for (int i = 0; i < nwarm; i++) mat.matVec(in, out, scale);
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < niter; i++) mat.matVec(in, out, scale);
    secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (ncalls) *ncalls = -1;
  }
  else
  {
    ProbeMat<dim> mat(da);
    mat.kind = kind;
    mat.Kref = Kref;
    mat.alpha = alpha;
    mat.dirichlet = dirichlet;
    for (int i = 0; i < nwarm; i++) mat.matVec(in, out, scale);
    mat.ncalls = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < niter; i++) mat.matVec(in, out, scale);
    secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (ncalls) *ncalls = mat.ncalls / (niter > 0 ? niter : 1);
  }
  return niter > 0 ? secs / niter : 0.0;
}

/* One call of HeatMat<dim>::elementalMatVec on an explicit element (for building K tables). */
template <unsigned dim>
void heat_elemental(DABase *b, const double *in, double *out, double *coords)
{
  ot::DA<dim> *da = static_cast<DAH<dim> *>(b)->da;
  HeatEq::HeatMat<dim> mat(da, 1);
  mat.setProblemDimensions(Point<dim>(-0.5, -0.5, -0.5), Point<dim>(0.5, 0.5, 0.5));
  mat.elementalMatVec(in, out, coords, 1.0);
}

#define DISPATCH(dimv, call2, call3, call4) ((dimv) == 2 ? (call2) : (dimv) == 3 ? (call3) : (call4))
} // namespace

/* libgomp is not linkable in this image; src/profiler.cpp only needs a wall clock. */
extern "C" double omp_get_wtime(void) { return MPI_Wtime(); }

extern "C"
{
  /* Must be called first: sets the reference's global max depth and loads the SFC tables. */
  int dktref_init(int dim, int maxDepth)
  {
    if (dim < 2 || dim > 4) return 1;
    m_uiMaxDepth = (unsigned)maxDepth;
    if (g_dim != dim) { _InitializeHcurve(dim); g_dim = dim; }
    return 0;
  }
  int dktref_is_hilbert(void)
  {
#ifdef HILBERT_ORDERING
    return 1;
#else
    return 0;
#endif
  }
  /* SFC tables as the reference holds them (include/hcurvedata.h:49-59). */
  int dktref_tables(int dim, char *rot, int *htab)
  {
    const int nrot = _KD_ROTATIONS_SIZE(dim), nh = _KD_HILBERT_TABLE_SIZE(dim);
    if (rot) memcpy(rot, rotations, nrot);
    if (htab) memcpy(htab, HILBERT_TABLE, sizeof(int) * nh);
    return nrot;
  }

  void *dktref_tree_from_points(int dim, const uint32_t *pts, long n, int maxPts, double flex, int balance)
  {
    return DISPATCH(dim, tree_from_points<2>(pts, n, maxPts, flex, balance), tree_from_points<3>(pts, n, maxPts, flex, balance),
                    tree_from_points<4>(pts, n, maxPts, flex, balance));
  }
  void *dktref_tree_from_elements(int dim, const uint32_t *xyz, const uint8_t *lev, long n, int sort)
  {
    return DISPATCH(dim, tree_from_elements<2>(xyz, lev, n, sort), tree_from_elements<3>(xyz, lev, n, sort),
                    tree_from_elements<4>(xyz, lev, n, sort));
  }
  void *dktref_tree_example(int dim, int which, int depth, int sort)
  {
    return DISPATCH(dim, tree_example<2>(which, depth, sort), tree_example<3>(which, depth, sort), tree_example<4>(which, depth, sort));
  }
  long dktref_tree_size(void *t)
  {
    TreeBase *b = (TreeBase *)t;
    return DISPATCH(b->m_dim, (long)static_cast<TreeH<2> *>(b)->v.size(), (long)static_cast<TreeH<3> *>(b)->v.size(),
                    (long)static_cast<TreeH<4> *>(b)->v.size());
  }
  void dktref_tree_export(void *t, uint32_t *xyz, uint8_t *lev)
  {
    TreeBase *b = (TreeBase *)t;
    DISPATCH(b->m_dim, tree_export<2>(b, xyz, lev), tree_export<3>(b, xyz, lev), tree_export<4>(b, xyz, lev));
  }
  void dktref_tree_destroy(void *t) { delete (TreeBase *)t; }

  void *dktref_da_create(void *t, int order)
  {
    TreeBase *b = (TreeBase *)t;
    return DISPATCH(b->m_dim, da_create<2>(b, order), da_create<3>(b, order), da_create<4>(b, order));
  }
  long dktref_da_num_nodes(void *d)
  {
    DABase *b = (DABase *)d;
    return DISPATCH(b->m_dim, (long)static_cast<DAH<2> *>(b)->da->getTotalNodalSz(), (long)static_cast<DAH<3> *>(b)->da->getTotalNodalSz(),
                    (long)static_cast<DAH<4> *>(b)->da->getTotalNodalSz());
  }
  void dktref_da_export_nodes(void *d, uint32_t *xyz, uint8_t *lev)
  {
    DABase *b = (DABase *)d;
    DISPATCH(b->m_dim, da_export_nodes<2>(b, xyz, lev), da_export_nodes<3>(b, xyz, lev), da_export_nodes<4>(b, xyz, lev));
  }
  long dktref_da_boundary(void *d, uint32_t *ids)
  {
    DABase *b = (DABase *)d;
    return DISPATCH(b->m_dim, da_boundary<2>(b, ids), da_boundary<3>(b, ids), da_boundary<4>(b, ids));
  }
  /* out[0..5] = getLocalNodalSz, getLocalNodeBegin, getTotalNodalSz, getNpesAll, getRankAll, getGlobalNodeSz (several ranks) */
  void dktref_da_local_info(void *d, long *out)
  {
    DABase *b = (DABase *)d;
#define DKT_INFO(D)                                                                                                    \
  {                                                                                                                    \
    ot::DA<D> *da = static_cast<DAH<D> *>(b)->da;                                                                      \
    out[0] = da->getLocalNodalSz(); out[1] = da->getLocalNodeBegin(); out[2] = da->getTotalNodalSz();                  \
    out[3] = da->getNpesAll(); out[4] = da->getRankAll(); out[5] = (long)da->getGlobalNodeSz();                       \
  }
    if (b->m_dim == 2) DKT_INFO(2) else if (b->m_dim == 3) DKT_INFO(3) else DKT_INFO(4)
#undef DKT_INFO
  }
  void dktref_da_destroy(void *d) { delete (DABase *)d; }

  /* 1-D operators of the reference's RefElement(dim, order), row-major M x M, M = order+1. */
  double dktref_refel(int dim, int order, double *ip0, double *ip1, double *Q, double *Dg, double *w)
  {
    RefElement r(dim, order);
    const int M = order + 1;
    memcpy(ip0, r.getIMChild0(), sizeof(double) * M * M);
    memcpy(ip1, r.getIMChild1(), sizeof(double) * M * M);
    memcpy(Q, r.getQ1d(), sizeof(double) * M * M);
    memcpy(Dg, r.getDg1d(), sizeof(double) * M * M);
    memcpy(w, r.getWgq(), sizeof(double) * M);
    return r.getElementSz();
  }

  /* v = A u through feMatrix::matVec.  Returns seconds per call (niter timed calls after nwarm). */
  double dktref_matvec(void *d, int kind, const double *Kref, double alpha, int dirichlet, const double *in, double *out,
                       double scale, int nwarm, int niter, long *ncalls)
  {
    DABase *b = (DABase *)d;
    return DISPATCH(b->m_dim, da_matvec<2>(b, kind, Kref, alpha, dirichlet, in, out, scale, nwarm, niter, ncalls),
                    da_matvec<3>(b, kind, Kref, alpha, dirichlet, in, out, scale, nwarm, niter, ncalls),
                    da_matvec<4>(b, kind, Kref, alpha, dirichlet, in, out, scale, nwarm, niter, ncalls));
  }
  void dktref_heat_elemental(void *d, const double *in, double *out, double *coords)
  {
    DABase *b = (DABase *)d;
    DISPATCH(b->m_dim, heat_elemental<2>(b, in, out, coords), heat_elemental<3>(b, in, out, coords), heat_elemental<4>(b, in, out, coords));
  }
}
