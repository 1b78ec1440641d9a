// feMatrix<LeafT,dim>::matVec over libdkt.so (FEM/include/feMatrix.h:15-259 of the reference).
//
// The reference calls the user's `elementalMatVec(in, out, coords, scale)` once per element from a
// recursive host traversal (FEM/include/matvec.h:468).  A host virtual cannot run per element on the
// GPU, so the callback is PROBED instead: on axis-aligned cells a translation-invariant elemental
// operator is K_e = h^alpha * K_ref, and both K_ref and alpha are recovered from N + N calls of the
// callback on two reference cells.  The recovered operator is then VERIFIED against the callback on
// sample elements of the actual tree (random input, 1e-12 relative); if the callback is not of that
// form, matVec throws instead of silently computing something else.  There is no host fallback.
#ifndef DKT_HOST_FEMATRIX_H
#define DKT_HOST_FEMATRIX_H

#include <cmath>
#include <random>
#include <vector>

#include "feMat.h"
#include "refel.h"
#include "tensor.h"

template <typename LeafT, unsigned int dim>
class feMatrix : public feMat<dim>
{
protected:
  static constexpr unsigned int m_uiDim = dim;
  unsigned int m_uiDof;

  // device form of elementalMatVec, valid for m_opScale
  std::vector<double> m_kref;
  double m_alpha = 0.0, m_opScale = 0.0;
  bool m_opReady = false;
  int m_lastIters = 0;
  std::vector<VECType> m_inGhosted, m_outGhosted;  // members, not function statics: re-entrant per object

public:
  feMatrix(ot::DA<dim> *da, unsigned int dof = 1) : feMat<dim>(da), m_uiDof(dof)
  {
    if (dof != 1) throw std::runtime_error("matvec only supports dof==1 (FEM/include/matvec.h:27)");
  }
  ~feMatrix() {}

  /** the user's elemental operator, same signature as FEM/include/feMatrix.h:56 */
  virtual void elementalMatVec(const VECType *in, VECType *out, double *coords, double scale) = 0;

  LeafT &asLeaf() { return static_cast<LeafT &>(*this); }
  /** CRTP hooks (feMatrix.h:91-122); the base versions do nothing */
  bool preMatVec(const VECType *, VECType *, double = 1.0) { return false; }
  bool postMatVec(const VECType *, VECType *, double = 1.0) { return false; }

  /** v = A u (feMatrix.h:190-259): copy to the ghosted buffer, preMatVec, [ghost read], the
   *  matvec on the GPU, [ghost write], copy back, postMatVec. */
  virtual void matVec(const VECType *in, VECType *out, double scale = 1.0)
  {
    ot::DA<dim> *da = feMat<dim>::m_uiOctDA;
    ensureDeviceOperator(scale);
    da->template createVector<VECType>(m_inGhosted, false, true, m_uiDof);
    da->template createVector<VECType>(m_outGhosted, false, true, m_uiDof);
    VECType *inG = m_inGhosted.data(), *outG = m_outGhosted.data();
    da->template nodalVecToGhostedNodal<VECType>(in, inG, true, m_uiDof);
    asLeaf().preMatVec(in, inG + da->getLocalNodeBegin(), scale);
    dkt_op op;
    op.kind = DKT_OP_DENSE;
    op.kref = m_kref.data();
    op.alpha = m_alpha;
    op.dirichlet = 0;
    op.terms = 0;
    dkt_host::check(dkt_matvec(da->handle(), &op, inG, outG, 1.0, DKT_VEC_HOST), "feMatrix::matVec");
    da->template ghostedNodalToNodalVec<VECType>(outG, out, true, m_uiDof);
    asLeaf().postMatVec(outG + da->getLocalNodeBegin(), out, scale);
  }

  /** HeatMat::cgSolve (FEM/examples/src/heatMat.cpp:165-325) with every vector resident on the GPU: x is
   *  the initial guess on entry and the solution on return, tol returns the achieved |r|_inf/|b|_inf.
   *  `dirichletRows` stands for the leaf's pre/postMatVec boundary zeroing (host hooks cannot run inside
   *  the resident loop).  Returns 0 if converged within max_iter, 1 otherwise. */
  int cgSolve(double *x, double *b, int max_iter, double &tol, bool dirichletRows = true, double scale = 1.0)
  {
    ot::DA<dim> *da = feMat<dim>::m_uiOctDA;
    ensureDeviceOperator(scale);
    dkt_op op;
    op.kind = DKT_OP_DENSE;
    op.kref = m_kref.data();
    op.alpha = m_alpha;
    op.dirichlet = dirichletRows ? 1 : 0;
    int iters = 0, status = 1;
    dkt_host::check(dkt_cg_solve(da->handle(), &op, x, b, max_iter, &tol, 1.0, DKT_VEC_HOST, &iters, &status), "feMatrix::cgSolve");
    m_lastIters = iters;
    return status;
  }
  int lastIterations() const { return m_lastIters; }

  /** the recovered reference-cell matrix and level exponent (after the first matVec) */
  const std::vector<double> &deviceKref() const { return m_kref; }
  double deviceAlpha() const { return m_alpha; }

protected:
  void probe(unsigned level, const unsigned *anchorCells, double scale, std::vector<double> &K)
  {
    ot::DA<dim> *da = feMat<dim>::m_uiOctDA;
    const unsigned N = da->getNumNodesPerElement(), M = da->getElementOrder() + 1;
    const double h = std::ldexp(1.0, -(int)level);
    std::vector<double> coords((size_t)N * dim), ein(N), eout(N);
    for (unsigned n = 0; n < N; n++)
    {
      unsigned r = n;
      for (unsigned d = 0; d < dim; d++)
      {
        coords[(size_t)n * dim + d] = h * (anchorCells[d] + double(r % M) / (M - 1));  // lexicographic, axis 0 fastest
        r /= M;
      }
    }
    K.assign((size_t)N * N, 0.0);
    for (unsigned j = 0; j < N; j++)
    {
      std::fill(ein.begin(), ein.end(), 0.0);
      ein[j] = 1.0;
      std::vector<double> c = coords;
      elementalMatVec(ein.data(), eout.data(), c.data(), scale);
      for (unsigned i = 0; i < N; i++) K[(size_t)i * N + j] = eout[i];
    }
  }

  void ensureDeviceOperator(double scale)
  {
    if (m_opReady && scale == m_opScale) return;
    ot::DA<dim> *da = feMat<dim>::m_uiOctDA;
    const unsigned N = da->getNumNodesPerElement();
    const unsigned zero[dim] = {};
    std::vector<double> K2, K3;
    probe(2, zero, scale, K2);
    probe(3, zero, scale, K3);
    // K(L) = 2^(-alpha L) K_ref: alpha from the largest entry, then every entry must agree
    size_t big = 0;
    for (size_t i = 0; i < K2.size(); i++)
      if (std::fabs(K2[i]) > std::fabs(K2[big])) big = i;
    if (K2[big] == 0.0 || K3[big] == 0.0 || K2[big] / K3[big] <= 0.0)
      throw std::runtime_error("feMatrix: elementalMatVec returned a zero/sign-changing operator on the probe cells");
    m_alpha = std::log2(K2[big] / K3[big]);
    const double f = std::exp2(-m_alpha);
    double tol = 0.0;
    for (size_t i = 0; i < K2.size(); i++) tol = std::max(tol, std::fabs(K3[i] - f * K2[i]));
    if (tol > 1e-12 * std::fabs(K3[big]))
      throw std::runtime_error("feMatrix: elementalMatVec is not of the form h^alpha * K_ref (level scaling differs between entries); "
                               "no device operator can represent it");
    m_kref.resize(K2.size());
    const double s = std::exp2(2.0 * m_alpha);
    for (size_t i = 0; i < K2.size(); i++) m_kref[i] = K2[i] * s;
    // verify on sample elements of the tree: random input, other positions and levels
    const ot::TreeNode<unsigned, dim> *tn = da->getTNCoords();
    std::mt19937_64 gen(12345);
    std::uniform_real_distribution<double> dist(-1.0, 1.0);
    const unsigned nSamples = 64;
    std::vector<double> ein(N), eout(N), want(N);
    for (unsigned sidx = 0; sidx < nSamples && da->getTotalNodalSz() > 0; sidx++)
    {
      const ot::TreeNode<unsigned, dim> &nd = tn[(size_t)(gen() % da->getTotalNodalSz())];
      unsigned level = std::max(1u, nd.getLevel());
      unsigned cell[dim];
      for (unsigned d = 0; d < dim; d++)
      {
        cell[d] = nd.getX(d) >> (m_uiMaxDepth - level);
        if (cell[d] >= (1u << level)) cell[d] = (1u << level) - 1;  // nodes on the upper domain boundary
      }
      std::vector<double> Kc;
      // one random vector through the callback on that cell
      {
        const unsigned M = da->getElementOrder() + 1;
        const double h = std::ldexp(1.0, -(int)level);
        std::vector<double> coords((size_t)N * dim);
        for (unsigned n = 0; n < N; n++)
        {
          unsigned r = n;
          for (unsigned d = 0; d < dim; d++) { coords[(size_t)n * dim + d] = h * (cell[d] + double(r % M) / (M - 1)); r /= M; }
        }
        for (unsigned i = 0; i < N; i++) ein[i] = dist(gen);
        elementalMatVec(ein.data(), eout.data(), coords.data(), scale);
        const double sc = std::exp2(-m_alpha * level);
        double mx = 0.0, err = 0.0;
        for (unsigned i = 0; i < N; i++)
        {
          double acc = 0.0;
          for (unsigned j = 0; j < N; j++) acc += m_kref[(size_t)i * N + j] * ein[j];
          want[i] = sc * acc;
          mx = std::max(mx, std::fabs(want[i]));
          err = std::max(err, std::fabs(want[i] - eout[i]));
        }
        if (err > 1e-12 * std::max(mx, 1e-300))
          throw std::runtime_error("feMatrix: elementalMatVec depends on the element position or is not level-scaled: the device "
                                   "operator recovered from the probe cells does not reproduce it on a sample element");
      }
    }
    m_opScale = scale;
    m_opReady = true;
  }
};
#endif
