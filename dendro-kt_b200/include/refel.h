// RefElement: the 1-D operators of the reference element of order p on [-1, 1] and the tensor-product parent <-> child
// interpolation (reference: FEM/include/refel.h:187-271, FEM/src/refel.cpp:21-282).
//
// The reference assembles its matrices from Jacobi-polynomial Vandermonde matrices and LAPACK dgesv; here every entry is written
// down directly from the Lagrange basis l_k on the p + 1 EQUISPACED nodes r_j = -1 + 2 j / p (the nodes Dendro-KT's elements
// use, include/nsort.tcc:313-389) and the (p + 1)-point Gauss-Legendre rule, so there is no linear solve and no LAPACK.  Layout
// as in the reference: M = p + 1, A[k * M + j] takes input (node) index k to output index j.
//   ip_1D_0 / ip_1D_1  parent nodes -> nodes of child 0 ([-1, 0]) / child 1 ([0, 1]):  l_k((r_j -+ 1) / 2)
//   quad_1D (Q)        nodes -> Gauss points g_j:  l_k(g_j);     Dg: l_k'(g_j);     Dr: l_k'(r_j);     w: Gauss weights
// Checked against the reference's own object at orders 1-3 in tests/test_cpp_api.py (1e-14).  getElementSz() is 2, the length of
// [-1, 1]; the reference returns 0 at order 2 because its GLL points are never filled there - a bug, not reproduced.
#ifndef DKT_HOST_REFEL_H
#define DKT_HOST_REFEL_H

#include <cassert>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <vector>

class RefElement
{
  int m_uiOrder = 1, m_uiDimension = 3, m_uiNrp = 2;
  std::vector<double> r, g, w, wgll;
  std::vector<double> ip_1D_0, ip_1D_1, ipT_1D_0, ipT_1D_1, quad_1D, quadT_1D, Dg, DgT, Dr, DrT;
  mutable std::vector<double> im_vec1, im_vec2;

  static double lagrange(const std::vector<double> &x, int k, double t)
  {
    double v = 1.0;
    for (int i = 0; i < (int)x.size(); i++)
      if (i != k) v *= (t - x[i]) / (x[k] - x[i]);
    return v;
  }
  static double dlagrange(const std::vector<double> &x, int k, double t)
  {
    double s = 0.0;
    for (int m = 0; m < (int)x.size(); m++)
    {
      if (m == k) continue;
      double v = 1.0 / (x[k] - x[m]);
      for (int i = 0; i < (int)x.size(); i++)
        if (i != k && i != m) v *= (t - x[i]) / (x[k] - x[i]);
      s += v;
    }
    return s;
  }
  // n-point Gauss-Legendre rule on [-1, 1]: Newton iteration on P_n
  static void gauss_legendre(int n, std::vector<double> &x, std::vector<double> &wt)
  {
    x.assign(n, 0.0);
    wt.assign(n, 0.0);
    for (int i = 0; i < n; i++)
    {
      double t = -std::cos(M_PI * (i + 0.75) / (n + 0.5)), dp = 1.0;
      for (int it = 0; it < 100; it++)
      {
        double p0 = 1.0, p1 = t;
        for (int k = 2; k <= n; k++)
        {
          const double pk = ((2.0 * k - 1.0) * t * p1 - (k - 1.0) * p0) / k;
          p0 = p1;
          p1 = pk;
        }
        if (n == 1) { p0 = 1.0; p1 = t; }
        dp = n * (t * p1 - p0) / (t * t - 1.0);
        const double dt = p1 / dp;
        t -= dt;
        if (std::fabs(dt) < 1e-16) break;
      }
      x[i] = t;
      double p0 = 1.0, p1 = t;
      for (int k = 2; k <= n; k++)
      {
        const double pk = ((2.0 * k - 1.0) * t * p1 - (k - 1.0) * p0) / k;
        p0 = p1;
        p1 = pk;
      }
      dp = n * (t * p1 - p0) / (t * t - 1.0);
      wt[i] = 2.0 / ((1.0 - t * t) * dp * dp);
    }
    for (int i = 0; i < n / 2; i++)  // exactly symmetric
    {
      const double a = 0.5 * (x[n - 1 - i] - x[i]), b = 0.5 * (wt[i] + wt[n - 1 - i]);
      x[i] = -a; x[n - 1 - i] = a; wt[i] = wt[n - 1 - i] = b;
    }
    if (n % 2) x[n / 2] = 0.0;
  }
  static std::vector<double> transpose(const std::vector<double> &A, int M)
  {
    std::vector<double> T(A.size());
    for (int k = 0; k < M; k++)
      for (int j = 0; j < M; j++) T[j * M + k] = A[k * M + j];
    return T;
  }
  // out = (A_{dim-1} x .. x A_0) in, axis d contracted with ops[d] (A[k*M+j]: in k -> out j)
  template <unsigned int dim>
  void kron_apply(const double *const ops[], const double *in, double *out) const
  {
    const int M = m_uiNrp;
    int total = 1;
    for (unsigned d = 0; d < dim; d++) total *= M;
    if ((int)im_vec1.size() < total) { im_vec1.assign(total, 0.0); im_vec2.assign(total, 0.0); }
    const double *src = in;
    for (unsigned d = 0; d < dim; d++)
    {
      double *dst = (d == dim - 1) ? out : ((d & 1) ? im_vec2.data() : im_vec1.data());
      if (dst == src) dst = (dst == im_vec1.data()) ? im_vec2.data() : im_vec1.data();
      int stride = 1;
      for (unsigned e = 0; e < d; e++) stride *= M;
      for (int base = 0; base < total; base++)
      {
        if ((base / stride) % M != 0) continue;
        for (int j = 0; j < M; j++)
        {
          double e = 0.0;
          for (int k = 0; k < M; k++) e += src[base + k * stride] * ops[d][k * M + j];
          dst[base + j * stride] = e;
        }
      }
      if (d == dim - 1 && dst != out) std::memcpy(out, dst, sizeof(double) * total);
      src = dst;
    }
  }

public:
  RefElement() : RefElement(3, 1) {}
  RefElement(unsigned int dim, unsigned int order) : m_uiOrder((int)order), m_uiDimension((int)dim), m_uiNrp((int)order + 1)
  {
    if (order < 1) throw std::invalid_argument("RefElement: order >= 1");
    const int M = m_uiNrp;
    r.resize(M);
    for (int j = 0; j < M; j++) r[j] = -1.0 + 2.0 * j / (double)order;
    gauss_legendre(M, g, w);
    ip_1D_0.resize(M * M); ip_1D_1.resize(M * M); quad_1D.resize(M * M); Dg.resize(M * M); Dr.resize(M * M);
    for (int k = 0; k < M; k++)
      for (int j = 0; j < M; j++)
      {
        ip_1D_0[k * M + j] = lagrange(r, k, 0.5 * (r[j] - 1.0));
        ip_1D_1[k * M + j] = lagrange(r, k, 0.5 * (r[j] + 1.0));
        quad_1D[k * M + j] = lagrange(r, k, g[j]);
        Dg[k * M + j] = dlagrange(r, k, g[j]);
        Dr[k * M + j] = dlagrange(r, k, r[j]);
      }
    ipT_1D_0 = transpose(ip_1D_0, M); ipT_1D_1 = transpose(ip_1D_1, M);
    quadT_1D = transpose(quad_1D, M); DgT = transpose(Dg, M); DrT = transpose(Dr, M);
    // nodal weights: integrals of the Lagrange basis (exact with the Gauss rule)
    wgll.assign(M, 0.0);
    for (int k = 0; k < M; k++)
      for (int j = 0; j < M; j++) wgll[k] += w[j] * quad_1D[k * M + j];
    int total = 1;
    for (unsigned d = 0; d < dim; d++) total *= M;
    im_vec1.assign(total, 0.0);
    im_vec2.assign(total, 0.0);
  }
  ~RefElement() {}

  inline int getOrder() const { return m_uiOrder; }
  inline int getDim() const { return m_uiDimension; }
  inline int get1DNumInterpolationPoints() { return m_uiNrp; }
  inline const double *getIMChild0() const { return ip_1D_0.data(); }
  inline const double *getIMChild1() const { return ip_1D_1.data(); }
  inline const double *getQ1d() const { return quad_1D.data(); }
  inline const double *getQT1d() const { return quadT_1D.data(); }
  inline const double *getDg1d() const { return Dg.data(); }
  inline const double *getDgT1d() const { return DgT.data(); }
  inline const double *getDr1d() const { return Dr.data(); }
  inline double *getImVec1() const { return im_vec1.data(); }
  inline double *getImVec2() const { return im_vec2.data(); }
  inline const double *getWgq() const { return w.data(); }
  inline const double *getWgll() const { return wgll.data(); }
  inline double getElementSz() const { return r.back() - r.front(); }

  /** parent nodal values -> values at the nodes of child `childNum` (Morton number: bit d selects the half along axis d);
   *  FEM/include/refel.h:214-240.  in == out is allowed. */
  template <unsigned int dim>
  inline void IKD_Parent2Child(const double *in, double *out, unsigned int childNum) const
  {
    assert(childNum < (1u << dim));
    const double *ops[dim];
    for (unsigned d = 0; d < dim; d++) ops[d] = ((childNum >> d) & 1u) ? ip_1D_1.data() : ip_1D_0.data();
    kron_apply<dim>(ops, in, out);
  }
  /** transpose of the above: child contributions -> parent nodes (FEM/include/refel.h:252-271) */
  template <unsigned int dim>
  inline void IKD_Child2Parent(const double *in, double *out, unsigned int childNum) const
  {
    assert(childNum < (1u << dim));
    const double *ops[dim];
    for (unsigned d = 0; d < dim; d++) ops[d] = ((childNum >> d) & 1u) ? ipT_1D_1.data() : ipT_1D_0.data();
    kron_apply<dim>(ops, in, out);
  }
};
#endif
