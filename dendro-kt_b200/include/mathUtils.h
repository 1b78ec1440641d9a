// The few helpers of the reference's include/mathUtils.h + include/parUtils.h that its FEM examples use (heatMat.cpp:165-325:
// normLInfty, dot, par::Mpi_Bcast, par::Mpi_Allreduce; intPow).  One process drives one GPU here: without MPI the
// collectives are the identity.
#ifndef DKT_HOST_MATHUTILS_H
#define DKT_HOST_MATHUTILS_H

#include <cmath>
#include <cstring>

#include "dendro.h"

template <typename T>
inline T intPow(T base, unsigned int e)
{
  T r = 1;
  for (unsigned int i = 0; i < e; i++) r *= base;
  return r;
}
template <typename T>
inline T normLInfty(const T *v, unsigned int n)
{
  T m = 0;
  for (unsigned int i = 0; i < n; i++) m = std::fabs(v[i]) > m ? std::fabs(v[i]) : m;
  return m;
}
template <typename T>
inline T normLInfty(const T *a, const T *b, unsigned int n)
{
  T m = 0;
  for (unsigned int i = 0; i < n; i++) m = std::fabs(a[i] - b[i]) > m ? std::fabs(a[i] - b[i]) : m;
  return m;
}
template <typename T>
inline T dot(const T *a, const T *b, unsigned int n)
{
  T s = 0;
  for (unsigned int i = 0; i < n; i++) s += a[i] * b[i];
  return s;
}

namespace par
{
#if defined(MPI_VERSION) || defined(DKT_HAVE_MPI)
template <typename T> inline int Mpi_Bcast(T *buf, int count, int root, MPI_Comm comm) { return MPI_Bcast(buf, count * (int)sizeof(T), MPI_BYTE, root, comm); }
inline int Mpi_Allreduce(const double *s, double *r, int count, MPI_Op op, MPI_Comm comm) { return MPI_Allreduce(s, r, count, MPI_DOUBLE, op, comm); }
#else
template <typename T> inline int Mpi_Bcast(T *, int, int, MPI_Comm) { return 0; }
template <typename T, typename Op> inline int Mpi_Allreduce(const T *s, T *r, int count, Op, MPI_Comm) { if (s != r) std::memcpy(r, s, sizeof(T) * count); return 0; }
#endif
} // namespace par

#if defined(MPI_VERSION) || defined(DKT_HAVE_MPI)
template <typename T> inline T normLInfty(const T *v, unsigned int n, MPI_Comm comm) { double l = (double)normLInfty(v, n), g = l; MPI_Allreduce(&l, &g, 1, MPI_DOUBLE, MPI_MAX, comm); return (T)g; }
template <typename T> inline T dot(const T *a, const T *b, unsigned int n, MPI_Comm comm) { double l = (double)dot(a, b, n), g = l; MPI_Allreduce(&l, &g, 1, MPI_DOUBLE, MPI_SUM, comm); return (T)g; }
#else
template <typename T> inline T normLInfty(const T *v, unsigned int n, MPI_Comm) { return normLInfty(v, n); }
template <typename T> inline T dot(const T *a, const T *b, unsigned int n, MPI_Comm) { return dot(a, b, n); }
#endif
#endif
