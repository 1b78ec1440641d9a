// Tensor-product contractions of an M x M 1-D operator with an elemental vector, the kernels behind the reference's
// sum-factorised operators (FEM/include/tensor.h, FEM/src/tensor.cpp:19-107).  Same names and argument meaning:
// A[k * M + j] takes input index k to output index j; axis 0 (x) is the fastest index of X.
#ifndef DKT_HOST_TENSOR_H
#define DKT_HOST_TENSOR_H

namespace dkt_host
{
// Y = X contracted with A along `axis` of an M^ndim array
inline void apply_axis(int M, int ndim, int axis, const double *A, const double *X, double *Y)
{
  int stride = 1, total = 1;
  for (int d = 0; d < axis; d++) stride *= M;
  for (int d = 0; d < ndim; d++) total *= M;
  for (int base = 0; base < total; base++)
  {
    if ((base / stride) % M != 0) continue;  // first entry of a line along the axis
    for (int j = 0; j < M; j++)
    {
      double e = 0.0;
      for (int k = 0; k < M; k++) e += X[base + k * stride] * A[k * M + j];
      Y[base + j * stride] = e;
    }
  }
}
} // namespace dkt_host

// 3-D
inline void DENDRO_TENSOR_IIAX_APPLY_ELEM(const int M, const double *A, const double *X, double *Y) { dkt_host::apply_axis(M, 3, 0, A, X, Y); }
inline void DENDRO_TENSOR_IAIX_APPLY_ELEM(const int M, const double *A, const double *X, double *Y) { dkt_host::apply_axis(M, 3, 1, A, X, Y); }
inline void DENDRO_TENSOR_AIIX_APPLY_ELEM(const int M, const double *A, const double *X, double *Y) { dkt_host::apply_axis(M, 3, 2, A, X, Y); }
// 2-D
inline void DENDRO_TENSOR_IAX_APPLY_ELEM_2D(const int M, const double *A, const double *X, double *Y) { dkt_host::apply_axis(M, 2, 0, A, X, Y); }
inline void DENDRO_TENSOR_AIX_APPLY_ELEM_2D(const int M, const double *A, const double *X, double *Y) { dkt_host::apply_axis(M, 2, 1, A, X, Y); }
// 4-D
inline void DENDRO_TENSOR_IIIAX_APPLY_ELEM(const int M, const double *A, const double *X, double *Y) { dkt_host::apply_axis(M, 4, 0, A, X, Y); }
inline void DENDRO_TENSOR_IIAIX_APPLY_ELEM(const int M, const double *A, const double *X, double *Y) { dkt_host::apply_axis(M, 4, 1, A, X, Y); }
inline void DENDRO_TENSOR_IAIIX_APPLY_ELEM(const int M, const double *A, const double *X, double *Y) { dkt_host::apply_axis(M, 4, 2, A, X, Y); }
inline void DENDRO_TENSOR_AIIIX_APPLY_ELEM(const int M, const double *A, const double *X, double *Y) { dkt_host::apply_axis(M, 4, 3, A, X, Y); }
#endif
