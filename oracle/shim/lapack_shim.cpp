/* Minimal dgesv_/dsyev_ for building the reference oracle (oracle/_ref).
 * TEST INFRASTRUCTURE - not part of the product path.
 *
 * The reference calls LAPACK only to build RefElement's 1-D operators
 * (FEM/src/refel.cpp:158-165 via FEM/include/lapac.h:22-23; FEM/src/basis.cpp:147).
 * LAPACK is an un-vendored, un-pinned dependency of the reference and is absent from this
 * image, so the two entry points are restated here: LU with partial pivoting, and a
 * cyclic Jacobi symmetric eigen-solver (ascending eigenvalues, like dsyev).  RefElement's
 * matrices are read back from the live object by both sides of every parity test, so
 * last-ulp differences against a vendor LAPACK cancel ("parity unpinned" at this boundary).
 */
#include <cmath>
#include <vector>
#include <algorithm>
#include <numeric>

extern "C" void dgesv_(int *n_, int *nrhs_, double *a, int *lda_, int *ipiv, double *b, int *ldb_, int *info)
{
  const int n = *n_, nrhs = *nrhs_, lda = *lda_, ldb = *ldb_;
  *info = 0;
  for (int k = 0; k < n; k++)
  {
    int p = k;
    double best = std::fabs(a[k + (long)k * lda]);
    for (int i = k + 1; i < n; i++)
      if (std::fabs(a[i + (long)k * lda]) > best) { best = std::fabs(a[i + (long)k * lda]); p = i; }
    ipiv[k] = p + 1;
    if (best == 0.0) { if (!*info) *info = k + 1; continue; }
    if (p != k)
    {
      for (int j = 0; j < n; j++) std::swap(a[k + (long)j * lda], a[p + (long)j * lda]);
      for (int j = 0; j < nrhs; j++) std::swap(b[k + (long)j * ldb], b[p + (long)j * ldb]);
    }
    const double piv = a[k + (long)k * lda];
    for (int i = k + 1; i < n; i++)
    {
      const double l = a[i + (long)k * lda] / piv;
      a[i + (long)k * lda] = l;
      for (int j = k + 1; j < n; j++) a[i + (long)j * lda] -= l * a[k + (long)j * lda];
      for (int j = 0; j < nrhs; j++) b[i + (long)j * ldb] -= l * b[k + (long)j * ldb];
    }
  }
  if (*info) return;
  for (int j = 0; j < nrhs; j++)
    for (int i = n - 1; i >= 0; i--)
    {
      double s = b[i + (long)j * ldb];
      for (int k = i + 1; k < n; k++) s -= a[i + (long)k * lda] * b[k + (long)j * ldb];
      b[i + (long)j * ldb] = s / a[i + (long)i * lda];
    }
}

extern "C" void dsyev_(char *, char *uplo, int *n_, double *a, int *lda_, double *w, double *work, int *lwork, int *info)
{
  const int n = *n_, lda = *lda_;
  *info = 0;
  if (*lwork == -1) { work[0] = (double)std::max(1, 3 * n); return; }
  const bool upper = (*uplo == 'U' || *uplo == 'u');
  std::vector<double> A((size_t)n * n), V((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++)
    {
      const int r = upper ? std::min(i, j) : std::max(i, j), c = upper ? std::max(i, j) : std::min(i, j);
      A[i + (size_t)j * n] = a[r + (long)c * lda];
    }
  for (int i = 0; i < n; i++) V[i + (size_t)i * n] = 1.0;
  for (int sweep = 0; sweep < 100; sweep++)
  {
    double off = 0.0;
    for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) off += A[p + (size_t)q * n] * A[p + (size_t)q * n];
    if (off == 0.0) break;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++)
      {
        const double apq = A[p + (size_t)q * n];
        if (apq == 0.0) continue;
        const double theta = (A[q + (size_t)q * n] - A[p + (size_t)p * n]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; k++)
        {
          const double akp = A[k + (size_t)p * n], akq = A[k + (size_t)q * n];
          A[k + (size_t)p * n] = c * akp - s * akq;
          A[k + (size_t)q * n] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++)
        {
          const double apk = A[p + (size_t)k * n], aqk = A[q + (size_t)k * n];
          A[p + (size_t)k * n] = c * apk - s * aqk;
          A[q + (size_t)k * n] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++)
        {
          const double vkp = V[k + (size_t)p * n], vkq = V[k + (size_t)q * n];
          V[k + (size_t)p * n] = c * vkp - s * vkq;
          V[k + (size_t)q * n] = s * vkp + c * vkq;
        }
      }
  }
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int x, int y) { return A[x + (size_t)x * n] < A[y + (size_t)y * n]; });
  for (int j = 0; j < n; j++)
  {
    w[j] = A[order[j] + (size_t)order[j] * n];
    for (int i = 0; i < n; i++) a[i + (long)j * lda] = V[i + (size_t)order[j] * n];
  }
}
