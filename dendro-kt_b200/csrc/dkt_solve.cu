// Device-resident conjugate gradients on top of the matvec: the hand-written CG of the reference's
// example operator, HeatMat::cgSolve (FEM/examples/src/heatMat.cpp:165-325), with all vectors kept
// in HBM between iterations (the reference's caller of the hot path).
// Same recurrences and the same stopping rule (max-norm of the residual relative to max-norm of b).
#include "dkt_internal.h"

// also compiles under -DDKT_EMU (tests/emu/cuda_emu.h) for the CPU test-suite; never part of libdkt.so that way
#ifdef DKT_EMU
#include "cuda_emu.h"
#define DKT_SOLVE_LAUNCH(kern, grid, stream) ::emu::make_launch(kern, (grid), 256, 0)
#else
#define DKT_SOLVE_LAUNCH(kern, grid, stream) kern<<<(grid), 256, 0, (stream)>>>
#endif

#include <cmath>
#include <cstring>

namespace dkt
{
#define CK(call)                                                                                     \
  do                                                                                                 \
  {                                                                                                  \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess)                                                                           \
    {                                                                                                \
      set_error(std::string(#call) + ": " + cudaGetErrorString(e_) + " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
      return DKT_ERR_CUDA;                                                                           \
    }                                                                                                \
  } while (0)

__device__ __forceinline__ double block_sum(double v)
{
  __shared__ double s[32];
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  v = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.0;
  if (threadIdx.x < 32)
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  return v;
}
__device__ __forceinline__ double block_max(double v)
{
  __shared__ double s[32];
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  v = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.0;
  if (threadIdx.x < 32)
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  return v;
}
__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v)
{  // non-negative doubles order like their bit patterns
  atomicMax((unsigned long long *)addr, (unsigned long long)__double_as_longlong(v));
}

// red[0] += a.b ; red[1] += c.d ; red[2] = max(red[2], |e|_inf)   (null pointers are skipped)
__global__ void k_reduce3(const double *a, const double *b, const double *c, const double *d, const double *e, uint64_t n,
                          double *red)
{
  double s0 = 0.0, s1 = 0.0, m = 0.0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
  {
    if (a) s0 += a[i] * b[i];
    if (c) s1 += c[i] * d[i];
    if (e) m = fmax(m, fabs(e[i]));
  }
  s0 = block_sum(s0);
  s1 = block_sum(s1);
  m = block_max(m);
  if (threadIdx.x == 0)
  {
    if (a) atomicAdd(red + 0, s0);
    if (c) atomicAdd(red + 1, s1);
    if (e) atomic_max_nonneg(red + 2, m);
  }
}
// r0 = b - Ax ; p = r0
__global__ void k_cg_init(const double *b, const double *Ax, uint64_t n, double *r0, double *p)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n)
  {
    const double r = b[i] - Ax[i];
    r0[i] = r;
    p[i] = r;
  }
}
// x += alpha p ; r1 = r0 - alpha Ap ; red[0] += r1.r1 ; red[2] = |r1|_inf
__global__ void k_cg_step1(double alpha, const double *p, const double *Ap, const double *r0, uint64_t n, double *x, double *r1,
                           double *red)
{
  double s = 0.0, m = 0.0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
  {
    x[i] += alpha * p[i];
    const double r = r0[i] - alpha * Ap[i];
    r1[i] = r;
    s += r * r;
    m = fmax(m, fabs(r));
  }
  s = block_sum(s);
  m = block_max(m);
  if (threadIdx.x == 0)
  {
    atomicAdd(red + 0, s);
    atomic_max_nonneg(red + 2, m);
  }
}
// p = r1 + beta p ; r0 = r1
__global__ void k_cg_step2(double beta, const double *r1, uint64_t n, double *p, double *r0)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n)
  {
    const double r = r1[i];
    p[i] = r + beta * p[i];
    r0[i] = r;
  }
}

int cg_solve(DA &da, Dist *dist, const dkt_op *op, double *d_x, const double *d_b, int max_iter, double *tol, double scale,
             unsigned flags, int *iters, int *status)
{
  const bool part = dist && dist->active;
  const uint64_t n = part ? dist->nOwned : da.nNodes;
  cudaStream_t s = da.stream;
  double *work = nullptr, *red = nullptr;
  // the chunk tables carry quirk Q1 statically: the conservative variant runs on the flat kernels (as in dkt_matvec)
  if (flags & DKT_NO_Q1_MASK) flags |= DKT_MV_FLAT;
  // CUDA failures below leave through `done` (buffers released)
#define CKD(call)                                                                                                      \
  do                                                                                                                   \
  {                                                                                                                    \
    cudaError_t e_ = (call);                                                                                           \
    if (e_ != cudaSuccess)                                                                                             \
    {                                                                                                                  \
      set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                                                   \
      rc = DKT_ERR_CUDA;                                                                                               \
      goto done;                                                                                                       \
    }                                                                                                                  \
  } while (0)
  int rc = DKT_OK;
  CK(cudaMalloc((void **)&work, std::max<uint64_t>(n, 1) * 4 * sizeof(double)));
  if (cudaMalloc((void **)&red, 4 * sizeof(double)) != cudaSuccess) { cudaFree(work); set_error("cg_solve: out of device memory"); return DKT_ERR_CUDA; }
  double *p = work, *Ap = work + n, *r0 = work + 2 * n, *r1 = work + 3 * n;
  const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 8), gridN = (unsigned)((n + 255) / 256);
  double h[4];
  auto mv = [&](const double *in, double *out) {
    return part ? run_matvec_dist(da, *dist, op, in, out, scale, flags)
                : ((flags & DKT_MV_FLAT) ? run_matvec(da, op, in, out, scale, flags) : run_matvec_chunked(da, op, in, out, scale, flags));
  };
  auto reduce = [&](int &rc) {  // red -> h (summed / maxed over ranks)
    rc = DKT_OK;
    if (part && dist->nranks > 1) rc = dist_allreduce(*dist, red, s);
    if (rc == DKT_OK && cudaMemcpyAsync(h, red, sizeof(h), cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = DKT_ERR_CUDA;
    if (rc == DKT_OK && cudaStreamSynchronize(s) != cudaSuccess) rc = DKT_ERR_CUDA;
  };
  *status = 1;
  *iters = 0;
  // normb, r0 = b - A x, p = r0
  CKD(cudaMemsetAsync(red, 0, 4 * sizeof(double), s));
  if (n) DKT_SOLVE_LAUNCH(k_reduce3, grid, s)(nullptr, nullptr, nullptr, nullptr, d_b, n, red);
  rc = mv(d_x, Ap);
  if (rc) goto done;
  if (n) DKT_SOLVE_LAUNCH(k_cg_init, gridN, s)(d_b, Ap, n, r0, p);
  g_launches += 2;
  reduce(rc);
  if (rc) goto done;
  {
    double normb = h[2];
    if (normb == 0.0) normb = 1.0;
    CKD(cudaMemsetAsync(red, 0, 4 * sizeof(double), s));
    if (n) DKT_SOLVE_LAUNCH(k_reduce3, grid, s)(r0, r0, nullptr, nullptr, r0, n, red);  // red[0] = r0.r0, red[2] = |r0|
    g_launches++;
    reduce(rc);
    if (rc) goto done;
    double rr = h[0], resid = h[2] / normb;
    if (resid <= *tol) { *tol = resid; *status = 0; goto done; }
    for (int i = 1; i <= max_iter; i++)
    {
      rc = mv(p, Ap);
      if (rc) goto done;
      CKD(cudaMemsetAsync(red, 0, 4 * sizeof(double), s));
      if (n) DKT_SOLVE_LAUNCH(k_reduce3, grid, s)(nullptr, nullptr, p, Ap, nullptr, n, red);  // red[1] = p.Ap
      g_launches++;
      reduce(rc);
      if (rc) goto done;
      const double alpha = rr / h[1];
      CKD(cudaMemsetAsync(red, 0, 4 * sizeof(double), s));
      if (n) DKT_SOLVE_LAUNCH(k_cg_step1, grid, s)(alpha, p, Ap, r0, n, d_x, r1, red);
      g_launches++;
      reduce(rc);
      if (rc) goto done;
      *iters = i;
      resid = h[2] / normb;
      if (resid <= *tol) { *status = 0; break; }
      const double beta = h[0] / rr;
      rr = h[0];
      if (n) DKT_SOLVE_LAUNCH(k_cg_step2, gridN, s)(beta, r1, n, p, r0);
      g_launches++;
    }
    *tol = resid;
  }
done:
#undef CKD
  cudaStreamSynchronize(s);
  cudaFree(work);
  cudaFree(red);
  if (rc == DKT_OK && cudaGetLastError() != cudaSuccess) { set_error("CUDA error in cg_solve"); rc = DKT_ERR_CUDA; }
  return rc;
}
} // namespace dkt
