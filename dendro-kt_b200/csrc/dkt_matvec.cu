// Matvec kernels (sm_100a): flat gather -> elemental operator -> scatter-add.
//
// Semantics = FEM/include/matvec.h:378-522 of the reference, executed from the tables built in
// dkt_build.cu instead of a recursive re-bucketing of all nodes at every level:
//   1. ein[r] = u[e2n[r]] for filled ranks                              (matvec.h:383-395)
//   2. hanging element: pin[q] = u[pnode[q]] (level L-1 nodes only, :417), ein[r] for unfilled
//      r = ((x)_d A_{c_d}) pin                                          (:450-456, refel.h:214)
//   3. eout = K_e ein                                                   (:468)
//   4. v[e2n[r]] += eout[r] for filled r                                (:471-480 + bottom_up)
//   5. hanging: w = eout with filled entries zeroed, t = ((x)_d A_{c_d}^T) w, and
//      v[pnode[q]] += t[q] for valid q with !filled[q]  (quirk Q1: the LEAF's fill flag indexed
//      by the PARENT's rank, :517)
// Regular elements [0,nReg) and hanging elements [nReg,nMv) are separate launches so neither
// diverges.  The operator matrix, the level scales and the 1-D interpolation matrices travel
// in the kernel parameter block (constant bank): every FMA of the dense product reads its
// matrix entry as an immediate constant operand.
#include "dkt_internal.h"

// also compiles under -DDKT_EMU (tests/emu/cuda_emu.h) for the CPU test-suite; never part of libdkt.so that way
#ifdef DKT_EMU
#include "cuda_emu.h"
#define DKT_FLAT_LAUNCH(kern, grid, stream) ::emu::make_launch_flat(kern, (grid), 128)
#else
#define DKT_FLAT_LAUNCH(kern, grid, stream) kern<<<(grid), 128, 0, (stream)>>>
#endif

#include <cstdio>
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

template <int DIM, int ORDER>
struct MvParams
{
  static constexpr int M = ORDER + 1;
  static constexpr int N = (DIM == 2 ? M * M : DIM == 3 ? M * M * M : M * M * M * M);
  const double *in;
  double *out;
  const uint32_t *e2n;
  const uint8_t *lev;
  const uint32_t *pnode;
  const uint8_t *child;
  const uint8_t *isbdy;
  uint32_t first, count;
  int q1mask;
  double lscale[32];
  double ip[2][M * M];
  double K[N * N];
};

__device__ __forceinline__ void red_add(double *p, double v) { atomicAdd(p, v); }

// one axis pass of the tensor-product interpolation, in registers.
// forward:   out[j] = sum_k A[k*M+j] in[k]   (parent -> child, FEM/src/tensor.cpp:41-59)
// transpose: out[j] = sum_k A[j*M+k] in[k]   (child -> parent, refel.cpp ipT = ip^T)
template <int DIM, int M, int AXIS, bool TRANSPOSE>
__device__ __forceinline__ void axis_pass(const double *A, double *v)
{
  constexpr int N = (DIM == 2 ? M * M : DIM == 3 ? M * M * M : M * M * M * M);
  int stride = 1;
#pragma unroll
  for (int d = 0; d < AXIS; d++) stride *= M;
#pragma unroll
  for (int base = 0; base < N; base++)
  {
    if ((base / stride) % M != 0) continue;  // first entry of each line along AXIS
    double line[M], res[M];
#pragma unroll
    for (int k = 0; k < M; k++) line[k] = v[base + k * stride];
#pragma unroll
    for (int j = 0; j < M; j++)
    {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < M; k++) acc = fma(TRANSPOSE ? A[j * M + k] : A[k * M + j], line[k], acc);
      res[j] = acc;
    }
#pragma unroll
    for (int j = 0; j < M; j++) v[base + j * stride] = res[j];
  }
}

template <int DIM, int M, bool TRANSPOSE>
__device__ __forceinline__ void tensor_interp(const double (&ip)[2][M * M], int child, double *v)
{
  // the reference applies the highest axis first going forward (tensor.h:181-203); the order
  // only changes rounding.
  double A[M * M];
#pragma unroll
  for (int i = 0; i < M * M; i++) A[i] = (child & 1) ? ip[1][i] : ip[0][i];
  axis_pass<DIM, M, 0, TRANSPOSE>(A, v);
#pragma unroll
  for (int i = 0; i < M * M; i++) A[i] = (child & 2) ? ip[1][i] : ip[0][i];
  axis_pass<DIM, M, 1, TRANSPOSE>(A, v);
  if (DIM >= 3)
  {
#pragma unroll
    for (int i = 0; i < M * M; i++) A[i] = (child & 4) ? ip[1][i] : ip[0][i];
    axis_pass<DIM, M, (DIM >= 3 ? 2 : 0), TRANSPOSE>(A, v);
  }
  if (DIM >= 4)
  {
#pragma unroll
    for (int i = 0; i < M * M; i++) A[i] = (child & 8) ? ip[1][i] : ip[0][i];
    axis_pass<DIM, M, (DIM >= 4 ? 3 : 0), TRANSPOSE>(A, v);
  }
}

template <int DIM, int ORDER, int OPKIND>
__device__ __forceinline__ void apply_op(const MvParams<DIM, ORDER> &p, int lev, const double *ein, double *eout)
{
  constexpr int N = MvParams<DIM, ORDER>::N;
  if (OPKIND == DKT_OP_IDENTITY)
  {
#pragma unroll
    for (int i = 0; i < N; i++) eout[i] = ein[i];
  }
  else
  {
    const double s = p.lscale[lev];
#pragma unroll
    for (int i = 0; i < N; i++)
    {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < N; j++) acc = fma(p.K[i * N + j], ein[j], acc);
      eout[i] = s * acc;
    }
  }
}

template <int DIM, int ORDER, int OPKIND, bool DIRI>
__global__ void __launch_bounds__(128) k_mv_regular(const __grid_constant__ MvParams<DIM, ORDER> p)
{
  constexpr int N = MvParams<DIM, ORDER>::N;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.count) return;
  const uint64_t e = (uint64_t)p.first + t;
  uint32_t idx[N];
  double ein[N], eout[N];
#pragma unroll
  for (int r = 0; r < N; r++) idx[r] = p.e2n[e * N + r];
#pragma unroll
  for (int r = 0; r < N; r++)
  {
    double v = p.in[idx[r]];
    if (DIRI && p.isbdy[idx[r]]) v = 0.0;
    ein[r] = v;
  }
  apply_op<DIM, ORDER, OPKIND>(p, p.lev[e], ein, eout);
#pragma unroll
  for (int r = 0; r < N; r++)
  {
    if (DIRI && p.isbdy[idx[r]]) continue;
    red_add(p.out + idx[r], eout[r]);
  }
}

template <int DIM, int ORDER, int OPKIND, bool DIRI>
__global__ void __launch_bounds__(128) k_mv_hanging(const __grid_constant__ MvParams<DIM, ORDER> p)
{
  constexpr int N = MvParams<DIM, ORDER>::N;
  constexpr int M = ORDER + 1;
  const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= p.count) return;
  const uint64_t e = (uint64_t)p.first + h;
  const int child = p.child[h];
  uint32_t idx[N], pn[N];
  double ein[N], eout[N], par[N];
#pragma unroll
  for (int r = 0; r < N; r++)
  {
    idx[r] = p.e2n[e * N + r];
    pn[r] = p.pnode[(uint64_t)h * N + r];
  }
#pragma unroll
  for (int r = 0; r < N; r++)
  {
    double v = 0.0;
    if (pn[r] != INVALID)
    {
      v = p.in[pn[r]];
      if (DIRI && p.isbdy[pn[r]]) v = 0.0;
    }
    par[r] = v;
  }
  tensor_interp<DIM, M, false>(p.ip, child, par);
#pragma unroll
  for (int r = 0; r < N; r++)
  {
    double v = par[r];
    if (idx[r] != INVALID)
    {
      v = p.in[idx[r]];
      if (DIRI && p.isbdy[idx[r]]) v = 0.0;
    }
    ein[r] = v;
  }
  apply_op<DIM, ORDER, OPKIND>(p, p.lev[e], ein, eout);
#pragma unroll
  for (int r = 0; r < N; r++)
  {
    if (idx[r] != INVALID)
    {
      if (!(DIRI && p.isbdy[idx[r]])) red_add(p.out + idx[r], eout[r]);
      eout[r] = 0.0;  // nullify prior to back-interpolation (matvec.h:497-499)
    }
  }
  tensor_interp<DIM, M, true>(p.ip, child, eout);
#pragma unroll
  for (int q = 0; q < N; q++)
  {
    if (pn[q] == INVALID) continue;
    if (p.q1mask && idx[q] != INVALID) continue;  // Q1 (matvec.h:517)
    if (DIRI && p.isbdy[pn[q]]) continue;
    red_add(p.out + pn[q], eout[q]);
  }
}

template <int DIM, int ORDER, int OPKIND, bool DIRI>
static int launch_mv(DA &da, const MvParams<DIM, ORDER> &base)
{
  MvParams<DIM, ORDER> p = base;
  const uint32_t chunk = 1u << 30;
  for (uint64_t first = 0; first < da.nReg; first += chunk)
  {
    p.first = (uint32_t)first;
    p.count = (uint32_t)std::min<uint64_t>(chunk, da.nReg - first);
    {
      auto kern = k_mv_regular<DIM, ORDER, OPKIND, DIRI>;
      DKT_FLAT_LAUNCH(kern, (p.count + 127) / 128, da.stream)(p);
    }
    g_launches++;
  }
  if (da.nHang)
  {
    p.first = (uint32_t)da.nReg;
    p.count = (uint32_t)da.nHang;
    {
      auto kern = k_mv_hanging<DIM, ORDER, OPKIND, DIRI>;
      DKT_FLAT_LAUNCH(kern, (p.count + 127) / 128, da.stream)(p);
    }
    g_launches++;
  }
  CK(cudaGetLastError());
  return DKT_OK;
}

// ------------------------------------------------------------------------------------------
// 81 nodes per element (4-D, order 2): the same algorithm with run-time loops over arrays in local memory and the
// operator matrix read from device memory (it does not fit the kernel parameter space).  Coverage, not speed: this
// combination has no shared-memory kernel yet.
// ------------------------------------------------------------------------------------------
struct MvBigParams
{
  const double *in;
  double *out;
  const uint32_t *e2n;
  const uint8_t *lev;
  const uint32_t *pnode;
  const uint8_t *child;
  const uint8_t *isbdy;
  const double *K;  // device, N*N row-major, or nullptr for the identity
  uint32_t first, count;
  int q1mask, dirichlet;
  double lscale[32];
  double ip[2][MAX_M * MAX_M];
};
template <int DIM, int M>
__device__ void interp_big(const double (&ip)[2][MAX_M * MAX_M], int child, bool transpose, double *v)
{
  constexpr int N = (DIM == 2 ? M * M : DIM == 3 ? M * M * M : M * M * M * M);
  int stride = 1;
  for (int d = 0; d < DIM; d++)
  {
    const double *A = ip[(child >> d) & 1];
    for (int base = 0; base < N; base++)
    {
      if ((base / stride) % M != 0) continue;
      double line[M], res[M];
      for (int k = 0; k < M; k++) line[k] = v[base + k * stride];
      for (int j = 0; j < M; j++)
      {
        double acc = 0.0;
        for (int k = 0; k < M; k++) acc = fma(transpose ? A[j * M + k] : A[k * M + j], line[k], acc);
        res[j] = acc;
      }
      for (int j = 0; j < M; j++) v[base + j * stride] = res[j];
    }
    stride *= M;
  }
}
template <int DIM, int ORDER, bool HANG>
__global__ void __launch_bounds__(128) k_mv_big(const __grid_constant__ MvBigParams p)
{
  constexpr int M = ORDER + 1;
  constexpr int N = (DIM == 2 ? M * M : DIM == 3 ? M * M * M : M * M * M * M);
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.count) return;
  const uint64_t e = (uint64_t)p.first + t;
  const uint32_t *idx = p.e2n + e * N;
  const uint32_t *pn = HANG ? p.pnode + (uint64_t)t * N : nullptr;
  const int child = HANG ? p.child[t] : 0;
  const bool diri = p.dirichlet != 0;
  double ein[N], eout[N];
  if (HANG)
  {
    for (int r = 0; r < N; r++)
    {
      double v = 0.0;
      if (pn[r] != INVALID)
      {
        v = p.in[pn[r]];
        if (diri && p.isbdy[pn[r]]) v = 0.0;
      }
      ein[r] = v;
    }
    interp_big<DIM, M>(p.ip, child, false, ein);
  }
  for (int r = 0; r < N; r++)
  {
    if (idx[r] == INVALID) continue;  // hanging: keep the interpolated value
    double v = p.in[idx[r]];
    if (diri && p.isbdy[idx[r]]) v = 0.0;
    ein[r] = v;
  }
  if (!p.K)
    for (int i = 0; i < N; i++) eout[i] = ein[i];
  else
  {
    const double s = p.lscale[p.lev[e]];
    for (int i = 0; i < N; i++)
    {
      double acc = 0.0;
      for (int j = 0; j < N; j++) acc = fma(p.K[i * N + j], ein[j], acc);
      eout[i] = s * acc;
    }
  }
  for (int r = 0; r < N; r++)
  {
    if (idx[r] == INVALID) continue;
    if (!(diri && p.isbdy[idx[r]])) red_add(p.out + idx[r], eout[r]);
    eout[r] = 0.0;  // nullify prior to back-interpolation (matvec.h:497-499)
  }
  if (HANG)
  {
    interp_big<DIM, M>(p.ip, child, true, eout);
    for (int q = 0; q < N; q++)
    {
      if (pn[q] == INVALID) continue;
      if (p.q1mask && idx[q] != INVALID) continue;  // Q1 (matvec.h:517)
      if (diri && p.isbdy[pn[q]]) continue;
      red_add(p.out + pn[q], eout[q]);
    }
  }
}
template <int DIM, int ORDER>
static int run_big(DA &da, const dkt_op *op, const double *d_in, double *d_out, double scale, unsigned flags)
{
  constexpr int M = ORDER + 1;
  constexpr int N = (DIM == 2 ? M * M : DIM == 3 ? M * M * M : M * M * M * M);
  MvBigParams p;
  p.in = d_in; p.out = d_out; p.e2n = da.d_e2n; p.lev = da.d_mv_lev; p.pnode = da.d_pnode; p.child = da.d_child; p.isbdy = da.d_node_isbdy;
  p.K = nullptr;
  p.q1mask = (flags & DKT_NO_Q1_MASK) ? 0 : 1;
  p.dirichlet = op->dirichlet != 0;
  for (int l = 0; l < 32; l++) p.lscale[l] = scale * std::pow(2.0, -op->alpha * l);
  for (int b = 0; b < 2; b++)
    for (int i = 0; i < M * M; i++) p.ip[b][i] = da.ip[b][i];
  if (op->kind == DKT_OP_DENSE)
  {
    if (!op->kref) { set_error("DKT_OP_DENSE needs kref"); return DKT_ERR_INVALID; }
    if (!da.d_kbuf) CK(cudaMalloc((void **)&da.d_kbuf, sizeof(double) * N * N));
    CK(cudaMemcpyAsync(da.d_kbuf, op->kref, sizeof(double) * N * N, cudaMemcpyHostToDevice, da.stream));
    CK(cudaStreamSynchronize(da.stream));  // kref is the caller's pageable memory
    p.K = da.d_kbuf;
  }
  else if (op->kind != DKT_OP_IDENTITY) { set_error("unknown operator kind"); return DKT_ERR_INVALID; }
  CK(cudaMemsetAsync(d_out, 0, da.nNodes * sizeof(double), da.stream));
  g_launches++;
  if (da.nReg)
  {
    p.first = 0;
    p.count = (uint32_t)da.nReg;
    auto kern = k_mv_big<DIM, ORDER, false>;
    DKT_FLAT_LAUNCH(kern, (p.count + 127) / 128, da.stream)(p);
    g_launches++;
  }
  if (da.nHang)
  {
    p.first = (uint32_t)da.nReg;
    p.count = (uint32_t)da.nHang;
    auto kern = k_mv_big<DIM, ORDER, true>;
    DKT_FLAT_LAUNCH(kern, (p.count + 127) / 128, da.stream)(p);
    g_launches++;
  }
  CK(cudaGetLastError());
  return DKT_OK;
}

template <int DIM, int ORDER>
static int run_typed(DA &da, const dkt_op *op, const double *d_in, double *d_out, double scale, unsigned flags)
{
  using P = MvParams<DIM, ORDER>;
  static thread_local P p;  // large (K up to 27x27); filled per call, copied into the launch; per host thread
  p.in = d_in;
  p.out = d_out;
  p.e2n = da.d_e2n;
  p.lev = da.d_mv_lev;
  p.pnode = da.d_pnode;
  p.child = da.d_child;
  p.isbdy = da.d_node_isbdy;
  p.q1mask = (flags & DKT_NO_Q1_MASK) ? 0 : 1;
  for (int l = 0; l < 32; l++) p.lscale[l] = scale * std::pow(2.0, -op->alpha * l);
  for (int b = 0; b < 2; b++)
    for (int i = 0; i < P::M * P::M; i++) p.ip[b][i] = da.ip[b][i];
  if (op->kind == DKT_OP_DENSE)
  {
    if (!op->kref) { set_error("DKT_OP_DENSE needs kref"); return DKT_ERR_INVALID; }
    std::memcpy(p.K, op->kref, sizeof(double) * P::N * P::N);
  }
  CK(cudaMemsetAsync(d_out, 0, da.nNodes * sizeof(double), da.stream));
  g_launches++;
  const bool diri = op->dirichlet != 0;
  if (op->kind == DKT_OP_IDENTITY)
    return diri ? launch_mv<DIM, ORDER, DKT_OP_IDENTITY, true>(da, p) : launch_mv<DIM, ORDER, DKT_OP_IDENTITY, false>(da, p);
  if (op->kind == DKT_OP_DENSE)
    return diri ? launch_mv<DIM, ORDER, DKT_OP_DENSE, true>(da, p) : launch_mv<DIM, ORDER, DKT_OP_DENSE, false>(da, p);
  set_error("unknown operator kind");
  return DKT_ERR_INVALID;
}

void kron_to_dense(const dkt_op *op, int dim, int M, std::vector<double> &K)
{
  int N = 1;
  for (int d = 0; d < dim; d++) N *= M;
  K.assign((size_t)N * N, 0.0);
  for (int t = 0; t < op->terms; t++)
    for (int i = 0; i < N; i++)      // output index
      for (int j = 0; j < N; j++)    // input index
      {
        double v = 1.0;
        for (int d = 0, ii = i, jj = j; d < dim; d++, ii /= M, jj /= M)
          v *= op->kref[((size_t)(t * dim + d) * M + (jj % M)) * M + (ii % M)];  // A[k*M + j]: in k -> out j
        K[(size_t)i * N + j] += v;
      }
}

int run_matvec(DA &da, const dkt_op *op, const double *d_in, double *d_out, double scale, unsigned flags)
{
  if (op->kind == DKT_OP_KRON)
  {
    // the flat kernels take the dense matrix the Kronecker terms stand for
    if (!op->kref || op->terms < 1 || op->terms > DKT_KRON_MAX_TERMS) { set_error("DKT_OP_KRON needs 1..5 terms in kref"); return DKT_ERR_INVALID; }
    static thread_local std::vector<double> K;
    kron_to_dense(op, da.dim, da.order + 1, K);
    dkt_op dense = *op;
    dense.kind = DKT_OP_DENSE;
    dense.kref = K.data();
    return run_matvec(da, &dense, d_in, d_out, scale, flags);
  }
  const int key = da.dim * 10 + da.order;
  switch (key)
  {
  case 21: return run_typed<2, 1>(da, op, d_in, d_out, scale, flags);
  case 22: return run_typed<2, 2>(da, op, d_in, d_out, scale, flags);
  case 31: return run_typed<3, 1>(da, op, d_in, d_out, scale, flags);
  case 32: return run_typed<3, 2>(da, op, d_in, d_out, scale, flags);
  case 41: return run_typed<4, 1>(da, op, d_in, d_out, scale, flags);
  case 42: return run_big<4, 2>(da, op, d_in, d_out, scale, flags);
  default:
    set_error("unsupported (dim, order): kernels exist for dim 2-4 with order 1,2");
    return DKT_ERR_UNSUPPORTED;
  }
}
} // namespace dkt
