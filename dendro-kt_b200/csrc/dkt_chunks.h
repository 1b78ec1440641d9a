// Declarations shared by the chunked-matvec translation units dkt_chunks.cu (tables, per-element kernels) and
// dkt_family.cu (sibling-family kernel).  Internal; also compiled under -DDKT_EMU (tests/emu/cuda_emu.h).
#ifndef DKT_CHUNKS_H
#define DKT_CHUNKS_H

#include "dkt_internal.h"

#ifdef DKT_EMU
#include "cuda_emu.h"
#else
#define DKT_LAUNCH(k, g, b, s, st) k<<<(g), (b), (s), (st)>>>
#define DKT_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#endif

#include <algorithm>
#include <cmath>
#include <cstdlib>
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

#ifndef DKT_MV3_LEAN
#define DKT_MV3_LEAN 1  // order-2 per-element kernels: node records read where needed instead of kept in registers (3 CTAs/SM)
#endif
constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;                       // 4096 slots per chunk
constexpr int SLOT_CAP = SORT_THREADS * SORT_ITEMS;
constexpr int MAX_LEN = 511;                         // run length of a node inside a chunk (9 bits)
constexpr uint32_t META_LEN = 0x1FFu;
constexpr uint32_t META_PRESENT = 0x2000u;  // node exists (its run may be empty: only read by this chunk)
constexpr uint32_t META_BDY = 0x4000u;
constexpr uint32_t META_SHARED = 0x8000u;
constexpr uint32_t REC_SHARED = 2u, REC_BDY = 1u;  // family sets: 4-byte node records gid << 2 | shared | boundary
constexpr int FAM_MAXRUN = 4;               // family sets: longest run of a chunk node (longer ones are cut, see k_chunk_build)
constexpr uint32_t SLOTW_BDY = 1u, SLOTW_ABSENT = 2u;  // family sets: slot words gid << 2 | flags
constexpr uint32_t SLOT_RO = 0x80000000u;   // unit slot table: read-only reference (node ids are < 2^31)

// Sibling-family sets.  The 3^dim lattice of a family lives in shared memory at  Ls[f * S + p0 + 3 p1 + SA p2 + SB p3];
// one thread (a "quad") handles the 4 children that differ in dimensions 0 and 1, the 2^(dim-2) quads of a family are
// neighbouring lanes.  S, SA, SB make every 8-byte access of the quad phase conflict-free (searched, half-warp model:
// tools/smem_sim.py validated that model against ncu in round 1).
template <int DIM>
struct Fam
{
  static constexpr int NL = DIM - 2;                 // dimensions spread over lanes
  static constexpr int QPF = 1 << NL;                // quads (threads) per family
  static constexpr int FPW = 32 / QPF;               // families per warp
  static constexpr int L = (DIM == 2 ? 9 : DIM == 3 ? 27 : 81);
  static constexpr int SA = (DIM == 4 ? 10 : 12), SB = 36;
  static constexpr int S = (DIM == 2 ? 9 : DIM == 3 ? 33 : 101);
  static constexpr int N = 1 << DIM;
  static constexpr int NS = 1 << NL;                 // (s2, s3) combinations: lattice points of a quad = 9 * NS
  static constexpr int TPB = 128;
  static constexpr int UPC = TPB / QPF;              // families per chunk (32 / 64 / 128: 512 elements)
};
__host__ __device__ constexpr int fam_L(int dim) { return dim == 2 ? 9 : dim == 3 ? 27 : 81; }
__host__ __device__ constexpr int fam_S(int dim) { return dim == 2 ? 9 : dim == 3 ? 33 : 101; }
__host__ __device__ constexpr int fam_SA(int dim) { return dim == 4 ? 10 : 12; }
__host__ __device__ constexpr int fam_UPC(int dim) { return 128 >> (dim - 2); }
// natural lattice index k = p0 + 3 p1 + 9 p2 + 27 p3  ->  offset inside the family's shared-memory lattice
__host__ __device__ constexpr int fam_laddr(int dim, int k) { return (k % 9) + fam_SA(dim) * ((k / 9) % 3) + 36 * (k / 27); }

// internal operator kind: K = (1/N) H diag(d) H with H the N x N Walsh-Hadamard matrix (N = 2^dim,
// order 1).  Every operator whose 1-D factors are 2x2 matrices of the form [[a,b],[b,a]] - mass,
// Laplacian and their combinations on axis-aligned cells - has this form; run_typed3 detects it
// from the dense kref on the host.  2*dim*N/2 add/sub pairs + N multiplies instead of N^2 FMAs.
constexpr int OP_HADAMARD = 100;

template <int N>
__device__ __forceinline__ void wht(double *v)
{
#pragma unroll
  for (int s = 1; s < N; s <<= 1)
  {
#pragma unroll
    for (int i = 0; i < N; i++)
    {
      if (i & s) continue;
      const double a = v[i], b = v[i + s];
      v[i] = a + b;
      v[i + s] = a - b;
    }
  }
}

#ifndef DKT_EMU
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc)
{
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc)
{
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
#else
inline void cp_async8(void *smem_dst, const void *gsrc) { memcpy(smem_dst, gsrc, 8); }
inline void cp_async4(void *smem_dst, const void *gsrc) { memcpy(smem_dst, gsrc, 4); }
inline void cp_async_commit() {}
inline void cp_async_wait_all() {}
#endif


// one sibling-family set of a matvec (dkt_family.cu).  opkind: DKT_OP_IDENTITY or OP_HADAMARD; Kdiag: 2^dim entries
// (diagonal of the Walsh-Hadamard form / 2^dim); lscale[32]: scale * 2^(-alpha level)
int launch_family_set(DA &da, const ChunkSet &cs, int opkind, bool dirichlet, const double *in, double *out, const double *lscale,
                      const double *Kdiag);
} // namespace dkt

#endif
