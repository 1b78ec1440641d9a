// Sibling-family matvec kernel (order 1, identity / Walsh-Hadamard operators, exact interpolation).
//
// One UNIT is a complete family - the 2^dim leaves of one parent - with its 3^dim node lattice (tables: dkt_chunks.cu,
// k_family_units / k_family_check / k_chunk_build).  One CTA per chunk of UPC families (512 elements), 128 threads, no loop
// over chunks: 3-4 CTAs share an SM and overlap each other's memory and compute phases.
//   L0  thread 0: cp.async.bulk (TMA 1-D bulk copies, mbarrier completion; SASS UBLKCP) of the chunk's contiguous tables -
//       node records, rk16, inv16, family records, jd|cnt - into shared memory
//   L1  all threads, once the records are there: cp.async 8-byte gathers un[n] <- u[gid[n]]               -- barrier A
//   Everything up to barrier B is warp-local: a warp owns FPW families, one thread per QUAD (the 4 children that differ in
//   dimensions 0, 1; the 2^(dim-2) quads of a family are neighbouring lanes).
//   F   the quads of a family share out its lattice points (static addressing, no index arithmetic):
//       Ls[f][p] = lscale(level f) * un[rk16[slot]]; then (__syncwarp) a hanging point takes the mean of the corners of
//       G(p) (exact order-1 interpolation from the parent's nodes = the family's corners)                 -- __syncwarp
//   Q   per child 2^dim conflict-free LDS with static offsets, the elemental operator (identity or Walsh-Hadamard form;
//       dimensions >= 2 are addressed XOR-permuted, with which both commute), Q1's scalar tau_c off its corner
//       (k_family_check), accumulation into the quad's 9 * 2^(dim-2) lattice points in registers; the points shared with
//       the other quads are summed with 1-2 shuffles each; in a family with hanging points the transposed interpolation
//       then runs on the registers, one dimension at a time (each lane owns the corner side of its XOR-permuted
//       dimensions, so no further communication); __syncwarp; the sums go back to Ls in place           -- barrier B
//   N   thread per chunk node: acc = sum_k Ls[inv16[jd[k] + n]], one plain store (node private to the chunk)
//       or one fp64 RED
// Also compiles under -DDKT_EMU (tests/emu): CPU test infrastructure only.
#include "dkt_chunks.h"

namespace dkt
{
#ifndef DKT_FAM_DIRECT
#define DKT_FAM_DIRECT 1   // 1: node values are gathered straight into the lattices; 0: staged per chunk node (un[]) and copied
#endif
#ifndef DKT_FAM_EXP
#define DKT_FAM_EXP 0      // timing experiments only (wrong results): 1 no node phase, 2 no hanging-node code, 4 no operator, 8 no gather, 16 no RED
#endif
#ifndef DKT_FAM_PF
#define DKT_FAM_PF 600   // > 0: every CTA prefetches the tables of chunk c + DKT_FAM_PF into L2 (cp.async.bulk.prefetch.L2)
#endif
#ifndef DKT_FAM_NB
#define DKT_FAM_NB 2     // nodes a thread of the node phase handles at a time
#endif
#ifndef DKT_FAM_TAU
#define DKT_FAM_TAU 1    // 1: quirk-Q1 scalar as FMAs with bit-masked weights instead of predicated adds
#endif
#ifndef DKT_FAM_GU
#define DKT_FAM_GU 2     // families whose slot words a lane loads before it issues their gathers
#endif
#ifndef DKT_FAM_TP
#define DKT_FAM_TP 1     // 1: transposed hanging-node passes as FMAs with bit-masked weights
#endif
#ifndef DKT_FAM_MINB
#define DKT_FAM_MINB 4   // resident CTAs per SM the family kernel is compiled for (128 registers: 64 bytes of spills in 4-D)
#endif

struct MvfParams
{
  const double *in;
  double *out;
  const uint16_t *rk16, *inv16, *jd;
  const uint32_t *frec, *rec, *nloc, *slotw;
  const uint64_t *node_off;
  uint32_t nSet, nChunks, jdStride, ncap;  // ncap: multiple of 4, >= the largest chunk's padded node count
  double lscale[32];
  double K[16];  // Walsh-Hadamard form: diagonal / N
};

#ifndef DKT_EMU
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes)
{
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(a), "r"(parity)
                 : "memory");
}
#else
// emulation: the barrier word counts [expected bytes + 1 | completed bytes]; a waiting fiber lets the others run
inline void mbar_init(uint64_t *bar, int) { *bar = 0; }
inline void mbar_init_fence() {}
inline void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { *bar += (uint64_t)bytes + 1; }
inline void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  memcpy(dst, src, bytes);
  *bar += (uint64_t)bytes << 32;
}
inline void bulk_prefetch_l2(const void *, uint32_t) {}
inline void mbar_wait(uint64_t *bar, uint32_t)
{
  while ((uint32_t)*bar == 0 || (*bar >> 32) + 1 < (uint32_t)*bar) emu::spin_yield();
}
#endif

// acc += a (acc = fma(a, w, acc)) under a predicate.  nvcc turns `if (p) acc += a` into an unconditional add and two FSELs;
// the PTX predicate keeps it ONE predicated instruction.
#ifndef DKT_EMU
__device__ __forceinline__ void add_if(double &acc, double a, uint32_t p)
{
  asm("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q add.rn.f64 %0, %0, %1;\n}\n" : "+d"(acc) : "d"(a), "r"(p));
}
__device__ __forceinline__ void fma_if(double &acc, double a, double w, uint32_t p)
{
  asm("{\n.reg .pred q;\nsetp.ne.u32 q, %3, 0;\n@q fma.rn.f64 %0, %1, %2, %0;\n}\n" : "+d"(acc) : "d"(a), "d"(w), "r"(p));
}
#else
inline void add_if(double &acc, double a, uint32_t p) { if (p) acc += a; }
inline void fma_if(double &acc, double a, double w, uint32_t p) { if (p) acc = fma(a, w, acc); }
#endif

// shared-memory layout of k_mvf (bytes; every section a multiple of 16)
template <int DIM>
struct FamSmem
{
  using F = Fam<DIM>;
  uint32_t oUn, oRec, oRk, oInv, oFrec, oJd, oBar, total;
  __host__ __device__ FamSmem(uint32_t ncap, uint32_t jdStride)
  {
    uint32_t o = F::UPC * F::S * 8;
    oUn = o; o += DKT_FAM_DIRECT ? 0 : (ncap + 2) * 8;
    oRec = o; o += ncap * 4;
    oRk = o; o += F::UPC * F::L * (DKT_FAM_DIRECT ? 4 : 2);  // direct: the slot words
    oInv = o; o += F::UPC * F::L * 2;
    oFrec = o; o += F::UPC * 16;
    oJd = o; o += 2 * jdStride * 2;
    oBar = o; o += 16;
    total = o;
  }
};

template <int DIM, int OPKIND, bool DIRI>
__global__ void __launch_bounds__(Fam<DIM>::TPB, DKT_FAM_MINB) k_mvf(const __grid_constant__ MvfParams p)
{
  using F = Fam<DIM>;
  constexpr int L = F::L, S = F::S, SA = F::SA, SB = F::SB, N = F::N, NS = F::NS, NL = F::NL, QPF = F::QPF, FPW = F::FPW, UPC = F::UPC,
                TPB = F::TPB;
  static_assert((UPC * S * 8) % 16 == 0 && (UPC * L * 2) % 16 == 0, "bulk copies need 16-byte sections");
  DKT_DYN_SMEM(double, sm);
  const FamSmem<DIM> lay(p.ncap, p.jdStride);
  char *smc = (char *)sm;
  double *Ls = sm;
  double *un = (double *)(smc + lay.oUn);
  uint32_t *rec = (uint32_t *)(smc + lay.oRec);
  uint16_t *rk = (uint16_t *)(smc + lay.oRk);
  uint16_t *inv = (uint16_t *)(smc + lay.oInv);
  uint32_t *frec = (uint32_t *)(smc + lay.oFrec);
  uint16_t *jd = (uint16_t *)(smc + lay.oJd);
  uint64_t *bar = (uint64_t *)(smc + lay.oBar);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t c = blockIdx.x;
  const uint32_t u0 = c * UPC;
  const int nfam = (int)min((uint32_t)UPC, p.nSet - u0);
  const uint64_t noff = p.node_off[c];
  const int nloc = (int)p.nloc[c];
  const uint32_t recBytes = (uint32_t)((nloc + 3) & ~3) * 4u;

  // ---- L0: bulk copies of the chunk's tables
  const uint32_t slotBytes = UPC * L * 2, frecBytes = UPC * 16, jdBytes = 2 * p.jdStride * 2;
  if (tid == 0)
  {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    mbar_init_fence();
  }
  __syncthreads();
#if DKT_FAM_DIRECT
  uint32_t *slw = (uint32_t *)rk;
  if (tid == 0)
  {
    mbar_expect_tx(bar, 2 * slotBytes + frecBytes);
    bulk_g2s(slw, p.slotw + (uint64_t)c * (UPC * L), 2 * slotBytes, bar);
    bulk_g2s(frec, p.frec + (uint64_t)c * (UPC * 4), frecBytes, bar);
    mbar_expect_tx(bar + 1, recBytes + slotBytes + jdBytes);
    bulk_g2s(rec, p.rec + noff, recBytes, bar + 1);
    bulk_g2s(inv, p.inv16 + (uint64_t)c * (UPC * L), slotBytes, bar + 1);
    bulk_g2s(jd, p.jd + (uint64_t)c * (2 * p.jdStride), jdBytes, bar + 1);
#if DKT_FAM_PF
    const uint32_t cp = c + DKT_FAM_PF;
    if (cp < p.nChunks)
    {
      bulk_prefetch_l2(p.slotw + (uint64_t)cp * (UPC * L), 2 * slotBytes);
      bulk_prefetch_l2(p.frec + (uint64_t)cp * (UPC * 4), frecBytes);
      bulk_prefetch_l2(p.inv16 + (uint64_t)cp * (UPC * L), slotBytes);
      bulk_prefetch_l2(p.rec + p.node_off[cp], (uint32_t)((p.nloc[cp] + 3) & ~3u) * 4u);
    }
#endif
  }
  (void)un;
  mbar_wait(bar, 0);  // slot words and family records
#else
  if (tid == 0)
  {
    mbar_expect_tx(bar, recBytes);
    bulk_g2s(rec, p.rec + noff, recBytes, bar);
    mbar_expect_tx(bar + 1, 2 * slotBytes + frecBytes + jdBytes);
    bulk_g2s(rk, p.rk16 + (uint64_t)c * (UPC * L), slotBytes, bar + 1);
    bulk_g2s(inv, p.inv16 + (uint64_t)c * (UPC * L), slotBytes, bar + 1);
    bulk_g2s(frec, p.frec + (uint64_t)c * (UPC * 4), frecBytes, bar + 1);
    bulk_g2s(jd, p.jd + (uint64_t)c * (2 * p.jdStride), jdBytes, bar + 1);
  }
  // ---- L1: gather the chunk's node values
  mbar_wait(bar, 0);
  for (int n = tid; n < nloc; n += TPB)
  {
    const uint32_t r = rec[n];
    if (DIRI && (r & REC_BDY)) un[n] = 0.0;
    else cp_async8(un + n, p.in + (r >> 2));
  }
  if (tid == 0) un[nloc] = 0.0;  // what the slots without a node read
  cp_async_commit();
  mbar_wait(bar + 1, 0);
  cp_async_wait_all();
  __syncthreads();  // A
#endif

  const int fw0 = warp * FPW;
  const int nfw = min(FPW, nfam - fw0);  // families of this warp
  if (nfw > 0)
  {
    const int fl = lane / QPF, j = lane % QPF;
    const bool act = fl < nfw;
    const int f = fw0 + (act ? fl : 0);
    const int c2 = j & 1, c3 = (j >> 1) & 1;
    double *Lf = Ls + f * S;
    const uint32_t *fr = frec + f * 4;
    // the quad's lattice points: 9 (i0, i1) x (s2, s3); s_d = 0: the corner side p_d = 2 c_d, s_d = 1: the middle p_d = 1
    int boff[NS];        // shared-memory offset of the points with (s2, s3)
    int kb[NS];          // their natural lattice index 9 p2 + 27 p3 (index into rk16)
    uint32_t g[NS];      // their hanging bits (bit i0 + 3 i1)
#pragma unroll
    for (int sg = 0; sg < NS; sg++)
    {
      const int s2 = sg & 1, s3 = sg >> 1;
      const int p2 = NL >= 1 ? (s2 ? 1 : 2 * c2) : 0, p3 = NL >= 2 ? (s3 ? 1 : 2 * c3) : 0;
      boff[sg] = SA * p2 + SB * p3;
      kb[sg] = 9 * p2 + 27 * p3;
      g[sg] = (fr[DIM == 4 ? p3 : 0] >> (9 * p2)) & 0x1FFu;
    }
    const bool hangfam = act && (fr[0] | fr[1] | fr[2]) != 0u && !(DKT_FAM_EXP & 2);

    // ---- F: fill the family's lattice.  A point with (s2, s3) is shared by the 2^|s| quads that differ in those dimensions;
    // they take its 9 (i0, i1) in turn (static, predicated code: no index arithmetic).
    {
      // real points.  A lane keeps the same lattice points k (at most KT of them: k0, k0 + 32, ..) for every family of the warp,
      // so the lattice address is computed once; the table reads of a family are coalesced.  Direct: one 8-byte cp.async per
      // slot, global -> lattice (a boundary node under Dirichlet rows gets 0, a hanging point is left to the passes below); no
      // barrier - the warp fills only its own families.  The level scale is applied to the output lattice instead (the
      // operator is linear).
      {
        constexpr int FPI = L >= 32 ? 1 : 32 / L;  // families per sweep of the warp
        constexpr int KT = (L + 31) / 32;          // lattice points per lane
        const int fsub = L >= 32 ? 0 : lane / L, k0 = L >= 32 ? lane : lane % L;
        int la[KT];
#pragma unroll
        for (int t = 0; t < KT; t++)
        {
          const int k = min(k0 + 32 * t, L - 1);
          la[t] = fam_laddr(DIM, k);
        }
#if DKT_FAM_DIRECT
        const uint32_t *sww = slw + fw0 * L;
#else
        const uint16_t *rkw = rk + fw0 * L;
#endif
#if DKT_FAM_DIRECT == 1 && DKT_FAM_GU > 1
        // DKT_FAM_GU families at a time: their slot words are loaded together, then the copies go out (the loads of one
        // family would otherwise wait for each other's shared-memory latency)
        if (fsub < FPI)
          for (int fl2 = fsub; fl2 < nfw; fl2 += FPI * DKT_FAM_GU)
          {
            uint32_t w[DKT_FAM_GU][KT];
#pragma unroll
            for (int u = 0; u < DKT_FAM_GU; u++)
#pragma unroll
              for (int t = 0; t < KT; t++)
              {
                const int k = k0 + 32 * t, f2 = fl2 + u * FPI;
                w[u][t] = (k < L && f2 < nfw) ? sww[f2 * L + k] : SLOTW_ABSENT;
              }
#pragma unroll
            for (int u = 0; u < DKT_FAM_GU; u++)
#pragma unroll
              for (int t = 0; t < KT; t++)
                if (!(w[u][t] & SLOTW_ABSENT))
                {
                  double *dst = Ls + (fw0 + fl2 + u * FPI) * S + la[t];
                  if ((DIRI && (w[u][t] & SLOTW_BDY)) || (DKT_FAM_EXP & 8)) *dst = 0.0;
                  else cp_async8(dst, p.in + (w[u][t] >> 2));
                }
          }
#else
        if (fsub < FPI)
          for (int fl2 = fsub; fl2 < nfw; fl2 += FPI)
          {
#pragma unroll
            for (int t = 0; t < KT; t++)
            {
              const int k = k0 + 32 * t;
              if (k >= L) continue;
              double *dst = Ls + (fw0 + fl2) * S + la[t];
#if DKT_FAM_DIRECT
              const uint32_t w = sww[fl2 * L + k];
              if (!(w & SLOTW_ABSENT))
              {
                if ((DIRI && (w & SLOTW_BDY)) || (DKT_FAM_EXP & 8)) *dst = 0.0;
                else cp_async8(dst, p.in + (w >> 2));
              }
#else
              *dst = un[rkw[fl2 * L + k]];
#endif
            }
          }
#endif
#if DKT_FAM_DIRECT
        cp_async_commit();
        cp_async_wait_all();
#endif
      }
      __syncwarp();
      // Hanging points: exact order-1 interpolation from the corners of G(p), one dimension at a time - the points whose LOWEST
      // odd coordinate is d take the mean of their two neighbours along d (corners, or hanging points of an earlier pass;
      // k_family_check's (C1) guarantees those hang too).  Quads that share a point compute the same value.
      if (NL >= 2)
      {
        if (hangfam)
        {
#pragma unroll
          for (int i1 = 0; i1 < 3; i1 += 2)
#pragma unroll
            for (int i0 = 0; i0 < 3; i0 += 2)
              if ((g[2] >> (i0 + 3 * i1)) & 1u)
              {
                const int o = i0 + 3 * i1 + SA * 2 * c2;
                Lf[boff[2] + i0 + 3 * i1] = 0.5 * (Lf[o] + Lf[o + 2 * SB]);
              }
        }
        __syncwarp();
      }
      if (NL >= 1)
      {
        if (hangfam)
        {
#pragma unroll
          for (int s3 = 0; s3 < (NL >= 2 ? 2 : 1); s3++)
#pragma unroll
            for (int i1 = 0; i1 < 3; i1 += 2)
#pragma unroll
              for (int i0 = 0; i0 < 3; i0 += 2)
                if ((g[1 + 2 * s3] >> (i0 + 3 * i1)) & 1u)
                {
                  const int o = i0 + 3 * i1 + (NL >= 2 ? SB * (s3 ? 1 : 2 * c3) : 0);
                  Lf[boff[1 + 2 * s3] + i0 + 3 * i1] = 0.5 * (Lf[o] + Lf[o + 2 * SA]);
                }
        }
        __syncwarp();
      }
      if (hangfam)
      {
#pragma unroll
        for (int sg = 0; sg < NS; sg++)
#pragma unroll
          for (int i0 = 0; i0 < 3; i0 += 2)
            if ((g[sg] >> (i0 + 3)) & 1u) Lf[boff[sg] + i0 + 3] = 0.5 * (Lf[boff[sg] + i0] + Lf[boff[sg] + i0 + 6]);
      }
      __syncwarp();
      if (hangfam)
      {
#pragma unroll
        for (int sg = 0; sg < NS; sg++)
#pragma unroll
          for (int i1 = 0; i1 < 3; i1++)
            if ((g[sg] >> (1 + 3 * i1)) & 1u) Lf[boff[sg] + 1 + 3 * i1] = 0.5 * (Lf[boff[sg] + 3 * i1] + Lf[boff[sg] + 2 + 3 * i1]);
      }
      __syncwarp();
    }

    // ---- Q: the quad's 4 children
    // Every lattice point is loaded ONCE, by the first child that touches it, and stays in a register until the last one has
    // used it (all indices are static: the live set peaks at the last child, where it is empty).
    double acc[9 * NS], lat[9 * NS];
#pragma unroll
    for (int cq = 0; cq < 4; cq++)
    {
      const int c0 = cq & 1, c1 = cq >> 1;
      double e[N];
#pragma unroll
      for (int r = 0; r < N; r++)
      {
        const int r0 = r & 1, r1 = (r >> 1) & 1;
        const int a = (c0 + r0) + 3 * (c1 + r1) + 9 * (r >> 2);
        if ((c0 == 0 || r0 == 1) && (c1 == 0 || r1 == 1)) lat[a] = Lf[boff[r >> 2] + (c0 + r0) + 3 * (c1 + r1)];
        e[r] = lat[a];
      }
      if (OPKIND == OP_HADAMARD && !(DKT_FAM_EXP & 4))
      {
        wht<N>(e);
#pragma unroll
        for (int i = 0; i < N; i++) e[i] *= p.K[i];
        wht<N>(e);
      }
      if (hangfam)
      {
        // quirk Q1 on a family (see k_family_check): tau = sum over the child's hanging ranks of 2^-|odd| eout, off its corner
#if DKT_FAM_TAU
        // the weight 2^-|odd| is a power of two: its bit pattern is all in the high word, which the hanging bit multiplies
        double tau = 0.0;
#pragma unroll
        for (int r = 0; r < N; r++)
        {
          const int i0 = c0 + (r & 1), i1 = c1 + ((r >> 1) & 1), sg = r >> 2;
          const int nodd = (i0 == 1) + (i1 == 1) + (sg & 1) + (sg >> 1);
          if (nodd == 0 || nodd == DIM) continue;  // the corner itself
          const uint32_t hi = ((g[sg] >> (i0 + 3 * i1)) & 1u) * (uint32_t)(0x3FF00000u - ((uint32_t)nodd << 20));
          tau = fma(__hiloint2double((int)hi, 0), e[r], tau);
        }
        e[c0 | (c1 << 1)] -= tau;
#else
        double tk[DIM] = {};  // sums of the hanging ranks with 1, 2, .. odd coordinates (the centre, all odd, never hangs)
#pragma unroll
        for (int r = 0; r < N; r++)
        {
          const int i0 = c0 + (r & 1), i1 = c1 + ((r >> 1) & 1), sg = r >> 2;
          const int nodd = (i0 == 1) + (i1 == 1) + (sg & 1) + (sg >> 1);
          if (nodd == 0 || nodd == DIM) continue;  // the corner itself
          add_if(tk[nodd - 1], e[r], (g[sg] >> (i0 + 3 * i1)) & 1u);
        }
        double tau = 0.0;
#pragma unroll
        for (int k = DIM - 2; k >= 0; k--) tau = fma(1.0 / (double)(2 << k), tk[k], tau);
        e[c0 | (c1 << 1)] -= tau;
#endif
      }
#pragma unroll
      for (int r = 0; r < N; r++)
      {
        const int r0 = r & 1, r1 = (r >> 1) & 1;
        const int a = (c0 + r0) + 3 * (c1 + r1) + 9 * (r >> 2);
        if ((c0 == 0 || r0 == 1) && (c1 == 0 || r1 == 1)) acc[a] = e[r];  // first child that touches the point
        else acc[a] += e[r];
      }
    }
    // the middle points of the lane dimensions are shared with the neighbouring quads
#pragma unroll
    for (int sg = 1; sg < NS; sg++)
#pragma unroll
      for (int i = 0; i < 9; i++)
      {
        double v = acc[i + 9 * sg];
        if (sg & 1) v += __shfl_xor_sync(0xffffffffu, v, 1);
        if (sg & 2) v += __shfl_xor_sync(0xffffffffu, v, 2);
        acc[i + 9 * sg] = v;
      }
    if (hangfam)
    {
      // transposed interpolation of the hanging points towards the corners, one dimension at a time
#if DKT_FAM_TP
      // weight 0.5 or 0 as a bit-masked high word: one FMA per target instead of an add and two selects (every lattice value
      // of a hanging family is finite, so 0 * value is 0; 0.5 * value is exact, so the rounding is that of the add)
#define DKT_HALF_IF(bit) __hiloint2double((int)((bit) * 0x3FE00000u), 0)
#pragma unroll
      for (int sg = 0; sg < NS; sg++)
#pragma unroll
        for (int i1 = 0; i1 < 3; i1++)
        {
          const double w = DKT_HALF_IF((g[sg] >> (1 + 3 * i1)) & 1u), m = acc[1 + 3 * i1 + 9 * sg];
          acc[0 + 3 * i1 + 9 * sg] = fma(m, w, acc[0 + 3 * i1 + 9 * sg]);
          acc[2 + 3 * i1 + 9 * sg] = fma(m, w, acc[2 + 3 * i1 + 9 * sg]);
        }
#pragma unroll
      for (int sg = 0; sg < NS; sg++)
#pragma unroll
        for (int i0 = 0; i0 < 3; i0++)
        {
          const double w = DKT_HALF_IF((g[sg] >> (i0 + 3)) & 1u), m = acc[i0 + 3 + 9 * sg];
          acc[i0 + 9 * sg] = fma(m, w, acc[i0 + 9 * sg]);
          acc[i0 + 6 + 9 * sg] = fma(m, w, acc[i0 + 6 + 9 * sg]);
        }
#pragma unroll
      for (int d = 0; d < NL; d++)
#pragma unroll
        for (int sg = 0; sg < NS; sg++)
        {
          if (!((sg >> d) & 1)) continue;
#pragma unroll
          for (int i = 0; i < 9; i++)
            acc[i + 9 * (sg ^ (1 << d))] = fma(acc[i + 9 * sg], DKT_HALF_IF((g[sg] >> i) & 1u), acc[i + 9 * (sg ^ (1 << d))]);
        }
#undef DKT_HALF_IF
#else
#pragma unroll
      for (int sg = 0; sg < NS; sg++)
#pragma unroll
        for (int i1 = 0; i1 < 3; i1++)
        {
          const uint32_t hb = (g[sg] >> (1 + 3 * i1)) & 1u;
          const double h = 0.5 * acc[1 + 3 * i1 + 9 * sg];
          add_if(acc[0 + 3 * i1 + 9 * sg], h, hb);
          add_if(acc[2 + 3 * i1 + 9 * sg], h, hb);
        }
#pragma unroll
      for (int sg = 0; sg < NS; sg++)
#pragma unroll
        for (int i0 = 0; i0 < 3; i0++)
        {
          const uint32_t hb = (g[sg] >> (i0 + 3)) & 1u;
          const double h = 0.5 * acc[i0 + 3 + 9 * sg];
          add_if(acc[i0 + 9 * sg], h, hb);
          add_if(acc[i0 + 6 + 9 * sg], h, hb);
        }
#pragma unroll
      for (int d = 0; d < NL; d++)
#pragma unroll
        for (int sg = 0; sg < NS; sg++)
        {
          if (!((sg >> d) & 1)) continue;
#pragma unroll
          for (int i = 0; i < 9; i++) fma_if(acc[i + 9 * (sg ^ (1 << d))], acc[i + 9 * sg], 0.5, (g[sg] >> i) & 1u);
        }
#endif
    }
    __syncwarp();  // every lane of the warp has read its lattice points
    if (act)
    {
      if (OPKIND != DKT_OP_IDENTITY)
      {
        const double sc = p.lscale[fr[3] & 31u];
#pragma unroll
        for (int i = 0; i < 9 * NS; i++) acc[i] *= sc;
      }
#pragma unroll
      for (int sg = 0; sg < NS; sg++)
      {
        // a shared point is stored by the quad with c_d == 0
        if (((sg & 1) && c2) || ((sg & 2) && c3)) continue;
#pragma unroll
        for (int i = 0; i < 9; i++) Lf[boff[sg] + i] = acc[i + 9 * sg];
      }
    }
  }
#if DKT_FAM_DIRECT
  mbar_wait(bar + 1, 0);
#endif
  __syncthreads();  // B

  // ---- N: the chunk's nodes.  Nodes are ranked by run length (descending) and cnt[k] = #nodes with a run longer than k,
  // so node n has a k-th contribution iff n < cnt[k]; every node of a family chunk has at least one and at most FAM_MAXRUN
  // (k_chunk_build cuts longer runs into pieces that are chunk nodes of their own).  A thread takes the nodes tid, tid + TPB,
  // ..., two at a time: their 2 x FAM_MAXRUN index loads go out together, then the value loads - two dependent steps for a
  // pair of nodes whatever their run lengths.  inv16 holds BYTE offsets into Ls.
  {
    static_assert(FAM_MAXRUN == 4, "the node phase is written for runs of at most four");
    const uint16_t *cnt = jd + p.jdStride;
    const int cnt0 = (DKT_FAM_EXP & 1) ? 0 : cnt[0];
    const int cn[4] = {cnt0, (int)cnt[1], (int)cnt[2], (int)cnt[3]};
    const int jo[4] = {0, (int)jd[1], (int)jd[2], (int)jd[3]};
    const char *Lb = (const char *)Ls;
    constexpr int NB = DKT_FAM_NB;
    for (int base = tid; base < cnt0; base += NB * TPB)
    {
      uint32_t ix[NB][4], rc[NB];
      double v[NB][4];
#pragma unroll
      for (int h = 0; h < NB; h++)
      {
        const int n = base + h * TPB;
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
          ix[h][k] = 0;
          if (n < cn[k]) ix[h][k] = inv[jo[k] + n];
        }
        rc[h] = n < cnt0 ? rec[n] : 0u;
      }
#pragma unroll
      for (int h = 0; h < NB; h++)
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
          const int n = base + h * TPB;
          v[h][k] = 0.0;
          if (n < cn[k]) v[h][k] = *(const double *)(Lb + ix[h][k]);
        }
#pragma unroll
      for (int h = 0; h < NB; h++)
      {
        const int n = base + h * TPB;
        if (n >= cnt0) continue;
        const double a = (v[h][0] + v[h][1]) + (v[h][2] + v[h][3]);
        const uint32_t r = rc[h];
        if (DIRI && (r & REC_BDY)) continue;
        double *dst = p.out + (r >> 2);
        if ((r & REC_SHARED) && !(DKT_FAM_EXP & 16)) atomicAdd(dst, a);
        else *dst = a;
      }
    }
  }
}

template <int DIM, int OPKIND, bool DIRI>
static int launch_family_one(DA &da, const ChunkSet &cs, MvfParams &p)
{
  using F = Fam<DIM>;
  if (cs.spu != F::L || (int)cs.elemsPerChunk != F::UPC) { set_error("internal: family set does not match its kernel"); return DKT_ERR_INVALID; }
  p.rk16 = cs.d_rk16; p.inv16 = cs.d_inv16; p.jd = cs.d_jd; p.frec = cs.d_frec; p.rec = (const uint32_t *)cs.d_rec; p.nloc = cs.d_nloc;
  p.node_off = cs.d_node_off; p.slotw = cs.d_slot;
  p.nSet = (uint32_t)cs.nElem; p.nChunks = cs.nChunks; p.jdStride = cs.jdStride;
  p.ncap = (cs.maxNloc + 3) & ~3u;
  const FamSmem<DIM> lay(p.ncap, p.jdStride);
  if (cs.maxLen > (uint32_t)FAM_MAXRUN || p.jdStride < 4) { set_error("internal: family set with runs longer than the kernel handles"); return DKT_ERR_INVALID; }
  auto kern = k_mvf<DIM, OPKIND, DIRI>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total));
  DKT_LAUNCH(kern, cs.nChunks, F::TPB, lay.total, da.cur ? da.cur : da.stream)(p);
  g_launches++;
  return DKT_OK;
}

template <int DIM>
static int launch_family_dim(DA &da, const ChunkSet &cs, int opkind, bool diri, MvfParams &p)
{
  if (opkind == OP_HADAMARD)
    return diri ? launch_family_one<DIM, OP_HADAMARD, true>(da, cs, p) : launch_family_one<DIM, OP_HADAMARD, false>(da, cs, p);
  if (opkind == DKT_OP_IDENTITY)
    return diri ? launch_family_one<DIM, DKT_OP_IDENTITY, true>(da, cs, p) : launch_family_one<DIM, DKT_OP_IDENTITY, false>(da, cs, p);
  set_error("internal: no sibling-family kernel for this operator");
  return DKT_ERR_UNSUPPORTED;
}

int launch_family_set(DA &da, const ChunkSet &cs, int opkind, bool dirichlet, const double *in, double *out, const double *lscale,
                      const double *Kdiag)
{
  static thread_local MvfParams p;
  p.in = in;
  p.out = out;
  for (int l = 0; l < 32; l++) p.lscale[l] = lscale[l];
  for (int i = 0; i < 16; i++) p.K[i] = i < (1 << da.dim) ? Kdiag[i] : 0.0;
  switch (da.dim)
  {
  case 2: return launch_family_dim<2>(da, cs, opkind, dirichlet, p);
  case 3: return launch_family_dim<3>(da, cs, opkind, dirichlet, p);
  case 4: return launch_family_dim<4>(da, cs, opkind, dirichlet, p);
  }
  set_error("unsupported dimension");
  return DKT_ERR_UNSUPPORTED;
}
} // namespace dkt
