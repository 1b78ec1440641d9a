// The production matvec: SFC-contiguous CHUNKS of the tree staged through shared memory.
//
// Why: the flat kernels (dkt_matvec.cu) issue (order+1)^dim scattered 8-byte gathers and fp64
// atomics per element straight to L1/L2 and are L1TEX/atomic bound at ~15 % of the HBM roofline
// (profiles/r01_*).  Shared-memory fp64 atomics are CAS spin loops on sm_100a
// (ATOMS.CAST.SPIN.64), so the in-chunk reduction is made atomic-free instead.
//
// Two table layouts feed two kernel families:
//
// (1) SIBLING FAMILIES (order 1, the default wherever it applies; k_mvf).  A complete family - the 2^dim leaves of
//     one parent - shares a 3^dim node lattice.  One UNIT is a family: 3^dim lattice slots (81 in 4-D) instead
//     of 2^dim x 2^dim element slots (256), 16-bit table entries, and no separate parent-lattice slots for hanging
//     elements: the parent's nodes ARE the corners of the family lattice, a hanging point is interpolated from them
//     when the lattice is filled, and the transposed interpolation runs on the summed lattice in registers.  Quirk Q1
//     of the reference (FEM/include/matvec.h:517) reduces, on such a family, to ONE scalar per child subtracted
//     from its corner node (derivation at k_family_check, which admits only families where it holds).  Elements
//     outside such families ("singles") keep layout (2).
//
// (2) PER-ELEMENT sets (every order; k_mv3): one unit is an element, regular and hanging elements in separate sets.
//
//   build (once per DA): every set of UNITS first gets a "unit slot table" U[unit][slot] = node id
//   (k_unit_slots_elem, k_family_units), then k_chunk_build (one CTA per chunk, cub::BlockRadixSort in shared
//   memory) turns it into the chunk tables:
//     * the chunk's slots are sorted by the node they touch -> unique nodes, run length `len` of each node
//     * nodes are re-ranked by (len descending, id ascending): jagged-diagonal storage.  The k-th
//       contribution to node n lives at position jd[k] + n; jd[k] = number of (node, j<k) pairs.
//     * per-element sets: every slot gets one 32-bit word  n | (jd[k] + n) << 16, stored slot-major; every node
//       its global id and a meta word: len | boundary bit | shared-with-another-chunk bit
//     * family sets: rk16[slot] = n, inv16[jd[k] + n] = shared-memory address of the contributing lattice point,
//       one 32-bit record gid | boundary | shared per node, [jd | cnt] with cnt[k] = #nodes with a run longer than k
//
//   per-element matvec (k_mv3, persistent CTAs looping over chunks c = blockIdx.x, +gridDim.x, ...):
//     T0  wait for the cp.async gather of this chunk's node values, barrier
//     T1  load the NEXT chunk's node ids / meta into registers (overlaps T2)
//     T2  thread per element: ein[r] = un[n] (LDS), K_e, X[pos] = eout[r] (STS; every slot owns
//         one position, so there are no write conflicts and no atomics)
//     T3  issue cp.async 8-byte gathers un_next[n] <- u[gid] for the next chunk, load its slot
//         words and jd table (overlaps T4)
//     T4  barrier; thread per node: acc = sum_k X[jd[k] + n]  (lanes read consecutive words;
//         neighbouring lanes have equal len, so no divergence) and ONE plain store (node private
//         to the chunk) or ONE fp64 RED (shared) per (chunk, node)
//   family matvec: see k_mvf.
//
// Global atomics drop from N per element to ~2-3 per element in 4-D (the chunk surface).
//
// Order-1 extras of the per-element sets (all build-time table tricks, no extra kernel work):
//   * XOR slot schedule: slot s of an element with Morton child number c holds lattice rank s ^ c, so
//     sibling elements read the SAME node in the same instruction (shared-memory broadcast).  The
//     identity and Walsh-Hadamard operator forms commute with that permutation, and the parent->child
//     interpolation becomes child-independent (exact form: subset-sum transforms).
//   * the first 16 jagged diagonals start at positions congruent to k modulo 16, so the k-th
//     contributions to one node written by siblings in one instruction fall into distinct bank pairs.
// Hanging elements live in their own chunk sets with two slot rows each (own lattice + parent lattice),
// so neither instantiation diverges, and they carry no predicates: an absent node reads the chunk's zero
// entry un[nloc] and writes to a trash position behind the diagonals; parent slots masked by quirk Q1
// (FEM/include/matvec.h:517) are read-only (node rank but no position, not counted in the run length);
// one 32-bit mask per element tells which own slots are filled.  Semantics are those of dkt_matvec.cu
// (reference: FEM/include/matvec.h:378-522); the Q1-free variant runs on the flat kernels.
// Partitioned DAs build two phases of sets (interior / boundary = touches a ghost node) so that dkt_dist.cu can run the
// ghost exchanges and the boundary elements beside the interior ones.
//
// This file also compiles under -DDKT_EMU with tests/emu/cuda_emu.h (fibers on the CPU) - that build exists
// ONLY so the CPU test-suite can execute the table construction and the kernels' logic against the oracle;
// it is never part of libdkt.so.
#include "dkt_chunks.h"

#ifndef DKT_EMU
#include <cub/block/block_radix_sort.cuh>
#include <cub/block/block_scan.cuh>
#endif

namespace dkt
{


// Rows (of N slots) per chunk: bounded by the block sort capacity and by ONE element per thread
// in the matvec kernels.
#ifndef DKT_ROWS
#define DKT_ROWS 256
#endif
#ifndef DKT_REG_MINB
#define DKT_REG_MINB 3   // resident CTAs per SM the regular kernel is compiled for (register budget)
#endif
#ifndef DKT_HANG_MINB
#define DKT_HANG_MINB 4
#endif
int rows_per_chunk(int N)
{
  int r = std::min(SLOT_CAP / N, DKT_ROWS);
  return r & ~1;
}
static inline unsigned nblk(uint64_t n) { return (unsigned)((n + 255) / 256); }

// ------------------------------------------------------------------------------------------
// build
// ------------------------------------------------------------------------------------------
// Number of WRITING references of every node over the unit slot tables of all sets.
__global__ void k_ref_count(const uint32_t *U, uint64_t n, uint32_t *cnt)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t k = U[i];
  if (k == INVALID || (k & SLOT_RO)) return;
  atomicAdd(cnt + k, 1u);
}
// bit s of fmask[h]: own slot s of hanging element h is filled (slot order, i.e. XOR-permuted at order 1)
__global__ void k_fmask(const uint32_t *e2n_hang, const uint8_t *child_hang, uint64_t nHang, int N, int xorperm, uint32_t *fmask)
{
  uint64_t h = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (h >= nHang) return;
  const int c = xorperm ? child_hang[h] : 0;
  uint32_t m = 0;
  for (int q = 0; q < N; q++)
    if (e2n_hang[h * N + q] != INVALID) m |= 1u << (q ^ c);
  fmask[h] = m;
}

// Unit slot table of a per-element set: slot q < N holds the element's lattice rank q ^ c (c = Morton child
// number with the XOR schedule, else 0), slots N..2N-1 (hanging sets) the parent-lattice rank (q-N) ^ c.  Own
// slots always write; a parent slot writes only if the element's own rank is unfilled (quirk Q1 is static in
// the chunk tables: FEM/include/matvec.h:517), otherwise it is a read-only reference.
__global__ void k_unit_slots_elem(const uint32_t *e2n, const uint32_t *pnode, const uint8_t *child, uint64_t nUnits, int N, int rows,
                                  int xorperm, uint32_t *U)
{
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const int spu = rows * N;
  if (i >= nUnits * spu) return;
  const uint64_t u = i / spu;
  const int q = (int)(i % spu);
  const int c = xorperm ? child[u] : 0;
  uint32_t key;
  if (q < N) key = e2n[u * N + (q ^ c)];
  else
  {
    const int r = (q - N) ^ c;
    key = pnode[u * N + r];
    if (key != INVALID && e2n[u * N + r] != INVALID) key |= SLOT_RO;
  }
  U[i] = key;
}

// One CTA per chunk of `upc` units with `spu` slots each (U is unit-major).  WRITE == false: only report the
// chunk's node count and longest run.  fam != 0 (family sets, fam = dim): the tables are rk16[slot] = node rank (unit-major,
// i.e. in U's order) and inv16[position] = shared-memory address of the lattice point that contributes at that
// jagged-diagonal position, the node records ONE 32-bit word gid | boundary bit | shared bit, and the jd table is followed by
// cnt[k] = #nodes with a run longer than k, which replaces the run length of the records.
template <bool WRITE, int ITEMS>
__global__ void __launch_bounds__(SORT_THREADS)
k_chunk_build(const uint32_t *U, uint64_t nUnits, int spu, int upc, int fam, const uint32_t *refcnt, const uint8_t *isbdy,
              const uint8_t *sent, const uint64_t *node_off, int jdStride, uint32_t *nloc_out, uint32_t *maxlen_out, uint32_t *slot, uint16_t *rk16,
              uint16_t *inv16, uint32_t *gid_out, uint16_t *meta_out, uint32_t *rec_out, uint16_t *jd_out)
{
  constexpr int CAP = SORT_THREADS * ITEMS;
  using SortPairs = cub::BlockRadixSort<uint32_t, SORT_THREADS, ITEMS, uint16_t>;
  using SortKeys = cub::BlockRadixSort<uint32_t, SORT_THREADS, ITEMS>;
  using Scan = cub::BlockScan<int, SORT_THREADS>;
  __shared__ union
  {
    typename SortPairs::TempStorage pairs;
    typename SortKeys::TempStorage keys;
    typename Scan::TempStorage scan;
    uint16_t newrank[CAP];                     // by gid-rank: rank after the (len desc) re-sort (after the sorts)
  } tmp;
  uint16_t *s_newrank = tmp.newrank;
  __shared__ uint32_t s_last[SORT_THREADS];
  __shared__ uint16_t s_start[CAP + 1];        // by gid-rank: first sorted position of the node's references
  __shared__ uint16_t s_cw[CAP + 1];           // by gid-rank: number of WRITING references sorted before the node
  __shared__ int s_hist[MAX_LEN + 2];
  __shared__ int s_jd[MAX_LEN + 2];
  __shared__ int s_cnt[MAX_LEN + 2];          // #nodes with run length > k
  __shared__ int s_total, s_P, s_maxlen;

  const uint64_t c = blockIdx.x;
  const uint64_t u0 = c * (uint64_t)upc;
  const int nu = (int)min((uint64_t)upc, nUnits - u0);
  const int nslots = nu * spu;

  // key = node id, val = slot index inside the chunk (unit-major: unit * spu + slot) | read-only bit 15
  uint32_t key[ITEMS];
  uint16_t val[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
  {
    const int s = threadIdx.x * ITEMS + i;
    key[i] = INVALID;
    val[i] = (uint16_t)s;
    if (s < nslots)
    {
      const uint32_t raw = U[u0 * spu + s];
      if (raw != INVALID)
      {
        key[i] = raw & ~SLOT_RO;
        if (raw & SLOT_RO) val[i] |= 0x8000u;
      }
    }
  }
  if (threadIdx.x == 0) { s_P = 0; s_maxlen = 0; }
  for (int i = threadIdx.x; i < MAX_LEN + 2; i += SORT_THREADS) s_hist[i] = 0;
  SortPairs(tmp.pairs).Sort(key, val);
  __syncthreads();
  // heads: first occurrence of each node id in sorted order (invalid keys sort last)
  s_last[threadIdx.x] = key[ITEMS - 1];
  __syncthreads();
  uint32_t prev = threadIdx.x ? s_last[threadIdx.x - 1] : INVALID;
  int head[ITEMS];
  int nheads = 0, lastvalid = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
  {
    const bool first = (threadIdx.x == 0 && i == 0);
    head[i] = (key[i] != INVALID) && (first || key[i] != prev);
    prev = key[i];
    nheads += head[i];
    if (key[i] != INVALID) lastvalid = threadIdx.x * ITEMS + i + 1;
  }
  int before = 0, total = 0;
  Scan(tmp.scan).ExclusiveSum(nheads, before, total);
  if (lastvalid) atomicMax(&s_P, lastvalid);
  uint16_t rank0[ITEMS];
  {
    int rk = before - 1;
#pragma unroll
    for (int i = 0; i < ITEMS; i++)
    {
      if (head[i])
      {
        rk++;
        s_start[rk] = (uint16_t)(threadIdx.x * ITEMS + i);
      }
      rank0[i] = (uint16_t)(rk < 0 ? 0 : rk);
    }
  }
  if (fam)
  {
    // Family sets: a run longer than FAM_MAXRUN is cut into pieces, each a chunk node of its own with the same global id (the
    // pieces are accumulated with RED like a node shared between chunks).  The kernel's node phase then needs at most
    // FAM_MAXRUN dependent steps per node instead of up to 2^dim.  Heads of the pieces: every FAM_MAXRUN-th reference.
    __syncthreads();
    nheads = 0;
#pragma unroll
    for (int i = 0; i < ITEMS; i++)
    {
      if (key[i] != INVALID)
      {
        const int kidx = threadIdx.x * ITEMS + i - (int)s_start[rank0[i]];
        head[i] = (kidx % FAM_MAXRUN) == 0;
      }
      nheads += head[i];
    }
    __syncthreads();
    Scan(tmp.scan).ExclusiveSum(nheads, before, total);
    int rk = before - 1;
#pragma unroll
    for (int i = 0; i < ITEMS; i++)
    {
      if (head[i])
      {
        rk++;
        s_start[rk] = (uint16_t)(threadIdx.x * ITEMS + i);
      }
      rank0[i] = (uint16_t)(rk < 0 ? 0 : rk);
    }
  }
  if (threadIdx.x == 0) s_total = total;
  // k-th WRITING reference of a node -> diagonal k
  int wr[ITEMS];
  int nwr = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
  {
    wr[i] = (key[i] != INVALID) && !(val[i] & 0x8000u);
    nwr += wr[i];
  }
  __syncthreads();
  int wbefore = 0, wtotal = 0;
  Scan(tmp.scan).ExclusiveSum(nwr, wbefore, wtotal);
  uint16_t cwi[ITEMS];
  {
    int acc = wbefore;
#pragma unroll
    for (int i = 0; i < ITEMS; i++)
    {
      cwi[i] = (uint16_t)acc;
      if (head[i]) s_cw[rank0[i]] = (uint16_t)acc;
      acc += wr[i];
    }
  }
  __syncthreads();
  const int nloc = s_total;
  if (threadIdx.x == 0)
  {
    s_start[nloc] = (uint16_t)s_P;
    s_cw[nloc] = (uint16_t)wtotal;
  }
  __syncthreads();
  // second sort: nodes by (len descending, gid-rank ascending)
  uint32_t k2[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
  {
    const int n = threadIdx.x * ITEMS + i;
    k2[i] = INVALID;
    if (n < nloc)
    {
      const int len = (int)s_cw[n + 1] - (int)s_cw[n];  // writing references only
      k2[i] = ((uint32_t)(MAX_LEN - min(len, MAX_LEN)) << 13) | (uint32_t)n;
      atomicAdd(&s_hist[min(len, MAX_LEN + 1)], 1);
      atomicMax(&s_maxlen, len);
    }
  }
  __syncthreads();
  if (!WRITE)
  {
    if (threadIdx.x == 0)
    {
      nloc_out[c] = (uint32_t)nloc;
      maxlen_out[c] = (uint32_t)s_maxlen;
    }
    return;
  }
  SortKeys(tmp.keys).Sort(k2, 0, 22);
  __syncthreads();  // tmp is re-used as newrank[]
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
  {
    const int j = threadIdx.x * ITEMS + i;
    if (k2[i] != INVALID) s_newrank[k2[i] & 0x1FFFu] = (uint16_t)j;
  }
  // jd[k] = sum_{j<k} count_j, count_j = #nodes with len > j
  if (threadIdx.x == 0)
  {
    const int ml = s_maxlen;
    int above = 0;
    for (int k = ml; k >= 0; k--)
    {
      s_jd[k] = above;  // #nodes with len > k
      s_cnt[k] = above;
      above += s_hist[k];
    }
    for (int k = ml + 1; k < MAX_LEN + 2; k++) s_cnt[k] = 0;
    // diagonal k starts at a position congruent to k modulo 16 (the number of 8-byte bank pairs):
    // the k-th contributions to one node - written by sibling elements in the same instruction
    // under the XOR slot schedule - then fall into distinct banks
    int acc = 0;
    for (int k = 0; k <= ml; k++)
    {
      const int cnt = s_jd[k];
      if (k < 16 && !fam)
        while ((acc & 15) != k) acc++;
      s_jd[k] = acc;
      acc += cnt;
    }
    for (int k = ml + 1; k < MAX_LEN + 2; k++) s_jd[k] = acc;
  }
  __syncthreads();
  const uint64_t noff = node_off[c];
  if (!fam)
    for (int k = threadIdx.x; k < jdStride; k += SORT_THREADS) jd_out[c * (uint64_t)jdStride + k] = (uint16_t)s_jd[min(k, MAX_LEN + 1)];
  else  // family sets: [jd | cnt] per chunk; the node records carry no run length, cnt[] gives it (nodes are ranked by it)
    for (int k = threadIdx.x; k < jdStride; k += SORT_THREADS)
    {
      jd_out[c * (uint64_t)(2 * jdStride) + k] = (uint16_t)s_jd[min(k, MAX_LEN + 1)];
      jd_out[c * (uint64_t)(2 * jdStride) + jdStride + k] = (uint16_t)s_cnt[min(k, MAX_LEN + 1)];
    }
  // absent node: read the chunk's zero entry un[nloc], write to the trash position behind the diagonals
  const uint32_t trash = (uint32_t)s_jd[MAX_LEN + 1];
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
  {
    const int sv = val[i] & 0x1FFF;
    if (sv >= nslots) continue;  // padding item of the sort
    // slot-major inside the chunk: a warp (consecutive units) reads consecutive words
    const int q = sv % spu;
    const int el = sv / spu;
    uint32_t nr = (uint32_t)nloc, pos = trash;
    if (key[i] != INVALID)
    {
      const int n0 = rank0[i];
      const int k = (int)cwi[i] - (int)s_cw[n0];
      nr = s_newrank[n0];
      if (wr[i]) pos = (uint32_t)s_jd[k] + nr;
    }
    if (!fam) slot[u0 * spu + (uint64_t)q * upc + el] = nr | (pos << 16);
    else
    {
      const uint64_t cbase = c * (uint64_t)upc * spu;
      rk16[cbase + sv] = (uint16_t)nr;
      // what the kernel gathers for the lattice slot: node id << 2 | boundary bit, or SLOTW_ABSENT (a hanging point)
      slot[cbase + sv] = key[i] != INVALID ? ((key[i] << 2) | (isbdy[key[i]] ? SLOTW_BDY : 0u)) : SLOTW_ABSENT;
      if (key[i] != INVALID) inv16[cbase + pos] = (uint16_t)(8 * (el * fam_S(fam) + fam_laddr(fam, q)));  // byte offset; every family slot writes
    }
    if (key[i] != INVALID && head[i])
    {
      const int n0 = rank0[i];
      const int len = (int)s_cw[n0 + 1] - (int)s_cw[n0];
      uint32_t m = (uint32_t)len | META_PRESENT;
      if (refcnt[key[i]] != (uint32_t)len || (sent && sent[key[i]])) m |= META_SHARED;
      if (isbdy[key[i]]) m |= META_BDY;
      if (!fam)
      {
        gid_out[noff + nr] = key[i];
        meta_out[noff + nr] = (uint16_t)m;
      }
      else rec_out[noff + nr] = (key[i] << 2) | ((m & META_SHARED) ? REC_SHARED : 0u) | ((m & META_BDY) ? REC_BDY : 0u);
    }
  }
}

__global__ void k_u32_widen(const uint32_t *in, uint64_t n, uint64_t *out)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}
__global__ void k_max_u32(const uint32_t *in, uint64_t n, uint32_t *out)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint32_t v = i < n ? in[i] : 0;
#ifndef DKT_EMU
  for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, v);
#else
  atomicMax(out, v);
#endif
}
__global__ void k_pad4_widen(const uint32_t *in, uint64_t n, uint64_t *out)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (in[i] + 3u) & ~3u;
}

// Chunk tables of one set from its unit slot table U (freed by the caller).
static int build_set(DA &da, ChunkSet &cs, const uint32_t *U, const uint32_t *refcnt)
{
  const uint64_t nSet = cs.nElem;
  if (nSet == 0) return DKT_OK;
  const int spu = cs.spu, upc = (int)cs.elemsPerChunk, fam = cs.kind == 2 ? da.dim : 0;
  cs.nChunks = (uint32_t)((nSet + upc - 1) / upc);
  uint32_t *nloc = nullptr, *mlen = nullptr;
  uint64_t *wide = nullptr, *off = nullptr;
  CK(cudaMalloc((void **)&nloc, (size_t)cs.nChunks * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&mlen, (size_t)cs.nChunks * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&wide, ((size_t)cs.nChunks + 1) * sizeof(uint64_t)));
  CK(cudaMalloc((void **)&off, ((size_t)cs.nChunks + 1) * sizeof(uint64_t)));
  auto kcount = k_chunk_build<false, SORT_ITEMS>;
  auto kwrite = k_chunk_build<true, SORT_ITEMS>;
  DKT_LAUNCH(kcount, cs.nChunks, SORT_THREADS, 0, da.stream)(U, nSet, spu, upc, fam, refcnt, da.d_node_isbdy, da.d_node_sent, nullptr, 0, nloc, mlen,
                                                              nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  g_launches++;
  CK(cudaMemsetAsync(wide, 0, ((size_t)cs.nChunks + 1) * sizeof(uint64_t), da.stream));
  // family sets: every chunk's node records start at a multiple of 16 bytes (cp.async.bulk)
  if (fam) DKT_LAUNCH(k_pad4_widen, nblk(cs.nChunks), 256, 0, da.stream)(nloc, cs.nChunks, wide);
  else DKT_LAUNCH(k_u32_widen, nblk(cs.nChunks), 256, 0, da.stream)(nloc, cs.nChunks, wide);
  g_launches++;
  int rc = device_exclusive_scan(da, wide, off, (uint64_t)cs.nChunks + 1);
  if (rc) return rc;
  uint64_t total = 0;
  CK(cudaMemcpy(&total, off + cs.nChunks, sizeof(uint64_t), cudaMemcpyDeviceToHost));
  uint32_t *dmax = nullptr, hmax[2] = {0, 0};
  CK(cudaMalloc((void **)&dmax, 2 * sizeof(uint32_t)));
  CK(cudaMemsetAsync(dmax, 0, 2 * sizeof(uint32_t), da.stream));
  DKT_LAUNCH(k_max_u32, nblk(cs.nChunks), 256, 0, da.stream)(nloc, cs.nChunks, dmax);
  DKT_LAUNCH(k_max_u32, nblk(cs.nChunks), 256, 0, da.stream)(mlen, cs.nChunks, dmax + 1);
  g_launches += 2;
  CK(cudaMemcpyAsync(hmax, dmax, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, da.stream));
  CK(cudaStreamSynchronize(da.stream));
  cs.maxNloc = hmax[0];
  cs.maxLen = hmax[1];
  if (cs.maxLen > (uint32_t)MAX_LEN) { set_error("internal: node referenced more than 511 times inside one chunk"); return DKT_ERR_UNSUPPORTED; }
  cs.jdStride = (cs.maxLen + 1 + 7) & ~7u;
  cs.totalNodes = total;
  cs.d_node_off = off;
  const size_t nslotsAll = (size_t)cs.nChunks * upc * spu;
  if (!fam)
  {
    CK(cudaMalloc((void **)&cs.d_slot, nslotsAll * sizeof(uint32_t)));
    CK(cudaMalloc((void **)&cs.d_gid, std::max<uint64_t>(total, 1) * sizeof(uint32_t)));
    CK(cudaMalloc((void **)&cs.d_meta, std::max<uint64_t>(total, 1) * sizeof(uint16_t)));
  }
  else
  {
    CK(cudaMalloc((void **)&cs.d_rk16, nslotsAll * sizeof(uint16_t)));
    CK(cudaMalloc((void **)&cs.d_slot, nslotsAll * sizeof(uint32_t)));
    CK(cudaMemsetAsync(cs.d_slot, 0xFF, nslotsAll * sizeof(uint32_t), da.stream));  // padding slots: absent
    CK(cudaMalloc((void **)&cs.d_inv16, nslotsAll * sizeof(uint16_t)));
    CK(cudaMalloc((void **)&cs.d_rec, std::max<uint64_t>(total, 4) * sizeof(uint32_t)));
    // bytes the kernel's bulk copies read but the build does not write (tail of the last chunk, padding records)
    CK(cudaMemsetAsync(cs.d_rk16, 0, nslotsAll * sizeof(uint16_t), da.stream));
    CK(cudaMemsetAsync(cs.d_inv16, 0, nslotsAll * sizeof(uint16_t), da.stream));
    CK(cudaMemsetAsync(cs.d_rec, 0, std::max<uint64_t>(total, 4) * sizeof(uint32_t), da.stream));
    cs.d_nloc = nloc;
  }
  CK(cudaMalloc((void **)&cs.d_jd, (size_t)cs.nChunks * cs.jdStride * (fam ? 2 : 1) * sizeof(uint16_t)));
  DKT_LAUNCH(kwrite, cs.nChunks, SORT_THREADS, 0, da.stream)(U, nSet, spu, upc, fam, refcnt, da.d_node_isbdy, da.d_node_sent, off, (int)cs.jdStride, nullptr,
                                                              nullptr, cs.d_slot, cs.d_rk16, cs.d_inv16, cs.d_gid, cs.d_meta, (uint32_t *)cs.d_rec,
                                                              cs.d_jd);
  g_launches++;
  CK(cudaStreamSynchronize(da.stream));
  CK(cudaGetLastError());
  if (!fam) cudaFree(nloc);
  cudaFree(mlen);
  cudaFree(wide);
  cudaFree(dmax);
  return DKT_OK;
}

void free_chunks(DA &da)
{
  cudaFree(da.d_mv_child);
  cudaFree(da.d_fmask);
  da.d_mv_child = nullptr;
  da.d_fmask = nullptr;
  for (std::vector<ChunkSet> *sets : {&da.sets, &da.sets_elem})
  {
    for (ChunkSet &cs : *sets)
    {
      cudaFree(cs.d_slot); cudaFree(cs.d_gid); cudaFree(cs.d_meta); cudaFree(cs.d_jd); cudaFree(cs.d_node_off);
      cudaFree(cs.d_rk16); cudaFree(cs.d_inv16); cudaFree(cs.d_frec); cudaFree(cs.d_nloc); cudaFree(cs.d_rec);
      for (void *p : cs.owned) cudaFree(p);
    }
    sets->clear();
  }
  da.families = false;
  if (da.ev_fork) cudaEventDestroy(da.ev_fork);
  da.ev_fork = nullptr;
  for (int a = 0; a < DA::MAX_AUX; a++)
  {
    if (da.aux[a]) cudaStreamDestroy(da.aux[a]);
    if (da.ev_join[a]) cudaEventDestroy(da.ev_join[a]);
    da.aux[a] = nullptr;
    da.ev_join[a] = nullptr;
  }
}

__global__ void k_child_numbers(const uint32_t *xyz, const uint8_t *lev, uint64_t n, int dim, int max_depth, uint8_t *child)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int L = lev[i];
  int c = 0;
  for (int d = 0; d < dim; d++) c |= ((xyz[i * dim + d] >> (max_depth - L)) & 1u) << d;
  child[i] = (uint8_t)(L ? c : 0);
}
// ---- sibling families: discovery -------------------------------------------------------------
// d_mv_src is the position of every visited element in the visit (SFC) order (minus mv_src0 on a partitioned
// DA); a complete family of leaves is 2^dim consecutive positions with one parent.
__global__ void k_invert_src(const uint32_t *src, uint64_t n, uint64_t base, uint32_t *inv)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) inv[src[i] - base] = (uint32_t)i;
}
__global__ void k_family_heads(const uint32_t *inv, const uint32_t *xyz, const uint8_t *lev, uint64_t n, int dim, int max_depth,
                               uint64_t *head)
{
  uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (j > n) return;
  if (j == n) { head[j] = 0; return; }
  const int nch = 1 << dim;
  uint64_t h = 0;
  if (j + nch <= n)
  {
    const uint32_t a = inv[j];
    const int L = lev[a];
    if (L >= 1)
    {
      const uint32_t pmask = ~((2u << (max_depth - L)) - 1u);
      uint32_t seen = 0;
      bool ok = true;
      for (int t = 0; t < nch && ok; t++)
      {
        const uint32_t b = inv[j + t];
        if (lev[b] != L) { ok = false; break; }
        int c = 0;
        for (int d = 0; d < dim; d++)
        {
          const uint32_t x = xyz[(uint64_t)b * dim + d];
          if ((x & pmask) != (xyz[(uint64_t)a * dim + d] & pmask)) ok = false;
          c |= ((x >> (max_depth - L)) & 1u) << d;
        }
        seen |= 1u << c;
      }
      h = (ok && seen == ((1u << nch) - 1u)) ? 1 : 0;
    }
  }
  head[j] = h;
}
// mem[f * 2^dim + c] = visited-element index of child c of family f
__global__ void k_family_members(const uint32_t *inv, const uint64_t *head, const uint64_t *fpos, const uint8_t *child, uint64_t n, int dim,
                                 uint32_t *mem)
{
  uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (j >= n || !head[j]) return;
  const int nch = 1 << dim;
  const uint64_t f = fpos[j];
  for (int t = 0; t < nch; t++)
  {
    const uint32_t b = inv[j + t];
    mem[f * nch + child[b]] = b;
  }
}
// class of a unit: bit 1 = hanging, bit 0 = boundary (touches a ghost node; partitioned DA with comm/compute overlap).
// A group is hanging / boundary if one of its 2^g members is.
__device__ __forceinline__ int elem_class(uint64_t e, uint64_t nReg, uint64_t nRegInt, uint64_t nHangInt, int phased)
{
  const int hang = e >= nReg;
  const int bdy = phased && (hang ? (e - nReg >= nHangInt) : (e >= nRegInt));
  return (hang << 1) | bdy;
}
// the elements outside complete families: class as above, 255 for family members
__global__ void k_single_class(const uint8_t *infam, uint64_t n, uint64_t nReg, uint64_t nRegInt, uint64_t nHangInt, int phased,
                               uint8_t *cls)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  cls[i] = infam[i] ? (uint8_t)255 : (uint8_t)elem_class(i, nReg, nRegInt, nHangInt, phased);
}
__global__ void k_class_flags(const uint8_t *cls, uint64_t n, int want, uint64_t *flag)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i > n) return;
  flag[i] = (i < n && cls[i] == want) ? 1 : 0;
}
__global__ void k_compact(const uint64_t *flag, const uint64_t *pos, uint64_t n, uint32_t *list)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n && flag[i]) list[pos[i]] = (uint32_t)i;
}
// compact copies of the per-element arrays of the elements in `list`
__global__ void k_gather_single(const uint32_t *list, uint64_t n, int N, uint64_t nReg, const uint32_t *e2n, const uint32_t *pnode,
                                const uint8_t *lev, const uint8_t *child, uint32_t *e2n_s, uint32_t *pnode_s, uint8_t *lev_s,
                                uint8_t *child_s)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t e = list[i];
  for (int r = 0; r < N; r++) e2n_s[i * N + r] = e2n[(uint64_t)e * N + r];
  if (pnode_s)
    for (int r = 0; r < N; r++) pnode_s[i * N + r] = pnode[((uint64_t)e - nReg) * N + r];
  lev_s[i] = lev[e];
  child_s[i] = child[e];
}

// ---- sibling families: unit slot tables ------------------------------------------------------
// Lattice point k = sum_d p_d 3^d (p_d in 0..2) of family f: the node every child touching it stores at that
// point, INVALID if they store none (a hanging point), bad[f] if the children disagree.
__global__ void k_family_units(const uint32_t *mem, uint64_t nFam, int dim, const uint32_t *e2n, uint32_t *U, uint32_t *hm, uint32_t *bad)
{
  const int L = fam_L(dim), N = 1 << dim;
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= nFam * L) return;
  const uint64_t f = i / L;
  const int k = (int)(i % L);
  int pd[4] = {0, 0, 0, 0};
  for (int d = 0, kk = k; d < dim; d++, kk /= 3) pd[d] = kk % 3;
  uint32_t key = INVALID;
  bool first = true, ok = true;
  for (int c = 0; c < N; c++)
  {
    int r = 0;
    bool touches = true;
    for (int d = 0; d < dim; d++)
    {
      const int rd = pd[d] - ((c >> d) & 1);
      if (rd < 0 || rd > 1) touches = false;
      else r |= rd << d;
    }
    if (!touches) continue;
    const uint32_t v = e2n[(uint64_t)mem[f * N + c] * N + r];
    if (first) { key = v; first = false; }
    else if (v != key) ok = false;
  }
  U[i] = key;
  if (!ok) atomicOr(bad + f, 1u);
  if (key == INVALID) atomicOr(hm + f * 4 + k / 27, 1u << (k % 27));
}

// Admits a family to the family kernel iff the lattice form of the hanging-node treatment equals the reference's
// per-element one (FEM/include/matvec.h:403-522) on it.  Per element the reference (a) interpolates every unfilled
// rank r of child c from the parent nodes q that are valid (level == L-1, matvec.h:417), (b) after the elemental
// operator adds sum_{r unfilled} A^T[q, r] eout_c[r] to parent node q - but only if the CHILD's own rank q is unfilled
// (quirk Q1, matvec.h:517).  On the family lattice (child c, rank r <-> point c + r; parent rank q <-> corner 2q):
//   * the weight of corner q for point p is 2^-|odd(p)| if q is a corner of G(p), the smallest face of the parent cell
//     that contains p (odd(p) = coordinates equal to 1), else 0 - independent of the child;
//   * child c's rank q is the point p' = c + q, which lies in the closure of G(p) for every hanging p of c that
//     gives q a non-zero weight, and is a corner of it only for q == c.
// So if (C1) every non-corner point of closed G(p) hangs whenever p hangs, Q1 drops exactly the terms with q == c,
// i.e. one scalar tau_c = sum_{r unfilled} 2^-|odd(c + r)| eout_c[r] per child, taken off its corner node, and the rest
// is the plain transposed interpolation of the SUMMED lattice.  (C1) holds on 2:1-balanced trees without
// domain-boundary hanging nodes (classes A/B of SURVEY 8a): p hangs on a coarser leaf E, closed G(p) lies in E's
// boundary, and a point of E's boundary that is not a corner of E carries no node.  Also checked: (C2) hanging points
// are neither corners nor the centre, (C3) for every hanging p, every child c touching it and every corner q of G(p)
// the reference's pnode[c][q] is valid and IS the lattice's corner node.  Families that fail (class P trees, quirks
// Q2/Q3) stay on the per-element sets.
__global__ void k_family_check(const uint32_t *mem, uint64_t nFam, int dim, uint64_t nReg, const uint32_t *pnode, const uint32_t *U,
                               const uint32_t *hm, uint32_t *bad)
{
  const int L = fam_L(dim), N = 1 << dim;
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= nFam * L) return;
  const uint64_t f = i / L;
  const int k = (int)(i % L);
  if (!((hm[f * 4 + k / 27] >> (k % 27)) & 1u)) return;
  int pd[4] = {0, 0, 0, 0}, p3[4] = {1, 3, 9, 27};
  int odd = 0, nodd = 0;
  for (int d = 0, kk = k; d < dim; d++, kk /= 3)
  {
    pd[d] = kk % 3;
    if (pd[d] == 1) { odd |= 1 << d; nodd++; }
  }
  bool ok = nodd >= 1 && nodd < dim;  // (C2)
  // (C1): the points of closed G(p): odd coordinates range over 0..2
  int npts = 1;
  for (int j = 0; j < nodd; j++) npts *= 3;
  for (int t = 0; t < npts && ok; t++)
  {
    int kk = 0, tt = t, no = 0;
    for (int d = 0; d < dim; d++)
    {
      int x = pd[d];
      if ((odd >> d) & 1) { x = tt % 3; tt /= 3; }
      if (x == 1) no++;
      kk += x * p3[d];
    }
    if (no == 0) { if (U[f * L + kk] == INVALID) ok = false; }                 // its corners are nodes
    else if (!((hm[f * 4 + kk / 27] >> (kk % 27)) & 1u)) ok = false;         // everything else hangs
  }
  // (C3)
  for (int c = 0; c < N && ok; c++)
  {
    bool touches = true;
    for (int d = 0; d < dim; d++)
    {
      const int rd = pd[d] - ((c >> d) & 1);
      if (rd < 0 || rd > 1) touches = false;
    }
    if (!touches) continue;
    const uint32_t e = mem[f * N + c];
    if (e < nReg) { ok = false; break; }  // a child with an unfilled rank is in the hanging list
    for (int sub = 0; sub < (1 << nodd) && ok; sub++)
    {
      int q = 0, kc = 0, j = 0;
      for (int d = 0; d < dim; d++)
      {
        int qd = pd[d] >> 1;
        if ((odd >> d) & 1) { qd = (sub >> j) & 1; j++; }
        q |= qd << d;
        kc += 2 * qd * p3[d];
      }
      if (pnode[((uint64_t)e - nReg) * N + q] != U[f * L + kc]) ok = false;
    }
  }
  if (!ok) atomicOr(bad + f, 1u);
}

// class of a family: 255 not admitted, else bit 0 = boundary (one member touches a ghost node; partitioned DA with overlap);
// infam[e] = 1 for the members of admitted families
__global__ void k_family_class(const uint32_t *mem, const uint32_t *bad, uint64_t nFam, int dim, uint64_t nReg, uint64_t nRegInt,
                               uint64_t nHangInt, int phased, uint8_t *cls, uint8_t *infam)
{
  uint64_t f = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (f >= nFam) return;
  if (bad[f]) { cls[f] = 255; return; }
  const int N = 1 << dim;
  int c = 0;
  for (int t = 0; t < N; t++)
  {
    const uint32_t e = mem[f * N + t];
    c |= elem_class(e, nReg, nRegInt, nHangInt, phased) & 1;
    infam[e] = 1;
  }
  cls[f] = (uint8_t)c;
}
// unit slot table, per-family records {hanging masks x3, level} of the selected families, in family order
__global__ void k_family_gather(const uint64_t *flag, const uint64_t *pos, uint64_t nFam, int dim, const uint32_t *mem, const uint8_t *lev,
                                const uint32_t *Uall, const uint32_t *hm, uint32_t *U, uint32_t *frec)
{
  const int L = fam_L(dim);
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= nFam * L) return;
  const uint64_t f = i / L;
  const int k = (int)(i % L);
  if (!flag[f]) return;
  const uint64_t u = pos[f];
  U[u * L + k] = Uall[i];
  if (k < 3) frec[u * 4 + k] = hm[f * 4 + k];
  if (k == 3) frec[u * 4 + 3] = lev[mem[f << dim]];
}
static int scan_total(DA &da, const uint64_t *flag, uint64_t *pos, uint64_t n, uint64_t &total)
{
  int rc = device_exclusive_scan(da, flag, pos, n + 1);
  if (rc) return rc;
  CK(cudaMemcpy(&total, pos + n, sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return DKT_OK;
}

struct PendingSet
{
  size_t idx;       // index into da.sets
  uint32_t *U;      // unit slot table (temporary)
};

// per-element set over contiguous per-element arrays (the DA's own, or compact copies of the ungrouped elements)
static int add_elem_set(DA &da, std::vector<ChunkSet> &sets, std::vector<PendingSet> &pend, const uint32_t *e2n, const uint32_t *pnode, const uint8_t *lev,
                        const uint8_t *child, const uint32_t *fmask, uint64_t n, int rows, int phase, uint64_t elem0, uint64_t hang0)
{
  if (n == 0) return DKT_OK;
  sets.emplace_back();
  ChunkSet &cs = sets.back();
  cs.rows = rows; cs.phase = phase; cs.elem0 = elem0; cs.hang0 = hang0; cs.nElem = n; cs.kind = 0;
  cs.xorperm = da.order == 1 ? 1 : 0;
  cs.spu = rows * da.N;
  cs.elemsPerChunk = rows_per_chunk(da.N) / rows;
  cs.lev = lev; cs.child = child; cs.fmask = fmask;
  uint32_t *U = nullptr;
  CK(cudaMalloc((void **)&U, n * cs.spu * sizeof(uint32_t)));
  DKT_LAUNCH(k_unit_slots_elem, nblk(n * cs.spu), 256, 0, da.stream)(e2n, pnode, child, n, da.N, rows, cs.xorperm, U);
  g_launches++;
  pend.push_back({sets.size() - 1, U});
  return DKT_OK;
}

// Inside a chunk the order of the families is free: the ones without hanging points go first, so that whole warps of the
// kernel (FPW families each) skip the hanging-node code.  src[u] = unit that moves to position u.
__global__ void k_family_order(const uint32_t *frec, uint64_t n, int upc, uint32_t *src)
{
  const uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t u0 = c * (uint64_t)upc;
  if (u0 >= n) return;
  const int nu = (int)min((uint64_t)upc, n - u0);
  int k = 0;
  for (int pass = 0; pass < 2; pass++)
    for (int i = 0; i < nu; i++)
    {
      const uint32_t *fr = frec + (u0 + i) * 4;
      const int hang = (fr[0] | fr[1] | fr[2]) != 0u;
      if (hang == pass) src[u0 + k++] = (uint32_t)(u0 + i);
    }
}
__global__ void k_family_permute(const uint32_t *src, uint64_t n, int L, const uint32_t *Uin, const uint32_t *frecIn, uint32_t *Uout,
                                 uint32_t *frecOut)
{
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n * L) return;
  const uint64_t u = i / L;
  const int k = (int)(i % L);
  const uint64_t s = src[u];
  Uout[i] = Uin[s * L + k];
  if (k < 4) frecOut[u * 4 + k] = frecIn[s * 4 + k];
}

// family set: units [a, b) of the gathered unit slot table / records of one class
static int add_family_set(DA &da, std::vector<ChunkSet> &sets, std::vector<PendingSet> &pend, const uint32_t *U, const uint32_t *frec, uint64_t n,
                          int phase)
{
  if (n == 0) return DKT_OK;
  sets.emplace_back();
  ChunkSet &cs = sets.back();
  cs.rows = 1; cs.phase = phase; cs.nElem = n; cs.kind = 2; cs.xorperm = 0;
  cs.spu = fam_L(da.dim);
  cs.elemsPerChunk = fam_UPC(da.dim);
  const uint32_t nChunks = (uint32_t)((n + cs.elemsPerChunk - 1) / cs.elemsPerChunk);
  uint32_t *Uc = nullptr, *src = nullptr;
  CK(cudaMalloc((void **)&Uc, n * cs.spu * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&src, n * sizeof(uint32_t)));
  // records padded to whole chunks: the kernel copies them chunk by chunk
  CK(cudaMalloc((void **)&cs.d_frec, (size_t)nChunks * cs.elemsPerChunk * 4 * sizeof(uint32_t)));
  CK(cudaMemsetAsync(cs.d_frec, 0, (size_t)nChunks * cs.elemsPerChunk * 4 * sizeof(uint32_t), da.stream));
  DKT_LAUNCH(k_family_order, nblk(nChunks), 256, 0, da.stream)(frec, n, (int)cs.elemsPerChunk, src);
  DKT_LAUNCH(k_family_permute, nblk(n * cs.spu), 256, 0, da.stream)(src, n, cs.spu, U, frec, Uc, cs.d_frec);
  g_launches += 2;
  CK(cudaStreamSynchronize(da.stream));
  cudaFree(src);
  pend.push_back({sets.size() - 1, Uc});
  return DKT_OK;
}

// per-element set of the elements in `list` (outside complete families): compact copies of their rows
static int add_single_set(DA &da, std::vector<ChunkSet> &sets, std::vector<PendingSet> &pend, const uint32_t *list, uint64_t nS, int hang,
                          int phase)
{
  if (nS == 0) return DKT_OK;
  const int N = da.N;
  uint32_t *e2n_s = nullptr, *pnode_s = nullptr, *fm_s = nullptr;
  uint8_t *lev_s = nullptr, *child_s = nullptr;
  CK(cudaMalloc((void **)&e2n_s, nS * N * sizeof(uint32_t)));
  if (hang) CK(cudaMalloc((void **)&pnode_s, nS * N * sizeof(uint32_t)));
  if (hang) CK(cudaMalloc((void **)&fm_s, nS * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&lev_s, nS));
  CK(cudaMalloc((void **)&child_s, nS));
  DKT_LAUNCH(k_gather_single, nblk(nS), 256, 0, da.stream)(list, nS, N, da.nReg, da.d_e2n, da.d_pnode, da.d_mv_lev, da.d_mv_child, e2n_s,
                                                           pnode_s, lev_s, child_s);
  g_launches++;
  if (hang)
  {
    DKT_LAUNCH(k_fmask, nblk(nS), 256, 0, da.stream)(e2n_s, child_s, nS, N, da.order == 1 ? 1 : 0, fm_s);
    g_launches++;
  }
  int rc = add_elem_set(da, sets, pend, e2n_s, pnode_s, lev_s, child_s, fm_s, nS, hang ? 2 : 1, phase, 0, 0);
  if (rc == DKT_OK)
  {
    ChunkSet &cs = sets.back();
    cs.owned.push_back(lev_s); cs.owned.push_back(child_s);
    if (fm_s) cs.owned.push_back(fm_s);
  }
  CK(cudaStreamSynchronize(da.stream));
  cudaFree(e2n_s);
  cudaFree(pnode_s);
  return rc;
}

// order 1: sibling-family sets unless DKT_FAMILIES=0
static bool families_wanted(const DA &da)
{
  if (da.order != 1) return false;
  const char *e = getenv("DKT_FAMILIES");
  return !(e && atoi(e) == 0);
}

// writing references of every node over all sets of a collection, then the chunk tables
static int finish_sets(DA &da, std::vector<ChunkSet> &sets, std::vector<PendingSet> &pend, int rc)
{
  uint32_t *refcnt = nullptr;
  CK(cudaMalloc((void **)&refcnt, std::max<uint64_t>(da.nNodes, 1) * sizeof(uint32_t)));
  CK(cudaMemsetAsync(refcnt, 0, std::max<uint64_t>(da.nNodes, 1) * sizeof(uint32_t), da.stream));
  for (const PendingSet &ps : pend)
  {
    const ChunkSet &cs = sets[ps.idx];
    const uint64_t ns = cs.nElem * cs.spu;
    if (rc == DKT_OK && ns)
    {
      DKT_LAUNCH(k_ref_count, nblk(ns), 256, 0, da.stream)(ps.U, ns, refcnt);
      g_launches++;
    }
  }
  for (const PendingSet &ps : pend)
    if (rc == DKT_OK) rc = build_set(da, sets[ps.idx], ps.U, refcnt);
  cudaStreamSynchronize(da.stream);
  for (const PendingSet &ps : pend) cudaFree(ps.U);
  cudaFree(refcnt);
  return rc;
}

// per-element sets of all visited elements: one regular + one hanging set, or (partitioned) two phases of each:
// interior (phase 0) / boundary = touches a ghost node (phase 1) - see run_matvec_dist
static int build_elem_sets(DA &da, std::vector<ChunkSet> &sets)
{
  const int N = da.N;
  std::vector<PendingSet> pend;
  int rc = DKT_OK;
  struct Range { uint64_t a, b; int phase; };
  std::vector<Range> rr, hr;
  if (da.phased)
  {
    const uint64_t ri = da.nRegInterior, hi = da.nHangInterior;
    rr = {{0, ri, 0}, {ri, da.nReg, 1}};
    hr = {{0, hi, 0}, {hi, da.nHang, 1}};
  }
  else
  {
    rr = {{0, da.nReg, 0}};
    hr = {{0, da.nHang, 0}};
  }
  for (const Range &r : rr)
    if (rc == DKT_OK && r.b > r.a)
      rc = add_elem_set(da, sets, pend, da.d_e2n + r.a * N, nullptr, da.d_mv_lev + r.a, da.d_mv_child + r.a, nullptr, r.b - r.a, 1, r.phase, r.a, 0);
  for (const Range &r : hr)
    if (rc == DKT_OK && r.b > r.a)
      rc = add_elem_set(da, sets, pend, da.d_e2n + (da.nReg + r.a) * N, da.d_pnode + r.a * N, da.d_mv_lev + da.nReg + r.a,
                        da.d_mv_child + da.nReg + r.a, da.d_fmask + r.a, r.b - r.a, 2, r.phase, da.nReg + r.a, r.a);
  return finish_sets(da, sets, pend, rc);
}

// family sets of the admitted complete families + per-element sets of everything else
static int build_family_sets(DA &da, std::vector<ChunkSet> &sets)
{
  const uint64_t n = da.nMv;
  const int nch = 1 << da.dim, L = fam_L(da.dim);
  const int phased = da.phased ? 1 : 0;
  std::vector<PendingSet> pend;
  int rc = DKT_OK;
  uint32_t *inv = nullptr, *mem = nullptr, *Uall = nullptr, *hm = nullptr, *bad = nullptr;
  uint64_t *flag = nullptr, *pos = nullptr;
  uint8_t *infam = nullptr, *cls = nullptr, *fcls = nullptr;
  CK(cudaMalloc((void **)&inv, std::max<uint64_t>(n, 1) * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&flag, (n + 1) * sizeof(uint64_t)));
  CK(cudaMalloc((void **)&pos, (n + 1) * sizeof(uint64_t)));
  CK(cudaMalloc((void **)&infam, std::max<uint64_t>(n, 1)));
  CK(cudaMalloc((void **)&cls, std::max<uint64_t>(n, 1)));
  CK(cudaMemsetAsync(infam, 0, std::max<uint64_t>(n, 1), da.stream));
  DKT_LAUNCH(k_invert_src, nblk(n), 256, 0, da.stream)(da.d_mv_src, n, da.mv_src0, inv);
  DKT_LAUNCH(k_family_heads, nblk(n + 1), 256, 0, da.stream)(inv, da.d_mv_xyz, da.d_mv_lev, n, da.dim, da.max_depth, flag);
  g_launches += 2;
  uint64_t nFam = 0;
  rc = scan_total(da, flag, pos, n, nFam);
  if (rc) return rc;
  std::vector<void *> temps;  // freed after the unit slot tables are built
  if (nFam)
  {
    CK(cudaMalloc((void **)&mem, nFam * nch * sizeof(uint32_t)));
    CK(cudaMalloc((void **)&Uall, nFam * L * sizeof(uint32_t)));
    CK(cudaMalloc((void **)&hm, nFam * 4 * sizeof(uint32_t)));
    CK(cudaMalloc((void **)&bad, nFam * sizeof(uint32_t)));
    CK(cudaMalloc((void **)&fcls, nFam));
    CK(cudaMemsetAsync(hm, 0, nFam * 4 * sizeof(uint32_t), da.stream));
    CK(cudaMemsetAsync(bad, 0, nFam * sizeof(uint32_t), da.stream));
    DKT_LAUNCH(k_family_members, nblk(n), 256, 0, da.stream)(inv, flag, pos, da.d_mv_child, n, da.dim, mem);
    DKT_LAUNCH(k_family_units, nblk(nFam * L), 256, 0, da.stream)(mem, nFam, da.dim, da.d_e2n, Uall, hm, bad);
    DKT_LAUNCH(k_family_check, nblk(nFam * L), 256, 0, da.stream)(mem, nFam, da.dim, da.nReg, da.d_pnode, Uall, hm, bad);
    DKT_LAUNCH(k_family_class, nblk(nFam), 256, 0, da.stream)(mem, bad, nFam, da.dim, da.nReg, da.nRegInterior, da.nHangInterior, phased, fcls,
                                                              infam);
    g_launches += 4;
    uint64_t *fflag = nullptr, *fpos = nullptr;
    CK(cudaMalloc((void **)&fflag, (nFam + 1) * sizeof(uint64_t)));
    CK(cudaMalloc((void **)&fpos, (nFam + 1) * sizeof(uint64_t)));
    for (int c = 0; c < 2 && rc == DKT_OK; c++)
    {
      DKT_LAUNCH(k_class_flags, nblk(nFam + 1), 256, 0, da.stream)(fcls, nFam, c, fflag);
      g_launches++;
      uint64_t cnt = 0;
      rc = scan_total(da, fflag, fpos, nFam, cnt);
      if (rc || !cnt) continue;
      uint32_t *Uc = nullptr, *frec = nullptr;
      CK(cudaMalloc((void **)&Uc, cnt * L * sizeof(uint32_t)));
      CK(cudaMalloc((void **)&frec, cnt * 4 * sizeof(uint32_t)));
      temps.push_back(Uc);
      temps.push_back(frec);
      DKT_LAUNCH(k_family_gather, nblk(nFam * L), 256, 0, da.stream)(fflag, fpos, nFam, da.dim, mem, da.d_mv_lev, Uall, hm, Uc, frec);
      g_launches++;
      // partitioned: the interior families (phase 0) run beside the exchange + boundary families (phase 1), see run_matvec_dist
      if (rc == DKT_OK) rc = add_family_set(da, sets, pend, Uc, frec, cnt, (phased && c == 1) ? 1 : 0);
    }
    CK(cudaStreamSynchronize(da.stream));
    cudaFree(fflag);
    cudaFree(fpos);
  }
  // the elements outside admitted families, by class: regular / hanging x interior / boundary
  DKT_LAUNCH(k_single_class, nblk(n), 256, 0, da.stream)(infam, n, da.nReg, da.nRegInterior, da.nHangInterior, phased, cls);
  g_launches++;
  for (int c = 0; c < 4 && rc == DKT_OK; c++)
  {
    DKT_LAUNCH(k_class_flags, nblk(n + 1), 256, 0, da.stream)(cls, n, c, flag);
    g_launches++;
    uint64_t cnt = 0;
    rc = scan_total(da, flag, pos, n, cnt);
    if (rc || !cnt) continue;
    uint32_t *list = nullptr;
    CK(cudaMalloc((void **)&list, cnt * sizeof(uint32_t)));
    temps.push_back(list);
    DKT_LAUNCH(k_compact, nblk(n), 256, 0, da.stream)(flag, pos, n, list);
    g_launches++;
    const int hang = (c >> 1) & 1, bdy = c & 1;
    if (rc == DKT_OK) rc = add_single_set(da, sets, pend, list, cnt, hang, (phased && bdy) ? 1 : 0);
  }
  CK(cudaStreamSynchronize(da.stream));
  for (void *t : temps) cudaFree(t);
  cudaFree(inv); cudaFree(mem); cudaFree(Uall); cudaFree(hm); cudaFree(bad); cudaFree(flag); cudaFree(pos); cudaFree(infam); cudaFree(cls);
  cudaFree(fcls);
  return finish_sets(da, sets, pend, rc);
}

int build_chunks(DA &da)
{
  // more than 27 nodes per element (4-D order 2): no shared-memory kernel; every matvec runs on the flat kernels
  if (da.N > MAX_NPE) return DKT_OK;
  if (da.nNodes >= 0x3FFFFFFFull) { set_error("more than 2^30 nodes on one rank"); return DKT_ERR_UNSUPPORTED; }
  CK(cudaMalloc((void **)&da.d_mv_child, std::max<uint64_t>(da.nMv, 1)));
  DKT_LAUNCH(k_child_numbers, nblk(da.nMv), 256, 0, da.stream)(da.d_mv_xyz, da.d_mv_lev, da.nMv, da.dim, da.max_depth, da.d_mv_child);
  g_launches++;
  if (da.nHang)
  {
    CK(cudaMalloc((void **)&da.d_fmask, da.nHang * sizeof(uint32_t)));
    DKT_LAUNCH(k_fmask, nblk(da.nHang), 256, 0, da.stream)(da.d_e2n + da.nReg * (uint64_t)da.N, da.d_mv_child + da.nReg, da.nHang, da.N,
                                                           da.order == 1 ? 1 : 0, da.d_fmask);
    g_launches++;
  }
  da.families = families_wanted(da);
  int rc = da.families ? build_family_sets(da, da.sets) : build_elem_sets(da, da.sets);
  if (rc) return rc;
  int dev = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&da.numSMs, cudaDevAttrMultiProcessorCount, dev));
  // the sets of one matvec are independent (nodes shared between sets are accumulated with RED): the small per-element sets
  // of a family DA run beside the family kernel on their own streams
  da.mvStreams = da.families ? 3 : 1;
  if (const char *e = getenv("DKT_MV_STREAMS")) da.mvStreams = std::max(1, std::min(atoi(e), DA::MAX_AUX + 1));
  if (da.mvStreams > 1 && !da.ev_fork)
  {
    CK(cudaEventCreateWithFlags(&da.ev_fork, cudaEventDisableTiming));
    for (int a = 0; a < da.mvStreams - 1; a++)
    {
      CK(cudaStreamCreateWithFlags(&da.aux[a], cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&da.ev_join[a], cudaEventDisableTiming));
    }
  }
  return DKT_OK;
}
// the per-element tables of a DA that runs on family sets, for the operators k_mvf does not serve
static int ensure_elem_sets(DA &da)
{
  if (!da.families || !da.sets_elem.empty() || da.nMv == 0) return DKT_OK;
  return build_elem_sets(da, da.sets_elem);
}
// ------------------------------------------------------------------------------------------
// matvec kernels
// ------------------------------------------------------------------------------------------
template <int DIM, int ORDER>
struct Mv3Params
{
  static constexpr int M = ORDER + 1;
  static constexpr int N = (DIM == 2 ? M * M : DIM == 3 ? M * M * M : M * M * M * M);
  const double *in;
  double *out;
  const uint32_t *slot;
  const uint32_t *gid;
  const uint16_t *meta;
  const uint16_t *jd;
  const uint64_t *node_off;
  const uint8_t *lev;    // level of the set's elements
  const uint8_t *child;  // Morton child numbers of the set's elements
  const uint32_t *fmask; // hanging set: filled own slots (slot order)
  uint32_t nSet, nChunks, elemsPerChunk, xcap, ncap, jdStride;
  int q1mask;
  int exact_ip;          // order 1: ip0/ip1 equal the exact interpolation to 1e-13
  double lscale[32];
  double ip[2][M * M];
  double ipx[2][M * M];  // ip[0] and J ip[1] J (XOR-permuted coordinates)
  double K[N * N];
  double kron[DKT_KRON_MAX_TERMS][DIM][M * M];  // DKT_OP_KRON: the 1-D factors
  int nterms;
};

template <int DIM, int M, int AXIS, bool TRANSPOSE>
__device__ __forceinline__ void axis_pass3(const double *A, double *v)
{
  constexpr int N = (DIM == 2 ? M * M : DIM == 3 ? M * M * M : M * M * M * M);
  int stride = 1;
#pragma unroll
  for (int d = 0; d < AXIS; d++) stride *= M;
#pragma unroll
  for (int base = 0; base < N; base++)
  {
    if ((base / stride) % M != 0) continue;
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
__device__ __forceinline__ void tensor_interp3(const double (&ip)[2][M * M], int child, double *v)
{
  double A[M * M];
#pragma unroll
  for (int i = 0; i < M * M; i++) A[i] = (child & 1) ? ip[1][i] : ip[0][i];
  axis_pass3<DIM, M, 0, TRANSPOSE>(A, v);
#pragma unroll
  for (int i = 0; i < M * M; i++) A[i] = (child & 2) ? ip[1][i] : ip[0][i];
  axis_pass3<DIM, M, 1, TRANSPOSE>(A, v);
  if (DIM >= 3)
  {
#pragma unroll
    for (int i = 0; i < M * M; i++) A[i] = (child & 4) ? ip[1][i] : ip[0][i];
    axis_pass3<DIM, M, (DIM >= 3 ? 2 : 0), TRANSPOSE>(A, v);
  }
  if (DIM >= 4)
  {
#pragma unroll
    for (int i = 0; i < M * M; i++) A[i] = (child & 8) ? ip[1][i] : ip[0][i];
    axis_pass3<DIM, M, (DIM >= 4 ? 3 : 0), TRANSPOSE>(A, v);
  }
}


// Exact order-1 parent->child interpolation in XOR-permuted coordinates: every axis uses
// A0 = [[1, 1/2], [0, 1/2]] (input k -> output j), i.e. child'[s] = 2^-|s| * sum_{t subset of s} parent'[t]
// (a subset-sum transform), and its transpose parent'[t] += sum_{s superset of t} 2^-|s| child'[s].
template <int N>
__device__ __forceinline__ void interp_exact(double *v)
{
#pragma unroll
  for (int b = 1; b < N; b <<= 1)
  {
#pragma unroll
    for (int i = 0; i < N; i++)
    {
      if (i & b) continue;
      v[i | b] = 0.5 * (v[i] + v[i | b]);
    }
  }
}
template <int N>
__device__ __forceinline__ void interp_exact_T(double *v)
{
#pragma unroll
  for (int b = 1; b < N; b <<= 1)
  {
#pragma unroll
    for (int i = 0; i < N; i++)
    {
      if (i & b) continue;
      const double h = 0.5 * v[i | b];
      v[i] += h;
      v[i | b] = h;
    }
  }
}

// eout = K_e ein  (ein is clobbered in the Hadamard form)
template <int DIM, int ORDER, int OPKIND>
__device__ __forceinline__ void apply_op3(const Mv3Params<DIM, ORDER> &p, int lev, double *ein, double *eout)
{
  constexpr int N = Mv3Params<DIM, ORDER>::N;
  constexpr int M = ORDER + 1;
  if (OPKIND == DKT_OP_IDENTITY)
  {
#pragma unroll
    for (int i = 0; i < N; i++) eout[i] = ein[i];
  }
  else if (OPKIND == OP_HADAMARD)
  {
    const double s = p.lscale[lev];
    wht<N>(ein);
#pragma unroll
    for (int i = 0; i < N; i++) eout[i] = ein[i] * (p.K[i] * s);
    wht<N>(eout);
  }
  else if (OPKIND == DKT_OP_KRON)
  {
    // sum-factorised: every term is DIM axis passes with an M x M matrix (3-FMA dependency chains, no N x N table)
    const double s = p.lscale[lev];
#pragma unroll
    for (int i = 0; i < N; i++) eout[i] = 0.0;
#pragma unroll 1
    for (int t = 0; t < p.nterms; t++)
    {
      double tmp[N];
#pragma unroll
      for (int i = 0; i < N; i++) tmp[i] = ein[i];
      {
        double A[M * M];
#pragma unroll
        for (int i = 0; i < M * M; i++) A[i] = p.kron[t][0][i];
        axis_pass3<DIM, M, 0, false>(A, tmp);
#pragma unroll
        for (int i = 0; i < M * M; i++) A[i] = p.kron[t][1][i];
        axis_pass3<DIM, M, 1, false>(A, tmp);
        if (DIM >= 3)
        {
#pragma unroll
          for (int i = 0; i < M * M; i++) A[i] = p.kron[t][DIM >= 3 ? 2 : 0][i];
          axis_pass3<DIM, M, (DIM >= 3 ? 2 : 0), false>(A, tmp);
        }
        if (DIM >= 4)
        {
#pragma unroll
          for (int i = 0; i < M * M; i++) A[i] = p.kron[t][DIM >= 4 ? 3 : 0][i];
          axis_pass3<DIM, M, (DIM >= 4 ? 3 : 0), false>(A, tmp);
        }
      }
#pragma unroll
      for (int i = 0; i < N; i++) eout[i] += tmp[i];
    }
#pragma unroll
    for (int i = 0; i < N; i++) eout[i] *= s;
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

// undo the XOR slot schedule in registers: v[r] <- v[r ^ c]
template <int DIM, int N, typename T>
__device__ __forceinline__ void xor_unpermute(T *v, int c)
{
#pragma unroll
  for (int d = 0; d < DIM; d++)
  {
    const bool f = (c >> d) & 1;
#pragma unroll
    for (int i = 0; i < N; i++)
    {
      if (i & (1 << d)) continue;
      const T a = v[i], b = v[i | (1 << d)];
      v[i] = f ? b : a;
      v[i | (1 << d)] = f ? a : b;
    }
  }
}

// TPB threads, one element per thread, NPT nodes per thread (chunk nodes <= NPT*TPB).
// Order 1: slot s of an element with child number c holds rank s ^ c.  The identity and
// Walsh-Hadamard operators commute with that permutation (H D H is a convolution on Z_2^dim) and
// the interpolation becomes child-independent up to the J-conjugated matrix ipx, so those paths
// work in permuted coordinates throughout; the dense path un-permutes the slot words first.
template <int DIM, int ORDER, int OPKIND, bool DIRI, bool HANG, int TPB, int NPT, bool EXIP>
__global__ void __launch_bounds__(TPB, ((Mv3Params<DIM, ORDER>::N <= 16 && OPKIND != DKT_OP_DENSE && OPKIND != DKT_OP_KRON) ? (HANG ? DKT_HANG_MINB : DKT_REG_MINB) * (256 / DKT_ROWS) : ((DKT_MV3_LEAN && Mv3Params<DIM, ORDER>::N > 16) ? 3 : 2))) k_mv3(const __grid_constant__ Mv3Params<DIM, ORDER> p)
{
  constexpr int N = Mv3Params<DIM, ORDER>::N;
  // Order 2 (27 nodes per element): the element phase alone needs ~110 registers, so the node records of the current and the
  // next chunk (2 x NPT ids + flags per thread) are NOT kept in registers: the gather and the node phase read them from
  // global memory where they are needed (the second read hits L2).  3 CTAs per SM instead of 2.
  constexpr bool LEAN = DKT_MV3_LEAN && N > 16;
  constexpr int M = ORDER + 1;
  constexpr int ROWS = HANG ? 2 : 1;
  DKT_DYN_SMEM(double, sm);
  double *X = sm;                                   // [xcap]
  double *unb = sm + p.xcap;                        // [2][ncap]
  int *jdb = (int *)(sm + p.xcap + 2 * p.ncap);     // [2][jdStride]

  const int tid = threadIdx.x;
  uint64_t c = blockIdx.x;
  if (c >= p.nChunks) return;
  const uint32_t E = p.elemsPerChunk;

  uint32_t gidC[LEAN ? 1 : NPT], gidN[LEAN ? 1 : NPT];
  uint16_t metaC[LEAN ? 1 : NPT], metaN[LEAN ? 1 : NPT];
  auto load_nodes = [&](uint64_t oa, uint64_t ob, uint32_t *g, uint16_t *m) {
    if (LEAN) return;
    const int nloc = (int)(ob - oa);
#pragma unroll
    for (int k = 0; k < NPT; k++)
    {
      const int n = tid + k * TPB;
      g[k] = 0;
      m[k] = 0;
      if (n < nloc)
      {
        g[k] = p.gid[oa + n];
        m[k] = p.meta[oa + n];
      }
    }
  };
  auto issue_gather_lean = [&](double *un, uint64_t oa, int nloc) {
    if (tid == 0) un[nloc] = 0.0;
#pragma unroll 4
    for (int n = tid; n < nloc; n += TPB)
    {
      const uint16_t m = p.meta[oa + n];
      if (!(m & META_PRESENT)) continue;
      if (DIRI && (m & META_BDY)) un[n] = 0.0;
      else cp_async8(un + n, p.in + p.gid[oa + n]);
    }
  };
  auto issue_gather = [&](double *un, const uint32_t *g, const uint16_t *m, int nloc) {
    if (tid == 0) un[nloc] = 0.0;  // the entry absent nodes read
#pragma unroll
    for (int k = 0; k < NPT; k++)
    {
      const int n = tid + k * TPB;
      if (!(m[k] & META_PRESENT)) continue;  // no such node
      if (DIRI && (m[k] & META_BDY)) un[n] = 0.0;
      else cp_async8(un + n, p.in + g[k]);
    }
  };
  uint32_t w[ROWS * N];
  uint32_t fm = 0;
  int lev = 0, child = 0;
  auto load_slots = [&](uint64_t cc) {
    const uint64_t e0 = cc * (uint64_t)E;
    const int ne = (int)min((uint64_t)E, (uint64_t)p.nSet - e0);
    if (tid < ne)
    {
      const uint32_t *sw = p.slot + e0 * (ROWS * N) + tid;
#pragma unroll
      for (int r = 0; r < ROWS * N; r++) w[r] = sw[(uint32_t)r * E];
      lev = p.lev[e0 + tid];
      if (HANG || (ORDER == 1 && OPKIND == DKT_OP_DENSE)) child = p.child[e0 + tid];
      if (HANG) fm = p.fmask[e0 + tid];
      if (ORDER == 1 && OPKIND == DKT_OP_DENSE)
      {  // natural rank order for the dense product
        xor_unpermute<DIM, N, uint32_t>(w, child);
        if (HANG)
        {
          xor_unpermute<DIM, N, uint32_t>(w + N, child);
          uint32_t nat = 0;
#pragma unroll
          for (int r = 0; r < N; r++) nat |= ((fm >> (r ^ child)) & 1u) << r;
          fm = nat;
        }
      }
    }
  };
  constexpr bool PERM = (ORDER == 1 && OPKIND != DKT_OP_DENSE);

  // ---- prologue: everything for the first chunk ------------------------------------------------
  uint64_t offA = p.node_off[c], offB = p.node_off[c + 1];
  load_nodes(offA, offB, gidC, metaC);
  if (LEAN) issue_gather_lean(unb, offA, (int)(offB - offA));
  else issue_gather(unb, gidC, metaC, (int)(offB - offA));
  cp_async_commit();
  for (int k = tid; k < (int)p.jdStride; k += TPB) jdb[k] = p.jd[c * (uint64_t)p.jdStride + k];
  load_slots(c);
  uint64_t cn = c + gridDim.x;
  bool hasN = cn < p.nChunks;
  uint64_t offNA = 0, offNB = 0;
  if (hasN) { offNA = p.node_off[cn]; offNB = p.node_off[cn + 1]; }
  int buf = 0;

  while (true)
  {
    double *un = unb + buf * p.ncap;
    const int *jd = jdb + buf * p.jdStride;
    cp_async_wait_all();
    __syncthreads();  // T0: un/jd of this chunk visible; X free again
    // T1: next chunk's node records -> registers; node_off of the chunk after it
    const uint64_t cnn = cn + gridDim.x;
    uint64_t offNNA = 0, offNNB = 0;
    if (hasN)
    {
      load_nodes(offNA, offNB, gidN, metaN);
      if (cnn < p.nChunks) { offNNA = p.node_off[cnn]; offNNB = p.node_off[cnn + 1]; }
    }
    // T2: elements of this chunk
    {
      const uint64_t e0 = c * (uint64_t)E;
      const int ne = (int)min((uint64_t)E, (uint64_t)p.nSet - e0);
      if (tid < ne)
      {
        if (!HANG)
        {
          double ein[N];
#pragma unroll
          for (int r = 0; r < N; r++) ein[r] = un[w[r] & 0xFFFFu];
          if (OPKIND == DKT_OP_IDENTITY)
          {
#pragma unroll
            for (int r = 0; r < N; r++) X[w[r] >> 16] = ein[r];
          }
          else if (OPKIND == OP_HADAMARD)
          {
            const double s = p.lscale[lev];
            wht<N>(ein);
#pragma unroll
            for (int i = 0; i < N; i++) ein[i] *= p.K[i] * s;
            wht<N>(ein);
#pragma unroll
            for (int r = 0; r < N; r++) X[w[r] >> 16] = ein[r];
          }
          else if (OPKIND == DKT_OP_KRON)
          {
            double eout[N];
            apply_op3<DIM, ORDER, OPKIND>(p, lev, ein, eout);
#pragma unroll
            for (int r = 0; r < N; r++) X[w[r] >> 16] = eout[r];
          }
          else
          {
            // each output row is stored as soon as it is complete: no eout[] array stays live
            const double s = p.lscale[lev];
#pragma unroll
            for (int i = 0; i < N; i++)
            {
              double acc = 0.0;
#pragma unroll
              for (int j = 0; j < N; j++) acc = fma(p.K[i * N + j], ein[j], acc);
              X[w[i] >> 16] = s * acc;
            }
          }
        }
        else
        {
          // absent nodes read un[nloc] == 0 and write to a trash position, read-only parent slots (static
          // Q1, matvec.h:517) write to trash too: no predicates on the gathers and scatters
          double ein[N], eout[N], par[N];
#pragma unroll
          for (int r = 0; r < N; r++) par[r] = un[w[N + r] & 0xFFFFu];
          if (EXIP && PERM) interp_exact<N>(par);
          else tensor_interp3<DIM, M, false>(PERM ? p.ipx : p.ip, child, par);
#pragma unroll
          for (int r = 0; r < N; r++)
          {
            const double own = un[w[r] & 0xFFFFu];
            ein[r] = ((fm >> r) & 1u) ? own : par[r];
          }
          apply_op3<DIM, ORDER, OPKIND>(p, lev, ein, eout);
#pragma unroll
          for (int r = 0; r < N; r++)
          {
            X[w[r] >> 16] = eout[r];
            if ((fm >> r) & 1u) eout[r] = 0.0;  // nullify prior to back-interpolation (matvec.h:497-499)
          }
          if (EXIP && PERM) interp_exact_T<N>(eout);
          else tensor_interp3<DIM, M, true>(PERM ? p.ipx : p.ip, child, eout);
#pragma unroll
          for (int q = 0; q < N; q++) X[w[N + q] >> 16] = eout[q];
        }
      }
    }
    // T3: start the next chunk's gather and index loads; they land during T4
    if (hasN)
    {
      if (LEAN) issue_gather_lean(unb + (buf ^ 1) * p.ncap, offNA, (int)(offNB - offNA));
      else issue_gather(unb + (buf ^ 1) * p.ncap, gidN, metaN, (int)(offNB - offNA));
      int *jdn = jdb + (buf ^ 1) * p.jdStride;
      for (int k = tid; k < (int)p.jdStride; k += TPB) jdn[k] = p.jd[cn * (uint64_t)p.jdStride + k];
      load_slots(cn);
    }
    cp_async_commit();
    __syncthreads();  // T4: X complete
    if (LEAN)
    {
      const int nlocC = (int)(offB - offA);
#pragma unroll 2
      for (int n = tid; n < nlocC; n += TPB)
      {
        const uint32_t m = p.meta[offA + n];
        const int len = m & META_LEN;
        if (len == 0) continue;  // absent, or only read by this chunk
        double acc = X[n];  // jd[0] == 0
        for (int j = 1; j < len; j++) acc += X[jd[j] + n];
        if (DIRI && (m & META_BDY)) continue;
        const uint32_t g = p.gid[offA + n];
        if (m & META_SHARED) atomicAdd(p.out + g, acc);
        else p.out[g] = acc;
      }
    }
    else
#pragma unroll
    for (int k = 0; k < NPT; k++)
    {
      const int n = tid + k * TPB;
      const int len = metaC[k] & META_LEN;
#ifdef DKT_PHASED_UNIFORM
      // warp-uniform trip count (nodes are sorted by run length, so lanes differ by little)
      const int lmax = __reduce_max_sync(0xffffffffu, len);
      double acc = 0.0;
      for (int j = 0; j < lmax; j++)
        if (j < len) acc += X[jd[j] + n];
      if (len == 0) continue;
#else
      if (len == 0) continue;  // absent, or only read by this chunk
      double acc = X[n];  // jd[0] == 0
      for (int j = 1; j < len; j++) acc += X[jd[j] + n];
#endif
      if (DIRI && (metaC[k] & META_BDY)) continue;
      if (metaC[k] & META_SHARED) atomicAdd(p.out + gidC[k], acc);
      else p.out[gidC[k]] = acc;
    }
    if (!hasN) break;
    c = cn;
    cn = cnn;
    hasN = cn < p.nChunks;
    if (!LEAN)
    {
#pragma unroll
      for (int k = 0; k < NPT; k++) { gidC[k] = gidN[k]; metaC[k] = metaN[k]; }
    }
    offA = offNA;
    offB = offNB;
    offNA = offNNA;
    offNB = offNNB;
    buf ^= 1;
  }
}

template <int DIM, int ORDER, int OPKIND, bool DIRI, bool HANG, int TPB, int NPT, bool EXIP>
static int launch_one(DA &da, const ChunkSet &cs, Mv3Params<DIM, ORDER> &p)
{
  constexpr int N = Mv3Params<DIM, ORDER>::N;
  p.slot = cs.d_slot; p.gid = cs.d_gid; p.meta = cs.d_meta; p.jd = cs.d_jd; p.node_off = cs.d_node_off;
  p.lev = cs.lev; p.child = cs.child; p.fmask = cs.fmask; p.nSet = (uint32_t)cs.nElem; p.nChunks = cs.nChunks; p.elemsPerChunk = cs.elemsPerChunk;
  p.xcap = (uint32_t)rows_per_chunk(N) * N + 258u;  // + padding of the first 16 diagonals + the trash position
  p.ncap = (cs.maxNloc + 2) & ~1u;
  p.jdStride = cs.jdStride;
  if (cs.maxNloc > (uint32_t)(NPT * TPB)) { set_error("internal: more nodes in a chunk than its kernel handles"); return DKT_ERR_UNSUPPORTED; }
  const size_t smem = ((size_t)p.xcap + 2 * (size_t)p.ncap) * sizeof(double) + 2 * (size_t)p.jdStride * sizeof(int);
  auto kern = k_mv3<DIM, ORDER, OPKIND, DIRI, HANG, TPB, NPT, EXIP>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int perSM = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, TPB, smem));
  if (perSM < 1) { set_error("chunk kernel does not fit on an SM"); return DKT_ERR_CUDA; }
  // Persistent CTAs normally fill every SM.  While a ghost exchange is in flight (partitioned DA, interior
  // phases) a few SMs are left free, otherwise the NCCL kernels could not start before this one ends and
  // nothing would overlap.
  int sms = da.numSMs;
  if (da.phased && cs.phase != 1) sms = std::max(1, da.numSMs - da.commSMs);
  const uint32_t grid = std::min<uint32_t>(cs.nChunks, (uint32_t)(perSM * sms));
  DKT_LAUNCH(kern, grid, TPB, smem, da.cur ? da.cur : da.stream)(p);
  g_launches++;
  return DKT_OK;
}
template <int DIM, int ORDER, int OPKIND, bool DIRI>
static int launch_mv3(DA &da, const std::vector<ChunkSet> &sets, Mv3Params<DIM, ORDER> &p, unsigned phaseMask)
{
  constexpr int N = Mv3Params<DIM, ORDER>::N;
  constexpr int TPB_R = (N == 27) ? 160 : DKT_ROWS;  // >= elements per chunk (one element per thread)
  constexpr int TPB_H = (N == 27) ? 96 : DKT_ROWS / 2;
  constexpr bool CAN_EXIP = (ORDER == 1 && OPKIND != DKT_OP_DENSE);
  // nodes per thread of the fall-back instantiations: enough for a chunk whose slots all touch different nodes
  // (scattered boundary elements of a partitioned DA at order 2: up to 150 x 27 nodes)
  constexpr int NPT_R = (SLOT_CAP + TPB_R - 1) / TPB_R, NPT_H = (SLOT_CAP + TPB_H - 1) / TPB_H;
  const bool exip = CAN_EXIP && p.exact_ip;
  // DKT_MV_STREAMS=n (opt-in): the sets of one call are independent (nodes shared between sets are accumulated with
  // RED), so they may run side by side on n streams - small sets then fill the tails of the big ones
  const int ns = std::min(da.mvStreams, DA::MAX_AUX + 1);
  int k = 0, used = 0;
  for (const ChunkSet &cs : sets)
  {
    if (!cs.nChunks || !((phaseMask >> cs.phase) & 1u)) continue;
    int rc = DKT_OK;
    da.cur = nullptr;
    if (ns > 1)
    {
      if (k == 0) CK(cudaEventRecord(da.ev_fork, da.stream));
      const int a = k % ns;
      if (a > 0)
      {
        da.cur = da.aux[a - 1];
        if (!((used >> a) & 1)) CK(cudaStreamWaitEvent(da.cur, da.ev_fork, 0));
        used |= 1 << a;
      }
      k++;
    }
    if (cs.kind == 2)
    {
      if constexpr (ORDER == 1 && (OPKIND == DKT_OP_IDENTITY || OPKIND == OP_HADAMARD))
        rc = launch_family_set(da, cs, OPKIND, DIRI, p.in, p.out, p.lscale, p.K);
      else
      {
        set_error("internal: no sibling-family kernel for this (order, operator)");
        rc = DKT_ERR_UNSUPPORTED;
      }
    }
    else if (cs.rows == 1)
    {
      if (cs.maxNloc <= 6u * TPB_R) rc = launch_one<DIM, ORDER, OPKIND, DIRI, false, TPB_R, 6, false>(da, cs, p);
      else rc = launch_one<DIM, ORDER, OPKIND, DIRI, false, TPB_R, NPT_R, false>(da, cs, p);
    }
    else if (cs.maxNloc <= 8u * TPB_H)
    {
      if (exip) rc = launch_one<DIM, ORDER, OPKIND, DIRI, true, TPB_H, 8, CAN_EXIP>(da, cs, p);
      else rc = launch_one<DIM, ORDER, OPKIND, DIRI, true, TPB_H, 8, false>(da, cs, p);
    }
    else rc = launch_one<DIM, ORDER, OPKIND, DIRI, true, TPB_H, NPT_H, false>(da, cs, p);
    da.cur = nullptr;
    if (rc) return rc;
  }
  for (int a = 1; a < ns; a++)
    if ((used >> a) & 1)
    {
      CK(cudaEventRecord(da.ev_join[a - 1], da.aux[a - 1]));
      CK(cudaStreamWaitEvent(da.stream, da.ev_join[a - 1], 0));
    }
  CK(cudaGetLastError());
  return DKT_OK;
}

template <int DIM, int ORDER>
static int run_typed3(DA &da, const dkt_op *op, const double *d_in, double *d_out, double scale, unsigned flags, unsigned phaseMask,
                      bool zeroOut)
{
  using P = Mv3Params<DIM, ORDER>;
  static thread_local P p;  // large; per host thread, so that independent DAs may be driven from different threads
  p.in = d_in;
  p.out = d_out;
  p.q1mask = (flags & DKT_NO_Q1_MASK) ? 0 : 1;
  for (int l = 0; l < 32; l++) p.lscale[l] = scale * std::pow(2.0, -op->alpha * l);
  for (int b = 0; b < 2; b++)
    for (int i = 0; i < P::M * P::M; i++) p.ip[b][i] = da.ip[b][i];
  for (int k = 0; k < P::M; k++)
    for (int j = 0; j < P::M; j++)
    {
      p.ipx[0][k * P::M + j] = da.ip[0][k * P::M + j];
      p.ipx[1][k * P::M + j] = da.ip[1][(P::M - 1 - k) * P::M + (P::M - 1 - j)];
    }
  p.exact_ip = 0;
  if (ORDER == 1 && !(flags & DKT_MV_NO_FASTPATH))
  {
    const double e0[4] = {1.0, 0.5, 0.0, 0.5}, e1[4] = {0.5, 0.0, 0.5, 1.0};
    double dev = 0.0;
    for (int i = 0; i < 4; i++) dev = std::max(dev, std::max(std::fabs(da.ip[0][i] - e0[i]), std::fabs(da.ip[1][i] - e1[i])));
    p.exact_ip = dev <= 1e-13;
  }
  bool hadamard = false;
  if (op->kind == DKT_OP_KRON)
  {
    if (!op->kref || op->terms < 1 || op->terms > DKT_KRON_MAX_TERMS) { set_error("DKT_OP_KRON needs 1..5 terms in kref"); return DKT_ERR_INVALID; }
    if (ORDER == 1)
    {
      // order 1: the dense matrix is tiny and opens the Walsh-Hadamard form / the family kernel
      static thread_local std::vector<double> Kd;
      kron_to_dense(op, DIM, P::M, Kd);
      dkt_op dense = *op;
      dense.kind = DKT_OP_DENSE;
      dense.kref = Kd.data();
      return run_typed3<DIM, ORDER>(da, &dense, d_in, d_out, scale, flags, phaseMask, zeroOut);
    }
    p.nterms = op->terms;
    for (int t = 0; t < op->terms; t++)
      for (int d = 0; d < DIM; d++)
        for (int i = 0; i < P::M * P::M; i++) p.kron[t][d][i] = op->kref[((size_t)(t * DIM + d)) * P::M * P::M + i];
  }
  if (op->kind == DKT_OP_DENSE)
  {
    if (!op->kref) { set_error("DKT_OP_DENSE needs kref"); return DKT_ERR_INVALID; }
    std::memcpy(p.K, op->kref, sizeof(double) * P::N * P::N);
    if (ORDER == 1 && !(flags & DKT_MV_NO_FASTPATH))
    {
      // D = H K H / N ; if it is diagonal the operator is applied in Walsh-Hadamard form
      constexpr int N = P::N;
      static thread_local double T[N * N], D[N * N];
      auto h = [](int i, int j) { return (__builtin_popcount((unsigned)(i & j)) & 1) ? -1.0 : 1.0; };
      for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++)
        {
          double a = 0.0;
          for (int k = 0; k < N; k++) a += h(i, k) * op->kref[k * N + j];
          T[i * N + j] = a;
        }
      double dmax = 0.0, omax = 0.0;
      for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++)
        {
          double a = 0.0;
          for (int k = 0; k < N; k++) a += T[i * N + k] * h(k, j);
          D[i * N + j] = a / N;
          if (i == j) dmax = std::max(dmax, std::fabs(a / N));
          else omax = std::max(omax, std::fabs(a / N));
        }
      if (omax <= 1e-14 * dmax)
      {
        hadamard = true;
        for (int i = 0; i < N; i++) p.K[i] = D[i * N + i] / N;
      }
    }
  }
  // family tables serve the fast forms only; everything else runs on the per-element tables of all elements (built on first use)
  const std::vector<ChunkSet> *sets = &da.sets;
  if (da.families && !((hadamard || op->kind == DKT_OP_IDENTITY) && p.exact_ip))
  {
    const int rc = ensure_elem_sets(da);
    if (rc) return rc;
    sets = &da.sets_elem;
  }
  if (zeroOut)
  {
    CK(cudaMemsetAsync(d_out, 0, da.nNodes * sizeof(double), da.stream));
    g_launches++;
  }
  const bool diri = op->dirichlet != 0;
  if constexpr (ORDER == 1)
  {
    if (hadamard)
      return diri ? launch_mv3<DIM, ORDER, OP_HADAMARD, true>(da, *sets, p, phaseMask) : launch_mv3<DIM, ORDER, OP_HADAMARD, false>(da, *sets, p, phaseMask);
  }
  if (op->kind == DKT_OP_IDENTITY)
    return diri ? launch_mv3<DIM, ORDER, DKT_OP_IDENTITY, true>(da, *sets, p, phaseMask) : launch_mv3<DIM, ORDER, DKT_OP_IDENTITY, false>(da, *sets, p, phaseMask);
  if (op->kind == DKT_OP_DENSE)
    return diri ? launch_mv3<DIM, ORDER, DKT_OP_DENSE, true>(da, *sets, p, phaseMask) : launch_mv3<DIM, ORDER, DKT_OP_DENSE, false>(da, *sets, p, phaseMask);
  if constexpr (ORDER == 2)
  {
    if (op->kind == DKT_OP_KRON)
      return diri ? launch_mv3<DIM, ORDER, DKT_OP_KRON, true>(da, *sets, p, phaseMask) : launch_mv3<DIM, ORDER, DKT_OP_KRON, false>(da, *sets, p, phaseMask);
  }
  set_error("unknown operator kind");
  return DKT_ERR_INVALID;
}

int run_matvec_chunked(DA &da, const dkt_op *op, const double *d_in, double *d_out, double scale, unsigned flags, unsigned phaseMask,
                       bool zeroOut)
{
  if (da.N > MAX_NPE) return run_matvec(da, op, d_in, d_out, scale, flags);  // 81 nodes per element: flat kernels only
  const int key = da.dim * 10 + da.order;
  switch (key)
  {
  case 21: return run_typed3<2, 1>(da, op, d_in, d_out, scale, flags, phaseMask, zeroOut);
  case 22: return run_typed3<2, 2>(da, op, d_in, d_out, scale, flags, phaseMask, zeroOut);
  case 31: return run_typed3<3, 1>(da, op, d_in, d_out, scale, flags, phaseMask, zeroOut);
  case 32: return run_typed3<3, 2>(da, op, d_in, d_out, scale, flags, phaseMask, zeroOut);
  case 41: return run_typed3<4, 1>(da, op, d_in, d_out, scale, flags, phaseMask, zeroOut);
  default:
    set_error("unsupported (dim, order)");
    return DKT_ERR_UNSUPPORTED;
  }
}
} // namespace dkt
