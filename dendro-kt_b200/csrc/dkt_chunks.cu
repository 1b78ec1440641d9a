// The production matvec: SFC-contiguous element CHUNKS staged through shared memory by
// persistent, software-pipelined CTAs.
//
// Why: the flat kernels (dkt_matvec.cu) issue (order+1)^dim scattered 8-byte gathers and fp64
// atomics per element straight to L1/L2 and are L1TEX/atomic bound at ~15 % of the HBM roofline
// (profiles/r01_*).  Shared-memory fp64 atomics are CAS spin loops on sm_100a
// (ATOMS.CAST.SPIN.64), so the in-chunk reduction is made atomic-free instead.
//
//   build (once per DA): every set of UNITS (elements, or sibling groups - see below) first gets a
//   "unit slot table" U[unit][slot] = node id (k_unit_slots_*), then k_chunk_build (one CTA per chunk,
//   cub::BlockRadixSort in shared memory) turns it into the chunk tables:
//     * the chunk's slots are sorted by the node they touch -> unique nodes, run length `len` of each node
//     * nodes are re-ranked by (len descending, id ascending): jagged-diagonal storage.  The k-th
//       contribution to node n lives at X[jd[k] + n]; jd[k] = number of (node, j<k) pairs.
//     * every slot gets one 32-bit word  n | (jd[k] + n) << 16 ; words are stored slot-major
//       inside the chunk so that a warp reads 32 consecutive words
//     * every node gets its global id and a meta word: len | boundary bit | shared-with-another-
//       chunk bit (global reference count != len)
//
//   matvec (k_mv3, persistent CTAs looping over chunks c = blockIdx.x, +gridDim.x, ...):
//     T0  wait for the cp.async gather of this chunk's node values, barrier
//     T1  load the NEXT chunk's node ids / meta into registers (overlaps T2)
//     T2  thread per element: ein[r] = un[n] (LDS), K_e, X[pos] = eout[r] (STS; every slot owns
//         one position, so there are no write conflicts and no atomics)
//     T3  issue cp.async 8-byte gathers un_next[n] <- u[gid] for the next chunk, load its slot
//         words and jd table (overlaps T4)
//     T4  barrier; thread per node: acc = sum_k X[jd[k] + n]  (lanes read consecutive words;
//         neighbouring lanes have equal len, so no divergence) and ONE plain store (node private
//         to the chunk) or ONE fp64 RED (shared) per (chunk, node)
//
// Global atomics drop from N per element to ~2-3 per element in 4-D (the chunk surface).
//
// Order-1 extras (all build-time table tricks, no extra kernel work):
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
// Partitioned DAs build three phases of sets (interior first half / boundary / interior second half)
// so that dkt_dist.cu can run the ghost exchanges beside the interior elements.
//
// SIBLING GROUPS (order 1, opt-in: environment DKT_GROUPS=g at DA construction; see k_mvg below): the
// 2^g leaves of a complete sibling family that agree in the child-number bits of the dimensions >= g are ONE
// unit handled by one thread.  They share a 3^g x 2^(dim-g) node lattice, so a quad (dim 4, g = 2) needs 36
// gathers and 36 scatters where four separate elements need 64 + 64, and a hanging group reads the 16 parent
// nodes once.  Elements outside complete families stay in per-element sets.
//
// This file also compiles under -DDKT_EMU with tests/emu/cuda_emu.h (fibers on the CPU) - that build exists
// ONLY so the CPU test-suite can execute the table construction and the kernels' logic against the oracle;
// it is never part of libdkt.so.
#include "dkt_internal.h"

#ifdef DKT_EMU
#include "cuda_emu.h"
#else
#include <cub/block/block_radix_sort.cuh>
#include <cub/block/block_scan.cuh>
#define DKT_LAUNCH(k, g, b, s, st) k<<<(g), (b), (s), (st)>>>
#define DKT_DYN_SMEM(type, name) extern __shared__ type name[]
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

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;                       // per-element sets: 4096 slots per chunk
constexpr int SORT_ITEMS_GRP = 20;                   // sibling-group sets: 5120 slots per chunk (128 quads of 36, 96 hanging quads of 52)
constexpr int SLOT_CAP = SORT_THREADS * SORT_ITEMS;
constexpr int SLOT_CAP_GRP = SORT_THREADS * SORT_ITEMS_GRP;
constexpr int MAX_LEN = 511;                         // run length of a node inside a chunk (9 bits)
constexpr uint32_t META_LEN = 0x1FFu;
constexpr uint32_t META_PRESENT = 0x2000u;  // node exists (its run may be empty: only read by this chunk)
constexpr uint32_t META_BDY = 0x4000u;
constexpr uint32_t META_SHARED = 0x8000u;
constexpr uint32_t REC_SHARED = 0x80000000u, REC_BDY = 0x40000000u, REC_GID = 0x3FFFFFFFu;  // group sets: 4-byte node records
constexpr uint32_t SLOT_RO = 0x80000000u;   // unit slot table: read-only reference (node ids are < 2^31)
#ifndef DKT_GRP_TPB
#define DKT_GRP_TPB 128  // threads (= units per chunk at most) of the group kernels; 96 lets three CTAs of regular quads share an SM
#endif
constexpr int GRP_TPB = DKT_GRP_TPB;
// units per chunk of a group set with `spu` slots per unit (groups of 2^g leaves): what the block sort holds, at most
// 512 elements in 4-D / 1024 below (the chunk's nodes are double-buffered in shared memory: with more, two CTAs no longer
// fit on an SM), rounded down to whole warps when that costs at most an eighth; threads of its kernel: the next multiple of 32
constexpr int grp_upc(int spu, int dim, int g)
{
  int u = SORT_THREADS * SORT_ITEMS_GRP / spu;
  if (u > GRP_TPB) u = GRP_TPB;
  const int cap = (dim >= 4 ? 512 : 1024) >> g;
  if (u > cap) u = cap;
  const int r = u & ~31;
  return (r > 0 && (u - r) * 8 <= u) ? r : u;
}
constexpr int grp_tpb(int spu, int dim, int g) { return (grp_upc(spu, dim, g) + 31) & ~31; }

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
// resident CTAs per SM the group kernels are compiled for.  ptxas, 4-D: 2 -> 220 (regular quads) / 255 (hanging quads) /
// 220 (hanging pairs) registers, no spills; 3 -> 168 registers: regular quads and hanging pairs without spills, hanging
// quads 0.6 KB of spills (profiles/r01_ptxas_groups.txt)
#ifndef DKT_GRP_RSHARE
#define DKT_GRP_RSHARE 1  // regular groups: shared butterflies along the XOR-permuted dimensions (see k_mvg)
#endif
#ifndef DKT_GRP_MINB_REG
#define DKT_GRP_MINB_REG 2
#endif
#ifndef DKT_GRP_MINB_HANG
#define DKT_GRP_MINB_HANG 2
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
// chunk's node count and longest run.  split16: the slot words are stored as two 16-bit arrays (node rank /
// position), two slots per 32-bit word, the node records as ONE 32-bit word gid | boundary bit | shared bit, and the jd table is
// followed by cnt[k] = #nodes with a run longer than k, which replaces the run length of the records (group sets).
template <bool WRITE, int ITEMS>
__global__ void __launch_bounds__(SORT_THREADS)
k_chunk_build(const uint32_t *U, uint64_t nUnits, int spu, int upc, int split16, const uint32_t *refcnt, const uint8_t *isbdy,
              const uint64_t *node_off, int jdStride, uint32_t *nloc_out, uint32_t *maxlen_out, uint32_t *slot, uint16_t *rk16,
              uint16_t *ps16, uint32_t *gid_out, uint16_t *meta_out, uint32_t *rec_out, uint16_t *jd_out)
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
      if (k < 16)
        while ((acc & 15) != k) acc++;
      s_jd[k] = acc;
      acc += cnt;
    }
    for (int k = ml + 1; k < MAX_LEN + 2; k++) s_jd[k] = acc;
  }
  __syncthreads();
  const uint64_t noff = node_off[c];
  if (!split16)
    for (int k = threadIdx.x; k < jdStride; k += SORT_THREADS) jd_out[c * (uint64_t)jdStride + k] = (uint16_t)s_jd[min(k, MAX_LEN + 1)];
  else  // group sets: [jd | cnt] per chunk; the node records carry no run length, cnt[] gives it (nodes are ranked by it)
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
    if (!split16) slot[u0 * spu + (uint64_t)q * upc + el] = nr | (pos << 16);
    else
    {
      const uint64_t w16 = (u0 * (uint64_t)(spu / 2) + (uint64_t)(q >> 1) * upc + el) * 2 + (q & 1);
      rk16[w16] = (uint16_t)nr;
      ps16[w16] = (uint16_t)pos;
    }
    if (key[i] != INVALID && head[i])
    {
      const int n0 = rank0[i];
      const int len = (int)s_cw[n0 + 1] - (int)s_cw[n0];
      uint32_t m = (uint32_t)len | META_PRESENT;
      if (refcnt[key[i]] != (uint32_t)len) m |= META_SHARED;
      if (isbdy[key[i]]) m |= META_BDY;
      if (!split16)
      {
        gid_out[noff + nr] = key[i];
        meta_out[noff + nr] = (uint16_t)m;
      }
      else rec_out[noff + nr] = key[i] | ((m & META_SHARED) ? REC_SHARED : 0u) | ((m & META_BDY) ? REC_BDY : 0u);
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

// Chunk tables of one set from its unit slot table U (freed by the caller).
static int build_set(DA &da, ChunkSet &cs, const uint32_t *U, const uint32_t *refcnt)
{
  const uint64_t nSet = cs.nElem;
  if (nSet == 0) return DKT_OK;
  const int spu = cs.spu, upc = (int)cs.elemsPerChunk, split16 = cs.kind == 1 ? 1 : 0;
  cs.nChunks = (uint32_t)((nSet + upc - 1) / upc);
  uint32_t *nloc = nullptr, *mlen = nullptr;
  uint64_t *wide = nullptr, *off = nullptr;
  CK(cudaMalloc((void **)&nloc, (size_t)cs.nChunks * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&mlen, (size_t)cs.nChunks * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&wide, ((size_t)cs.nChunks + 1) * sizeof(uint64_t)));
  CK(cudaMalloc((void **)&off, ((size_t)cs.nChunks + 1) * sizeof(uint64_t)));
  auto kcount = split16 ? k_chunk_build<false, SORT_ITEMS_GRP> : k_chunk_build<false, SORT_ITEMS>;
  auto kwrite = split16 ? k_chunk_build<true, SORT_ITEMS_GRP> : k_chunk_build<true, SORT_ITEMS>;
  DKT_LAUNCH(kcount, cs.nChunks, SORT_THREADS, 0, da.stream)(U, nSet, spu, upc, split16, refcnt, da.d_node_isbdy, nullptr, 0, nloc, mlen,
                                                              nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  g_launches++;
  CK(cudaMemsetAsync(wide, 0, ((size_t)cs.nChunks + 1) * sizeof(uint64_t), da.stream));
  DKT_LAUNCH(k_u32_widen, nblk(cs.nChunks), 256, 0, da.stream)(nloc, cs.nChunks, wide);
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
  if (!split16)
  {
    CK(cudaMalloc((void **)&cs.d_slot, nslotsAll * sizeof(uint32_t)));
    CK(cudaMalloc((void **)&cs.d_gid, std::max<uint64_t>(total, 1) * sizeof(uint32_t)));
    CK(cudaMalloc((void **)&cs.d_meta, std::max<uint64_t>(total, 1) * sizeof(uint16_t)));
  }
  else
  {
    CK(cudaMalloc((void **)&cs.d_rk16, nslotsAll * sizeof(uint16_t)));
    CK(cudaMalloc((void **)&cs.d_ps16, nslotsAll * sizeof(uint16_t)));
    CK(cudaMalloc((void **)&cs.d_rec, std::max<uint64_t>(total, 1) * sizeof(uint32_t)));
  }
  CK(cudaMalloc((void **)&cs.d_jd, (size_t)cs.nChunks * cs.jdStride * (split16 ? 2 : 1) * sizeof(uint16_t)));
  DKT_LAUNCH(kwrite, cs.nChunks, SORT_THREADS, 0, da.stream)(U, nSet, spu, upc, split16, refcnt, da.d_node_isbdy, off, (int)cs.jdStride, nullptr,
                                                              nullptr, cs.d_slot, cs.d_rk16, cs.d_ps16, cs.d_gid, cs.d_meta, (uint32_t *)cs.d_rec,
                                                              cs.d_jd);
  g_launches++;
  CK(cudaStreamSynchronize(da.stream));
  CK(cudaGetLastError());
  cudaFree(nloc);
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
  for (ChunkSet &cs : da.sets)
  {
    cudaFree(cs.d_slot); cudaFree(cs.d_gid); cudaFree(cs.d_meta); cudaFree(cs.d_jd); cudaFree(cs.d_node_off);
    cudaFree(cs.d_rk16); cudaFree(cs.d_ps16); cudaFree(cs.d_rec);
    for (void *p : cs.owned) cudaFree(p);
  }
  da.sets.clear();
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

// ---- sibling groups: discovery ---------------------------------------------------------------
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
// mem[f * 2^dim + c] = visited-element index of child c of family f; infam[e] = 1
__global__ void k_family_members(const uint32_t *inv, const uint64_t *head, const uint64_t *fpos, const uint8_t *child, uint64_t n, int dim,
                                 uint32_t *mem, uint8_t *infam)
{
  uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (j >= n || !head[j]) return;
  const int nch = 1 << dim;
  const uint64_t f = fpos[j];
  for (int t = 0; t < nch; t++)
  {
    const uint32_t b = inv[j + t];
    mem[f * nch + child[b]] = b;
    infam[b] = 1;
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
// mode 0: plain.  mode 1 (groups of 2^g with smaller hanging groups to follow): hanging groups get class 255.
// mode 2 (the smaller groups, gOuter > g): groups inside a REGULAR outer group get class 255 - that one has them.
__global__ void k_group_class(const uint32_t *mem, uint64_t nGroups, int dim, int g, int gOuter, int mode, uint64_t nReg, uint64_t nRegInt,
                              uint64_t nHangInt, int phased, uint8_t *cls)
{
  uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (u >= nGroups) return;
  const int NR = 1 << (dim - g), NC = 1 << g;
  const uint64_t f = u / NR;
  const int cR = (int)(u % NR);
  int c = 0;
  for (int cG = 0; cG < NC; cG++) c |= elem_class(mem[(f << dim) + ((cR << g) | cG)], nReg, nRegInt, nHangInt, phased);
  if (mode == 1 && (c & 2)) c = 255;
  if (mode == 2)
  {
    const int cRo = cR >> (gOuter - g);
    bool outerHang = false;
    for (int cG = 0; cG < (1 << gOuter); cG++) outerHang |= mem[(f << dim) + ((cRo << gOuter) | cG)] >= nReg;
    if (!outerHang) c = 255;
  }
  cls[u] = (uint8_t)c;
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
// member lists of the selected groups, in (family, cR) order
__global__ void k_group_list(const uint32_t *mem, const uint64_t *flag, const uint64_t *pos, uint64_t nGroups, int dim, int g,
                             uint32_t *list)
{
  uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (u >= nGroups || !flag[u]) return;
  const int NR = 1 << (dim - g), NC = 1 << g;
  const uint64_t f = u / NR;
  const int cR = (int)(u % NR);
  uint32_t *dst = list + pos[u] * NC;
  for (int cG = 0; cG < NC; cG++) dst[cG] = mem[(f << dim) + ((cR << g) | cG)];
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
// Unit slot table of a group set.  Own slot l = xg + 3^g * sR: xg = sum_{d<g} x_d 3^d is the position in the
// family's 3-point lattice of the grouped dimensions, sR the XOR-permuted rank bits of the others (lattice
// bit s_d ^ c_d, the group's members share c_d for d >= g).  Slots LP.. (hanging groups): parent-lattice
// rank (qG natural, tR ^ cR).  The table is padded to an even number of slots per unit.
__global__ void k_unit_slots_group(const uint32_t *list, uint64_t nUnits, int dim, int g, int hang, uint64_t nReg, const uint32_t *e2n,
                                   const uint32_t *pnode, const uint8_t *child, const uint8_t *lev, uint32_t *U, uint8_t *lev_g,
                                   unsigned long long *fmask64)
{
  const int N = 1 << dim, NC = 1 << g, NR = 1 << (dim - g);
  int L3 = 1;
  for (int d = 0; d < g; d++) L3 *= 3;
  const int LP = L3 * NR, spu = (LP + (hang ? N : 0) + 1) & ~1;
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= nUnits * spu) return;
  const uint64_t u = i / spu;
  const int q = (int)(i % spu);
  const uint32_t *m = list + u * NC;
  const int cR = child[m[0]] >> g;
  uint32_t key = INVALID;
  if (q < LP)
  {
    const int xg = q % L3, sR = q / L3;
    for (int cG = 0; cG < NC; cG++)
    {
      int rG = 0, rem = xg;
      bool ok = true;
      for (int d = 0; d < g; d++)
      {
        const int r = (rem % 3) - ((cG >> d) & 1);
        rem /= 3;
        if (r < 0 || r > 1) ok = false;
        else rG |= r << d;
      }
      if (!ok) continue;
      const uint32_t k = e2n[(uint64_t)m[cG] * N + (rG | ((sR ^ cR) << g))];
      if (k != INVALID) key = k;
    }
  }
  else if (q < LP + (hang ? N : 0))
  {
    const int t = q - LP;
    const int pr = (t & (NC - 1)) | (((t >> g) ^ cR) << g);
    for (int cG = 0; cG < NC; cG++)
      if (m[cG] >= nReg) { key = pnode[((uint64_t)m[cG] - nReg) * N + pr]; break; }
  }
  U[i] = key;
  if (q == 0) lev_g[u] = lev[m[0]];
  if (hang && q < LP && key != INVALID) atomicOr(fmask64 + u, 1ull << q);
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
static int add_elem_set(DA &da, std::vector<PendingSet> &pend, const uint32_t *e2n, const uint32_t *pnode, const uint8_t *lev,
                        const uint8_t *child, const uint32_t *fmask, uint64_t n, int rows, int phase, uint64_t elem0, uint64_t hang0)
{
  if (n == 0) return DKT_OK;
  da.sets.emplace_back();
  ChunkSet &cs = da.sets.back();
  cs.rows = rows; cs.phase = phase; cs.elem0 = elem0; cs.hang0 = hang0; cs.nElem = n; cs.kind = 0;
  cs.xorperm = da.order == 1 ? 1 : 0;
  cs.spu = rows * da.N;
  cs.elemsPerChunk = rows_per_chunk(da.N) / rows;
  cs.lev = lev; cs.child = child; cs.fmask = fmask;
  uint32_t *U = nullptr;
  CK(cudaMalloc((void **)&U, n * cs.spu * sizeof(uint32_t)));
  DKT_LAUNCH(k_unit_slots_elem, nblk(n * cs.spu), 256, 0, da.stream)(e2n, pnode, child, n, da.N, rows, cs.xorperm, U);
  g_launches++;
  pend.push_back({da.sets.size() - 1, U});
  return DKT_OK;
}

static int add_group_set(DA &da, std::vector<PendingSet> &pend, const uint32_t *list, uint64_t n, int g, int hang, int phase)
{
  if (n == 0) return DKT_OK;
  da.sets.emplace_back();
  ChunkSet &cs = da.sets.back();
  int L3 = 1;
  for (int d = 0; d < g; d++) L3 *= 3;
  const int LP = L3 << (da.dim - g);
  cs.rows = hang ? 2 : 1; cs.phase = phase; cs.nElem = n; cs.kind = 1; cs.g = g; cs.xorperm = 1;
  cs.spu = (LP + (hang ? da.N : 0) + 1) & ~1;
  cs.elemsPerChunk = grp_upc(cs.spu, da.dim, g);
  uint32_t *U = nullptr;
  uint8_t *lev_g = nullptr;
  unsigned long long *fm = nullptr;
  CK(cudaMalloc((void **)&U, n * cs.spu * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&lev_g, n));
  CK(cudaMalloc((void **)&fm, n * sizeof(unsigned long long)));
  CK(cudaMemsetAsync(fm, 0, n * sizeof(unsigned long long), da.stream));
  DKT_LAUNCH(k_unit_slots_group, nblk(n * cs.spu), 256, 0, da.stream)(list, n, da.dim, g, hang, da.nReg, da.d_e2n, da.d_pnode, da.d_mv_child,
                                                                        da.d_mv_lev, U, lev_g, fm);
  g_launches++;
  cs.lev = lev_g; cs.fmask64 = (const uint64_t *)fm;
  cs.owned.push_back(lev_g); cs.owned.push_back(fm);
  pend.push_back({da.sets.size() - 1, U});
  return DKT_OK;
}

// per-element set of the elements in `list` (outside complete families): compact copies of their rows
static int add_single_set(DA &da, std::vector<PendingSet> &pend, const uint32_t *list, uint64_t nS, int hang, int phase)
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
  int rc = add_elem_set(da, pend, e2n_s, pnode_s, lev_s, child_s, fm_s, nS, hang ? 2 : 1, phase, 0, 0);
  if (rc == DKT_OK)
  {
    ChunkSet &cs = da.sets.back();
    cs.owned.push_back(lev_s); cs.owned.push_back(child_s);
    if (fm_s) cs.owned.push_back(fm_s);
  }
  CK(cudaStreamSynchronize(da.stream));
  cudaFree(e2n_s);
  cudaFree(pnode_s);
  return rc;
}

// DKT_GROUPS=g: group the leaves of complete sibling families (see the file header).  0 / unset: off.
// "g" or "gR,gH" (gH <= gR): regular groups of 2^gR leaves; groups of that size with a hanging member are split into
// groups of 2^gH leaves, which are regular or hanging in their turn (smaller hanging groups need fewer registers).
static bool group_kernel_exists(int dim, int g)
{
  return (dim == 4 && g >= 1 && g <= 3) || (dim == 3 && (g == 2 || g == 3)) || (dim == 2 && g == 2);
}
static int groups_requested(const DA &da, int &gH)
{
  gH = 0;
  const char *e = getenv("DKT_GROUPS");
  if (!e) return 0;
  const int g = atoi(e);
  gH = g;
  if (const char *c = strchr(e, ',')) gH = atoi(c + 1);
  if (g <= 0 || da.order != 1 || gH <= 0 || gH > g) return 0;
  if (!group_kernel_exists(da.dim, g) || !group_kernel_exists(da.dim, gH)) return 0;
  return g;
}

int build_chunks(DA &da)
{
  // more than 27 nodes per element (4-D order 2): no shared-memory kernel; every matvec runs on the flat kernels
  if (da.N > MAX_NPE) return DKT_OK;
  if (da.nNodes >= 0x7FFFFFFFull) { set_error("more than 2^31 nodes on one rank"); return DKT_ERR_UNSUPPORTED; }
  if (getenv("DKT_GROUPS") && da.nNodes >= 0x3FFFFFFFull) { set_error("DKT_GROUPS: more than 2^30 nodes on one rank"); return DKT_ERR_UNSUPPORTED; }
  CK(cudaMalloc((void **)&da.d_mv_child, std::max<uint64_t>(da.nMv, 1)));
  DKT_LAUNCH(k_child_numbers, nblk(da.nMv), 256, 0, da.stream)(da.d_mv_xyz, da.d_mv_lev, da.nMv, da.dim, da.max_depth, da.d_mv_child);
  g_launches++;
  const int N = da.N;
  const int xorperm = da.order == 1 ? 1 : 0;
  if (da.nHang)
  {
    CK(cudaMalloc((void **)&da.d_fmask, da.nHang * sizeof(uint32_t)));
    DKT_LAUNCH(k_fmask, nblk(da.nHang), 256, 0, da.stream)(da.d_e2n + da.nReg * (uint64_t)N, da.d_mv_child + da.nReg, da.nHang, N, xorperm,
                                                           da.d_fmask);
    g_launches++;
  }
  std::vector<PendingSet> pend;
  int rc = DKT_OK;
  int gHang = 0;
  const int g = groups_requested(da, gHang);
  da.groups = g;
  if (!g)
  {
    // element ranges: one regular + one hanging set, or (partitioned) three phases of each:
    // interior first half / boundary / interior second half - see run_matvec_dist
    struct Range { uint64_t a, b; int phase; };
    std::vector<Range> rr, hr;
    if (da.phased)
    {
      const uint64_t ri = da.nRegInterior, hi = da.nHangInterior;
      rr = {{0, ri / 2, 0}, {ri, da.nReg, 1}, {ri / 2, ri, 2}};
      hr = {{0, hi / 2, 0}, {hi, da.nHang, 1}, {hi / 2, hi, 2}};
    }
    else
    {
      rr = {{0, da.nReg, 0}};
      hr = {{0, da.nHang, 0}};
    }
    for (const Range &r : rr)
      if (rc == DKT_OK && r.b > r.a)
        rc = add_elem_set(da, pend, da.d_e2n + r.a * N, nullptr, da.d_mv_lev + r.a, da.d_mv_child + r.a, nullptr, r.b - r.a, 1, r.phase, r.a, 0);
    for (const Range &r : hr)
      if (rc == DKT_OK && r.b > r.a)
        rc = add_elem_set(da, pend, da.d_e2n + (da.nReg + r.a) * N, da.d_pnode + r.a * N, da.d_mv_lev + da.nReg + r.a,
                          da.d_mv_child + da.nReg + r.a, da.d_fmask + r.a, r.b - r.a, 2, r.phase, da.nReg + r.a, r.a);
  }
  else
  {
    const uint64_t n = da.nMv;
    const int nch = 1 << da.dim;
    const int phased = da.phased ? 1 : 0;
    uint32_t *inv = nullptr, *mem = nullptr;
    uint64_t *flag = nullptr, *pos = nullptr;
    uint8_t *infam = nullptr, *cls = nullptr;
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
    CK(cudaMalloc((void **)&mem, std::max<uint64_t>(nFam, 1) * nch * sizeof(uint32_t)));
    DKT_LAUNCH(k_family_members, nblk(n), 256, 0, da.stream)(inv, flag, pos, da.d_mv_child, n, da.dim, mem, infam);
    g_launches++;
    std::vector<uint32_t *> lists;  // freed after the unit slot tables are built
    // interior units run in two halves around the boundary ones (see run_matvec_dist); unpartitioned: one set
    auto add_sets = [&](const uint32_t *list, uint64_t cnt, int width, int c, int gsz) -> int {
      const int hang = (c >> 1) & 1, bdy = c & 1;
      struct Sub { uint64_t a, b; int phase; };
      std::vector<Sub> subs;
      if (!phased) subs = {{0, cnt, 0}};
      else if (bdy) subs = {{0, cnt, 1}};
      else subs = {{0, cnt / 2, 0}, {cnt / 2, cnt, 2}};
      for (const Sub &s : subs)
      {
        const int r = gsz ? add_group_set(da, pend, list + s.a * width, s.b - s.a, gsz, hang, s.phase)
                          : add_single_set(da, pend, list + s.a, s.b - s.a, hang, s.phase);
        if (r) return r;
      }
      return DKT_OK;
    };
    // groups by class: regular / hanging x interior / boundary; with DKT_GROUPS=gR,gH a second pass makes the
    // smaller groups out of the size-2^gR groups that have a hanging member
    for (int pass = 0; pass < (gHang < g ? 2 : 1) && rc == DKT_OK; pass++)
    {
      const int gp = pass == 0 ? g : gHang;
      const int mode = gHang < g ? pass + 1 : 0;
      const int NCp = 1 << gp;
      const uint64_t nGroups = nFam << (da.dim - gp);
      if (!nGroups) continue;
      uint64_t *gflag = nullptr, *gpos = nullptr;
      uint8_t *gcls = nullptr;
      CK(cudaMalloc((void **)&gflag, (nGroups + 1) * sizeof(uint64_t)));
      CK(cudaMalloc((void **)&gpos, (nGroups + 1) * sizeof(uint64_t)));
      CK(cudaMalloc((void **)&gcls, nGroups));
      DKT_LAUNCH(k_group_class, nblk(nGroups), 256, 0, da.stream)(mem, nGroups, da.dim, gp, g, mode, da.nReg, da.nRegInterior,
                                                                  da.nHangInterior, phased, gcls);
      g_launches++;
      for (int c = 0; c < 4 && rc == DKT_OK; c++)
      {
        DKT_LAUNCH(k_class_flags, nblk(nGroups + 1), 256, 0, da.stream)(gcls, nGroups, c, gflag);
        g_launches++;
        uint64_t cnt = 0;
        rc = scan_total(da, gflag, gpos, nGroups, cnt);
        if (rc || !cnt) continue;
        uint32_t *list = nullptr;
        CK(cudaMalloc((void **)&list, cnt * NCp * sizeof(uint32_t)));
        lists.push_back(list);
        DKT_LAUNCH(k_group_list, nblk(nGroups), 256, 0, da.stream)(mem, gflag, gpos, nGroups, da.dim, gp, list);
        g_launches++;
        rc = add_sets(list, cnt, NCp, c, gp);
      }
      CK(cudaStreamSynchronize(da.stream));
      cudaFree(gflag); cudaFree(gpos); cudaFree(gcls);
    }
    // the elements outside complete families, by the same classes
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
      lists.push_back(list);
      DKT_LAUNCH(k_compact, nblk(n), 256, 0, da.stream)(flag, pos, n, list);
      g_launches++;
      rc = add_sets(list, cnt, 1, c, 0);
    }
    CK(cudaStreamSynchronize(da.stream));
    for (uint32_t *l : lists) cudaFree(l);
    cudaFree(inv); cudaFree(mem); cudaFree(flag); cudaFree(pos); cudaFree(infam); cudaFree(cls);
  }
  // writing references of every node over all sets, then the chunk tables
  uint32_t *refcnt = nullptr;
  CK(cudaMalloc((void **)&refcnt, std::max<uint64_t>(da.nNodes, 1) * sizeof(uint32_t)));
  CK(cudaMemsetAsync(refcnt, 0, std::max<uint64_t>(da.nNodes, 1) * sizeof(uint32_t), da.stream));
  for (const PendingSet &ps : pend)
  {
    const ChunkSet &cs = da.sets[ps.idx];
    const uint64_t ns = cs.nElem * cs.spu;
    if (rc == DKT_OK && ns)
    {
      DKT_LAUNCH(k_ref_count, nblk(ns), 256, 0, da.stream)(ps.U, ns, refcnt);
      g_launches++;
    }
  }
  for (const PendingSet &ps : pend)
    if (rc == DKT_OK) rc = build_set(da, da.sets[ps.idx], ps.U, refcnt);
  cudaStreamSynchronize(da.stream);
  for (const PendingSet &ps : pend) cudaFree(ps.U);
  cudaFree(refcnt);
  int dev = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&da.numSMs, cudaDevAttrMultiProcessorCount, dev));
  da.mvStreams = 1;
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
  return rc;
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
  const uint32_t *rk16, *ps16;   // group sets: 16-bit node ranks / positions, two slots per word
  const uint32_t *rec;           // group sets: gid | boundary bit 30 | shared bit 31 per chunk node
  const uint64_t *fmask64;       // hanging group sets: filled own lattice slots
  uint32_t nSet, nChunks, elemsPerChunk, xcap, ncap, jdStride;
  int q1mask;
  int exact_ip;          // order 1: ip0/ip1 equal the exact interpolation to 1e-13
  double lscale[32];
  double ip[2][M * M];
  double ipx[2][M * M];  // ip[0] and J ip[1] J (XOR-permuted coordinates)
  double K[N * N];
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
__global__ void __launch_bounds__(TPB, ((Mv3Params<DIM, ORDER>::N <= 16 && OPKIND != DKT_OP_DENSE) ? (HANG ? DKT_HANG_MINB : DKT_REG_MINB) * (256 / DKT_ROWS) : 2)) k_mv3(const __grid_constant__ Mv3Params<DIM, ORDER> p)
{
  constexpr int N = Mv3Params<DIM, ORDER>::N;
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

  uint32_t gidC[NPT], gidN[NPT];
  uint16_t metaC[NPT], metaN[NPT];
  auto load_nodes = [&](uint64_t oa, uint64_t ob, uint32_t *g, uint16_t *m) {
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
  issue_gather(unb, gidC, metaC, (int)(offB - offA));
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
      issue_gather(unb + (buf ^ 1) * p.ncap, gidN, metaN, (int)(offNB - offNA));
      int *jdn = jdb + (buf ^ 1) * p.jdStride;
      for (int k = tid; k < (int)p.jdStride; k += TPB) jdn[k] = p.jd[cn * (uint64_t)p.jdStride + k];
      load_slots(cn);
    }
    cp_async_commit();
    __syncthreads();  // T4: X complete
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
#pragma unroll
    for (int k = 0; k < NPT; k++) { gidC[k] = gidN[k]; metaC[k] = metaN[k]; }
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

// ------------------------------------------------------------------------------------------
// sibling-group kernel (order 1, identity / Walsh-Hadamard operators, exact interpolation)
// ------------------------------------------------------------------------------------------
// One thread per GROUP: the 2^G leaves of a complete sibling family that share the child-number bits of the
// dimensions >= G.  Their nodes form a 3^G x 2^(DIM-G) lattice (LP slots; the dimensions >= G keep the XOR
// schedule, so the 2^(DIM-G) groups of a family still read the same node in the same instruction).  Per group:
// LP gathers, 2^G elemental operators on register-resident values with STATIC indices, LP scatters; a hanging
// group additionally reads the 2^DIM parent nodes once, interpolates them to the whole lattice (exact order-1
// interpolation: midpoints) and scatters one masked transposed sum.  Quirk Q1 (FEM/include/matvec.h:517) is
// applied per child at run time from the 64-bit fill mask, so every parent slot is a writing slot here.
// Pipeline per chunk (two barriers): the 4-byte node records are staged in shared memory by cp.async - every
// thread copies, consumes and later overwrites only its OWN records -, the gather of chunk i+1 is issued right
// after barrier A and stays in flight during the whole of chunk i.
template <int DIM, int G>
struct Grp
{
  static constexpr int N = 1 << DIM, NC = 1 << G, NR = 1 << (DIM - G);
  static constexpr int L3 = (G == 1 ? 3 : G == 2 ? 9 : G == 3 ? 27 : 81);
  static constexpr int LP = L3 * NR;
  // lattice slot of rank r (grouped bits natural, the others XOR-permuted) of child cG
  __host__ __device__ static constexpr int lat(int cG, int r)
  {
    int xg = 0, p3 = 1;
    for (int d = 0; d < G; d++)
    {
      xg += (((cG >> d) & 1) + ((r >> d) & 1)) * p3;
      p3 *= 3;
    }
    return xg + L3 * (r >> G);
  }
  // the lattice slots of child cG as a bit mask
  __host__ __device__ static constexpr unsigned long long child_mask(int cG)
  {
    unsigned long long m = 0;
    for (int r = 0; r < N; r++) m |= 1ull << lat(cG, r);
    return m;
  }
  // child cG is the first (smallest) child touching the lattice point of its rank r
  __host__ __device__ static constexpr bool first(int cG, int r) { return ((cG & ~r) & (NC - 1)) == 0; }
};

// parent values (index: grouped bits natural | other bits already interpolated) -> lattice, one grouped dimension
// at a time: src index a + 3^K * b (b's lowest bit = parent corner along dimension K), dst index a + 3^K * x + 3^(K+1) * (b >> 1)
template <int DIM, int G, int K>
__device__ __forceinline__ void grp_expand(const double *src, double *dst)
{
  constexpr int P3 = (K == 0 ? 1 : K == 1 ? 3 : K == 2 ? 9 : 27);
  constexpr int NB = 1 << (DIM - K - 1);
#pragma unroll
  for (int b = 0; b < NB; b++)
#pragma unroll
    for (int a = 0; a < P3; a++)
    {
      const double lo = src[a + P3 * (2 * b)], hi = src[a + P3 * (2 * b + 1)];
      dst[a + 3 * P3 * b] = lo;
      dst[a + P3 + 3 * P3 * b] = 0.5 * (lo + hi);
      dst[a + 2 * P3 + 3 * P3 * b] = hi;
    }
}

// butterflies of the strides LO, 2 LO, .. < HI
template <int N, int LO, int HI>
__device__ __forceinline__ void wht_range(double *v)
{
#pragma unroll
  for (int s = LO; s < HI; s <<= 1)
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
// Walsh-Hadamard butterflies along the XOR-permuted dimensions (>= G) of a group lattice, index xg + 3^G * sR
template <int DIM, int G>
__device__ __forceinline__ void grp_wht_lattice(double *v)
{
  using GP = Grp<DIM, G>;
#pragma unroll
  for (int k = 0; k < DIM - G; k++)
  {
    const int st = GP::L3 << k;
#pragma unroll
    for (int l = 0; l < GP::LP; l++)
    {
      if ((l / st) & 1) continue;
      const double a = v[l], b = v[l + st];
      v[l] = a + b;
      v[l + st] = a - b;
    }
  }
}

template <int DIM, int G, int OPKIND, bool DIRI, bool HANG, int TPB>
__global__ void __launch_bounds__(TPB, HANG ? DKT_GRP_MINB_HANG : DKT_GRP_MINB_REG) k_mvg(const __grid_constant__ Mv3Params<DIM, 1> p)
{
  using GP = Grp<DIM, G>;
  constexpr int N = GP::N, LP = GP::LP, NC = GP::NC;
  constexpr int SPU = (LP + (HANG ? N : 0) + 1) & ~1;
  constexpr int NW = SPU / 2;
  DKT_DYN_SMEM(double, sm);
  double *X = sm;                                      // [xcap]
  double *unb = sm + p.xcap;                           // [2][ncap]
  int *jdb = (int *)(sm + p.xcap + 2 * p.ncap);        // [2][jd[jdStride] | cnt[jdStride]]
  uint32_t *recb = (uint32_t *)(jdb + 4 * p.jdStride); // [2][ncap]

  const int tid = threadIdx.x;
  const uint32_t E = p.elemsPerChunk;
  uint64_t c = blockIdx.x;
  if (c >= p.nChunks) return;

  auto issue_rec = [&](uint64_t oa, int nloc, int b) {
    uint32_t *dst = recb + b * p.ncap;
    for (int n = tid; n < nloc; n += TPB) cp_async4(dst + n, p.rec + oa + n);
  };
  auto issue_gather = [&](uint64_t cc, int nloc, int b) {
    double *un = unb + b * p.ncap;
    const uint32_t *rec = recb + b * p.ncap;
    if (tid == 0) un[nloc] = 0.0;  // the entry absent nodes read
    for (int n = tid; n < nloc; n += TPB)
    {
      const uint32_t r = rec[n];
      if (DIRI && (r & REC_BDY)) un[n] = 0.0;
      else cp_async8(un + n, p.in + (r & REC_GID));
    }
    int *jdn = jdb + b * 2 * p.jdStride;
    for (int k = tid; k < 2 * (int)p.jdStride; k += TPB) jdn[k] = p.jd[cc * (uint64_t)(2 * p.jdStride) + k];
  };
  uint32_t wr[NW];
  int levU = 0;
  uint64_t fm = 0;
  auto load_ranks = [&](uint64_t cc) {
    const uint64_t u0 = cc * (uint64_t)E;
    const int nu = (int)min((uint64_t)E, (uint64_t)p.nSet - u0);
    if (tid < nu)
    {
      const uint32_t *sw = p.rk16 + u0 * NW + tid;
#pragma unroll
      for (int j = 0; j < NW; j++) wr[j] = sw[(uint32_t)j * E];
      levU = p.lev[u0 + tid];
      if (HANG) fm = p.fmask64[u0 + tid];
    }
  };
#define GRK(l) ((wr[(l) >> 1] >> (((l) & 1) * 16)) & 0xFFFFu)
#define GPS(l) ((ps[(l) >> 1] >> (((l) & 1) * 16)) & 0xFFFFu)

  // ---- prologue ---------------------------------------------------------------------------------
  const uint64_t stride = gridDim.x;
  uint64_t oa = p.node_off[c], ob = p.node_off[c + 1];
  int nlocC = (int)(ob - oa);
  issue_rec(oa, nlocC, 0);
  cp_async_commit();
  cp_async_wait_all();
  issue_gather(c, nlocC, 0);
  uint64_t cn = c + stride;
  bool hasN = cn < p.nChunks;
  int nlocN = 0;
  if (hasN)
  {
    const uint64_t a = p.node_off[cn], b = p.node_off[cn + 1];
    nlocN = (int)(b - a);
    issue_rec(a, nlocN, 1);
  }
  cp_async_commit();
  load_ranks(c);
  int buf = 0;

  while (true)
  {
    double *un = unb + buf * p.ncap;
    const int *jd = jdb + buf * 2 * p.jdStride;
    cp_async_wait_all();
    __syncthreads();  // A: un/jd of this chunk visible, own records of the next chunk landed; X and un[buf^1] free
    const uint64_t cnn = cn + stride;
    const bool hasNN = hasN && cnn < p.nChunks;
    uint64_t oaNN = 0;
    int nlocNN = 0;
    if (hasN)
    {
      issue_gather(cn, nlocN, buf ^ 1);
      if (hasNN)
      {
        oaNN = p.node_off[cnn];
        nlocNN = (int)(p.node_off[cnn + 1] - oaNN);
      }
    }
    cp_async_commit();
    // ---- T2: the groups of this chunk
    {
      const uint64_t u0 = c * (uint64_t)E;
      const int nu = (int)min((uint64_t)E, (uint64_t)p.nSet - u0);
      if (tid < nu)
      {
        uint32_t ps[NW];
        {
          const uint32_t *sw = p.ps16 + u0 * NW + tid;
#pragma unroll
          for (int j = 0; j < NW; j++) ps[j] = sw[(uint32_t)j * E];
        }
        const double s = p.lscale[levU];
        double v[LP], o[LP];
        if (!HANG)
        {
#pragma unroll
          for (int l = 0; l < LP; l++) v[l] = un[GRK(l)];
          // Walsh-Hadamard form: the butterflies along the XOR-permuted dimensions are the same for every child of the
          // group, so they run ONCE on the lattice (forward here, backward on the summed output); the children then only
          // transform along the grouped dimensions.  (H D H commutes with the XOR permutation: its signs in the
          // frequency domain cancel around the diagonal D.)
          constexpr bool RSHARE = (OPKIND == OP_HADAMARD) && (DKT_GRP_RSHARE != 0) && (G < DIM);
          if (RSHARE) grp_wht_lattice<DIM, G>(v);
#pragma unroll
          for (int cG = 0; cG < NC; cG++)
          {
            double e[N];
#pragma unroll
            for (int r = 0; r < N; r++) e[r] = v[GP::lat(cG, r)];
            if (OPKIND == OP_HADAMARD)
            {
              wht_range<N, 1, (RSHARE ? NC : N)>(e);
#pragma unroll
              for (int i = 0; i < N; i++) e[i] *= p.K[i] * s;
              wht_range<N, 1, (RSHARE ? NC : N)>(e);
            }
#pragma unroll
            for (int r = 0; r < N; r++)
            {
              if (GP::first(cG, r)) o[GP::lat(cG, r)] = e[r];
              else o[GP::lat(cG, r)] += e[r];
            }
          }
          if (RSHARE) grp_wht_lattice<DIM, G>(o);
#pragma unroll
          for (int l = 0; l < LP; l++) X[GPS(l)] = o[l];
        }
        else
        {
          {
            // parent nodes -> whole lattice: subset sums along the XOR-permuted dimensions, midpoints along the grouped ones
            double par[N];
#pragma unroll
            for (int t = 0; t < N; t++) par[t] = un[GRK(LP + t)];
#pragma unroll
            for (int b = NC; b < N; b <<= 1)
#pragma unroll
              for (int i = 0; i < N; i++)
              {
                if (i & b) continue;
                par[i | b] = 0.5 * (par[i] + par[i | b]);
              }
            if constexpr (G == 1) grp_expand<DIM, G, 0>(par, v);
            else if constexpr (G == 2)
            {
              double t1[3 * (N / 2)];
              grp_expand<DIM, G, 0>(par, t1);
              grp_expand<DIM, G, 1>(t1, v);
            }
            else
            {
              static_assert(G <= 3, "grouped dimensions");
              double t1[3 * (N / 2)], t2[9 * (N / 4)];
              grp_expand<DIM, G, 0>(par, t1);
              grp_expand<DIM, G, 1>(t1, t2);
              grp_expand<DIM, G, 2>(t2, v);
            }
          }
#pragma unroll
          for (int l = 0; l < LP; l++)
          {
            const double own = un[GRK(l)];
            if ((fm >> l) & 1ull) v[l] = own;
          }
          double ta[N];
#pragma unroll
          for (int q = 0; q < N; q++) ta[q] = 0.0;
#pragma unroll
          for (int cG = 0; cG < NC; cG++)
          {
            double e[N];
#pragma unroll
            for (int r = 0; r < N; r++) e[r] = v[GP::lat(cG, r)];
            if (OPKIND == OP_HADAMARD)
            {
              wht<N>(e);
#pragma unroll
              for (int i = 0; i < N; i++) e[i] *= p.K[i] * s;
              wht<N>(e);
            }
#pragma unroll
            for (int r = 0; r < N; r++)
            {
              if (GP::first(cG, r)) o[GP::lat(cG, r)] = e[r];
              else o[GP::lat(cG, r)] += e[r];
            }
            // a child without hanging nodes sends nothing to the parent nodes.  Neighbouring groups hang on the same
            // coarse face, so the branch is mostly warp-uniform.
            if ((~fm & GP::child_mask(cG)) == 0ull) continue;
#pragma unroll
            for (int r = 0; r < N; r++)
              if ((fm >> GP::lat(cG, r)) & 1ull) e[r] = 0.0;  // nullify prior to back-interpolation (matvec.h:497-499)
            // transposed interpolation of this child: A0^T along permuted dimensions and grouped ones with bit 0, A1^T otherwise
#pragma unroll
            for (int d = 0; d < DIM; d++)
            {
              const int b = 1 << d;
              const bool a1 = (d < G) && ((cG >> d) & 1);
#pragma unroll
              for (int i = 0; i < N; i++)
              {
                if (i & b) continue;
                if (!a1)
                {
                  const double h = 0.5 * e[i | b];
                  e[i] += h;
                  e[i | b] = h;
                }
                else
                {
                  const double h = 0.5 * e[i];
                  e[i | b] += h;
                  e[i] = h;
                }
              }
            }
            // quirk Q1: the LEAF's fill flag of rank q masks the contribution to the PARENT's rank q (matvec.h:517)
#pragma unroll
            for (int q = 0; q < N; q++)
            {
              ta[q] += ((fm >> GP::lat(cG, q)) & 1ull) ? 0.0 : e[q];
            }
          }
#pragma unroll
          for (int l = 0; l < LP; l++) X[GPS(l)] = o[l];
#pragma unroll
          for (int t = 0; t < N; t++) X[GPS(LP + t)] = ta[t];
        }
      }
    }
    if (hasN) load_ranks(cn);
    __syncthreads();  // B: X complete
    // ---- T4: own nodes of this chunk
    {
      // four nodes per thread at a time: independent accumulation chains, one jd[j] / cnt[j] load for the four.  Nodes
      // are ranked by run length (descending) and cnt[j] = #nodes with a run longer than j, so node n has a j-th
      // contribution iff n < cnt[j]; nodes >= cnt[0] are only read by this chunk.
      const uint32_t *rec = recb + buf * p.ncap;
      const int *cnt = jd + p.jdStride;
      const int cnt0 = cnt[0];
      for (int base = tid; base < cnt0; base += 4 * TPB)
      {
        uint32_t r[4];
        double acc[4];
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
          const int n = base + i * TPB;
          r[i] = 0u;
          acc[i] = 0.0;
          if (n < cnt0)
          {
            r[i] = rec[n];
            acc[i] = X[n];  // jd[0] == 0
          }
        }
        for (int j = 1; base < cnt[j]; j++)
        {
          const int off = jd[j] + base, cj = cnt[j];
#pragma unroll
          for (int i = 0; i < 4; i++)
            if (base + i * TPB < cj) acc[i] += X[off + i * TPB];
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
          if (base + i * TPB >= cnt0) continue;
          if (DIRI && (r[i] & REC_BDY)) continue;
          if (r[i] & REC_SHARED) atomicAdd(p.out + (r[i] & REC_GID), acc[i]);
          else p.out[r[i] & REC_GID] = acc[i];
        }
      }
    }
    if (!hasN) break;
    if (hasNN) issue_rec(oaNN, nlocNN, buf);  // own records of the chunk after the next one
    cp_async_commit();
    c = cn;
    cn = cnn;
    hasN = hasNN;
    nlocC = nlocN;
    nlocN = nlocNN;
    buf ^= 1;
  }
#undef GRK
#undef GPS
}

template <int DIM, int G, int OPKIND, bool DIRI, bool HANG>
static int launch_group_one(DA &da, const ChunkSet &cs, Mv3Params<DIM, 1> &p)
{
  using GP = Grp<DIM, G>;
  constexpr int SPU = (GP::LP + (HANG ? GP::N : 0) + 1) & ~1;
  constexpr int TPB = grp_tpb(SPU, DIM, G);
  if (cs.spu != SPU || (int)cs.elemsPerChunk > TPB) { set_error("internal: group set does not match its kernel"); return DKT_ERR_INVALID; }
  p.rk16 = (const uint32_t *)cs.d_rk16; p.ps16 = (const uint32_t *)cs.d_ps16; p.rec = (const uint32_t *)cs.d_rec; p.jd = cs.d_jd;
  p.node_off = cs.d_node_off; p.lev = cs.lev; p.fmask64 = cs.fmask64;
  p.nSet = (uint32_t)cs.nElem; p.nChunks = cs.nChunks; p.elemsPerChunk = cs.elemsPerChunk;
  p.xcap = (uint32_t)cs.elemsPerChunk * SPU + 258u;  // + padding of the first 16 diagonals + the trash position
  p.ncap = (cs.maxNloc + 2) & ~1u;
  p.jdStride = cs.jdStride;
  const size_t smem = ((size_t)p.xcap + 2 * (size_t)p.ncap) * sizeof(double) + 2 * (size_t)p.ncap * sizeof(uint32_t) +
                      4 * (size_t)p.jdStride * sizeof(int);
  auto kern = k_mvg<DIM, G, OPKIND, DIRI, HANG, TPB>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int perSM = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, TPB, smem));
  if (perSM < 1) { set_error("group kernel does not fit on an SM"); return DKT_ERR_CUDA; }
  const uint32_t grid = std::min<uint32_t>(cs.nChunks, (uint32_t)(perSM * da.numSMs));
  DKT_LAUNCH(kern, grid, TPB, smem, da.cur ? da.cur : da.stream)(p);
  g_launches++;
  return DKT_OK;
}

template <int DIM, int ORDER, int OPKIND, bool DIRI>
static int launch_group(DA &da, const ChunkSet &cs, Mv3Params<DIM, ORDER> &p)
{
  if constexpr (ORDER == 1 && (OPKIND == DKT_OP_IDENTITY || OPKIND == OP_HADAMARD))
  {
    const bool hang = cs.rows == 2;
#define GRP_CASE(D, GG)                                                                      \
  if constexpr (DIM == D)                                                                    \
  {                                                                                          \
    if (cs.g == GG)                                                                          \
      return hang ? launch_group_one<DIM, GG, OPKIND, DIRI, true>(da, cs, p)                 \
                  : launch_group_one<DIM, GG, OPKIND, DIRI, false>(da, cs, p);               \
  }
    GRP_CASE(4, 1)
    GRP_CASE(4, 2)
    GRP_CASE(4, 3)
    GRP_CASE(3, 2)
    GRP_CASE(3, 3)
    GRP_CASE(2, 2)
#undef GRP_CASE
  }
  set_error("internal: no sibling-group kernel for this (dim, g, operator)");
  return DKT_ERR_UNSUPPORTED;
}

template <int DIM, int ORDER, int OPKIND, bool DIRI>
static int launch_mv3(DA &da, Mv3Params<DIM, ORDER> &p, unsigned phaseMask)
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
  for (const ChunkSet &cs : da.sets)
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
    if (cs.kind == 1) rc = launch_group<DIM, ORDER, OPKIND, DIRI>(da, cs, p);
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
  if (da.groups && !((hadamard || op->kind == DKT_OP_IDENTITY) && p.exact_ip))
    return run_matvec(da, op, d_in, d_out, scale, flags);  // group tables serve the fast forms only (DKT_GROUPS is opt-in)
  if (zeroOut)
  {
    CK(cudaMemsetAsync(d_out, 0, da.nNodes * sizeof(double), da.stream));
    g_launches++;
  }
  const bool diri = op->dirichlet != 0;
  if constexpr (ORDER == 1)
  {
    if (hadamard)
      return diri ? launch_mv3<DIM, ORDER, OP_HADAMARD, true>(da, p, phaseMask) : launch_mv3<DIM, ORDER, OP_HADAMARD, false>(da, p, phaseMask);
  }
  if (op->kind == DKT_OP_IDENTITY)
    return diri ? launch_mv3<DIM, ORDER, DKT_OP_IDENTITY, true>(da, p, phaseMask) : launch_mv3<DIM, ORDER, DKT_OP_IDENTITY, false>(da, p, phaseMask);
  if (op->kind == DKT_OP_DENSE)
    return diri ? launch_mv3<DIM, ORDER, DKT_OP_DENSE, true>(da, p, phaseMask) : launch_mv3<DIM, ORDER, DKT_OP_DENSE, false>(da, p, phaseMask);
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
