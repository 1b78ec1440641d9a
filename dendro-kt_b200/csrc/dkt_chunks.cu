// The production matvec: SFC-contiguous element CHUNKS staged through shared memory by
// persistent, software-pipelined CTAs.
//
// Why: the flat kernels (dkt_matvec.cu) issue (order+1)^dim scattered 8-byte gathers and fp64
// atomics per element straight to L1/L2 and are L1TEX/atomic bound at ~15 % of the HBM roofline
// (profiles/r01_*).  Shared-memory fp64 atomics are CAS spin loops on sm_100a
// (ATOMS.CAST.SPIN.64), so the in-chunk reduction is made atomic-free instead.
//
//   build (once per DA, k_chunk_build, one CTA per chunk, cub::BlockRadixSort in shared memory):
//     * the chunk's slots (element, rank) are sorted by the node they touch -> unique nodes,
//       run length `len` of each node
//     * nodes are re-ranked by (len descending, id ascending): jagged-diagonal storage.  The k-th
//       contribution to node n lives at X[jd[k] + n]; jd[k] = number of (node, j<k) pairs.
//     * every slot gets one 32-bit word  n | (jd[k] + n) << 16 ; words are stored rank-major
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
#include "dkt_internal.h"

#include <cub/block/block_radix_sort.cuh>
#include <cub/block/block_scan.cuh>

#include <algorithm>
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

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;
constexpr int SLOT_CAP = SORT_THREADS * SORT_ITEMS;  // 4096 slots per chunk
constexpr int MAX_LEN = 511;                         // run length of a node inside a chunk (9 bits)
constexpr uint32_t META_LEN = 0x1FFu;
constexpr uint32_t META_PRESENT = 0x2000u;  // node exists (its run may be empty: only read by this chunk)
constexpr uint32_t META_BDY = 0x4000u;
constexpr uint32_t META_SHARED = 0x8000u;

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

// ------------------------------------------------------------------------------------------
// build
// ------------------------------------------------------------------------------------------
// Number of WRITING references of every node.  Own-lattice slots always write; a parent-lattice slot q
// of a hanging element writes only if the element's own rank q is unfilled (quirk Q1 is static in the
// chunk tables: FEM/include/matvec.h:517) - `own` is the element's e2n row, null for the own rows.
__global__ void k_ref_count(const uint32_t *ids, const uint32_t *own, uint64_t n, uint32_t *cnt)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n || ids[i] == INVALID) return;
  if (own && own[i] != INVALID) return;  // read-only parent slot
  atomicAdd(cnt + ids[i], 1u);
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

// One CTA per chunk.  WRITE == false: only report the chunk's node count and longest run.
template <bool WRITE>
__global__ void __launch_bounds__(SORT_THREADS)
k_chunk_build(const uint32_t *e2n, const uint32_t *pnode, const uint32_t *mv_xyz, const uint8_t *mv_lev, int dim, int max_depth,
              int xorperm, uint64_t elem0, uint64_t nSet, int N, int rows, int elemsPerChunk, const uint32_t *refcnt,
              const uint8_t *isbdy, const uint64_t *node_off, int jdStride, uint32_t *nloc_out, uint32_t *maxlen_out,
              uint32_t *slot, uint32_t *gid_out, uint16_t *meta_out, uint16_t *jd_out)
{
  using SortPairs = cub::BlockRadixSort<uint32_t, SORT_THREADS, SORT_ITEMS, uint16_t>;
  using SortKeys = cub::BlockRadixSort<uint32_t, SORT_THREADS, SORT_ITEMS>;
  using Scan = cub::BlockScan<int, SORT_THREADS>;
  __shared__ union
  {
    typename SortPairs::TempStorage pairs;
    typename SortKeys::TempStorage keys;
    typename Scan::TempStorage scan;
    uint16_t newrank[SLOT_CAP];                // by gid-rank: rank after the (len desc) re-sort (after the sorts)
  } tmp;
  uint16_t *s_newrank = tmp.newrank;
  __shared__ uint32_t s_last[SORT_THREADS];
  __shared__ uint16_t s_start[SLOT_CAP + 1];   // by gid-rank: first sorted position of the node's references
  __shared__ uint16_t s_cw[SLOT_CAP + 1];      // by gid-rank: number of WRITING references sorted before the node
  __shared__ int s_hist[MAX_LEN + 2];
  __shared__ int s_jd[MAX_LEN + 2];
  __shared__ int s_total, s_P, s_maxlen;

  const uint64_t c = blockIdx.x;
  const uint64_t e0 = c * (uint64_t)elemsPerChunk;
  const int ne = (int)min((uint64_t)elemsPerChunk, nSet - e0);
  const int spe = rows * N;  // slots per element
  const int nslots = ne * spe;

  uint32_t key[SORT_ITEMS];
  uint16_t val[SORT_ITEMS];
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++)
  {
    const int s = threadIdx.x * SORT_ITEMS + i;
    key[i] = INVALID;
    val[i] = (uint16_t)s;
    if (s < nslots)
    {
      const uint64_t e = e0 + s / spe;  // element index inside the set
      const int q = s % spe;
      key[i] = q < N ? e2n[(elem0 + e) * N + q] : pnode[e * (uint64_t)N + (q - N)];
    }
  }
  if (threadIdx.x == 0) { s_P = 0; s_maxlen = 0; }
  for (int i = threadIdx.x; i < MAX_LEN + 2; i += SORT_THREADS) s_hist[i] = 0;
  SortPairs(tmp.pairs).Sort(key, val);
  __syncthreads();
  // heads: first occurrence of each node id in sorted order (invalid keys sort last)
  s_last[threadIdx.x] = key[SORT_ITEMS - 1];
  __syncthreads();
  uint32_t prev = threadIdx.x ? s_last[threadIdx.x - 1] : INVALID;
  int head[SORT_ITEMS];
  int nheads = 0, lastvalid = 0;
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++)
  {
    const bool first = (threadIdx.x == 0 && i == 0);
    head[i] = (key[i] != INVALID) && (first || key[i] != prev);
    prev = key[i];
    nheads += head[i];
    if (key[i] != INVALID) lastvalid = threadIdx.x * SORT_ITEMS + i + 1;
  }
  int before = 0, total = 0;
  Scan(tmp.scan).ExclusiveSum(nheads, before, total);
  if (lastvalid) atomicMax(&s_P, lastvalid);
  uint16_t rank0[SORT_ITEMS];
  {
    int rk = before - 1;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++)
    {
      if (head[i])
      {
        rk++;
        s_start[rk] = (uint16_t)(threadIdx.x * SORT_ITEMS + i);
      }
      rank0[i] = (uint16_t)(rk < 0 ? 0 : rk);
    }
  }
  if (threadIdx.x == 0) s_total = total;
  // writing references: every valid own-lattice slot; parent-lattice slots only where the element's
  // own rank is unfilled (static Q1).  k-th WRITING reference of a node -> diagonal k.
  int wr[SORT_ITEMS];
  int nwr = 0;
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++)
  {
    wr[i] = 0;
    if (key[i] != INVALID)
    {
      const int q = val[i] % spe, el = val[i] / spe;
      wr[i] = (q < N) ? 1 : (e2n[(elem0 + e0 + el) * N + (q - N)] == INVALID);
    }
    nwr += wr[i];
  }
  __syncthreads();
  int wbefore = 0, wtotal = 0;
  Scan(tmp.scan).ExclusiveSum(nwr, wbefore, wtotal);
  uint16_t cwi[SORT_ITEMS];
  {
    int acc = wbefore;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++)
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
  uint32_t k2[SORT_ITEMS];
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++)
  {
    const int n = threadIdx.x * SORT_ITEMS + i;
    k2[i] = INVALID;
    if (n < nloc)
    {
      const int len = (int)s_cw[n + 1] - (int)s_cw[n];  // writing references only
      k2[i] = ((uint32_t)(MAX_LEN - min(len, MAX_LEN)) << 12) | (uint32_t)n;
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
  SortKeys(tmp.keys).Sort(k2, 0, 21);
  __syncthreads();  // tmp is re-used as newrank[]
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++)
  {
    const int j = threadIdx.x * SORT_ITEMS + i;
    if (k2[i] != INVALID) s_newrank[k2[i] & 0xFFFu] = (uint16_t)j;
  }
  // jd[k] = sum_{j<k} count_j, count_j = #nodes with len > j
  if (threadIdx.x == 0)
  {
    const int ml = s_maxlen;
    int above = 0;
    for (int k = ml; k >= 0; k--)
    {
      s_jd[k] = above;  // #nodes with len > k
      above += s_hist[k];
    }
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
  for (int k = threadIdx.x; k < jdStride; k += SORT_THREADS) jd_out[c * (uint64_t)jdStride + k] = (uint16_t)s_jd[min(k, MAX_LEN + 1)];
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++)
  {
    const int posn = threadIdx.x * SORT_ITEMS + i;
    // rank-major inside the chunk; with the XOR schedule slot s of an element with Morton child
    // number c holds lattice rank s ^ c (both rows), so that siblings touch the SAME node in the
    // same instruction (shared-memory broadcast) - see k_mv3
    int q = val[i] % spe;
    const int el = val[i] / spe;
    if (xorperm && val[i] < nslots)
    {
      const uint64_t ge = elem0 + e0 + el;
      const int L = mv_lev[ge];
      int cnum = 0;
      for (int d = 0; d < dim; d++) cnum |= ((mv_xyz[ge * dim + d] >> (max_depth - L)) & 1u) << d;
      q = (q < N) ? (q ^ cnum) : (N + ((q - N) ^ cnum));
    }
    const uint64_t dst = e0 * spe + (uint64_t)q * elemsPerChunk + el;
    // absent node: read the chunk's zero entry un[nloc], write to the trash position behind the diagonals
    const uint32_t trash = (uint32_t)s_jd[MAX_LEN + 1];
    if (key[i] == INVALID)
    {
      if (val[i] < nslots) slot[dst] = (uint32_t)nloc | (trash << 16);
      continue;
    }
    const int n0 = rank0[i];
    const int k = (int)cwi[i] - (int)s_cw[n0];
    const int nr = s_newrank[n0];
    slot[dst] = (uint32_t)nr | ((wr[i] ? (uint32_t)(s_jd[k] + nr) : trash) << 16);
    (void)posn;
    if (head[i])
    {
      const int len = (int)s_cw[n0 + 1] - (int)s_cw[n0];
      uint32_t m = (uint32_t)len | META_PRESENT;
      if (refcnt[key[i]] != (uint32_t)len) m |= META_SHARED;
      if (isbdy[key[i]]) m |= META_BDY;
      gid_out[noff + nr] = key[i];
      meta_out[noff + nr] = (uint16_t)m;
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
  for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, v);
}

static int build_set(DA &da, ChunkSet &cs, uint64_t elem0, uint64_t hang0, uint64_t nSet, int rows, int phase, const uint32_t *refcnt)
{
  cs = ChunkSet();
  cs.rows = rows;
  cs.phase = phase;
  cs.elem0 = elem0;
  cs.hang0 = hang0;
  cs.nElem = nSet;
  if (nSet == 0) return DKT_OK;
  const int N = da.N;
  cs.elemsPerChunk = rows_per_chunk(N) / rows;
  cs.nChunks = (uint32_t)((nSet + cs.elemsPerChunk - 1) / cs.elemsPerChunk);
  uint32_t *nloc = nullptr, *mlen = nullptr;
  uint64_t *wide = nullptr, *off = nullptr;
  CK(cudaMalloc((void **)&nloc, (size_t)cs.nChunks * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&mlen, (size_t)cs.nChunks * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&wide, ((size_t)cs.nChunks + 1) * sizeof(uint64_t)));
  CK(cudaMalloc((void **)&off, ((size_t)cs.nChunks + 1) * sizeof(uint64_t)));
  const int xorperm = da.order == 1 ? 1 : 0;
  cs.xorperm = xorperm;
  k_chunk_build<false><<<cs.nChunks, SORT_THREADS, 0, da.stream>>>(da.d_e2n, da.d_pnode + hang0 * N, da.d_mv_xyz, da.d_mv_lev, da.dim, da.max_depth,
                                                                    xorperm, elem0, nSet, N, rows, cs.elemsPerChunk, refcnt,
                                                                    da.d_node_isbdy, nullptr, 0, nloc, mlen, nullptr, nullptr, nullptr,
                                                                    nullptr);
  g_launches++;
  CK(cudaMemsetAsync(wide, 0, ((size_t)cs.nChunks + 1) * sizeof(uint64_t), da.stream));
  k_u32_widen<<<(cs.nChunks + 255) / 256, 256, 0, da.stream>>>(nloc, cs.nChunks, wide);
  g_launches++;
  int rc = device_exclusive_scan(da, wide, off, (uint64_t)cs.nChunks + 1);
  if (rc) return rc;
  uint64_t total = 0;
  CK(cudaMemcpy(&total, off + cs.nChunks, sizeof(uint64_t), cudaMemcpyDeviceToHost));
  uint32_t *dmax = nullptr, hmax[2] = {0, 0};
  CK(cudaMalloc((void **)&dmax, 2 * sizeof(uint32_t)));
  CK(cudaMemsetAsync(dmax, 0, 2 * sizeof(uint32_t), da.stream));
  k_max_u32<<<(cs.nChunks + 255) / 256, 256, 0, da.stream>>>(nloc, cs.nChunks, dmax);
  k_max_u32<<<(cs.nChunks + 255) / 256, 256, 0, da.stream>>>(mlen, cs.nChunks, dmax + 1);
  g_launches += 2;
  CK(cudaMemcpyAsync(hmax, dmax, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, da.stream));
  CK(cudaStreamSynchronize(da.stream));
  cs.maxNloc = hmax[0];
  cs.maxLen = hmax[1];
  if (cs.maxLen > (uint32_t)MAX_LEN) { set_error("internal: node referenced more than 511 times inside one chunk"); return DKT_ERR_UNSUPPORTED; }
  cs.jdStride = (cs.maxLen + 1 + 7) & ~7u;
  cs.totalNodes = total;
  cs.d_node_off = off;
  CK(cudaMalloc((void **)&cs.d_slot, (size_t)cs.nChunks * cs.elemsPerChunk * rows * N * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&cs.d_gid, std::max<uint64_t>(total, 1) * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&cs.d_meta, std::max<uint64_t>(total, 1) * sizeof(uint16_t)));
  CK(cudaMalloc((void **)&cs.d_jd, (size_t)cs.nChunks * cs.jdStride * sizeof(uint16_t)));
  k_chunk_build<true><<<cs.nChunks, SORT_THREADS, 0, da.stream>>>(da.d_e2n, da.d_pnode + hang0 * N, da.d_mv_xyz, da.d_mv_lev, da.dim, da.max_depth,
                                                                   xorperm, elem0, nSet, N, rows, cs.elemsPerChunk, refcnt,
                                                                   da.d_node_isbdy, off, (int)cs.jdStride, nullptr, nullptr, cs.d_slot,
                                                                   cs.d_gid, cs.d_meta, cs.d_jd);
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
  }
  da.sets.clear();
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

int build_chunks(DA &da)
{
  CK(cudaMalloc((void **)&da.d_mv_child, std::max<uint64_t>(da.nMv, 1)));
  k_child_numbers<<<(unsigned)((da.nMv + 255) / 256), 256, 0, da.stream>>>(da.d_mv_xyz, da.d_mv_lev, da.nMv, da.dim, da.max_depth,
                                                                           da.d_mv_child);
  g_launches++;
  uint32_t *refcnt = nullptr;
  CK(cudaMalloc((void **)&refcnt, std::max<uint64_t>(da.nNodes, 1) * sizeof(uint32_t)));
  CK(cudaMemsetAsync(refcnt, 0, std::max<uint64_t>(da.nNodes, 1) * sizeof(uint32_t), da.stream));
  const uint64_t n1 = da.nMv * (uint64_t)da.N, n2 = da.nHang * (uint64_t)da.N;
  if (n1) { k_ref_count<<<(unsigned)((n1 + 255) / 256), 256, 0, da.stream>>>(da.d_e2n, nullptr, n1, refcnt); g_launches++; }
  if (n2)
  {
    k_ref_count<<<(unsigned)((n2 + 255) / 256), 256, 0, da.stream>>>(da.d_pnode, da.d_e2n + da.nReg * (uint64_t)da.N, n2, refcnt);
    CK(cudaMalloc((void **)&da.d_fmask, da.nHang * sizeof(uint32_t)));
    k_fmask<<<(unsigned)((da.nHang + 255) / 256), 256, 0, da.stream>>>(da.d_e2n + da.nReg * (uint64_t)da.N, da.d_mv_child + da.nReg, da.nHang,
                                                                       da.N, da.order == 1 ? 1 : 0, da.d_fmask);
    g_launches += 2;
  }
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
  int rc = DKT_OK;
  for (const Range &r : rr)
  {
    if (rc != DKT_OK || r.b <= r.a) continue;
    da.sets.emplace_back();
    rc = build_set(da, da.sets.back(), r.a, 0, r.b - r.a, 1, r.phase, refcnt);
  }
  for (const Range &r : hr)
  {
    if (rc != DKT_OK || r.b <= r.a) continue;
    da.sets.emplace_back();
    rc = build_set(da, da.sets.back(), da.nReg + r.a, r.a, r.b - r.a, 2, r.phase, refcnt);
  }
  cudaFree(refcnt);
  int dev = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&da.numSMs, cudaDevAttrMultiProcessorCount, dev));
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

__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc)
{
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

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
  extern __shared__ double sm[];
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
static int launch_one(DA &da, const ChunkSet &cs, Mv3Params<DIM, ORDER> &p, const uint8_t *lev, const uint8_t *child)
{
  constexpr int N = Mv3Params<DIM, ORDER>::N;
  p.slot = cs.d_slot; p.gid = cs.d_gid; p.meta = cs.d_meta; p.jd = cs.d_jd; p.node_off = cs.d_node_off;
  p.lev = lev; p.child = child; p.nSet = (uint32_t)cs.nElem; p.nChunks = cs.nChunks; p.elemsPerChunk = cs.elemsPerChunk;
  p.xcap = (uint32_t)rows_per_chunk(N) * N + 258u;  // + padding of the first 16 diagonals + the trash position
  p.ncap = (cs.maxNloc + 2) & ~1u;
  p.jdStride = cs.jdStride;
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
  kern<<<grid, TPB, smem, da.stream>>>(p);
  g_launches++;
  return DKT_OK;
}

template <int DIM, int ORDER, int OPKIND, bool DIRI>
static int launch_mv3(DA &da, Mv3Params<DIM, ORDER> &p, unsigned phaseMask)
{
  constexpr int N = Mv3Params<DIM, ORDER>::N;
  constexpr int TPB_R = (N == 27) ? 160 : DKT_ROWS;  // >= elements per chunk (one element per thread)
  constexpr int TPB_H = (N == 27) ? 96 : DKT_ROWS / 2;
  constexpr bool CAN_EXIP = (ORDER == 1 && OPKIND != DKT_OP_DENSE);
  const bool exip = CAN_EXIP && p.exact_ip;
  for (const ChunkSet &cs : da.sets)
  {
    if (!cs.nChunks || !((phaseMask >> cs.phase) & 1u)) continue;
    const uint8_t *lev = da.d_mv_lev + cs.elem0, *child = da.d_mv_child + cs.elem0;
    p.fmask = da.d_fmask ? da.d_fmask + cs.hang0 : nullptr;
    int rc = DKT_OK;
    if (cs.rows == 1)
    {
      if (cs.maxNloc <= 6u * TPB_R) rc = launch_one<DIM, ORDER, OPKIND, DIRI, false, TPB_R, 6, false>(da, cs, p, lev, child);
      else rc = launch_one<DIM, ORDER, OPKIND, DIRI, false, TPB_R, 16, false>(da, cs, p, lev, child);
    }
    else if (cs.maxNloc <= 8u * TPB_H)
    {
      if (exip) rc = launch_one<DIM, ORDER, OPKIND, DIRI, true, TPB_H, 8, CAN_EXIP>(da, cs, p, lev, child);
      else rc = launch_one<DIM, ORDER, OPKIND, DIRI, true, TPB_H, 8, false>(da, cs, p, lev, child);
    }
    else rc = launch_one<DIM, ORDER, OPKIND, DIRI, true, TPB_H, 32, false>(da, cs, p, lev, child);
    if (rc) return rc;
  }
  CK(cudaGetLastError());
  return DKT_OK;
}

template <int DIM, int ORDER>
static int run_typed3(DA &da, const dkt_op *op, const double *d_in, double *d_out, double scale, unsigned flags, unsigned phaseMask,
                      bool zeroOut)
{
  using P = Mv3Params<DIM, ORDER>;
  static P p;
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
      static double T[N * N], D[N * N];
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
