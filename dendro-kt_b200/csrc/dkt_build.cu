// Construction of the flat matvec tables on the GPU.
//
// Replaces, for one rank, what the reference does on the host at DA construction and then
// re-discovers inside every matvec:
//   * tree order ............ SFC_Tree::locTreeSort (include/tsort.tcc:17-80,191-334)
//   * lattice nodes ......... Element::appendNodes (include/nsort.tcc:313-389)
//   * CG node set ........... SFC_NodeSort::countCGNodes + resolveInterface_lowOrder
//                             (include/nsort.tcc:879-944,1267-1313); boundary nodes skip the
//                             hanging test (quirk Q2, :913-924)
//   * CG node ORDER ......... countCGNodes_impl / bucketByHyperplane / locTreeSortAsPoints
//                             (include/nsort.tcc:1155-1260, 997-1105), src/oda.cpp:90-121,
//                             expressed as ONE composite radix key per node
//   * element -> node ....... the bucketing of FEM/include/matvec.h:75-199 + get_lexNodeRank
//                             (include/nsort.tcc:284-303), done once instead of per matvec
//   * hanging tables ........ FEM/include/matvec.h:403-457 (parent nodes of level L-1)
//   * phantom children ...... FEM/include/matvec.h:102-107,340-362 (quirk Q3)
// Sorting/scans use CUB (construction is one-off; the hot path is in dkt_matvec.cu).
// Like dkt_chunks.cu this file also compiles under -DDKT_EMU (tests/emu/cuda_emu.h) so that the CPU test-suite can run
// the table construction against the reference's golden fixtures; the product is never built that way.
#include "dkt_internal.h"

#ifdef DKT_EMU
#include "cuda_emu.h"
#else
#include <cub/cub.cuh>
#endif

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>

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

// SFC tables of the DA being built (build_da is synchronous, one at a time per process).
__constant__ uint8_t c_rot_inv[192 * 16];
__constant__ uint8_t c_htab[192 * 16];

template <typename T>
struct Buf
{
  T *p = nullptr;
  size_t n = 0;
  ~Buf() { release(); }
  void release()
  {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  cudaError_t alloc(size_t count)
  {
    release();
    n = count;
    return cudaMalloc((void **)&p, std::max<size_t>(count, 1) * sizeof(T));
  }
  T *take()
  {
    T *r = p;
    p = nullptr;
    n = 0;
    return r;
  }
};

static inline unsigned nblk(uint64_t n, unsigned t = 256) { return (unsigned)((n + t - 1) / t); }
#ifdef DKT_EMU
#define DKT_BUILD_LAUNCH(kern, grid, stream) ::emu::make_launch_flat(kern, (grid), 256)
#else
#define DKT_BUILD_LAUNCH(kern, grid, stream) kern<<<(grid), 256, 0, (stream)>>>
#endif
#define LAUNCH(kern, n, ...)                                                       \
  do                                                                               \
  {                                                                                \
    if ((n) > 0)                                                                   \
    {                                                                              \
      DKT_BUILD_LAUNCH(kern, nblk(n), da.stream)(__VA_ARGS__);                     \
      g_launches++;                                                                \
    }                                                                              \
  } while (0)

struct Geo
{
  int dim, order, M, N, max_depth, lmax, lk, shift, bits, nch;
};

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------

// Key of a cell/point for the tree order: SFC digits of levels 1..lmax, `dim` bits each, most
// significant first; digits deeper than `own_lev` are zero (they never decide a comparison
// between leaves).  `depth` is the number of coordinate bits (max_depth, or max_depth+1 for
// doubled probe coordinates).
__device__ inline uint64_t tree_key(const uint32_t *x, int dim, int depth, int lmax, int own_lev)
{
  const int nch = 1 << dim;
  uint64_t key = 0;
  int rot = 0;
  for (int l = 1; l <= lmax; l++)
  {
    int dgt = 0;
    if (l <= own_lev)
    {
      int m = 0;
      for (int d = 0; d < dim; d++) m |= ((x[d] >> (depth - l)) & 1u) << d;
      dgt = c_rot_inv[rot * nch + m];
      rot = c_htab[rot * nch + m];
    }
    key = (key << dim) | (uint64_t)dgt;
  }
  return key;
}

__device__ inline uint64_t pack_coords(const uint32_t *x, const Geo &g)
{
  uint64_t k = 0;
  for (int d = 0; d < g.dim; d++) k |= (uint64_t)(x[d] >> g.shift) << (g.bits * d);
  return k;
}
__device__ inline void unpack_coords(uint64_t k, const Geo &g, uint32_t *x)
{
  const uint64_t mask = (g.bits >= 64) ? ~0ull : ((1ull << g.bits) - 1);
  for (int d = 0; d < g.dim; d++) x[d] = (uint32_t)(((k >> (g.bits * d)) & mask) << g.shift);
}

// lattice coordinate of rank r of a cell (Element::appendNodes, include/nsort.tcc:313-333)
__device__ inline void lattice_point(const uint32_t *a, uint32_t len, int r, const Geo &g, uint32_t *x)
{
  for (int d = 0; d < g.dim; d++)
  {
    const int i = r % g.M;
    r /= g.M;
    x[d] = a[d] + (uint32_t)(((uint64_t)len * (uint64_t)i) / (uint64_t)g.order);
  }
}
__device__ inline bool rank_is_interior(int r, const Geo &g)
{
  for (int d = 0; d < g.dim; d++)
  {
    const int i = r % g.M;
    r /= g.M;
    if (i == 0 || i == g.order) return false;
  }
  return true;
}

// coordinate -> index into the sorted unique location table, or -1
__device__ inline int64_t find_location(const uint32_t *x, const Geo &g, const uint64_t *ukey, uint64_t nU)
{
  const uint32_t full = 1u << g.max_depth;
  for (int d = 0; d < g.dim; d++)
    if (x[d] > full || ((x[d] >> g.shift) << g.shift) != x[d]) return -1;
  const uint64_t k = pack_coords(x, g);
  uint64_t lo = 0, hi = nU;
  while (lo < hi)
  {
    const uint64_t mid = (lo + hi) >> 1;
    if (ukey[mid] < k) lo = mid + 1;
    else hi = mid;
  }
  return (lo < nU && ukey[lo] == k) ? (int64_t)lo : -1;
}

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------

__global__ void k_max_level(const uint8_t *lev, uint64_t n, int *out)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  int v = i < n ? lev[i] : 0;
#ifndef DKT_EMU
  for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(out, v);
#else
  if (v > 0) atomicMax(out, v);
#endif
}

__global__ void k_elem_keys(const uint32_t *xyz, const uint8_t *lev, uint64_t n, Geo g, uint64_t *key, uint32_t *idx)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t x[4];
  for (int d = 0; d < g.dim; d++) x[d] = xyz[i * g.dim + d];
  key[i] = tree_key(x, g.dim, g.max_depth, g.lmax, lev[i]);
  idx[i] = (uint32_t)i;
}

__global__ void k_gather_elems(const uint32_t *xyz_in, const uint8_t *lev_in, const uint32_t *perm, uint64_t n, int dim,
                               uint32_t *xyz, uint8_t *lev)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = perm[i];
  for (int d = 0; d < dim; d++) xyz[i * dim + d] = xyz_in[(uint64_t)s * dim + d];
  lev[i] = lev_in[s];
}

// one thread per (element, rank): packed coordinate key + instance id
__global__ void k_instances(const uint32_t *xyz, const uint8_t *lev, uint64_t nInst, Geo g, uint64_t *key, uint32_t *inst)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= nInst) return;
  const uint64_t e = i / g.N;
  const int r = (int)(i % g.N);
  uint32_t a[4], x[4];
  for (int d = 0; d < g.dim; d++) a[d] = xyz[e * g.dim + d];
  lattice_point(a, 1u << (g.max_depth - lev[e]), r, g, x);
  key[i] = pack_coords(x, g);
  inst[i] = (uint32_t)i;
}

// per unique location: the reference's keep / level / boundary decision
// flags: bit0 keep, bit1 boundary, bit2 q2 (boundary node hanging on a coarser leaf), bit3 element-interior
__global__ void k_resolve(const uint64_t *ukey, const uint32_t *ucnt, const uint64_t *uoff, uint64_t nU, const uint32_t *inst,
                          const uint8_t *elev, Geo g, uint8_t *ulev, uint8_t *uflag)
{
  uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (s >= nU) return;
  const uint32_t cnt = ucnt[s];
  const uint64_t o = uoff[s];
  int minlev = 255, maxlev = 0;
  bool interior = false;
  for (uint32_t j = 0; j < cnt; j++)
  {
    const uint32_t id = inst[o + j];
    const int l = elev[id / g.N];
    minlev = min(minlev, l);
    maxlev = max(maxlev, l);
    if (g.order > 1 && rank_is_interior((int)(id % g.N), g)) interior = true;
  }
  uint32_t x[4];
  unpack_coords(ukey[s], g, x);
  const uint32_t full = 1u << g.max_depth;
  int nb = 0, cdim = 0;
  const uint32_t lmask = (1u << (g.max_depth - minlev)) - 1u;
  for (int d = 0; d < g.dim; d++)
  {
    nb += (x[d] == 0 || x[d] == full);
    cdim += (x[d] & lmask) != 0;
  }
  const bool mixed = minlev != maxlev;
  const bool rule = mixed || cnt == (1u << (g.dim - cdim));             // resolveInterface_lowOrder
  const int ex = g.dim - cdim - nb;
  const bool complete = mixed || cnt == (1u << (ex > 0 ? ex : 0));      // every touching leaf owns it
  const bool bdy = nb > 0;
  uint8_t f = 0;
  if (bdy || rule) f |= 1;
  if (bdy) f |= 2;
  if (bdy && !complete) f |= 4;
  if (interior) f |= 8;
  ulev[s] = (uint8_t)minlev;
  uflag[s] = f;
}

// 128-bit append helper
struct Key128
{
  uint64_t hi, lo;
  __device__ inline void push(uint32_t sym, int nb)
  {
    hi = (hi << nb) | (lo >> (64 - nb));
    lo = (lo << nb) | (uint64_t)sym;
  }
};

// Composite order key of a kept node (right-aligned in 128 bits):
//   non-boundary exterior: 0 | d_1..d_lc | T | hp | d_{lc+1}..d_lk
//   boundary:              1 | rot_inv0[bit_maxDepth(x)] | d_1..d_lk | 0
//   element-interior:      2 | element index | rank
// symbols are (dim+1) bits; T = 2^dim sorts after every child digit (children before own
// interface, include/nsort.tcc:1194-1198); hp = first incident hyperplane (:138-151).
__global__ void k_order_keys(const uint64_t *ukey, const uint8_t *uflag, const uint64_t *uoff, const uint32_t *inst,
                             const uint64_t *kpos, uint64_t nU, Geo g, uint64_t *khi, uint64_t *klo, uint32_t *kseg)
{
  uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (s >= nU) return;
  const uint8_t f = uflag[s];
  if (!(f & 1)) return;
  const uint64_t k = kpos[s];
  const int sb = g.dim + 1;
  const int nsym = g.lk + 2;
  const int field = max(nsym * sb, 40);  // width of the part below the 2-bit flag
  Key128 key{0, 0};
  if (f & 8)
  {
    const uint32_t id = inst[uoff[s]];
    key.push(2u, 2);
    for (int i = 0; i < field - 40; i += 8) key.push(0u, min(8, field - 40 - i));
    key.push(id / g.N, 32);
    key.push(id % g.N, 8);
  }
  else
  {
    uint32_t x[4];
    unpack_coords(ukey[s], g, x);
    const int nch = 1 << g.dim;
    if (f & 2)
    {
      key.push(1u, 2);
      int m0 = 0;
      for (int d = 0; d < g.dim; d++) m0 |= ((x[d] >> g.max_depth) & 1u) << d;
      key.push(c_rot_inv[m0], sb);
      int rot = 0;
      for (int l = 1; l <= g.lk; l++)
      {
        int m = 0;
        for (int d = 0; d < g.dim; d++) m |= ((x[d] >> (g.max_depth - l)) & 1u) << d;
        key.push(c_rot_inv[rot * nch + m], sb);
        rot = c_htab[rot * nch + m];
      }
      key.push(0u, sb);
    }
    else
    {
      key.push(0u, 2);
      int tz = 0;  // max over axes of the number of trailing zeros
      for (int d = 0; d < g.dim; d++) tz = max(tz, __ffs((int)x[d]) - 1);
      const int lc = g.max_depth - tz - 1;  // level of the finest open container
      const uint32_t hmask = (1u << (g.max_depth - (lc + 1))) - 1u;
      int hp = 0;
      for (int d = g.dim - 1; d >= 0; d--)
        if ((x[d] & hmask) == 0) hp = d;
      int rot = 0;
      for (int l = 1; l <= g.lk; l++)
      {
        if (l == lc + 1)
        {
          key.push((uint32_t)nch, sb);
          key.push((uint32_t)hp, sb);
        }
        int m = 0;
        for (int d = 0; d < g.dim; d++) m |= ((x[d] >> (g.max_depth - l)) & 1u) << d;
        key.push(c_rot_inv[rot * nch + m], sb);
        rot = c_htab[rot * nch + m];
      }
    }
    for (int i = 0; i < field - nsym * sb; i += 8) key.push(0u, min(8, field - nsym * sb - i));
  }
  khi[k] = key.hi;
  klo[k] = key.lo;
  kseg[k] = (uint32_t)s;
}

__global__ void k_gather_u64(const uint64_t *src, const uint32_t *perm, uint64_t n, uint64_t *dst)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[perm[i]];
}
__global__ void k_iota(uint32_t *p, uint64_t n)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = (uint32_t)i;
}

// final node arrays; perm[i] = kept index at DA position i
__global__ void k_fill_nodes(const uint32_t *perm, const uint32_t *kseg, const uint64_t *ukey, const uint8_t *ulev,
                             const uint8_t *uflag, uint64_t nK, Geo g, uint32_t *node_xyz, uint8_t *node_lev,
                             uint8_t *node_isbdy, uint32_t *unode)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= nK) return;
  const uint32_t s = kseg[perm[i]];
  uint32_t x[4];
  unpack_coords(ukey[s], g, x);
  for (int d = 0; d < g.dim; d++) node_xyz[i * g.dim + d] = x[d];
  node_lev[i] = ulev[s];
  node_isbdy[i] = (uflag[s] & 2) ? 1 : 0;
  unode[s] = (uint32_t)i;
}

__global__ void k_fill_e2n(const uint32_t *ucnt, const uint64_t *uoff, const uint32_t *unode, const uint32_t *inst, uint64_t nU,
                           uint32_t *e2n)
{
  uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (s >= nU) return;
  const uint32_t node = unode[s];
  const uint64_t o = uoff[s];
  const uint32_t cnt = ucnt[s];
  for (uint32_t j = 0; j < cnt; j++) e2n[inst[o + j]] = node;
}

// q2 nodes: probe the 2^dim diagonal neighbours half a finest cell away; a containing leaf that
// is coarser than the node holds a node of level > its own in its closed cell -> the reference
// descends into its (non-existent) children (FEM/include/matvec.h:102-107).
__global__ void k_mark_split(const uint64_t *ukey, const uint8_t *ulev, const uint8_t *uflag, uint64_t nU, Geo g,
                             const uint64_t *ekey, const uint8_t *elev, uint64_t nE, uint8_t *split)
{
  uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t s = t >> g.dim;
  const int corner = (int)(t & ((1u << g.dim) - 1));
  if (s >= nU || !(uflag[s] & 4)) return;
  uint32_t x[4], p[4];
  unpack_coords(ukey[s], g, x);
  const uint32_t full2 = 2u << g.max_depth;
  for (int d = 0; d < g.dim; d++)
  {
    const int64_t v = 2 * (int64_t)x[d] + (((corner >> d) & 1) ? 1 : -1);
    if (v <= 0 || v >= (int64_t)full2) return;
    p[d] = (uint32_t)v;
  }
  const uint64_t k = tree_key(p, g.dim, g.max_depth + 1, g.lmax, g.lmax);
  uint64_t lo = 0, hi = nE;  // upper_bound
  while (lo < hi)
  {
    const uint64_t mid = (lo + hi) >> 1;
    if (ekey[mid] <= k) lo = mid + 1;
    else hi = mid;
  }
  if (lo == 0) return;
  const uint64_t e = lo - 1;
  if (elev[e] < ulev[s]) split[e] = 1;
}

// children of split leaves: lattice lookups instead of instance scatter
__global__ void k_children(const uint32_t *sp, uint64_t nSp, const uint32_t *exyz, const uint8_t *elev, Geo g,
                           const uint64_t *ukey, const uint32_t *unode, uint64_t nU, uint32_t *cxyz, uint8_t *clev, uint32_t *ce2n,
                           uint8_t *cnonempty)
{
  uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (t >= nSp * g.nch) return;
  const uint32_t e = sp[t / g.nch];
  const int c = (int)(t % g.nch);
  const int L = elev[e] + 1;
  const uint32_t len = 1u << (g.max_depth - L);
  uint32_t a[4], x[4];
  for (int d = 0; d < g.dim; d++)
  {
    a[d] = exyz[(uint64_t)e * g.dim + d] + (((c >> d) & 1) ? len : 0u);
    cxyz[t * g.dim + d] = a[d];
  }
  clev[t] = (uint8_t)L;
  bool any = false;
  for (int r = 0; r < g.N; r++)
  {
    lattice_point(a, len, r, g, x);
    const int64_t s = find_location(x, g, ukey, nU);
    const uint32_t node = s >= 0 ? unode[s] : INVALID;
    ce2n[t * g.N + r] = node;
    any |= node != INVALID;
  }
  cnonempty[t] = any ? 1 : 0;
}

__global__ void k_hang_flags(const uint32_t *e2n, uint64_t n, int N, uint8_t *hang)
{
  uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (e >= n) return;
  bool h = false;
  for (int r = 0; r < N; r++) h |= e2n[e * N + r] == INVALID;
  hang[e] = h ? 1 : 0;
}

// candidate list -> visited list with regular elements first (stable), hanging after
__global__ void k_partition_mv(const uint32_t *cxyz, const uint8_t *clev, const uint32_t *ce2n, const uint8_t *hang,
                               const uint64_t *hpos, uint64_t n, uint64_t nReg, Geo g, uint32_t *mxyz, uint8_t *mlev,
                               uint32_t *me2n, uint32_t *msrc)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t dst = hang[i] ? nReg + hpos[i] : i - hpos[i];
  for (int d = 0; d < g.dim; d++) mxyz[dst * g.dim + d] = cxyz[i * g.dim + d];
  mlev[dst] = clev[i];
  msrc[dst] = (uint32_t)i;
  for (int r = 0; r < g.N; r++) me2n[dst * g.N + r] = ce2n[i * g.N + r];
}

// parent-lattice nodes of level L-1 (FEM/include/matvec.h:415-431) + class-U detection (:439-447)
__global__ void k_hang_tables(const uint32_t *mxyz, const uint8_t *mlev, const uint32_t *me2n, uint64_t nReg, uint64_t nHang,
                              Geo g, const uint64_t *ukey, const uint32_t *unode, uint64_t nU, const uint8_t *node_lev,
                              uint32_t *pnode, uint8_t *child, int *undefined)
{
  uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (t >= nHang * g.N) return;
  const uint64_t h = t / g.N;
  const int q = (int)(t % g.N);
  const uint64_t e = nReg + h;
  const int L = mlev[e];
  const uint32_t len = 1u << (g.max_depth - L);
  uint32_t a[4], pa[4], x[4];
  int c = 0;
  for (int d = 0; d < g.dim; d++)
  {
    a[d] = mxyz[e * g.dim + d];
    c |= ((a[d] >> (g.max_depth - L)) & 1u) << d;
    pa[d] = a[d] & ~((len << 1) - 1u);
  }
  lattice_point(pa, len << 1, q, g, x);
  const int64_t s = find_location(x, g, ukey, nU);
  uint32_t node = s >= 0 ? unode[s] : INVALID;
  if (node != INVALID && (int)node_lev[node] != L - 1) node = INVALID;
  pnode[t] = node;
  if (q == 0) child[h] = (uint8_t)c;
  if (node == INVALID && me2n[e * g.N + q] == INVALID) atomicOr(undefined, 1);
}

__global__ void k_collect_bdy(const uint8_t *isbdy, const uint64_t *pos, uint64_t n, uint32_t *out)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n && isbdy[i]) out[pos[i]] = (uint32_t)i;
}

template <typename T>
__global__ void k_widen(const T *in, uint64_t n, uint64_t *out, int mask)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (in[i] & mask) ? 1 : 0;
}
__global__ void k_u32_to_u64(const uint32_t *in, uint64_t n, uint64_t *out)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}
__global__ void k_compact_idx(const uint8_t *flag, const uint64_t *pos, uint64_t n, int mask, int want, uint32_t *out)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n && (((flag[i] & mask) != 0) == (want != 0))) out[pos[i]] = (uint32_t)i;
}
__global__ void k_not(const uint8_t *in, uint64_t n, uint8_t *out)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] ? 0 : 1;
}
__global__ void k_gather_cand(const uint32_t *idx, uint64_t n, const uint32_t *xyz, const uint8_t *lev, const uint32_t *e2n, Geo g,
                              uint32_t *oxyz, uint8_t *olev, uint32_t *oe2n)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t s = idx[i];
  for (int d = 0; d < g.dim; d++) oxyz[i * g.dim + d] = xyz[s * g.dim + d];
  olev[i] = lev[s];
  for (int r = 0; r < g.N; r++) oe2n[i * g.N + r] = e2n[s * g.N + r];
}

// ------------------------------------------------------------------------------------------
// host helpers around CUB
// ------------------------------------------------------------------------------------------
static int sort_pairs(DA &da, uint64_t *kin, uint64_t *kout, uint32_t *vin, uint32_t *vout, uint64_t n, int bit0, int bit1)
{
  if (n == 0) return DKT_OK;
  if (n > 0x7fffffffull * 2) { set_error("too many items for one radix sort"); return DKT_ERR_UNSUPPORTED; }
  size_t tmp = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, kin, kout, vin, vout, (int64_t)n, bit0, bit1, da.stream));
  Buf<char> t;
  CK(t.alloc(tmp));
  CK(cub::DeviceRadixSort::SortPairs(t.p, tmp, kin, kout, vin, vout, (int64_t)n, bit0, bit1, da.stream));
  g_launches += 8;
  CK(cudaStreamSynchronize(da.stream));
  return DKT_OK;
}
static int exclusive_scan(DA &da, const uint64_t *in, uint64_t *out, uint64_t n)
{
  if (n == 0) return DKT_OK;
  size_t tmp = 0;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, (int64_t)n, da.stream));
  Buf<char> t;
  CK(t.alloc(tmp));
  CK(cub::DeviceScan::ExclusiveSum(t.p, tmp, in, out, (int64_t)n, da.stream));
  g_launches += 2;
  CK(cudaStreamSynchronize(da.stream));
  return DKT_OK;
}
// flags (uint8, masked) -> exclusive positions + total
static int flag_positions(DA &da, const uint8_t *flag, int mask, uint64_t n, Buf<uint64_t> &pos, uint64_t &total)
{
  total = 0;
  CK(pos.alloc(n + 1));
  if (n == 0) return DKT_OK;
  Buf<uint64_t> w;
  CK(w.alloc(n + 1));
  CK(cudaMemsetAsync(w.p, 0, (n + 1) * sizeof(uint64_t), da.stream));
  LAUNCH(k_widen<uint8_t>, n, flag, n, w.p, mask);
  int rc = exclusive_scan(da, w.p, pos.p, n + 1);
  if (rc) return rc;
  CK(cudaMemcpy(&total, pos.p + n, sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return DKT_OK;
}

int device_exclusive_scan(DA &da, const uint64_t *in, uint64_t *out, uint64_t n) { return exclusive_scan(da, in, out, n); }

void free_da(DA &da)
{
  free_chunks(da);
  cudaFree(da.d_elem_xyz); cudaFree(da.d_elem_lev); cudaFree(da.d_node_xyz); cudaFree(da.d_node_lev);
  cudaFree(da.d_bdy); cudaFree(da.d_node_isbdy); cudaFree(da.d_node_sent); cudaFree(da.d_e2n); cudaFree(da.d_mv_lev); cudaFree(da.d_mv_src);
  cudaFree(da.d_mv_xyz); cudaFree(da.d_pnode); cudaFree(da.d_child); cudaFree(da.d_ukey); cudaFree(da.d_unode);
  cudaFree(da.d_in); cudaFree(da.d_out); cudaFree(da.d_kbuf); cudaFree(da.d_dof);
  if (da.ev0) cudaEventDestroy(da.ev0);
  if (da.ev1) cudaEventDestroy(da.ev1);
  if (da.own_stream) cudaStreamDestroy(da.own_stream);
  da = DA();
}

int build_da(DA &da, const uint32_t *elem_xyz, const uint8_t *elem_lev, uint64_t n, unsigned flags)
{
  const int dim = da.dim;
  CK(cudaGetDevice(&da.device));
  CK(cudaStreamCreateWithFlags(&da.own_stream, cudaStreamNonBlocking));
  da.stream = da.own_stream;
  CK(cudaEventCreate(&da.ev0));
  CK(cudaEventCreate(&da.ev1));
  if (n == 0) { set_error("empty tree"); return DKT_ERR_INVALID; }
  // device-resident input may still be in flight on a stream of the caller's (e.g. torch's current stream);
  // construction is a one-off, so simply wait for the device
  if (flags & DKT_ELEMS_ON_DEVICE) CK(cudaDeviceSynchronize());
  if (n * (uint64_t)da.N >= 0xFFFFFFFFull) { set_error("n_elem * nodes_per_elem must be < 2^32"); return DKT_ERR_UNSUPPORTED; }

  SfcTables tab;
  make_sfc_tables(dim, da.sfc_mode, tab);
  if (tab.nrot > 192) { set_error("internal: too many SFC rotations"); return DKT_ERR_UNSUPPORTED; }
  // c_rot_inv / c_htab are per-process __constant__ symbols read by the key kernels below: two host threads building
  // DAs of different (dim, sfc_mode) must not interleave the upload and the kernels that read it.  Construction
  // synchronises the device before it returns (the size download), so holding the lock to the end is enough.
  static std::mutex sfc_tables_mutex;
  std::lock_guard<std::mutex> sfc_tables_lock(sfc_tables_mutex);
  CK(cudaMemcpyToSymbol(c_rot_inv, tab.rot_inv.data(), tab.rot_inv.size()));
  CK(cudaMemcpyToSymbol(c_htab, tab.htab.data(), tab.htab.size()));

  // ---- elements on the device -------------------------------------------------------------
  Buf<uint32_t> in_xyz;
  Buf<uint8_t> in_lev;
  const uint32_t *src_xyz = elem_xyz;
  const uint8_t *src_lev = elem_lev;
  if (!(flags & DKT_ELEMS_ON_DEVICE))
  {
    CK(in_xyz.alloc(n * dim));
    CK(in_lev.alloc(n));
    CK(cudaMemcpy(in_xyz.p, elem_xyz, n * dim * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(in_lev.p, elem_lev, n, cudaMemcpyHostToDevice));
    src_xyz = in_xyz.p;
    src_lev = in_lev.p;
  }
  Buf<int> dmax;
  CK(dmax.alloc(1));
  CK(cudaMemsetAsync(dmax.p, 0, sizeof(int), da.stream));
  LAUNCH(k_max_level, n, src_lev, n, dmax.p);
  int lmax = 0;
  CK(cudaMemcpyAsync(&lmax, dmax.p, sizeof(int), cudaMemcpyDeviceToHost, da.stream));
  CK(cudaStreamSynchronize(da.stream));

  Geo g;
  g.dim = dim; g.order = da.order; g.M = da.M; g.N = da.N; g.max_depth = da.max_depth; g.nch = 1 << dim;
  g.lmax = lmax;
  g.lk = lmax + (da.order == 2 ? 1 : 0);
  g.shift = da.max_depth - g.lk;
  g.bits = g.lk + 1;
  da.finest_level = lmax; da.lk = g.lk; da.shift = g.shift; da.bits = g.bits;
  if (g.lk > da.max_depth) { set_error("order-2 elements at max_depth have no room for mid-side nodes"); return DKT_ERR_INVALID; }
  if (g.bits * dim > 64 || lmax * dim > 64)
  {
    set_error("unsupported depth: (finest lattice level + 1) * dim must be <= 64");
    return DKT_ERR_UNSUPPORTED;
  }
  if (2 + (g.lk + 2) * (dim + 1) > 128) { set_error("unsupported depth for the 128-bit node order key"); return DKT_ERR_UNSUPPORTED; }

  // ---- tree order -----------------------------------------------------------------------------
  da.nElem = n;
  CK(cudaMalloc((void **)&da.d_elem_xyz, n * dim * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&da.d_elem_lev, n));
  Buf<uint64_t> ekey;
  CK(ekey.alloc(n));
  {
    Buf<uint64_t> k0;
    Buf<uint32_t> i0, i1;
    CK(k0.alloc(n)); CK(i0.alloc(n)); CK(i1.alloc(n));
    LAUNCH(k_elem_keys, n, src_xyz, src_lev, n, g, k0.p, i0.p);
    if (flags & DKT_ELEMS_PRESORTED)
    {
      CK(cudaMemcpyAsync(ekey.p, k0.p, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, da.stream));
      CK(cudaMemcpyAsync(da.d_elem_xyz, src_xyz, n * dim * sizeof(uint32_t), cudaMemcpyDeviceToDevice, da.stream));
      CK(cudaMemcpyAsync(da.d_elem_lev, src_lev, n, cudaMemcpyDeviceToDevice, da.stream));
    }
    else
    {
      int rc = sort_pairs(da, k0.p, ekey.p, i0.p, i1.p, n, 0, std::max(1, lmax * dim));
      if (rc) return rc;
      LAUNCH(k_gather_elems, n, src_xyz, src_lev, i1.p, n, dim, da.d_elem_xyz, da.d_elem_lev);
    }
    CK(cudaStreamSynchronize(da.stream));
  }
  in_xyz.release();
  in_lev.release();

  // ---- lattice instances -> unique locations -----------------------------------------------------
  const uint64_t nInst = n * (uint64_t)da.N;
  Buf<uint64_t> ikey;   // sorted instance keys
  Buf<uint32_t> inst;   // instance ids in sorted order
  {
    Buf<uint64_t> k0;
    Buf<uint32_t> i0;
    CK(k0.alloc(nInst)); CK(i0.alloc(nInst)); CK(ikey.alloc(nInst)); CK(inst.alloc(nInst));
    LAUNCH(k_instances, nInst, da.d_elem_xyz, da.d_elem_lev, nInst, g, k0.p, i0.p);
    int rc = sort_pairs(da, k0.p, ikey.p, i0.p, inst.p, nInst, 0, g.bits * dim);
    if (rc) return rc;
  }
  Buf<uint64_t> ukey, uoff;
  Buf<uint32_t> ucnt;
  uint64_t nU = 0;
  {
    Buf<uint64_t> uk;
    Buf<uint32_t> uc;
    Buf<uint64_t> nruns;
    CK(uk.alloc(nInst)); CK(uc.alloc(nInst)); CK(nruns.alloc(1));
    size_t tmp = 0;
    CK(cub::DeviceRunLengthEncode::Encode(nullptr, tmp, ikey.p, uk.p, uc.p, nruns.p, (int64_t)nInst, da.stream));
    Buf<char> t;
    CK(t.alloc(tmp));
    CK(cub::DeviceRunLengthEncode::Encode(t.p, tmp, ikey.p, uk.p, uc.p, nruns.p, (int64_t)nInst, da.stream));
    g_launches += 2;
    CK(cudaMemcpyAsync(&nU, nruns.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, da.stream));
    CK(cudaStreamSynchronize(da.stream));
    ikey.release();
    CK(ukey.alloc(nU)); CK(ucnt.alloc(nU)); CK(uoff.alloc(nU + 1));
    CK(cudaMemcpyAsync(ukey.p, uk.p, nU * sizeof(uint64_t), cudaMemcpyDeviceToDevice, da.stream));
    CK(cudaMemcpyAsync(ucnt.p, uc.p, nU * sizeof(uint32_t), cudaMemcpyDeviceToDevice, da.stream));
    Buf<uint64_t> w;
    CK(w.alloc(nU + 1));
    CK(cudaMemsetAsync(w.p, 0, (nU + 1) * sizeof(uint64_t), da.stream));
    LAUNCH(k_u32_to_u64, nU, ucnt.p, nU, w.p);
    int rc = exclusive_scan(da, w.p, uoff.p, nU + 1);
    if (rc) return rc;
  }
  da.nU = nU;

  Buf<uint8_t> ulev, uflag;
  CK(ulev.alloc(nU)); CK(uflag.alloc(nU));
  LAUNCH(k_resolve, nU, ukey.p, ucnt.p, uoff.p, nU, inst.p, da.d_elem_lev, g, ulev.p, uflag.p);

  // ---- node order -----------------------------------------------------------------------------
  Buf<uint64_t> kpos;
  uint64_t nK = 0;
  {
    int rc = flag_positions(da, uflag.p, 1, nU, kpos, nK);
    if (rc) return rc;
  }
  da.nNodes = nK;
  if (nK >= 0xFFFFFFFEull) { set_error("too many nodes for 32-bit local ids"); return DKT_ERR_UNSUPPORTED; }
  Buf<uint32_t> unode;
  CK(unode.alloc(nU));
  CK(cudaMemsetAsync(unode.p, 0xFF, nU * sizeof(uint32_t), da.stream));
  CK(cudaMalloc((void **)&da.d_node_xyz, std::max<uint64_t>(nK, 1) * dim * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&da.d_node_lev, std::max<uint64_t>(nK, 1)));
  CK(cudaMalloc((void **)&da.d_node_isbdy, std::max<uint64_t>(nK, 1)));
  {
    Buf<uint64_t> khi, klo, tmpk, tmpk2;
    Buf<uint32_t> kseg, p0, p1, p2;
    CK(khi.alloc(nK)); CK(klo.alloc(nK)); CK(kseg.alloc(nK)); CK(p0.alloc(nK)); CK(p1.alloc(nK)); CK(tmpk.alloc(nK));
    LAUNCH(k_order_keys, nU, ukey.p, uflag.p, uoff.p, inst.p, kpos.p, nU, g, khi.p, klo.p, kseg.p);
    LAUNCH(k_iota, nK, p0.p, nK);
    const int total_bits = 2 + std::max((g.lk + 2) * (dim + 1), 40);
    int rc = sort_pairs(da, klo.p, tmpk.p, p0.p, p1.p, nK, 0, std::min(64, total_bits));
    if (rc) return rc;
    uint32_t *perm = p1.p;
    if (total_bits > 64)
    {
      CK(tmpk2.alloc(nK)); CK(p2.alloc(nK));
      LAUNCH(k_gather_u64, nK, khi.p, p1.p, nK, tmpk.p);
      rc = sort_pairs(da, tmpk.p, tmpk2.p, p1.p, p2.p, nK, 0, total_bits - 64);
      if (rc) return rc;
      perm = p2.p;
    }
    LAUNCH(k_fill_nodes, nK, perm, kseg.p, ukey.p, ulev.p, uflag.p, nK, g, da.d_node_xyz, da.d_node_lev, da.d_node_isbdy, unode.p);
    CK(cudaStreamSynchronize(da.stream));
  }
  // boundary ids
  {
    Buf<uint64_t> bpos;
    uint64_t nB = 0;
    int rc = flag_positions(da, da.d_node_isbdy, 1, nK, bpos, nB);
    if (rc) return rc;
    da.nBdy = nB;
    CK(cudaMalloc((void **)&da.d_bdy, std::max<uint64_t>(nB, 1) * sizeof(uint32_t)));
    LAUNCH(k_collect_bdy, nK, da.d_node_isbdy, bpos.p, nK, da.d_bdy);
  }

  // ---- element -> node ---------------------------------------------------------------------------
  Buf<uint32_t> e2n;
  CK(e2n.alloc(nInst));
  LAUNCH(k_fill_e2n, nU, ucnt.p, uoff.p, unode.p, inst.p, nU, e2n.p);

  // ---- class P: split leaves and phantom children ---------------------------------------------------
  Buf<uint8_t> split;
  CK(split.alloc(n));
  CK(cudaMemsetAsync(split.p, 0, n, da.stream));
  LAUNCH(k_mark_split, nU << dim, ukey.p, ulev.p, uflag.p, nU, g, ekey.p, da.d_elem_lev, n, split.p);
  Buf<uint64_t> spos;
  uint64_t nSplit = 0;
  {
    int rc = flag_positions(da, split.p, 1, n, spos, nSplit);
    if (rc) return rc;
  }
  da.nSplit = nSplit;
  inst.release(); ucnt.release(); uoff.release(); ekey.release(); kpos.release();

  // candidate list: non-split tree elements (tree order) followed by non-empty phantom children
  Buf<uint32_t> cand_xyz, cand_e2n;
  Buf<uint8_t> cand_lev;
  uint64_t nCand = n;
  const uint32_t *cxyz = da.d_elem_xyz;
  const uint8_t *clev = da.d_elem_lev;
  const uint32_t *ce2n = e2n.p;
  if (nSplit > 0)
  {
    Buf<uint32_t> sp, keepidx, ch_xyz, ch_e2n, chidx;
    Buf<uint8_t> ch_lev, ch_ne;
    const uint64_t nCh = nSplit << dim;
    CK(sp.alloc(nSplit)); CK(keepidx.alloc(n - nSplit));
    CK(ch_xyz.alloc(nCh * dim)); CK(ch_lev.alloc(nCh)); CK(ch_e2n.alloc(nCh * da.N)); CK(ch_ne.alloc(nCh));
    LAUNCH(k_compact_idx, n, split.p, spos.p, n, 1, 1, sp.p);
    // positions of the kept ones: i - spos[i]
    Buf<uint8_t> nsplit;
    Buf<uint64_t> kp;
    uint64_t nKeep = 0;
    CK(nsplit.alloc(n));
    LAUNCH(k_not, n, split.p, n, nsplit.p);
    int rc = flag_positions(da, nsplit.p, 1, n, kp, nKeep);
    if (rc) return rc;
    LAUNCH(k_compact_idx, n, nsplit.p, kp.p, n, 1, 1, keepidx.p);
    LAUNCH(k_children, nCh, sp.p, nSplit, da.d_elem_xyz, da.d_elem_lev, g, ukey.p, unode.p, nU, ch_xyz.p, ch_lev.p, ch_e2n.p, ch_ne.p);
    Buf<uint64_t> cp;
    uint64_t nChKeep = 0;
    rc = flag_positions(da, ch_ne.p, 1, nCh, cp, nChKeep);
    if (rc) return rc;
    CK(chidx.alloc(nChKeep));
    LAUNCH(k_compact_idx, nCh, ch_ne.p, cp.p, nCh, 1, 1, chidx.p);
    nCand = nKeep + nChKeep;
    CK(cand_xyz.alloc(nCand * dim)); CK(cand_lev.alloc(nCand)); CK(cand_e2n.alloc(nCand * da.N));
    LAUNCH(k_gather_cand, nKeep, keepidx.p, nKeep, da.d_elem_xyz, da.d_elem_lev, e2n.p, g, cand_xyz.p, cand_lev.p, cand_e2n.p);
    LAUNCH(k_gather_cand, nChKeep, chidx.p, nChKeep, ch_xyz.p, ch_lev.p, ch_e2n.p, g, cand_xyz.p + nKeep * dim, cand_lev.p + nKeep,
           cand_e2n.p + nKeep * da.N);
    CK(cudaStreamSynchronize(da.stream));
    cxyz = cand_xyz.p; clev = cand_lev.p; ce2n = cand_e2n.p;
  }

  // ---- regular first, hanging after ------------------------------------------------------------------
  Buf<uint8_t> hang;
  CK(hang.alloc(nCand));
  LAUNCH(k_hang_flags, nCand, ce2n, nCand, da.N, hang.p);
  Buf<uint64_t> hpos;
  uint64_t nHang = 0;
  {
    int rc = flag_positions(da, hang.p, 1, nCand, hpos, nHang);
    if (rc) return rc;
  }
  da.nMv = nCand; da.nHang = nHang; da.nReg = nCand - nHang;
  CK(cudaMalloc((void **)&da.d_mv_xyz, nCand * dim * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&da.d_mv_lev, nCand));
  CK(cudaMalloc((void **)&da.d_mv_src, nCand * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&da.d_e2n, nCand * da.N * sizeof(uint32_t)));
  LAUNCH(k_partition_mv, nCand, cxyz, clev, ce2n, hang.p, hpos.p, nCand, da.nReg, g, da.d_mv_xyz, da.d_mv_lev, da.d_e2n, da.d_mv_src);
  CK(cudaMalloc((void **)&da.d_pnode, std::max<uint64_t>(nHang, 1) * da.N * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&da.d_child, std::max<uint64_t>(nHang, 1)));
  Buf<int> undef;
  CK(undef.alloc(1));
  CK(cudaMemsetAsync(undef.p, 0, sizeof(int), da.stream));
  LAUNCH(k_hang_tables, nHang * da.N, da.d_mv_xyz, da.d_mv_lev, da.d_e2n, da.nReg, nHang, g, ukey.p, unode.p, nU, da.d_node_lev,
         da.d_pnode, da.d_child, undef.p);
  int h_undef = 0;
  CK(cudaMemcpyAsync(&h_undef, undef.p, sizeof(int), cudaMemcpyDeviceToHost, da.stream));
  CK(cudaStreamSynchronize(da.stream));
  da.tree_class = h_undef ? DKT_CLASS_U : nSplit ? DKT_CLASS_P : nHang ? DKT_CLASS_B : DKT_CLASS_A;
  da.d_ukey = ukey.take();
  da.d_unode = unode.take();
  CK(cudaGetLastError());
  if (da.tree_class == DKT_CLASS_U && !(flags & DKT_ALLOW_UNDEFINED))
  {
    set_error("class-U tree: some hanging lattice node has no parent node of level L-1; the reference reads undefined "
              "values here (FEM/include/matvec.h:439-447). Pass DKT_ALLOW_UNDEFINED to build it anyway.");
    return DKT_ERR_UNDEFINED_TREE;
  }
  return DKT_OK;
}
} // namespace dkt
