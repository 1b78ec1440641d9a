// Trees from points on the GPU: construction with at most maxPtsPerRegion points per leaf, and 2:1 balancing.
//
// Replaces, for one rank (paths relative to the reference repository):
//   * SFC_Tree::locTreeConstruction / distTreeConstruction   src/tsort.cpp:566-716
//   * SFC_Tree::propagateNeighbours                           src/tsort.cpp:775-823
//   * SFC_Tree::distTreeBalancing                             src/tsort.cpp:862-877
//   * TreeNode::appendAllNeighbours / getNeighbour1d          include/treeNode.tcc:453-585
// The reference recurses top-down, bucketing the points at every level (construction), then walks the levels bottom-up
// pushing each node's parent and the parent's 3^dim - 1 neighbours into std::vectors that it sorts and de-duplicates per
// level, and finally runs the construction again on that seed set with one seed per region.  What comes out is a SET of
// leaves that has a closed form (oracle/tree.py is the literal restatement the tests compare against):
//
//   Q_l = the nodes of level l that the construction splits = ancestors (1 <= l <= maxDepth-1) of more than maxPts points
//   P_l = Q_l  u  parents(P_{l+1})  u  parents(neighbours(P_{l+1}))        (balancing; P_l = Q_l without), P_0 = {root}
//   leaves = the children of every node of P that are not in P themselves.
//
// Why: the seed set S of propagateNeighbours satisfies S_l = T0_l u P_l u N(P_l) with P_l = parents(S_{l+1}), every
// parent of a seed is a seed, and the final construction splits a region iff it holds two seeds of level >= its own,
// i.e. iff it is the parent of a seed.  parents(neighbours(p)) has at most 2^dim distinct members: along each axis only
// the neighbour on the far side of p from its sibling leaves p's parent.
//
// Here: Morton keys of the points, ONE radix sort, the deepest split ancestor of each point from the common prefixes of
// windows of maxPts+1 consecutive keys, each split node emitted once (by the first point inside it); then one
// sort + unique per level bottom-up (<= maxDepth of them), a binary search per child for the leaves, and the library's
// tree-order key (Morton or Hilbert, same tables as dkt_build.cu) for the output order.  A level-l node is the Morton
// interleave of its anchor >> (maxDepth - l): dim*l bits, parent = key >> dim.  Needs dim * maxDepth <= 56.
// Also compiles under -DDKT_EMU (tests/emu/cuda_emu.h) for the CPU test-suite; the product is never built that way.
#include "dkt_internal.h"

#ifdef DKT_EMU
#include "cuda_emu.h"
#else
#include <cub/cub.cuh>
#endif

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

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

namespace
{
template <typename T>
struct TBuf
{
  T *p = nullptr;
  ~TBuf() { release(); }
  void release()
  {
    if (p) cudaFree(p);
    p = nullptr;
  }
  cudaError_t alloc(size_t count)
  {
    release();
    return cudaMalloc((void **)&p, std::max<size_t>(count, 1) * sizeof(T));
  }
  T *take()
  {
    T *r = p;
    p = nullptr;
    return r;
  }
};

inline unsigned tblk(uint64_t n) { return (unsigned)((n + 255) / 256); }
#ifdef DKT_EMU
#define TREE_LAUNCH(kern, n, stream) ::emu::make_launch_flat(kern, tblk(n), 256)
#else
#define TREE_LAUNCH(kern, n, stream) kern<<<tblk(n), 256, 0, (stream)>>>
#endif
#define TLAUNCH(kern, n, ...)                          \
  do                                                   \
  {                                                    \
    if ((n) > 0)                                       \
    {                                                  \
      TREE_LAUNCH(kern, n, stream)(__VA_ARGS__);       \
      g_launches++;                                    \
    }                                                  \
  } while (0)

// ---- keys -------------------------------------------------------------------------------------------------------------------
// Morton interleave of the top `lev` bits (of `depth`) of each coordinate: digit of level 1 most significant.
__host__ __device__ inline uint64_t node_key(const uint32_t *x, int dim, int depth, int lev)
{
  uint64_t k = 0;
  for (int l = 1; l <= lev; l++)
  {
    uint64_t dg = 0;
    for (int d = 0; d < dim; d++) dg |= (uint64_t)((x[d] >> (depth - l)) & 1u) << d;
    k = (k << dim) | dg;
  }
  return k;
}
// coordinates (in units of the level's cell size) of a level-`lev` key
__host__ __device__ inline void key_coords(uint64_t k, int dim, int lev, uint32_t *x)
{
  for (int d = 0; d < dim; d++) x[d] = 0;
  for (int l = 0; l < lev; l++)
  {
    const uint64_t dg = (k >> (dim * l)) & ((1u << dim) - 1);
    for (int d = 0; d < dim; d++) x[d] |= (uint32_t)((dg >> d) & 1u) << l;
  }
}
// number of leading levels two full-depth keys share
__device__ inline int common_level(uint64_t a, uint64_t b, int dim, int depth)
{
  const uint64_t x = a ^ b;
  if (x == 0) return depth;
#ifdef DKT_EMU
  int h = 63;
  while (!((x >> h) & 1ull)) h--;
#else
  const int h = 63 - __clzll((long long)x);
#endif
  return depth - 1 - h / dim;
}

__global__ void k_point_keys(const uint32_t *pts, uint64_t n, int dim, int depth, uint64_t *keys, int *outside)
{
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int d = 0; d < dim; d++)
    if ((uint64_t)pts[i * dim + d] >> depth) *outside = 1;  // benign race: every writer stores 1
  keys[i] = node_key(pts + i * dim, dim, depth, depth);
}

// Deepest split ancestor of every point (src/tsort.cpp:612-626: a child bucket with more than maxPts points recurses while
// it is coarser than maxDepth) and the levels it is the FIRST point of: (lo, hi].
__global__ void k_split_depth(const uint64_t *keys, uint64_t n, uint64_t m, int dim, int depth, uint8_t *lo, uint8_t *hi, uint64_t *cnt)
{
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int L = 0;
  if (n > m && m <= 32)
  {
    // windows of m + 1 consecutive keys that contain i: the common prefix of a window's ends is shared by all of it
    const uint64_t j0 = i >= m ? i - m : 0, j1 = (i + m < n) ? i : n - 1 - m;
    for (uint64_t j = j0; j <= j1; j++) L = max(L, common_level(keys[j], keys[j + m], dim, depth));
  }
  else if (n > m)
  {
    // large maxPts: count the points under the level-l ancestor by two binary searches; the count falls with l
    int lo = 0, hi = depth - 1;  // the root holds all n > m points
    while (lo < hi)
    {
      const int l = (lo + hi + 1) >> 1, sh = dim * (depth - l);
      const uint64_t first = (keys[i] >> sh) << sh, last = first + ((1ull << sh) - 1);
      uint64_t a = 0, b = n;  // lower bound of first
      while (a < b) { const uint64_t mid = (a + b) >> 1; if (keys[mid] < first) a = mid + 1; else b = mid; }
      uint64_t c = a, d = n;  // upper bound of last
      while (c < d) { const uint64_t mid = (c + d) >> 1; if (keys[mid] <= last) c = mid + 1; else d = mid; }
      if (c - a > m) lo = l;
      else hi = l - 1;
    }
    L = lo;
  }
  L = min(L, depth - 1);
  const int l0 = i > 0 ? min(common_level(keys[i - 1], keys[i], dim, depth), L) : 0;
  lo[i] = (uint8_t)l0;
  hi[i] = (uint8_t)L;
  cnt[i] = (uint64_t)(L - l0);
}
__global__ void k_emit_split(const uint64_t *keys, uint64_t n, const uint8_t *lo, const uint8_t *hi, const uint64_t *off, int dim, int depth,
                             uint64_t *out)
{
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t o = off[i];
  for (int l = lo[i] + 1; l <= hi[i]; l++) out[o++] = ((uint64_t)l << 56) | (keys[i] >> (dim * (depth - l)));
}
__global__ void k_level_count(const uint64_t *tagged, uint64_t n, unsigned long long *hist)
{
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == 0 || (tagged[i] >> 56) != (tagged[i - 1] >> 56)) hist[tagged[i] >> 56] = i;  // start of the level's run
}
__global__ void k_untag(const uint64_t *tagged, uint64_t n, uint64_t *out)
{
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = tagged[i] & ((1ull << 56) - 1);
}
// parents of p and of its neighbours (TreeNode::appendAllNeighbours, clipped at the domain like getNeighbour1d): for every
// subset s of the axes, the node one step from p on the side away from its sibling along the axes of s
__global__ void k_parent_candidates(const uint64_t *P, uint64_t n, int dim, int lev, uint64_t *out)
{
  const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const int nch = 1 << dim;
  if (t >= n * nch) return;
  const uint64_t p = P[t / nch];
  const int s = (int)(t % nch);
  uint32_t x[4];
  key_coords(p, dim, lev, x);
  const uint32_t size = 1u << lev;
  bool ok = true;
  for (int d = 0; d < dim; d++)
    if ((s >> d) & 1)
    {
      if (x[d] & 1u)
      {
        if (x[d] + 1 >= size) ok = false;
        else x[d] += 1;
      }
      else
      {
        if (x[d] == 0) ok = false;
        else x[d] -= 1;
      }
    }
  out[t] = ok ? (node_key(x, dim, lev, lev) >> dim) : (p >> dim);
}
__global__ void k_head_flags(const uint64_t *sorted, uint64_t n, uint64_t *flag)
{
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  flag[i] = (i == 0 || sorted[i] != sorted[i - 1]) ? 1 : 0;
}
__global__ void k_compact(const uint64_t *sorted, uint64_t n, const uint64_t *flag, const uint64_t *pos, uint64_t *out)
{
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flag[i]) out[pos[i]] = sorted[i];
}
__device__ inline bool contains(const uint64_t *a, uint64_t n, uint64_t k)
{
  uint64_t lo = 0, hi = n;
  while (lo < hi)
  {
    const uint64_t mid = (lo + hi) >> 1;
    if (a[mid] < k) lo = mid + 1;
    else hi = mid;
  }
  return lo < n && a[lo] == k;
}
// the children of the split nodes of level `lev` that are not split themselves are leaves (level lev + 1)
__global__ void k_leaf_flags(const uint64_t *P, uint64_t n, const uint64_t *Pnext, uint64_t nNext, int dim, uint64_t *flag)
{
  const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const int nch = 1 << dim;
  if (t >= n * nch) return;
  const uint64_t c = (P[t / nch] << dim) | (t % nch);
  flag[t] = contains(Pnext, nNext, c) ? 0 : 1;
}
__global__ void k_emit_leaves(const uint64_t *P, uint64_t n, const uint64_t *flag, const uint64_t *pos, int dim, int lev, int depth,
                              uint32_t *xyz, uint8_t *olev)
{
  const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const int nch = 1 << dim;
  if (t >= n * nch || !flag[t]) return;
  const uint64_t c = (P[t / nch] << dim) | (t % nch);
  uint32_t x[4];
  key_coords(c, dim, lev + 1, x);
  const uint64_t o = pos[t];
  for (int d = 0; d < dim; d++) xyz[o * dim + d] = x[d] << (depth - lev - 1);
  olev[o] = (uint8_t)(lev + 1);
}
// tree-order key of a leaf: SFC digits of levels 1..lmax, zero below the leaf's own level (as dkt_build.cu's tree_key,
// with the rotation tables in global memory)
__global__ void k_leaf_order_keys(const uint32_t *xyz, const uint8_t *lev, uint64_t n, int dim, int depth, int lmax, const uint8_t *rot_inv,
                                  const uint8_t *htab, uint64_t *key, uint32_t *idx)
{
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int nch = 1 << dim;
  uint64_t k = 0;
  int rot = 0;
  for (int l = 1; l <= lmax; l++)
  {
    int dg = 0;
    if (l <= lev[i])
    {
      int m = 0;
      for (int d = 0; d < dim; d++) m |= ((xyz[i * dim + d] >> (depth - l)) & 1u) << d;
      dg = rot_inv[rot * nch + m];
      rot = htab[rot * nch + m];
    }
    k = (k << dim) | (uint64_t)dg;
  }
  key[i] = k;
  idx[i] = (uint32_t)i;
}
__global__ void k_gather_leaves(const uint32_t *xyz, const uint8_t *lev, const uint32_t *idx, uint64_t n, int dim, uint32_t *oxyz, uint8_t *olev)
{
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = idx[i];
  for (int d = 0; d < dim; d++) oxyz[i * dim + d] = xyz[(uint64_t)s * dim + d];
  olev[i] = lev[s];
}

// ---- host helpers around CUB --------------------------------------------------------------------------------------------------
int sort_keys(cudaStream_t stream, uint64_t *kin, uint64_t *kout, uint64_t n, int bits)
{
  if (n == 0) return DKT_OK;
  size_t tmp = 0;
  CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp, kin, kout, (int64_t)n, 0, std::max(1, bits), stream));
  TBuf<char> t;
  CK(t.alloc(tmp));
  CK(cub::DeviceRadixSort::SortKeys(t.p, tmp, kin, kout, (int64_t)n, 0, std::max(1, bits), stream));
  g_launches += 8;
  return DKT_OK;
}
int scan(cudaStream_t stream, const uint64_t *in, uint64_t *out, uint64_t n)
{
  if (n == 0) return DKT_OK;
  size_t tmp = 0;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, (int64_t)n, stream));
  TBuf<char> t;
  CK(t.alloc(tmp));
  CK(cub::DeviceScan::ExclusiveSum(t.p, tmp, in, out, (int64_t)n, stream));
  g_launches += 2;
  return DKT_OK;
}
// total of a 0/1 (or count) array given its exclusive scan
int scan_total(cudaStream_t stream, const uint64_t *val, const uint64_t *pos, uint64_t n, uint64_t &total)
{
  total = 0;
  if (n == 0) return DKT_OK;
  uint64_t a = 0, b = 0;
  CK(cudaMemcpyAsync(&a, val + n - 1, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
  CK(cudaMemcpyAsync(&b, pos + n - 1, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  total = a + b;
  return DKT_OK;
}
// sorted unique copy of `cand` (destroyed) in a fresh buffer
int sort_unique(cudaStream_t stream, TBuf<uint64_t> &cand, uint64_t n, int bits, TBuf<uint64_t> &out, uint64_t &nOut)
{
  nOut = 0;
  if (n == 0) return out.alloc(0) == cudaSuccess ? DKT_OK : DKT_ERR_CUDA;
  TBuf<uint64_t> sorted, flag, pos;
  CK(sorted.alloc(n)); CK(flag.alloc(n)); CK(pos.alloc(n));
  int rc = sort_keys(stream, cand.p, sorted.p, n, bits);
  if (rc) return rc;
  TLAUNCH(k_head_flags, n, sorted.p, n, flag.p);
  rc = scan(stream, flag.p, pos.p, n);
  if (rc) return rc;
  rc = scan_total(stream, flag.p, pos.p, n, nOut);
  if (rc) return rc;
  CK(out.alloc(nOut));
  TLAUNCH(k_compact, n, sorted.p, n, flag.p, pos.p, out.p);
  return DKT_OK;
}
}  // namespace

void free_tree(Tree &t)
{
  cudaFree(t.d_xyz);
  cudaFree(t.d_lev);
  t.d_xyz = nullptr;
  t.d_lev = nullptr;
}

int build_tree(Tree &tr, const uint32_t *pts, uint64_t n, uint64_t maxPts, bool balance, unsigned flags)
{
  const int dim = tr.dim, depth = tr.max_depth;
  if (dim < 2 || dim > 4 || depth < 1) { set_error("dkt_tree_from_points: dim must be 2..4 and max_depth >= 1"); return DKT_ERR_INVALID; }
  if (dim * depth > 56) { set_error("dkt_tree_from_points: dim * max_depth must be <= 56"); return DKT_ERR_UNSUPPORTED; }
  if (maxPts < 1) { set_error("dkt_tree_from_points: max_pts_per_region must be >= 1"); return DKT_ERR_INVALID; }
  if (n == 0) { set_error("dkt_tree_from_points: no points (the reference returns an empty tree)"); return DKT_ERR_INVALID; }
  if (n >= 0xFFFFFFFFull) { set_error("dkt_tree_from_points: too many points"); return DKT_ERR_UNSUPPORTED; }
  cudaStream_t stream = 0;
  // DKT_TREE_TIMING=1: host-side stage times (diagnostics; adds a device synchronisation per stage)
  const bool timing = getenv("DKT_TREE_TIMING") && atoi(getenv("DKT_TREE_TIMING"));
  auto tprev = std::chrono::steady_clock::now();
  auto stage = [&](const char *what) {
    if (!timing) return;
    cudaDeviceSynchronize();
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[dkt tree] %-28s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(now - tprev).count());
    tprev = now;
  };
  if (flags & DKT_ELEMS_ON_DEVICE) CK(cudaDeviceSynchronize());  // the points may still be in flight on a stream of the caller's
  TBuf<uint32_t> in;
  const uint32_t *src = pts;
  if (!(flags & DKT_ELEMS_ON_DEVICE))
  {
    CK(in.alloc(n * dim));
    CK(cudaMemcpy(in.p, pts, n * dim * sizeof(uint32_t), cudaMemcpyHostToDevice));
    src = in.p;
  }

  // ---- Q: the nodes the construction splits -------------------------------------------------------------------------------
  TBuf<uint64_t> keys;
  {
    TBuf<uint64_t> k0;
    CK(k0.alloc(n)); CK(keys.alloc(n));
    TBuf<int> outside;
    CK(outside.alloc(1));
    CK(cudaMemsetAsync(outside.p, 0, sizeof(int), stream));
    TLAUNCH(k_point_keys, n, src, n, dim, depth, k0.p, outside.p);
    int h_out = 0;
    CK(cudaMemcpyAsync(&h_out, outside.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (h_out) { set_error("dkt_tree_from_points: a point coordinate is outside [0, 2^max_depth)"); return DKT_ERR_INVALID; }
    int rc = sort_keys(stream, k0.p, keys.p, n, dim * depth);
    if (rc) return rc;
  }
  in.release();
  stage("point keys + sort");
  std::vector<uint64_t> qBegin(depth + 1, 0), qEnd(depth + 1, 0);  // Q_l = Qs[qBegin[l], qEnd[l])
  TBuf<uint64_t> Qs;
  uint64_t nQ = 0;
  {
    TBuf<uint8_t> lo, hi;
    TBuf<uint64_t> cnt, off;
    CK(lo.alloc(n)); CK(hi.alloc(n)); CK(cnt.alloc(n)); CK(off.alloc(n));
    TLAUNCH(k_split_depth, n, keys.p, n, maxPts, dim, depth, lo.p, hi.p, cnt.p);
    int rc = scan(stream, cnt.p, off.p, n);
    if (rc) return rc;
    rc = scan_total(stream, cnt.p, off.p, n, nQ);
    if (rc) return rc;
    TBuf<uint64_t> tagged, tsorted;
    CK(tagged.alloc(nQ)); CK(tsorted.alloc(nQ)); CK(Qs.alloc(nQ));
    TLAUNCH(k_emit_split, n, keys.p, n, lo.p, hi.p, off.p, dim, depth, tagged.p);
    rc = sort_keys(stream, tagged.p, tsorted.p, nQ, 64);
    if (rc) return rc;
    TBuf<unsigned long long> hist;
    CK(hist.alloc(64));
    CK(cudaMemsetAsync(hist.p, 0xFF, 64 * sizeof(unsigned long long), stream));
    TLAUNCH(k_level_count, nQ, tsorted.p, nQ, hist.p);
    TLAUNCH(k_untag, nQ, tsorted.p, nQ, Qs.p);
    unsigned long long h[64];
    CK(cudaMemcpyAsync(h, hist.p, sizeof(h), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    uint64_t end = nQ;
    for (int l = depth; l >= 0; l--)
    {
      if (h[l] == ~0ull) { qBegin[l] = qEnd[l] = end; continue; }
      qBegin[l] = h[l];
      qEnd[l] = end;
      end = h[l];
    }
  }
  keys.release();
  stage("split nodes Q");

  // ---- P: bottom-up over the levels ----------------------------------------------------------------------------------------
  const int nch = 1 << dim;
  std::vector<TBuf<uint64_t>> P(depth + 1);
  std::vector<uint64_t> nP(depth + 1, 0);
  for (int l = depth - 1; l >= 0; l--)
  {
    const uint64_t nq = qEnd[l] - qBegin[l];
    const uint64_t nUp = (balance && l + 1 <= depth - 1) ? nP[l + 1] * nch : 0;
    const uint64_t nCand = nq + nUp + (l == 0 ? 1 : 0);
    TBuf<uint64_t> cand;
    CK(cand.alloc(nCand));
    if (nq) CK(cudaMemcpyAsync(cand.p, Qs.p + qBegin[l], nq * sizeof(uint64_t), cudaMemcpyDeviceToDevice, stream));
    if (nUp) TLAUNCH(k_parent_candidates, nUp, P[l + 1].p, nP[l + 1], dim, l + 1, cand.p + nq);
    if (l == 0) CK(cudaMemsetAsync(cand.p + nq + nUp, 0, sizeof(uint64_t), stream));  // the root is always split (src/tsort.cpp:702-706)
    int rc = sort_unique(stream, cand, nCand, dim * l, P[l], nP[l]);
    if (rc) return rc;
  }
  Qs.release();
  stage("levels bottom-up (P)");

  // ---- leaves ----------------------------------------------------------------------------------------------------------------
  uint64_t nLeaves = 0;
  int lmax = 0;
  std::vector<TBuf<uint64_t>> flag(depth), pos(depth);
  std::vector<uint64_t> nLeaf(depth, 0);
  for (int l = 0; l < depth; l++)
  {
    const uint64_t nt = nP[l] * nch;
    if (!nt) continue;
    CK(flag[l].alloc(nt)); CK(pos[l].alloc(nt));
    TLAUNCH(k_leaf_flags, nt, P[l].p, nP[l], P[l + 1].p, nP[l + 1], dim, flag[l].p);
    int rc = scan(stream, flag[l].p, pos[l].p, nt);
    if (rc) return rc;
    rc = scan_total(stream, flag[l].p, pos[l].p, nt, nLeaf[l]);
    if (rc) return rc;
    nLeaves += nLeaf[l];
    if (nLeaf[l]) lmax = l + 1;
  }
  if (nLeaves >= 0xFFFFFFFFull) { set_error("dkt_tree_from_points: more than 2^32 leaves"); return DKT_ERR_UNSUPPORTED; }
  TBuf<uint32_t> lxyz;
  TBuf<uint8_t> llev;
  CK(lxyz.alloc(nLeaves * dim)); CK(llev.alloc(nLeaves));
  {
    uint64_t o = 0;
    for (int l = 0; l < depth; l++)
    {
      const uint64_t nt = nP[l] * nch;
      if (!nt) continue;
      TLAUNCH(k_emit_leaves, nt, P[l].p, nP[l], flag[l].p, pos[l].p, dim, l, depth, lxyz.p + o * dim, llev.p + o);
      o += nLeaf[l];
    }
  }

  stage("leaves");

  // ---- tree order (SFC_Tree::locTreeSort, include/tsort.tcc:17-80) ----------------------------------------------------------
  SfcTables tab;
  make_sfc_tables(dim, tr.sfc_mode, tab);
  TBuf<uint8_t> dRot, dHt;
  CK(dRot.alloc(tab.rot_inv.size())); CK(dHt.alloc(tab.htab.size()));
  CK(cudaMemcpyAsync(dRot.p, tab.rot_inv.data(), tab.rot_inv.size(), cudaMemcpyHostToDevice, stream));
  CK(cudaMemcpyAsync(dHt.p, tab.htab.data(), tab.htab.size(), cudaMemcpyHostToDevice, stream));
  TBuf<uint64_t> k0, k1;
  TBuf<uint32_t> i0, i1;
  CK(k0.alloc(nLeaves)); CK(k1.alloc(nLeaves)); CK(i0.alloc(nLeaves)); CK(i1.alloc(nLeaves));
  TLAUNCH(k_leaf_order_keys, nLeaves, lxyz.p, llev.p, nLeaves, dim, depth, lmax, dRot.p, dHt.p, k0.p, i0.p);
  {
    size_t tmp = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, k0.p, k1.p, i0.p, i1.p, (int64_t)nLeaves, 0, std::max(1, lmax * dim), stream));
    TBuf<char> t;
    CK(t.alloc(tmp));
    CK(cub::DeviceRadixSort::SortPairs(t.p, tmp, k0.p, k1.p, i0.p, i1.p, (int64_t)nLeaves, 0, std::max(1, lmax * dim), stream));
    g_launches += 8;
    CK(cudaStreamSynchronize(stream));
  }
  CK(cudaMalloc((void **)&tr.d_xyz, std::max<uint64_t>(nLeaves, 1) * dim * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&tr.d_lev, std::max<uint64_t>(nLeaves, 1)));
  TLAUNCH(k_gather_leaves, nLeaves, lxyz.p, llev.p, i1.p, nLeaves, dim, tr.d_xyz, tr.d_lev);
  CK(cudaStreamSynchronize(stream));
  CK(cudaGetLastError());
  stage("tree order");
  tr.n = nLeaves;
  tr.finest_level = lmax;
  return DKT_OK;
}
}  // namespace dkt
