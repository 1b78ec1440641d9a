// Multi-GPU matvec: SFC-contiguous element ranges per rank + NCCL ghost exchange.
//
// Replaces the reference's distributed DA for the matvec path:
//   * partition ............ SFC_Tree::distTreePartition (src/tsort.cpp:229-508): contiguous ranges
//                            of the tree order, equal element counts
//   * node ownership ....... the rank of the first (lowest tree position) element touching the node
//                            (the reference: owner of the node's SFC key, include/nsort.tcc:715-786)
//   * ghosted vector ....... [owned | ghosts grouped by owner rank]  (reference: [pre | local | post],
//                            src/oda.cpp:115-121)
//   * readFromGhostBegin/End (include/oda.tcc:212-315): owners send the values other ranks ghost
//   * writeToGhostsBegin/End (include/oda.tcc:319-435): ghost partial sums go back and are added
// as ncclSend/ncclRecv groups over NVLink on the DA's stream.
//
// Round-1 construction strategy ("replicated build, partitioned matvec"): every rank builds the
// GLOBAL tables on its own GPU with the single-rank pipeline (dkt_build.cu) - identical on all
// ranks, so ownership and both sides of every send/recv list are derived without communication -
// then keeps only its element range, renumbers the nodes it touches and builds its chunk tables.
// The global tables are freed afterwards.  (A build whose memory scales with the local part only
// is future work; the matvec itself is fully partitioned.)
#include "dkt_internal.h"
#include "dkt_p2p.cuh"

// also compiles under -DDKT_EMU (tests/emu/cuda_emu.h): the CPU test-suite runs partition_da and the phased matvec of
// several ranks in one process; never part of libdkt.so that way
#ifdef DKT_EMU
#define DKT_DIST_LAUNCH(kern, grid, block, stream) ::emu::make_launch(kern, (grid), (block), 0)
#else
#include <cub/cub.cuh>
#define DKT_DIST_LAUNCH(kern, grid, block, stream) kern<<<(grid), (block), 0, (stream)>>>
#endif

#include <dlfcn.h>

#include <algorithm>
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

// ---- NCCL, resolved at run time (libnccl.so.2 ships with torch; no link-time dependency) ----------
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void *ncclComm_p;
struct NcclApi
{
  void *lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId_t *) = nullptr;
  int (*CommInitRank)(ncclComm_p *, int, ncclUniqueId_t, int) = nullptr;
  int (*CommDestroy)(ncclComm_p) = nullptr;
  int (*Send)(const void *, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, ncclComm_p, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
constexpr int NCCL_FLOAT64 = 8;
constexpr int NCCL_UINT8 = 1;

static int load_nccl()
{
#ifdef DKT_EMU
  return DKT_OK;  // dry-run partitions only
#endif
  if (g_nccl.lib) return DKT_OK;
  const char *cands[] = {getenv("DKT_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char *c : cands)
  {
    if (!c) continue;
    g_nccl.lib = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) { set_error(std::string("cannot load libnccl.so.2 (set DKT_NCCL_LIB): ") + dlerror()); return DKT_ERR_NCCL; }
#define SYM(field, name)                                                       \
  *(void **)(&g_nccl.field) = dlsym(g_nccl.lib, name);                         \
  if (!g_nccl.field) { set_error(std::string("libnccl lacks ") + name); return DKT_ERR_NCCL; }
  SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
  SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString") SYM(AllReduce, "ncclAllReduce") SYM(AllGather, "ncclAllGather")
#undef SYM
  return DKT_OK;
}
#define NCK(call)                                                                        \
  do                                                                                     \
  {                                                                                      \
    int r_ = (call);                                                                     \
    if (r_ != 0)                                                                         \
    {                                                                                    \
      set_error(std::string(#call) + ": " + g_nccl.GetErrorString(r_));                  \
      return DKT_ERR_NCCL;                                                               \
    }                                                                                    \
  } while (0)

int nccl_unique_id(void *out128)
{
  int rc = load_nccl();
  if (rc) return rc;
  NCK(g_nccl.GetUniqueId((ncclUniqueId_t *)out128));
  return DKT_OK;
}

// ---- kernels -------------------------------------------------------------------------------------
struct Bounds
{
  uint32_t b[66];  // b[p] = first tree position of rank p, b[nranks] = nMv
  int nranks;
};
__device__ __forceinline__ int rank_of(uint32_t src, const Bounds &B)
{
  int p = 0;
  while (p + 1 < B.nranks && src >= B.b[p + 1]) p++;
  return p;
}

// per node: smallest tree position of a referencing element, and the set of ranks referencing it
__global__ void k_node_refs(const uint32_t *ids, uint64_t n, int N, const uint32_t *src, uint64_t elem0, Bounds B, uint32_t *minsrc,
                            unsigned long long *refmask)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t node = ids[i];
  if (node == INVALID) return;
  const uint32_t s = src[elem0 + i / N];
  atomicMin(minsrc + node, s);
  atomicOr(refmask + node, 1ull << rank_of(s, B));
}

// flags for one selection pass: mode 0 = owned by me; 1 = my ghost owned by peer; 2 = owned by me and
// referenced by peer (send list)
__global__ void k_select(const uint32_t *minsrc, const unsigned long long *refmask, uint64_t n, Bounds B, int me, int peer, int mode,
                         uint8_t *flag)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int owner = rank_of(minsrc[i], B);
  bool f;
  if (mode == 0) f = owner == me;
  else if (mode == 1) f = owner == peer && ((refmask[i] >> me) & 1ull);
  else f = owner == me && ((refmask[i] >> peer) & 1ull);
  flag[i] = f ? 1 : 0;
}
__global__ void k_assign_local(const uint8_t *flag, const uint64_t *pos, uint64_t n, uint32_t base, uint32_t *g2l, uint32_t *l2g)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n || !flag[i]) return;
  const uint32_t l = base + (uint32_t)pos[i];
  g2l[i] = l;
  l2g[l] = (uint32_t)i;
}
__global__ void k_send_list(const uint8_t *flag, const uint64_t *pos, uint64_t n, const uint32_t *g2l, uint32_t *out)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n && flag[i]) out[pos[i]] = g2l[i];
}
__global__ void k_u8_widen(const uint8_t *in, uint64_t n, uint64_t *out)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}
__global__ void k_remap(const uint32_t *in, uint64_t n, const uint32_t *g2l, uint32_t *out)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t v = in[i];
  out[i] = v == INVALID ? INVALID : g2l[v];
}
__global__ void k_gather_rows32(const uint32_t *src, const uint32_t *idx, uint64_t n, int width, uint32_t *dst)
{
  const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (t >= n * width) return;
  dst[t] = src[(uint64_t)idx[t / width] * width + t % width];
}
__global__ void k_gather_u8(const uint8_t *in, const uint32_t *idx, uint64_t n, uint8_t *out)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[idx[i]];
}
__global__ void k_lower_bound(const uint32_t *a, uint64_t n, uint32_t key, uint64_t *out)
{
  if (blockIdx.x || threadIdx.x) return;
  uint64_t lo = 0, hi = n;
  while (lo < hi)
  {
    const uint64_t mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  *out = lo;
}
// weight of the element at each tree position (visited list = [regular | hanging], d_mv_src = position)
__global__ void k_weights(const uint32_t *src, uint64_t nMv, uint64_t nReg, uint64_t wr, uint64_t wh, uint64_t *w)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < nMv) w[src[i]] = i < nReg ? wr : wh;
}
__global__ void k_lower_bound64(const uint64_t *a, uint64_t n, uint64_t key, uint64_t *out)
{
  if (blockIdx.x || threadIdx.x) return;
  uint64_t lo = 0, hi = n;
  while (lo < hi)
  {
    const uint64_t mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  *out = lo;
}
// boundary element = touches a ghost node (local id >= nOwned) through its lattice or parent-lattice slots
__global__ void k_is_boundary(const uint32_t *e2n, const uint32_t *pnode, uint64_t n, int N, uint32_t nOwned, uint8_t *flag)
{
  uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (e >= n) return;
  bool b = false;
  for (int r = 0; r < N; r++)
  {
    const uint32_t a = e2n[e * N + r];
    b |= (a != INVALID && a >= nOwned);
    if (pnode)
    {
      const uint32_t q = pnode[e * N + r];
      b |= (q != INVALID && q >= nOwned);
    }
  }
  flag[e] = b ? 1 : 0;
}
// stable partition destination: interior elements first, boundary elements after
__global__ void k_partition_dst(const uint8_t *flag, const uint64_t *bpos, uint64_t n, uint64_t nInterior, uint32_t *dst)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (uint32_t)(flag[i] ? nInterior + bpos[i] : i - bpos[i]);
}
template <typename T>
__global__ void k_permute_rows(const T *in, const uint32_t *dst, uint64_t n, int width, T *out)
{
  uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (t >= n * width) return;
  const uint64_t i = t / width;
  out[(uint64_t)dst[i] * width + t % width] = in[t];
}
__global__ void k_pack(const double *v, const uint32_t *idx, uint64_t n, double *buf)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) buf[i] = v[idx[i]];
}
// several peers may return contributions to the same owned node -> atomic
__global__ void k_unpack_add(double *v, const uint32_t *idx, uint64_t n, const double *buf)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(v + idx[i], buf[i]);
}

__global__ void k_mark_u8(const uint32_t *idx, uint64_t n, uint8_t *flag)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) flag[idx[i]] = 1;
}
#define LAUNCHS(kern, n, stream, ...)                                          \
  do                                                                           \
  {                                                                            \
    if ((n) > 0)                                                               \
    {                                                                          \
      DKT_DIST_LAUNCH(kern, (unsigned)(((n) + 255) / 256), 256, stream)(__VA_ARGS__); \
      g_launches++;                                                            \
    }                                                                          \
  } while (0)

// exclusive positions of set flags + total (host)
static int positions(DA &g, const uint8_t *flag, uint64_t n, uint64_t *wide, uint64_t *pos, uint64_t &total)
{
  total = 0;
  if (n == 0) return DKT_OK;
  CK(cudaMemsetAsync(wide + n, 0, sizeof(uint64_t), g.stream));
  LAUNCHS(k_u8_widen, n, g.stream, flag, n, wide);
  int rc = device_exclusive_scan(g, wide, pos, n + 1);
  if (rc) return rc;
  CK(cudaMemcpy(&total, pos + n, sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return DKT_OK;
}

int dist_allreduce(Dist &d, double *red, cudaStream_t s)
{
  if (!d.comm) { set_error("no communicator"); return DKT_ERR_NCCL; }
  NCK(g_nccl.AllReduce(red, red, 2, NCCL_FLOAT64, 0 /* ncclSum */, (ncclComm_p)d.comm, s));
  NCK(g_nccl.AllReduce(red + 2, red + 2, 1, NCCL_FLOAT64, 2 /* ncclMax */, (ncclComm_p)d.comm, s));
  return DKT_OK;
}

void free_dist(Dist &d)
{
  if (d.timing && d.tcount)
    fprintf(stderr, "[dkt rank %d] %d overlapped matvecs, min ms from start: ghost values in %.4f, interior done %.4f, boundary done %.4f, end %.4f\n",
            d.rank, d.tcount, d.tsum[0], d.tsum[1], d.tsum[2], d.tsum[3]);
  for (int i = 0; i < 5; i++)
    if (d.tev[i]) cudaEventDestroy(d.tev[i]);
  if (d.comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_p)d.comm);
  for (int i = 0; i < 4; i++)
    if (d.ev[i]) cudaEventDestroy(d.ev[i]);
  if (d.comm_stream) cudaStreamDestroy(d.comm_stream);
  cudaFree(d.d_send_idx); cudaFree(d.d_send_buf); cudaFree(d.d_recv_buf); cudaFree(d.d_in_local); cudaFree(d.d_out_local);
  cudaFree(d.d_owned_gid);
  for (void *b : d.peer_base)
    if (b) cudaIpcCloseMemHandle(b);
  cudaFree(d.xbuf); cudaFree(d.d_peer_xr); cudaFree(d.d_peer_xw); cudaFree(d.d_peer_flagR); cudaFree(d.d_peer_flagW);
  cudaFree(d.d_send_off); cudaFree(d.d_recv_off); cudaFree(d.d_p2p_err);
  d = Dist();
}

// Peer-memory exchange: allocate the exchange buffer, trade IPC handles and segment offsets through the (already
// initialised) NCCL communicator, map the peers' buffers and build the device pointer tables.
static int p2p_alloc(Dist &d, P2PInfo &mine)
{
  const int R = d.nranks;
  if (R > P2P_MAX_RANKS) { set_error("DKT_DIST_P2P: at most 64 ranks"); return DKT_ERR_UNSUPPORTED; }
  const uint64_t nGhost = d.recv_off[R], totalSend = d.send_off[R];
  const size_t bytes = P2P_FLAG_BYTES + (nGhost + totalSend + 1) * sizeof(double);
  CK(cudaMalloc((void **)&d.xbuf, bytes));
  CK(cudaMemset(d.xbuf, 0, bytes));
  std::memset(&mine, 0, sizeof(mine));
  mine.nGhost = nGhost;
  mine.totalSend = totalSend;
  for (int p = 0; p <= R; p++) { mine.recv_off[p] = d.recv_off[p]; mine.send_off[p] = d.send_off[p]; }
  return DKT_OK;
}
// base[p]: rank p's xbuf as seen from this process (nullptr: no exchange with p), info[p]: its segment offsets
static int p2p_wire(Dist &d, const std::vector<P2PInfo> &info, const std::vector<void *> &base)
{
  const int R = d.nranks, me = d.rank;
  std::vector<double *> xr, xw;
  std::vector<uint32_t *> fr, fw;
  std::string why;
  if (!p2p_tables(me, R, d.send_off.data(), d.recv_off.data(), info.data(), base.data(), xr, xw, fr, fw, why))
  {
    set_error("DKT_DIST_P2P: " + why);
    return DKT_ERR_INVALID;
  }
  CK(cudaMalloc((void **)&d.d_peer_xr, R * sizeof(double *)));
  CK(cudaMalloc((void **)&d.d_peer_xw, R * sizeof(double *)));
  CK(cudaMalloc((void **)&d.d_peer_flagR, R * sizeof(uint32_t *)));
  CK(cudaMalloc((void **)&d.d_peer_flagW, R * sizeof(uint32_t *)));
  CK(cudaMalloc((void **)&d.d_send_off, (R + 1) * sizeof(uint64_t)));
  CK(cudaMalloc((void **)&d.d_recv_off, (R + 1) * sizeof(uint64_t)));
  CK(cudaMalloc((void **)&d.d_p2p_err, sizeof(int)));
  CK(cudaMemset(d.d_p2p_err, 0, sizeof(int)));
  CK(cudaMemcpy(d.d_peer_xr, xr.data(), R * sizeof(double *), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d.d_peer_xw, xw.data(), R * sizeof(double *), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d.d_peer_flagR, fr.data(), R * sizeof(uint32_t *), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d.d_peer_flagW, fw.data(), R * sizeof(uint32_t *), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d.d_send_off, d.send_off.data(), (R + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d.d_recv_off, d.recv_off.data(), (R + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice));
  d.p2p = true;
  return DKT_OK;
}
static int setup_p2p(DA &g, Dist &d)
{
  const int R = d.nranks, me = d.rank;
  std::vector<P2PInfo> info(R);
  int rc = p2p_alloc(d, info[me]);
  if (rc) return rc;
  CK(cudaIpcGetMemHandle(&info[me].handle, d.xbuf));
  char *dinfo = nullptr;
  CK(cudaMalloc((void **)&dinfo, sizeof(P2PInfo) * R));
  CK(cudaMemcpyAsync(dinfo + sizeof(P2PInfo) * me, &info[me], sizeof(P2PInfo), cudaMemcpyHostToDevice, g.stream));
  NCK(g_nccl.AllGather(dinfo + sizeof(P2PInfo) * me, dinfo, sizeof(P2PInfo), NCCL_UINT8, (ncclComm_p)d.comm, g.stream));
  CK(cudaMemcpyAsync(info.data(), dinfo, sizeof(P2PInfo) * R, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  cudaFree(dinfo);
  d.peer_base.assign(R, nullptr);
  std::vector<void *> base(R, nullptr);
  for (int p = 0; p < R; p++)
  {
    if (p == me) continue;
    if (d.send_off[p + 1] == d.send_off[p] && d.recv_off[p + 1] == d.recv_off[p]) continue;
    CK(cudaIpcOpenMemHandle(&base[p], info[p].handle, cudaIpcMemLazyEnablePeerAccess));
    d.peer_base[p] = base[p];
  }
  return p2p_wire(d, info, base);
}
// Single-process variant for tests: the R ranks of one partition live in this process (dry-run DAs, possibly all on
// one GPU) and see each other's exchange buffers directly - the same kernels and protocol without IPC and NCCL.
int p2p_attach_local(Dist **ranks, int R)
{
  std::vector<P2PInfo> info(R);
  std::vector<void *> base(R, nullptr);
  for (int p = 0; p < R; p++)
  {
    if (!ranks[p] || !ranks[p]->active || ranks[p]->nranks != R || ranks[p]->rank != p || ranks[p]->p2p)
    {
      set_error("p2p_attach_local: pass the dry-run DAs of ranks 0..R-1 of one partition, once");
      return DKT_ERR_INVALID;
    }
    int rc = p2p_alloc(*ranks[p], info[p]);
    if (rc) return rc;
    base[p] = ranks[p]->xbuf;
  }
  for (int p = 0; p < R; p++)
  {
    int rc = p2p_wire(*ranks[p], info, base);
    if (rc) return rc;
  }
  return DKT_OK;
}

// g: the global single-rank DA (already built, no chunk tables).  On success `g` has been turned into
// the LOCAL DA of this rank (its global tables released) and `dist` describes the exchange.
int partition_da(DA &g, Dist &dist, int rank, int nranks, const void *nccl_id)
{
  if (nranks < 1 || nranks > 64 || rank < 0 || rank >= nranks) { set_error("bad rank/nranks (1..64)"); return DKT_ERR_INVALID; }
  int rc = load_nccl();
  if (rc) return rc;
  dist.rank = rank;
  dist.nranks = nranks;
  const int N = g.N, dim = g.dim;
  const uint64_t nMv = g.nMv, nReg = g.nReg, nHang = g.nHang, nNodes = g.nNodes;
  dist.nGlobalNodes = nNodes;
  dist.nGlobalElems = g.nElem;
  // ---- SFC-contiguous ranges of equal WEIGHT: a hanging element (two slot rows, interpolation both
  //      ways) costs HANG_WEIGHT/REG_WEIGHT regular ones in the chunked kernels
  Bounds B;
  B.nranks = nranks;
  {
    constexpr uint64_t REG_WEIGHT = 3, HANG_WEIGHT = 5;  // measured: 45 vs 75 ns per element (profiles/)
    uint64_t *wsrc = nullptr, *wscan = nullptr, *dbound = nullptr;
    CK(cudaMalloc((void **)&wsrc, (nMv + 1) * sizeof(uint64_t)));
    CK(cudaMalloc((void **)&wscan, (nMv + 1) * sizeof(uint64_t)));
    CK(cudaMalloc((void **)&dbound, (nranks + 1) * sizeof(uint64_t)));
    CK(cudaMemsetAsync(wsrc, 0, (nMv + 1) * sizeof(uint64_t), g.stream));
    LAUNCHS(k_weights, nMv, g.stream, g.d_mv_src, nMv, nReg, REG_WEIGHT, HANG_WEIGHT, wsrc);
    rc = device_exclusive_scan(g, wsrc, wscan, nMv + 1);
    if (rc) return rc;
    uint64_t W = 0;
    CK(cudaMemcpy(&W, wscan + nMv, sizeof(uint64_t), cudaMemcpyDeviceToHost));
    for (int p = 0; p <= nranks; p++)
    {
      DKT_DIST_LAUNCH(k_lower_bound64, 1, 1, g.stream)(wscan, nMv + 1, (W * (uint64_t)p) / (uint64_t)nranks, dbound + p);
      g_launches++;
    }
    std::vector<uint64_t> hb(nranks + 1);
    // g.stream is a non-blocking stream: a plain cudaMemcpy would not wait for the kernels above
    CK(cudaMemcpyAsync(hb.data(), dbound, (nranks + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    for (int p = 0; p <= nranks; p++) B.b[p] = (uint32_t)std::min<uint64_t>(hb[p], nMv);
    B.b[0] = 0;
    B.b[nranks] = (uint32_t)nMv;
    cudaFree(wsrc); cudaFree(wscan); cudaFree(dbound);
  }

  // ---- ownership and reference masks ------------------------------------------------------------------
  uint32_t *minsrc = nullptr;
  unsigned long long *refmask = nullptr;
  CK(cudaMalloc((void **)&minsrc, std::max<uint64_t>(nNodes, 1) * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&refmask, std::max<uint64_t>(nNodes, 1) * sizeof(unsigned long long)));
  CK(cudaMemsetAsync(minsrc, 0xFF, nNodes * sizeof(uint32_t), g.stream));
  CK(cudaMemsetAsync(refmask, 0, nNodes * sizeof(unsigned long long), g.stream));
  LAUNCHS(k_node_refs, nMv * N, g.stream, g.d_e2n, nMv * N, N, g.d_mv_src, 0, B, minsrc, refmask);
  LAUNCHS(k_node_refs, nHang * N, g.stream, g.d_pnode, nHang * N, N, g.d_mv_src, nReg, B, minsrc, refmask);

  // ---- local numbering: [owned (global order) | ghosts by owner rank (global order inside)] -----------
  uint8_t *flag = nullptr;
  uint64_t *wide = nullptr, *pos = nullptr;
  uint32_t *g2l = nullptr, *l2g = nullptr;
  CK(cudaMalloc((void **)&flag, std::max<uint64_t>(nNodes, 1)));
  CK(cudaMalloc((void **)&wide, (nNodes + 1) * sizeof(uint64_t)));
  CK(cudaMalloc((void **)&pos, (nNodes + 1) * sizeof(uint64_t)));
  CK(cudaMalloc((void **)&g2l, std::max<uint64_t>(nNodes, 1) * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&l2g, std::max<uint64_t>(nNodes, 1) * sizeof(uint32_t)));
  CK(cudaMemsetAsync(g2l, 0xFF, nNodes * sizeof(uint32_t), g.stream));
  uint64_t nOwned = 0;
  LAUNCHS(k_select, nNodes, g.stream, minsrc, refmask, nNodes, B, rank, rank, 0, flag);
  rc = positions(g, flag, nNodes, wide, pos, nOwned);
  if (rc) return rc;
  LAUNCHS(k_assign_local, nNodes, g.stream, flag, pos, nNodes, 0u, g2l, l2g);
  dist.nOwned = nOwned;
  dist.recv_off.assign(nranks + 1, 0);
  dist.send_off.assign(nranks + 1, 0);
  uint64_t nLocal = nOwned;
  for (int p = 0; p < nranks; p++)
  {
    dist.recv_off[p] = nLocal - nOwned;
    if (p == rank) continue;
    uint64_t cnt = 0;
    LAUNCHS(k_select, nNodes, g.stream, minsrc, refmask, nNodes, B, rank, p, 1, flag);
    rc = positions(g, flag, nNodes, wide, pos, cnt);
    if (rc) return rc;
    LAUNCHS(k_assign_local, nNodes, g.stream, flag, pos, nNodes, (uint32_t)nLocal, g2l, l2g);
    nLocal += cnt;
  }
  dist.recv_off[nranks] = nLocal - nOwned;
  dist.nGhost = nLocal - nOwned;
  // ---- send lists ------------------------------------------------------------------------------------------
  std::vector<uint64_t> scnt(nranks, 0);
  uint64_t totalSend = 0;
  for (int p = 0; p < nranks; p++)
  {
    if (p == rank) continue;
    LAUNCHS(k_select, nNodes, g.stream, minsrc, refmask, nNodes, B, rank, p, 2, flag);
    rc = positions(g, flag, nNodes, wide, pos, scnt[p]);
    if (rc) return rc;
    totalSend += scnt[p];
  }
  CK(cudaMalloc((void **)&dist.d_send_idx, std::max<uint64_t>(totalSend, 1) * sizeof(uint32_t)));
  {
    uint64_t off = 0;
    for (int p = 0; p < nranks; p++)
    {
      dist.send_off[p] = off;
      if (p == rank || scnt[p] == 0) continue;
      uint64_t c2 = 0;
      LAUNCHS(k_select, nNodes, g.stream, minsrc, refmask, nNodes, B, rank, p, 2, flag);
      rc = positions(g, flag, nNodes, wide, pos, c2);
      if (rc) return rc;
      LAUNCHS(k_send_list, nNodes, g.stream, flag, pos, nNodes, g2l, dist.d_send_idx + off);
      off += scnt[p];
    }
    dist.send_off[nranks] = off;
  }
  CK(cudaMalloc((void **)&dist.d_send_buf, std::max<uint64_t>(totalSend, 1) * sizeof(double)));
  CK(cudaMalloc((void **)&dist.d_recv_buf, std::max<uint64_t>(totalSend, 1) * sizeof(double)));
  CK(cudaMalloc((void **)&dist.d_in_local, std::max<uint64_t>(nLocal, 1) * sizeof(double)));
  CK(cudaMalloc((void **)&dist.d_out_local, std::max<uint64_t>(nLocal, 1) * sizeof(double)));
  CK(cudaMalloc((void **)&dist.d_owned_gid, std::max<uint64_t>(nOwned, 1) * sizeof(uint32_t)));
  CK(cudaMemcpyAsync(dist.d_owned_gid, l2g, nOwned * sizeof(uint32_t), cudaMemcpyDeviceToDevice, g.stream));

  // ---- my element ranges (regular and hanging lists are each sorted by tree position) ----------------------
  uint64_t *dlb = nullptr, lb[4];
  CK(cudaMalloc((void **)&dlb, 4 * sizeof(uint64_t)));
  DKT_DIST_LAUNCH(k_lower_bound, 1, 1, g.stream)(g.d_mv_src, nReg, B.b[rank], dlb + 0);
  DKT_DIST_LAUNCH(k_lower_bound, 1, 1, g.stream)(g.d_mv_src, nReg, B.b[rank + 1], dlb + 1);
  DKT_DIST_LAUNCH(k_lower_bound, 1, 1, g.stream)(g.d_mv_src + nReg, nHang, B.b[rank], dlb + 2);
  DKT_DIST_LAUNCH(k_lower_bound, 1, 1, g.stream)(g.d_mv_src + nReg, nHang, B.b[rank + 1], dlb + 3);
  g_launches += 4;
  CK(cudaMemcpyAsync(lb, dlb, sizeof(lb), cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  cudaFree(dlb);
  const uint64_t r0 = lb[0], r1 = lb[1], h0 = lb[2], h1 = lb[3];
  const uint64_t nRegL = r1 - r0, nHangL = h1 - h0, nMvL = nRegL + nHangL;

  // ---- local tables ---------------------------------------------------------------------------------------------
  uint32_t *e2nL = nullptr, *pnodeL = nullptr, *xyzL = nullptr, *srcL = nullptr;
  uint8_t *levL = nullptr, *childL = nullptr, *bdyL = nullptr;
  CK(cudaMalloc((void **)&e2nL, std::max<uint64_t>(nMvL, 1) * N * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&pnodeL, std::max<uint64_t>(nHangL, 1) * N * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&xyzL, std::max<uint64_t>(nMvL, 1) * dim * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&srcL, std::max<uint64_t>(nMvL, 1) * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&levL, std::max<uint64_t>(nMvL, 1)));
  CK(cudaMalloc((void **)&childL, std::max<uint64_t>(nHangL, 1)));
  CK(cudaMalloc((void **)&bdyL, std::max<uint64_t>(nLocal, 1)));
  LAUNCHS(k_remap, nRegL * N, g.stream, g.d_e2n + r0 * N, nRegL * N, g2l, e2nL);
  LAUNCHS(k_remap, nHangL * N, g.stream, g.d_e2n + (nReg + h0) * N, nHangL * N, g2l, e2nL + nRegL * N);
  LAUNCHS(k_remap, nHangL * N, g.stream, g.d_pnode + h0 * N, nHangL * N, g2l, pnodeL);
  CK(cudaMemcpyAsync(xyzL, g.d_mv_xyz + r0 * dim, nRegL * dim * sizeof(uint32_t), cudaMemcpyDeviceToDevice, g.stream));
  CK(cudaMemcpyAsync(xyzL + nRegL * dim, g.d_mv_xyz + (nReg + h0) * dim, nHangL * dim * sizeof(uint32_t), cudaMemcpyDeviceToDevice, g.stream));
  CK(cudaMemcpyAsync(levL, g.d_mv_lev + r0, nRegL, cudaMemcpyDeviceToDevice, g.stream));
  CK(cudaMemcpyAsync(levL + nRegL, g.d_mv_lev + nReg + h0, nHangL, cudaMemcpyDeviceToDevice, g.stream));
  CK(cudaMemcpyAsync(srcL, g.d_mv_src + r0, nRegL * sizeof(uint32_t), cudaMemcpyDeviceToDevice, g.stream));
  CK(cudaMemcpyAsync(srcL + nRegL, g.d_mv_src + nReg + h0, nHangL * sizeof(uint32_t), cudaMemcpyDeviceToDevice, g.stream));
  CK(cudaMemcpyAsync(childL, g.d_child + h0, nHangL, cudaMemcpyDeviceToDevice, g.stream));
  LAUNCHS(k_gather_u8, nLocal, g.stream, g.d_node_isbdy, l2g, nLocal, bdyL);
  // node coordinates / levels of the local vector [owned | ghosts] (DA::getTNCoords of a rank, src/oda.cpp:131-140) and the owned
  // nodes on the domain boundary (DA::getBoundaryNodeIndices): what dkt_da_export_nodes / _boundary return on a partitioned DA
  uint32_t *nodeXyzL = nullptr, *bdyIdsL = nullptr;
  uint8_t *nodeLevL = nullptr;
  uint64_t nBdyOwned = 0;
  CK(cudaMalloc((void **)&nodeXyzL, std::max<uint64_t>(nLocal, 1) * dim * sizeof(uint32_t)));
  CK(cudaMalloc((void **)&nodeLevL, std::max<uint64_t>(nLocal, 1)));
  LAUNCHS(k_gather_rows32, nLocal * dim, g.stream, g.d_node_xyz, l2g, nLocal, dim, nodeXyzL);
  LAUNCHS(k_gather_u8, nLocal, g.stream, g.d_node_lev, l2g, nLocal, nodeLevL);
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaGetLastError());
  {
    std::vector<uint8_t> hb(nOwned);
    if (nOwned) CK(cudaMemcpy(hb.data(), bdyL, nOwned, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> ids;
    for (uint64_t i = 0; i < nOwned; i++)
      if (hb[i]) ids.push_back((uint32_t)i);
    nBdyOwned = ids.size();
    CK(cudaMalloc((void **)&bdyIdsL, std::max<uint64_t>(nBdyOwned, 1) * sizeof(uint32_t)));
    if (nBdyOwned) CK(cudaMemcpy(bdyIdsL, ids.data(), nBdyOwned * sizeof(uint32_t), cudaMemcpyHostToDevice));
  }

  // ---- [interior | boundary] order inside the regular and the hanging list, so that interior elements can
  //      run while the ghost exchanges are in flight --------------------------------------------------------------
  // Measured on 8 B200s (profiles/README.md): the split pays at 8 ranks (0.95 vs 1.00 ms) and costs at 2
  // (three kernel pairs instead of one); default: on for more than 2 ranks, DKT_DIST_OVERLAP=0/1 overrides.
  bool wantOverlap = nranks > 1;
  if (const char *e = getenv("DKT_DIST_OVERLAP")) wantOverlap = atoi(e) != 0;
  uint64_t nRegInt = nRegL, nHangInt = nHangL;
  for (int pass = 0; pass < 2 && wantOverlap; pass++)
  {
    const uint64_t cnt = pass == 0 ? nRegL : nHangL, first = pass == 0 ? 0 : nRegL;
    if (cnt == 0) continue;
    uint8_t *bflag = nullptr;
    uint64_t *bw = nullptr, *bp = nullptr;
    uint32_t *dst = nullptr;
    CK(cudaMalloc((void **)&bflag, cnt));
    CK(cudaMalloc((void **)&bw, (cnt + 1) * sizeof(uint64_t)));
    CK(cudaMalloc((void **)&bp, (cnt + 1) * sizeof(uint64_t)));
    CK(cudaMalloc((void **)&dst, cnt * sizeof(uint32_t)));
    LAUNCHS(k_is_boundary, cnt, g.stream, e2nL + first * N, pass == 0 ? (const uint32_t *)nullptr : (const uint32_t *)pnodeL, cnt, N,
            (uint32_t)nOwned, bflag);
    uint64_t nB = 0;
    rc = positions(g, bflag, cnt, bw, bp, nB);
    if (rc) return rc;
    const uint64_t nI = cnt - nB;
    LAUNCHS(k_partition_dst, cnt, g.stream, bflag, bp, cnt, nI, dst);
    auto permute32 = [&](uint32_t *&arr, uint64_t off, int width) -> int {
      uint32_t *tmp = nullptr;
      CK(cudaMalloc((void **)&tmp, cnt * width * sizeof(uint32_t)));
      LAUNCHS(k_permute_rows<uint32_t>, cnt * width, g.stream, arr + off * width, dst, cnt, width, tmp);
      CK(cudaMemcpyAsync(arr + off * width, tmp, cnt * width * sizeof(uint32_t), cudaMemcpyDeviceToDevice, g.stream));
      CK(cudaStreamSynchronize(g.stream));
      cudaFree(tmp);
      return DKT_OK;
    };
    auto permute8 = [&](uint8_t *&arr, uint64_t off) -> int {
      uint8_t *tmp = nullptr;
      CK(cudaMalloc((void **)&tmp, cnt));
      LAUNCHS(k_permute_rows<uint8_t>, cnt, g.stream, arr + off, dst, cnt, 1, tmp);
      CK(cudaMemcpyAsync(arr + off, tmp, cnt, cudaMemcpyDeviceToDevice, g.stream));
      CK(cudaStreamSynchronize(g.stream));
      cudaFree(tmp);
      return DKT_OK;
    };
    if ((rc = permute32(e2nL, first, N))) return rc;
    if ((rc = permute32(xyzL, first, dim))) return rc;
    if ((rc = permute32(srcL, first, 1))) return rc;
    if ((rc = permute8(levL, first))) return rc;
    if (pass == 1)
    {
      if ((rc = permute32(pnodeL, 0, N))) return rc;
      if ((rc = permute8(childL, 0))) return rc;
      nHangInt = nI;
    }
    else
      nRegInt = nI;
    cudaFree(bflag); cudaFree(bw); cudaFree(bp); cudaFree(dst);
  }
  g.phased = wantOverlap;
  g.commSMs = 0;  // measured at 8 ranks: reserving 16 SMs for NCCL is slower (0.987 ms) than not (0.951 ms)
  if (const char *e = getenv("DKT_COMM_SMS")) g.commSMs = atoi(e);
  g.nRegInterior = nRegInt;
  g.nHangInterior = nHangInt;

  // ---- swap the global tables for the local ones -------------------------------------------------------------------
  cudaFree(g.d_e2n); cudaFree(g.d_pnode); cudaFree(g.d_mv_xyz); cudaFree(g.d_mv_src); cudaFree(g.d_mv_lev); cudaFree(g.d_child);
  cudaFree(g.d_node_isbdy); cudaFree(g.d_ukey); cudaFree(g.d_unode);
  cudaFree(g.d_node_xyz); cudaFree(g.d_node_lev); cudaFree(g.d_bdy);
  g.d_node_xyz = nodeXyzL; g.d_node_lev = nodeLevL; g.d_bdy = bdyIdsL; g.nBdy = nBdyOwned;
  g.d_ukey = nullptr; g.d_unode = nullptr;
  g.d_e2n = e2nL; g.d_pnode = pnodeL; g.d_mv_xyz = xyzL; g.d_mv_src = srcL; g.d_mv_lev = levL; g.d_child = childL;
  g.d_node_isbdy = bdyL;
  g.nMv = nMvL; g.nReg = nRegL; g.nHang = nHangL;
  g.mv_src0 = B.b[rank];  // the rank's elements are the visit-order positions [b[rank], b[rank+1])
  g.nNodes = nLocal;  // the chunk tables and the kernels work on the local (owned + ghost) vector
  if (wantOverlap && totalSend)
  {
    // owned nodes that other ranks ghost: RED in every chunk (see DA::d_node_sent)
    CK(cudaMalloc((void **)&g.d_node_sent, nLocal));
    CK(cudaMemsetAsync(g.d_node_sent, 0, nLocal, g.stream));
    LAUNCHS(k_mark_u8, totalSend, g.stream, dist.d_send_idx, totalSend, g.d_node_sent);
    CK(cudaStreamSynchronize(g.stream));
  }
  cudaFree(minsrc); cudaFree(refmask); cudaFree(flag); cudaFree(wide); cudaFree(pos); cudaFree(g2l); cudaFree(l2g);

  // ---- communicator ------------------------------------------------------------------------------------------------------
  if (nranks > 1 && nccl_id)
  {
    ncclUniqueId_t id;
    std::memcpy(&id, nccl_id, sizeof(id));
    ncclComm_p comm = nullptr;
    NCK(g_nccl.CommInitRank(&comm, nranks, id, rank));
    dist.comm = comm;
    // Ghost exchange: grouped ncclSend/Recv by default.  DKT_DIST_P2P=1 selects peer-memory puts + epoch flags instead (a rank on
    // which the IPC mapping fails makes ALL ranks fall back to NCCL).  Measured on 8 B200s with the two-stream schedule of
    // run_matvec_dist (round 2): NCCL 0.488 ms per matvec, peer memory 0.534 ms; on 2 B200s both 0.462 ms.
    const char *e = getenv("DKT_DIST_P2P");
    if (e && atoi(e) != 0)
    {
      const int mine = setup_p2p(g, dist) == DKT_OK ? 1 : 0;
      double *agree = nullptr;
      CK(cudaMalloc((void **)&agree, sizeof(double)));
      const double h = (double)mine;
      CK(cudaMemcpyAsync(agree, &h, sizeof(double), cudaMemcpyHostToDevice, g.stream));
      NCK(g_nccl.AllReduce(agree, agree, 1, NCCL_FLOAT64, 3 /* ncclMin */, (ncclComm_p)dist.comm, g.stream));
      double all = 0.0;
      CK(cudaMemcpyAsync(&all, agree, sizeof(double), cudaMemcpyDeviceToHost, g.stream));
      CK(cudaStreamSynchronize(g.stream));
      cudaFree(agree);
      if (all < 0.5)
      {
        dist.p2p = false;  // the buffers are released with the DA
        set_error("");
      }
    }
  }
  {
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&dist.comm_stream, cudaStreamNonBlocking, hi));  // exchanges first
  }
  for (int i = 0; i < 4; i++) CK(cudaEventCreateWithFlags(&dist.ev[i], cudaEventDisableTiming));
  if (const char *e = getenv("DKT_DIST_TIMING"))
    if (atoi(e) != 0)
    {
      dist.timing = true;
      for (int i = 0; i < 5; i++) CK(cudaEventCreate(&dist.tev[i]));
    }
  dist.active = true;
  return DKT_OK;
}

// run_matvec_chunked of one phase on another stream (the exchange stream): single-stream launch, the DA's aux streams and
// their events belong to the call on the main stream
static int chunked_on(DA &da, cudaStream_t st, const dkt_op *op, const double *in, double *out, double scale, unsigned flags,
                      unsigned phaseMask)
{
  cudaStream_t keep = da.stream;
  const int keepStreams = da.mvStreams;
  da.stream = st;
  da.mvStreams = 1;
  const int rc = run_matvec_chunked(da, op, in, out, scale, flags, phaseMask, false);
  da.stream = keep;
  da.mvStreams = keepStreams;
  return rc;
}

// Schedule of a partitioned matvec with comm/compute overlap (da.phased), both exchanges:
//   main stream s:      zero out_local ............ interior elements (one launch per set) ............ join, add what came back
//   exchange stream cs: (high priority) ghost read -> boundary elements -> ghost write-back
// The interior elements touch no ghost node, the boundary ones run as soon as the ghost values are there, and both kinds
// accumulate shared nodes with RED (a node private to one chunk belongs to one set), so the two streams need no ordering
// between them.  The owners add the returned partial sums after the join: that kernel must not race with plain stores.
//
// The peer-memory exchange (DKT_DIST_P2P=1): puts into the peers' buffers + epoch flags instead of NCCL send/recv.
// `stages` (bit mask) exists for the single-process emulation of several ranks, where a rank cannot wait for a flag
// another rank has not yet had the chance to raise: 1 = put + signal (+ interior elements), 2 = wait for the ghost
// values, boundary (or all) elements, write-back put + signal, 4 = wait + accumulate.
static int run_matvec_dist_p2p(DA &da, Dist &d, const dkt_op *op, const double *d_in, double *d_out, double *in_local, double *out_local,
                               double scale, unsigned flags, bool overlap, bool ghosted, unsigned stages = 7u)
{
  cudaStream_t s = da.stream;
  cudaStream_t cs = overlap ? d.comm_stream : s;
  const int R = d.nranks;
  const uint64_t nOwned = d.nOwned, totalSend = d.send_off[R], nGhost = d.recv_off[R];
  if (stages & 1u) ++d.epoch;
  const uint32_t epoch = d.epoch;
  double *xr = (double *)(d.xbuf + P2P_FLAG_BYTES), *xw = xr + nGhost;
  const volatile uint32_t *flagR = (const volatile uint32_t *)d.xbuf, *flagW = flagR + P2P_MAX_RANKS;
  int rc = DKT_OK;
  if (stages & 1u)
  {
    if (overlap)
    {
      CK(cudaMemsetAsync(out_local, 0, (nOwned + nGhost) * sizeof(double), s));
      g_launches++;
      CK(cudaEventRecord(d.ev[0], s));
      CK(cudaStreamWaitEvent(cs, d.ev[0], 0));
    }
    // readFromGhost: owned values other ranks ghost -> their xr, then publish the epoch
    LAUNCHS(k_p2p_put, totalSend, cs, d_in, d.d_send_idx, totalSend, d.d_send_off, d.d_peer_xr, R);
    DKT_DIST_LAUNCH(k_p2p_signal, 1, P2P_MAX_RANKS, cs)(d.d_peer_flagR, d.d_send_off, R, epoch);
    g_launches++;
    if (overlap)
    {
      rc = run_matvec_chunked(da, op, in_local, out_local, scale, flags, 1u << 0, false);  // interior elements, on s
      if (rc) return rc;
    }
  }
  if (stages & 2u)
  {
    if (nGhost)
    {
      DKT_DIST_LAUNCH(k_p2p_gate, 1, P2P_MAX_RANKS, cs)(flagR, d.d_recv_off, R, epoch, d.d_p2p_err);
      g_launches++;
    }
    LAUNCHS(k_p2p_wait_copy, nGhost, cs, flagR, d.d_recv_off, R, epoch, xr, in_local + nOwned, nGhost, d.d_p2p_err);
    if (overlap) rc = chunked_on(da, cs, op, in_local, out_local, scale, flags, 1u << 1);  // boundary elements
    else rc = (flags & DKT_MV_FLAT) ? run_matvec(da, op, in_local, out_local, scale, flags)
                                    : run_matvec_chunked(da, op, in_local, out_local, scale, flags);
    if (rc) return rc;
    // writeToGhosts: ghost partial sums -> the owners' xw, publish
    LAUNCHS(k_p2p_put, nGhost, cs, out_local + nOwned, (const uint32_t *)nullptr, nGhost, d.d_recv_off, d.d_peer_xw, R);
    DKT_DIST_LAUNCH(k_p2p_signal, 1, P2P_MAX_RANKS, cs)(d.d_peer_flagW, d.d_recv_off, R, epoch);
    g_launches++;
  }
  if (!(stages & 4u)) return DKT_OK;
  // add what came back for the owned nodes.  With overlap this runs on the exchange stream while the interior elements may still
  // be at work: the nodes concerned are accumulated with RED everywhere (DA::d_node_sent), so the order does not matter.
  if (totalSend)
  {
    DKT_DIST_LAUNCH(k_p2p_gate, 1, P2P_MAX_RANKS, cs)(flagW, d.d_send_off, R, epoch, d.d_p2p_err);
    g_launches++;
  }
  LAUNCHS(k_p2p_wait_add, totalSend, cs, flagW, d.d_send_off, R, epoch, xw, out_local, d.d_send_idx, totalSend, d.d_p2p_err);
  if (overlap)
  {
    CK(cudaEventRecord(d.ev[3], cs));
    CK(cudaStreamWaitEvent(s, d.ev[3], 0));
  }
  if (!ghosted) CK(cudaMemcpyAsync(d_out, out_local, nOwned * sizeof(double), cudaMemcpyDeviceToDevice, s));
  if (getenv("DKT_P2P_CHECK"))
  {
    int err = 0;
    CK(cudaMemcpyAsync(&err, d.d_p2p_err, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (err) { set_error("DKT_DIST_P2P: a peer's epoch flag did not arrive within the time limit"); return DKT_ERR_NCCL; }
  }
  return DKT_OK;
}

// The ghost exchanges of ot::DA on their own (include/oda.tcc:212-435), on a DEVICE vector in the ghosted layout
// [owned | ghosts grouped by owner].  begin: the exchange is queued on the exchange stream behind the work already queued on the
// DA's stream; end: the DA's stream waits for it.  read: owners -> ghosts.  write: ghost entries -> owners, ACCUMULATED into the
// owned entries (the reference's writeToGhostsEnd adds, oda.tcc:420-431).  which: 0 read, 1 write.
int ghost_exchange_begin(DA &da, Dist &d, double *vec, int which)
{
  if (d.nranks <= 1) return DKT_OK;
  if (!d.comm) { set_error("ghost exchange: no communicator (dry-run partition)"); return DKT_ERR_INVALID; }
  cudaStream_t s = da.stream, cs = d.comm_stream;
  const uint64_t nOwned = d.nOwned, totalSend = d.send_off[d.nranks];
  CK(cudaEventRecord(d.ev[0], s));
  CK(cudaStreamWaitEvent(cs, d.ev[0], 0));
  if (which == 0) LAUNCHS(k_pack, totalSend, cs, vec, d.d_send_idx, totalSend, d.d_send_buf);
  NCK(g_nccl.GroupStart());
  for (int p = 0; p < d.nranks; p++)
  {
    if (p == d.rank) continue;
    const uint64_t sc = d.send_off[p + 1] - d.send_off[p], rcv = d.recv_off[p + 1] - d.recv_off[p];
    if (which == 0)
    {
      if (sc) NCK(g_nccl.Send(d.d_send_buf + d.send_off[p], sc, NCCL_FLOAT64, p, (ncclComm_p)d.comm, cs));
      if (rcv) NCK(g_nccl.Recv(vec + nOwned + d.recv_off[p], rcv, NCCL_FLOAT64, p, (ncclComm_p)d.comm, cs));
    }
    else
    {
      if (rcv) NCK(g_nccl.Send(vec + nOwned + d.recv_off[p], rcv, NCCL_FLOAT64, p, (ncclComm_p)d.comm, cs));
      if (sc) NCK(g_nccl.Recv(d.d_recv_buf + d.send_off[p], sc, NCCL_FLOAT64, p, (ncclComm_p)d.comm, cs));
    }
  }
  NCK(g_nccl.GroupEnd());
  g_launches++;
  if (which == 1) LAUNCHS(k_unpack_add, totalSend, cs, vec, d.d_send_idx, totalSend, d.d_recv_buf);
  CK(cudaEventRecord(d.ev[1], cs));
  return DKT_OK;
}
int ghost_exchange_end(DA &da, Dist &d)
{
  if (d.nranks <= 1) return DKT_OK;
  CK(cudaStreamWaitEvent(da.stream, d.ev[1], 0));
  return DKT_OK;
}

// v = A u on the partition: in/out are DEVICE vectors of the nOwned owned nodes
int run_matvec_dist(DA &da, Dist &d, const dkt_op *op, const double *d_in, double *d_out, double scale, unsigned flags)
{
  return run_matvec_dist_stages(da, d, op, d_in, d_out, scale, flags, 7u);
}
int run_matvec_dist_stages(DA &da, Dist &d, const dkt_op *op, const double *d_in, double *d_out, double scale, unsigned flags,
                           unsigned stages)
{
  cudaStream_t s = da.stream;
  const uint64_t nOwned = d.nOwned;
  const uint64_t totalSend = d.send_off[d.nranks], nGhost = d.recv_off[d.nranks];
  // DKT_VEC_GHOSTED: the caller's vectors already have room for the ghost segment ([owned | ghosts],
  // like the reference's ghosted vectors) and are used in place - no staging copies
  if (da.N > MAX_NPE) flags |= DKT_MV_FLAT;  // 81 nodes per element: flat kernels only (no phases)
  const bool ghosted = (flags & DKT_VEC_GHOSTED) != 0;
  double *in_local = ghosted ? const_cast<double *>(d_in) : d.d_in_local;
  double *out_local = ghosted ? d_out : d.d_out_local;
  if (!ghosted && (stages & 1u)) CK(cudaMemcpyAsync(in_local, d_in, nOwned * sizeof(double), cudaMemcpyDeviceToDevice, s));
  const bool overlap = d.nranks > 1 && da.phased && !(flags & DKT_MV_FLAT);
  if (d.p2p && d.nranks > 1) return run_matvec_dist_p2p(da, d, op, d_in, d_out, in_local, out_local, scale, flags, overlap, ghosted, stages);
  if (stages != 7u) { set_error("staged execution exists for the peer-memory exchange only"); return DKT_ERR_INVALID; }
  cudaStream_t cs = overlap ? d.comm_stream : s;  // the exchanges and the boundary elements run beside the interior elements
  int rc = DKT_OK;
  const bool tm = d.timing && overlap;
  if (tm) CK(cudaEventRecord(d.tev[0], s));
  if (overlap)
  {
    CK(cudaMemsetAsync(out_local, 0, (nOwned + nGhost) * sizeof(double), s));
    g_launches++;
    CK(cudaEventRecord(d.ev[0], s));
    CK(cudaStreamWaitEvent(cs, d.ev[0], 0));
  }
  if (d.nranks > 1)
  {
    // readFromGhost: owners -> ghosts, received straight into the ghost segments of the local vector
    LAUNCHS(k_pack, totalSend, cs, d_in, d.d_send_idx, totalSend, d.d_send_buf);
    NCK(g_nccl.GroupStart());
    for (int p = 0; p < d.nranks; p++)
    {
      if (p == d.rank) continue;
      const uint64_t sc = d.send_off[p + 1] - d.send_off[p], rcv = d.recv_off[p + 1] - d.recv_off[p];
      if (sc) NCK(g_nccl.Send(d.d_send_buf + d.send_off[p], sc, NCCL_FLOAT64, p, (ncclComm_p)d.comm, cs));
      if (rcv) NCK(g_nccl.Recv(in_local + nOwned + d.recv_off[p], rcv, NCCL_FLOAT64, p, (ncclComm_p)d.comm, cs));
    }
    NCK(g_nccl.GroupEnd());
    g_launches++;
  }
  if (!overlap)
  {
    rc = (flags & DKT_MV_FLAT) ? run_matvec(da, op, in_local, out_local, scale, flags)
                               : run_matvec_chunked(da, op, in_local, out_local, scale, flags);
    if (rc) return rc;
  }
  else
  {
    // interior elements touch no ghost node: they run on s while the ghost values arrive and the boundary elements run on cs
    if (tm) CK(cudaEventRecord(d.tev[1], cs));  // ghost values have arrived
    rc = run_matvec_chunked(da, op, in_local, out_local, scale, flags, 1u << 0, false);
    if (rc) return rc;
    if (tm) CK(cudaEventRecord(d.tev[2], s));   // interior elements done
    rc = chunked_on(da, cs, op, in_local, out_local, scale, flags, 1u << 1);
    if (rc) return rc;
    if (tm) CK(cudaEventRecord(d.tev[3], cs));  // boundary elements done
  }
  if (d.nranks > 1)
  {
    // writeToGhosts: ghost partial sums (complete after the boundary elements) -> owners, accumulated
    NCK(g_nccl.GroupStart());
    for (int p = 0; p < d.nranks; p++)
    {
      if (p == d.rank) continue;
      const uint64_t sc = d.send_off[p + 1] - d.send_off[p], rcv = d.recv_off[p + 1] - d.recv_off[p];
      if (rcv) NCK(g_nccl.Send(out_local + nOwned + d.recv_off[p], rcv, NCCL_FLOAT64, p, (ncclComm_p)d.comm, cs));
      if (sc) NCK(g_nccl.Recv(d.d_recv_buf + d.send_off[p], sc, NCCL_FLOAT64, p, (ncclComm_p)d.comm, cs));
    }
    NCK(g_nccl.GroupEnd());
    g_launches++;
    LAUNCHS(k_unpack_add, totalSend, cs, out_local, d.d_send_idx, totalSend, d.d_recv_buf);  // RED-accumulated nodes: see d_node_sent
    if (overlap)
    {
      CK(cudaEventRecord(d.ev[3], cs));
      CK(cudaStreamWaitEvent(s, d.ev[3], 0));
    }
  }
  if (!ghosted) CK(cudaMemcpyAsync(d_out, out_local, nOwned * sizeof(double), cudaMemcpyDeviceToDevice, s));
  if (tm)
  {
    CK(cudaEventRecord(d.tev[4], s));
    CK(cudaEventSynchronize(d.tev[4]));
    CK(cudaEventSynchronize(d.tev[3]));
    for (int i = 0; i < 4; i++)
    {
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, d.tev[0], d.tev[i + 1]));
      d.tsum[i] = d.tcount ? std::min(d.tsum[i], (double)ms) : (double)ms;  // minimum over the calls
    }
    d.tcount++;
  }
  return DKT_OK;
}
} // namespace dkt
