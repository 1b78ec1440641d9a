// Internal declarations shared by the translation units of libdkt.so (not installed).
#ifndef DKT_INTERNAL_H
#define DKT_INTERNAL_H

#include "../../include/dkt.h"

#include <cstdint>
#include <string>
#include <vector>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#else
typedef struct CUstream_st *cudaStream_t;
typedef struct CUevent_st *cudaEvent_t;
#endif

namespace dkt
{
constexpr uint32_t INVALID = 0xFFFFFFFFu;
constexpr int MAX_NPE = 27;   // nodes per element handled by the register kernels (3-D p=2)
constexpr int MAX_M = 3;      // order + 1

struct SfcTables
{
  int dim = 0, nch = 0, nrot = 0;
  std::vector<uint8_t> rot_perm, rot_inv, htab;
};
void make_sfc_tables(int dim, int mode, SfcTables &t);

void set_error(const std::string &msg);
extern uint64_t g_launches;

// One set of element chunks for the shared-memory matvec (dkt_chunks.cu).
struct ChunkSet
{
  int rows = 1;                  // 1 regular, 2 hanging (per-element sets: slot rows = own + parent lattice)
  int phase = 0;                 // partitioned DA: 0 interior, 1 boundary (touches a ghost node)
  uint64_t elem0 = 0;            // first visited element of the set (index into d_mv_*, d_e2n)
  uint64_t hang0 = 0;            // hanging sets: first hanging-local element (index into d_pnode, d_fmask)
  int xorperm = 0;               // slot s of an element with child number c holds rank s ^ c (order 1)
  uint64_t nElem = 0;
  uint32_t nChunks = 0, elemsPerChunk = 0, maxNloc = 0, maxLen = 0, jdStride = 0;
  uint64_t totalNodes = 0;
  uint32_t *d_slot = nullptr;    // [nChunks*elemsPerChunk*rows*N] node rank | position << 16, rank-major per chunk
                                 // (family sets: [nChunks*upc*spu] node id << 2 | boundary bit, or "absent", unit-major)
  uint32_t *d_gid = nullptr;     // [totalNodes] global node id, chunk by chunk, (len desc) order inside a chunk
  uint16_t *d_meta = nullptr;    // [totalNodes] run length | boundary bit | shared bit
  uint16_t *d_jd = nullptr;      // [nChunks*jdStride] jagged-diagonal offsets
  uint64_t *d_node_off = nullptr;// [nChunks+1]
  // per-unit attributes (the DA's own arrays, or copies listed in `owned`)
  const uint8_t *lev = nullptr;    // level
  const uint8_t *child = nullptr;  // per-element sets: Morton child number
  const uint32_t *fmask = nullptr; // hanging per-element sets: filled own slots
  // sibling-family sets (kind == 2, see dkt_chunks.cu): one unit = a complete family of 2^dim leaves with its 3^dim node lattice
  int kind = 0;
  int spu = 0;                     // slots per unit
  uint16_t *d_rk16 = nullptr;      // [nChunks*upc*spu] node rank of every lattice slot (unit-major); nloc = no node (hanging point)
  uint16_t *d_inv16 = nullptr;     // [nChunks*upc*spu] shared-memory lattice address of the contribution at jagged-diagonal position i
  uint32_t *d_frec = nullptr;      // [nChunks*upc*4] per family: hanging-point masks (27 lattice points per word) x3, level
  uint32_t *d_nloc = nullptr;      // [nChunks] nodes of every chunk (d_node_off is padded to multiples of 4 for the bulk copies)
  void *d_rec = nullptr;           // [totalNodes] uint32: gid | boundary bit 30 | shared bit 31 (d_jd then holds [jd | cnt] per chunk)
  std::vector<void *> owned;       // device buffers freed with the set
};

// Everything one rank needs on the device.  All d_* pointers are device memory.
struct DA
{
  int dim = 0, order = 0, max_depth = 0, sfc_mode = 0;
  int N = 0, M = 0;
  int finest_level = 0;  // finest element level
  int lk = 0;            // finest lattice level (finest_level + 1 for order 2)
  int shift = 0, bits = 0;
  int tree_class = 0;
  int device = 0;
  int numSMs = 0;
  uint64_t nElem = 0, nMv = 0, nReg = 0, nHang = 0, nNodes = 0, nExtNodes = 0, nBdy = 0, nSplit = 0, nU = 0;

  uint32_t *d_elem_xyz = nullptr;  // [nElem*dim] AoS, tree order
  uint8_t *d_elem_lev = nullptr;
  uint32_t *d_node_xyz = nullptr;  // [nNodes*dim] DA order
  uint8_t *d_node_lev = nullptr;
  uint32_t *d_bdy = nullptr;       // [nBdy]
  uint8_t *d_node_isbdy = nullptr; // [nNodes]
  // partitioned DA: [nNodes] 1 = owned node that other ranks ghost.  Its chunk entries are accumulated with RED even when a
  // single chunk touches it, so that the returned partial sums may be added while the elements still run (dkt_dist.cu)
  uint8_t *d_node_sent = nullptr;

  // visited elements: regular ones first [0,nReg), hanging ones after [nReg,nMv)
  uint32_t *d_e2n = nullptr;       // [nMv*N]
  uint32_t *d_mv_xyz = nullptr;    // [nMv*dim]
  uint8_t *d_mv_lev = nullptr;     // [nMv]
  uint32_t *d_mv_src = nullptr;    // [nMv] position in SFC visit order (for export)
  uint64_t mv_src0 = 0;            // partitioned DA: first visit-order position of this rank (d_mv_src - mv_src0 is a permutation)
  uint32_t *d_pnode = nullptr;     // [nHang*N]
  uint8_t *d_child = nullptr;      // [nHang]
  uint8_t *d_mv_child = nullptr;   // [nMv] Morton child number of every visited element
  uint32_t *d_fmask = nullptr;     // [nHang] filled own slots of every hanging element (slot order)

  // coordinate -> node lookup: sorted unique packed keys of all lattice locations
  uint64_t *d_ukey = nullptr;      // [nU]
  uint32_t *d_unode = nullptr;     // [nU] node id or INVALID (hanging location)

  double ip[2][MAX_M * MAX_M];     // parent->child 1-D matrices, A[k*M+j]

  // chunked tables.  Order 1: {families, regular singles, hanging singles}; otherwise {regular, hanging}; partitioned: x3 phases
  std::vector<ChunkSet> sets;
  // order 1 with family sets: the per-element tables of ALL elements, built on first use by an operator the family
  // kernel does not serve (general dense K_ref, caller-supplied interpolation)
  std::vector<ChunkSet> sets_elem;
  bool families = false;           // `sets` holds family sets
  // partitioned DA: regular elements are ordered [interior | boundary], boundary = touches a ghost node
  uint64_t nRegInterior = 0, nHangInterior = 0;
  bool phased = false;
  int commSMs = 0;                 // SMs left to the NCCL kernels during the interior phases
  // DKT_MV_STREAMS=n (opt-in): the chunk sets of one matvec call run on n streams (dkt_chunks.cu launch_mv3)
  static constexpr int MAX_AUX = 3;
  int mvStreams = 1;
  cudaStream_t aux[MAX_AUX] = {nullptr, nullptr, nullptr};
  cudaStream_t cur = nullptr;      // stream of the set being launched (nullptr: `stream`)
  cudaEvent_t ev_fork = nullptr, ev_join[MAX_AUX] = {nullptr, nullptr, nullptr};

  double *d_in = nullptr, *d_out = nullptr;  // staging for host-pointer matvecs
  double *d_dof = nullptr;         // dkt_matvec_dof: component-major copies [in: dof x n | out: dof x n | host staging: n x dof]
  size_t dof_cap = 0;              // doubles allocated in d_dof
  double *d_kbuf = nullptr;        // operator matrix of the 81-node (4-D order 2) flat kernels
  cudaStream_t stream = nullptr;      // stream in use (own_stream or the caller's)
  cudaStream_t own_stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  float last_ms = 0.f;
};

// Exchange plan of one rank of a partitioned DA (dkt_dist.cu).
struct Dist
{
  bool active = false;
  int rank = 0, nranks = 1;
  void *comm = nullptr;                 // ncclComm_t
  cudaStream_t comm_stream = nullptr;   // the exchanges run here, overlapped with interior elements
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  uint64_t nOwned = 0, nGhost = 0, nGlobalNodes = 0, nGlobalElems = 0;
  std::vector<uint64_t> send_off, recv_off;  // [nranks+1] offsets into the send list / the ghost segment
  uint32_t *d_send_idx = nullptr;       // local ids of owned nodes other ranks ghost, grouped by peer
  double *d_send_buf = nullptr, *d_recv_buf = nullptr;
  double *d_in_local = nullptr, *d_out_local = nullptr;  // [owned | ghosts by owner rank]
  uint32_t *d_owned_gid = nullptr;      // global (single-rank DA order) id of each owned node
  // peer-memory exchange (opt-in DKT_DIST_P2P=1, dkt_dist.cu): kernels store straight into the peers' receive
  // buffers over NVLink and signal epoch flags; no NCCL kernel, no second stream
  bool p2p = false;
  char *xbuf = nullptr;                 // [flags 1 KiB | xr: nGhost doubles | xw: totalSend doubles], IPC-exported
  std::vector<void *> peer_base;        // the peers' xbuf mapped into this process (nullptr: not needed / self)
  double **d_peer_xr = nullptr, **d_peer_xw = nullptr;          // [nranks] where this rank's segment starts on peer p
  uint32_t **d_peer_flagR = nullptr, **d_peer_flagW = nullptr;  // [nranks] this rank's flag on peer p
  uint64_t *d_send_off = nullptr, *d_recv_off = nullptr;        // device copies [nranks+1]
  int *d_p2p_err = nullptr;             // set by a wait kernel that timed out
  uint32_t epoch = 0;
  // DKT_DIST_TIMING=1 (diagnostics): per-call device times of the overlapped schedule, printed when the DA is destroyed
  bool timing = false;
  cudaEvent_t tev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  double tsum[4] = {0, 0, 0, 0};
  int tcount = 0;
};
int partition_da(DA &g, Dist &dist, int rank, int nranks, const void *nccl_id);
int run_matvec_dist(DA &da, Dist &d, const dkt_op *op, const double *d_in, double *d_out, double scale, unsigned flags);
// the same in stages (bit mask, see run_matvec_dist_p2p): only for the single-process emulation of several ranks
int run_matvec_dist_stages(DA &da, Dist &d, const dkt_op *op, const double *d_in, double *d_out, double scale, unsigned flags,
                           unsigned stages);
int ghost_exchange_begin(DA &da, Dist &d, double *vec, int which);  // which: 0 read (owners -> ghosts), 1 write (ghosts -> owners, added)
int ghost_exchange_end(DA &da, Dist &d);
void free_dist(Dist &d);
int nccl_unique_id(void *out128);
int p2p_attach_local(Dist **ranks, int R);
// red[0], red[1]: sum over ranks; red[2]: max over ranks (device buffer of 4 doubles)
int dist_allreduce(Dist &d, double *red, cudaStream_t s);
int cg_solve(DA &da, Dist *dist, const dkt_op *op, double *d_x, const double *d_b, int max_iter, double *tol, double scale,
             unsigned flags, int *iters, int *status);

// DKT_OP_KRON -> the dense reference matrix it stands for (N x N row-major)
void kron_to_dense(const dkt_op *op, int dim, int M, std::vector<double> &K);
int build_da(DA &da, const uint32_t *elem_xyz, const uint8_t *elem_lev, uint64_t n, unsigned flags);

// A linear tree built on the device from points (dkt_tree.cu): leaves in tree order.
struct Tree
{
  int dim = 0, max_depth = 0, sfc_mode = 0, device = 0, finest_level = 0;
  uint64_t n = 0;
  uint32_t *d_xyz = nullptr;  // [n * dim] anchors
  uint8_t *d_lev = nullptr;   // [n]
};
int build_tree(Tree &t, const uint32_t *pts, uint64_t n, uint64_t maxPts, bool balance, unsigned flags);
void free_tree(Tree &t);
void free_da(DA &da);
int run_matvec(DA &da, const dkt_op *op, const double *d_in, double *d_out, double scale, unsigned flags);
int build_chunks(DA &da);
void free_chunks(DA &da);
int run_matvec_chunked(DA &da, const dkt_op *op, const double *d_in, double *d_out, double scale, unsigned flags,
                       unsigned phaseMask = 7u, bool zeroOut = true);
int device_exclusive_scan(DA &da, const uint64_t *in, uint64_t *out, uint64_t n);
} // namespace dkt

#endif
