// Host entry of the emulated chunk path (tests only): fills a dkt::DA from tables the ORACLE built
// (oracle/flat.py), runs dkt_chunks.cu's build_chunks + run_matvec_chunked under tests/emu/cuda_emu.h and
// returns the vector.  Nothing here is linked into libdkt.so.
#include "dkt_internal.h"
#include "cuda_emu.h"

#include <string>

namespace dkt { extern std::string g_emu_err; }

template <typename T>
static T *dup(const T *src, size_t n)
{
  T *p = nullptr;
  cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
  if (src && n) memcpy(p, src, n * sizeof(T));
  return p;
}


// info: per set 8 numbers {kind, rows, slots per unit, units, chunks, units per chunk, max nodes per chunk, total chunk nodes}, up to 8 sets
extern "C" int emu_matvec(int dim, int order, int max_depth, uint64_t nMv, uint64_t nReg, uint64_t nNodes, const uint32_t *e2n,
                          const uint32_t *pnode, const uint32_t *mv_xyz, const uint8_t *mv_lev, const uint32_t *mv_src,
                          const uint8_t *isbdy, const double *ip0, const double *ip1, int op_kind, const double *kref, double alpha,
                          int dirichlet, const double *in, double *out, double scale, unsigned flags, uint64_t *info, int phased,
                          uint64_t nRegInt, uint64_t nHangInt, uint64_t src0)
{
  using namespace dkt;
  DA da;
  da.dim = dim; da.order = order; da.max_depth = max_depth;
  da.M = order + 1;
  da.N = 1;
  for (int i = 0; i < dim; i++) da.N *= da.M;
  da.nMv = nMv; da.nReg = nReg; da.nHang = nMv - nReg; da.nNodes = nNodes; da.nElem = nMv;
  const int N = da.N;
  da.d_e2n = dup(e2n, nMv * N);
  da.d_pnode = dup(pnode, da.nHang * N);
  da.d_mv_xyz = dup(mv_xyz, nMv * dim);
  da.d_mv_lev = dup(mv_lev, nMv);
  da.d_mv_src = dup(mv_src, nMv);
  da.d_node_isbdy = dup(isbdy, nNodes);
  for (int i = 0; i < da.M * da.M; i++) { da.ip[0][i] = ip0[i]; da.ip[1][i] = ip1[i]; }
  da.phased = phased != 0; da.nRegInterior = nRegInt; da.nHangInterior = nHangInt; da.mv_src0 = src0;
  int rc = build_chunks(da);
  if (rc == DKT_OK)
  {
    for (int i = 0; i < 128; i++) info[i] = 0;
    size_t k = 0;
    for (const ChunkSet &cs : da.sets)
    {
      if (k >= 16) break;
      uint64_t *o = info + 8 * k++;
      o[0] = cs.kind; o[1] = cs.rows; o[2] = cs.spu; o[3] = cs.nElem; o[4] = cs.nChunks; o[5] = cs.elemsPerChunk; o[6] = cs.maxNloc;
      o[7] = cs.totalNodes | ((uint64_t)cs.phase << 56);
    }
    dkt_op op;
    op.kind = op_kind & 15; op.kref = kref; op.alpha = alpha; op.dirichlet = dirichlet; op.terms = op_kind >> 4;  // DKT_OP_KRON: terms in the high bits
    double *din = dup(in, nNodes), *dout = dup((const double *)nullptr, nNodes);
    if (!phased) rc = run_matvec_chunked(da, &op, din, dout, scale, flags);
    else  // the three phases of run_matvec_dist, without the exchanges
      for (int ph = 0; ph < 3 && rc == DKT_OK; ph++) rc = run_matvec_chunked(da, &op, din, dout, scale, flags, 1u << ph, ph == 0);
    memcpy(out, dout, nNodes * sizeof(double));
    cudaFree(din);
    cudaFree(dout);
  }
  free_chunks(da);
  cudaFree(da.d_e2n); cudaFree(da.d_pnode); cudaFree(da.d_mv_xyz); cudaFree(da.d_mv_lev); cudaFree(da.d_mv_src); cudaFree(da.d_node_isbdy);
  return rc;
}
