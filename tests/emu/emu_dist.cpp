// Several ranks of a partitioned DA in one process under the CUDA-on-CPU emulation (tests only): every rank runs the
// library's own build_da + partition_da (dry run) + build_chunks, then the harness performs the ghost read / write-back
// with the library's send and receive lists (what ncclSend/Recv or the peer-memory kernels do between GPUs) around the
// library's phased chunk matvec.  Checks ownership, local numbering, exchange lists, the [interior | boundary] element
// order and the phased chunk sets (sibling-family and per-element) on REAL partitions.
#include "dkt_internal.h"
#include "cuda_emu.h"

#include <memory>
#include <string>

namespace dkt { extern std::string g_emu_err; }
#define g_err g_emu_err
using namespace dkt;


struct Rank
{
  DA da;
  Dist dist;
  std::vector<double> in, out;
};

// u, v: vectors in the single-rank DA order (n_global entries).  info[r*8..]: nOwned, nGhost, nMv, nHang, nRegInterior,
// nHangInterior, phased, number of chunk sets.
extern "C" int emu_dist_matvec(int dim, int order, int max_depth, int sfc, const uint32_t *xyz, const uint8_t *lev, uint64_t n,
                               const double *ip0, const double *ip1, int R, int op_kind, const double *kref, double alpha, int dirichlet,
                               double scale, const double *u, double *v, uint64_t n_global, uint64_t *info, int p2p, int ghosted)
{
  std::vector<std::unique_ptr<Rank>> ranks;
  int rc = DKT_OK;
  for (int r = 0; r < R && rc == DKT_OK; r++)
  {
    ranks.emplace_back(new Rank);
    DA &da = ranks.back()->da;
    da.dim = dim; da.order = order; da.max_depth = max_depth; da.sfc_mode = sfc;
    da.M = order + 1;
    da.N = 1;
    for (int i = 0; i < dim; i++) da.N *= da.M;
    for (int i = 0; i < da.M * da.M; i++) { da.ip[0][i] = ip0[i]; da.ip[1][i] = ip1[i]; }
    rc = build_da(da, xyz, lev, n, 0);
    if (rc == DKT_OK && da.nNodes != n_global) { set_error("node count differs from the caller's"); rc = DKT_ERR_INVALID; }
    if (rc == DKT_OK) rc = partition_da(da, ranks.back()->dist, r, R, nullptr);
    if (rc == DKT_OK) rc = build_chunks(da);
  }
  dkt_op op;
  op.kind = op_kind; op.kref = kref; op.alpha = alpha; op.dirichlet = dirichlet; op.terms = 0;
  if (rc == DKT_OK && p2p)
  {
    // the library's own peer-memory flow (run_matvec_dist_p2p), stage by stage over all ranks, twice (buffer re-use)
    std::vector<Dist *> dd;
    for (auto &rk : ranks) dd.push_back(&rk->dist);
    rc = p2p_attach_local(dd.data(), R);
    std::vector<double *> din(R, nullptr), dout(R, nullptr);
    for (int r = 0; r < R && rc == DKT_OK; r++)
    {
      Rank &me = *ranks[r];
      // ghosted: the caller's vectors hold [owned | ghosts] and are used in place (DKT_VEC_GHOSTED, what bench.py does)
      const uint64_t len = me.dist.nOwned + (ghosted ? me.dist.nGhost : 0);
      cudaMalloc(&din[r], std::max<uint64_t>(len, 1) * sizeof(double));
      cudaMalloc(&dout[r], std::max<uint64_t>(len, 1) * sizeof(double));
      uint64_t *o = info + 8 * r;
      o[0] = me.dist.nOwned; o[1] = me.dist.nGhost; o[2] = me.da.nMv; o[3] = me.da.nHang; o[4] = me.da.nRegInterior; o[5] = me.da.nHangInterior;
      o[6] = me.da.phased ? 1 : 0; o[7] = me.da.sets.size();
    }
    for (int epoch = 0; epoch < 2 && rc == DKT_OK; epoch++)
    {
      for (int r = 0; r < R; r++)
        for (uint64_t j = 0; j < ranks[r]->dist.nOwned; j++) din[r][j] = (epoch == 0 ? 0.5 : 1.0) * u[ranks[r]->dist.d_owned_gid[j]];
      for (unsigned stage = 1; stage <= 4 && rc == DKT_OK; stage <<= 1)
        for (int r = 0; r < R && rc == DKT_OK; r++)
          rc = run_matvec_dist_stages(ranks[r]->da, ranks[r]->dist, &op, din[r], dout[r], scale, DKT_VEC_DEVICE | (ghosted ? DKT_VEC_GHOSTED : 0u), stage);
    }
    if (rc == DKT_OK)
    {
      for (uint64_t i = 0; i < n_global; i++) v[i] = std::nan("");
      for (int r = 0; r < R; r++)
      {
        for (uint64_t j = 0; j < ranks[r]->dist.nOwned; j++) v[ranks[r]->dist.d_owned_gid[j]] = dout[r][j];
        int err = *ranks[r]->dist.d_p2p_err;
        if (err) { set_error("a peer-memory wait timed out"); rc = DKT_ERR_NCCL; }
      }
    }
    for (int r = 0; r < R; r++) { cudaFree(din[r]); cudaFree(dout[r]); }
  }
  else if (rc == DKT_OK)
  {
    // owned values in, ghost read with the send / receive lists
    for (int r = 0; r < R; r++)
    {
      Rank &me = *ranks[r];
      me.in.assign(me.da.nNodes, std::nan(""));
      me.out.assign(me.da.nNodes, std::nan(""));
      for (uint64_t j = 0; j < me.dist.nOwned; j++) me.in[j] = u[me.dist.d_owned_gid[j]];
      uint64_t *o = info + 8 * r;
      o[0] = me.dist.nOwned; o[1] = me.dist.nGhost; o[2] = me.da.nMv; o[3] = me.da.nHang; o[4] = me.da.nRegInterior; o[5] = me.da.nHangInterior;
      o[6] = me.da.phased ? 1 : 0; o[7] = me.da.sets.size();
    }
    for (int r = 0; r < R; r++)
      for (int p = 0; p < R; p++)
      {
        if (p == r) continue;
        Rank &src = *ranks[r], &dst = *ranks[p];
        const uint64_t a = src.dist.send_off[p], b = src.dist.send_off[p + 1];
        if (b - a != dst.dist.recv_off[r + 1] - dst.dist.recv_off[r]) { set_error("send and receive counts disagree"); rc = DKT_ERR_INVALID; }
        for (uint64_t i = a; i < b && rc == DKT_OK; i++)
          dst.in[dst.dist.nOwned + dst.dist.recv_off[r] + (i - a)] = src.in[src.dist.d_send_idx[i]];
      }
  }
  for (int r = 0; r < R && rc == DKT_OK && !p2p; r++)
  {
    Rank &me = *ranks[r];
    double *din = nullptr, *dout = nullptr;
    cudaMalloc(&din, me.da.nNodes * sizeof(double));
    cudaMalloc(&dout, me.da.nNodes * sizeof(double));
    memcpy(din, me.in.data(), me.da.nNodes * sizeof(double));
    if (!me.da.phased) rc = run_matvec_chunked(me.da, &op, din, dout, scale, 0);
    else
      for (int ph = 0; ph < 3 && rc == DKT_OK; ph++) rc = run_matvec_chunked(me.da, &op, din, dout, scale, 0, 1u << ph, ph == 0);
    memcpy(me.out.data(), dout, me.da.nNodes * sizeof(double));
    cudaFree(din);
    cudaFree(dout);
  }
  if (rc == DKT_OK && !p2p)
  {
    // ghost partial sums back to the owners, accumulated; then gather the owned entries
    for (int p = 0; p < R; p++)
      for (int r = 0; r < R; r++)
      {
        if (p == r) continue;
        Rank &gh = *ranks[p], &own = *ranks[r];
        const uint64_t a = gh.dist.recv_off[r], b = gh.dist.recv_off[r + 1];
        for (uint64_t i = a; i < b; i++) own.out[own.dist.d_send_idx[own.dist.send_off[p] + (i - a)]] += gh.out[gh.dist.nOwned + i];
      }
    for (uint64_t i = 0; i < n_global; i++) v[i] = std::nan("");
    for (int r = 0; r < R; r++)
      for (uint64_t j = 0; j < ranks[r]->dist.nOwned; j++) v[ranks[r]->dist.d_owned_gid[j]] = ranks[r]->out[j];
  }
  for (auto &rk : ranks)
  {
    const std::string keep = g_err;
    free_dist(rk->dist);
    free_da(rk->da);
    g_err = keep;
  }
  return rc;
}
