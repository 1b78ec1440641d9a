// Emulated run of the peer-memory ghost exchange (dendro-kt_b200/csrc/dkt_p2p.cuh) with all ranks in one process:
// pointer tables (p2p_tables) + put / signal / wait kernels, in the order the protocol guarantees (every rank puts and
// signals before anybody waits).  Tests only.
#include "cuda_emu.h"
#include "dkt_p2p.cuh"

#include <string>

static std::string g_perr;
extern "C" const char *emu_p2p_error() { return g_perr.c_str(); }

// Flattened per-rank inputs: off[r*(R+1) + p] offsets; send_idx / in_local / out_local concatenated with *_base[r] starts.
// Runs `epochs` exchanges; in_local's ghost segments and out_local's owned entries are updated in place.
extern "C" int emu_p2p_exchange(int R, const uint64_t *send_off, const uint64_t *recv_off, const uint64_t *nOwned, const uint32_t *send_idx,
                                const uint64_t *sidx_base, double *in_local, double *out_local, const uint64_t *vec_base, int epochs)
{
  using namespace dkt;
  std::vector<P2PInfo> info(R);
  std::vector<void *> base(R, nullptr);
  for (int r = 0; r < R; r++)
  {
    memset(&info[r], 0, sizeof(P2PInfo));
    info[r].nGhost = recv_off[r * (R + 1) + R];
    info[r].totalSend = send_off[r * (R + 1) + R];
    for (int p = 0; p <= R; p++) { info[r].recv_off[p] = recv_off[r * (R + 1) + p]; info[r].send_off[p] = send_off[r * (R + 1) + p]; }
    char *x = nullptr;
    cudaMalloc(&x, P2P_FLAG_BYTES + (info[r].nGhost + info[r].totalSend + 1) * sizeof(double));
    memset(x, 0, P2P_FLAG_BYTES);
    base[r] = x;
  }
  struct Tab { std::vector<double *> xr, xw; std::vector<uint32_t *> fr, fw; };
  std::vector<Tab> tab(R);
  for (int r = 0; r < R; r++)
    if (!p2p_tables(r, R, send_off + r * (R + 1), recv_off + r * (R + 1), info.data(), base.data(), tab[r].xr, tab[r].xw, tab[r].fr, tab[r].fw, g_perr))
      return 1;
  int err = 0;
  auto nblk = [](uint64_t n) { return (unsigned)((n + 255) / 256); };
  for (int e = 1; e <= epochs; e++)
  {
    // readFromGhost: every rank puts + signals, then every rank waits and copies
    for (int r = 0; r < R; r++)
    {
      const uint64_t *so = send_off + r * (R + 1);
      const uint64_t ts = so[R];
      if (ts) DKT_LAUNCH(k_p2p_put, nblk(ts), 256, 0, 0)(in_local + vec_base[r], send_idx + sidx_base[r], ts, so, tab[r].xr.data(), R);
      DKT_LAUNCH(k_p2p_signal, 1, P2P_MAX_RANKS, 0, 0)(tab[r].fr.data(), so, R, (uint32_t)e);
    }
    for (int r = 0; r < R; r++)
    {
      const uint64_t *ro = recv_off + r * (R + 1);
      const uint64_t ng = ro[R];
      const volatile uint32_t *flagR = (const volatile uint32_t *)base[r];
      double *xr = (double *)((char *)base[r] + P2P_FLAG_BYTES);
      if (ng) DKT_LAUNCH(k_p2p_wait_copy, nblk(ng), 256, 0, 0)(flagR, ro, R, (uint32_t)e, xr, in_local + vec_base[r] + nOwned[r], ng, &err);
    }
    // writeToGhosts: ghost partial sums back to the owners, accumulated
    for (int r = 0; r < R; r++)
    {
      const uint64_t *ro = recv_off + r * (R + 1);
      const uint64_t ng = ro[R];
      if (ng) DKT_LAUNCH(k_p2p_put, nblk(ng), 256, 0, 0)(out_local + vec_base[r] + nOwned[r], (const uint32_t *)nullptr, ng, ro, tab[r].xw.data(), R);
      DKT_LAUNCH(k_p2p_signal, 1, P2P_MAX_RANKS, 0, 0)(tab[r].fw.data(), ro, R, (uint32_t)e);
    }
    for (int r = 0; r < R; r++)
    {
      const uint64_t *so = send_off + r * (R + 1);
      const uint64_t ts = so[R];
      const volatile uint32_t *flagW = (const volatile uint32_t *)base[r] + P2P_MAX_RANKS;
      double *xw = (double *)((char *)base[r] + P2P_FLAG_BYTES) + info[r].nGhost;
      if (ts) DKT_LAUNCH(k_p2p_wait_add, nblk(ts), 256, 0, 0)(flagW, so, R, (uint32_t)e, xw, out_local + vec_base[r], send_idx + sidx_base[r], ts, &err);
    }
  }
  for (int r = 0; r < R; r++) cudaFree(base[r]);
  if (err) { g_perr = "a wait timed out"; return 2; }
  return 0;
}
