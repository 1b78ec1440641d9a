// Peer-memory ghost exchange (DKT_DIST_P2P=1): the kernels and the pointer tables, shared by dkt_dist.cu and by the CPU
// emulation of the protocol (tests/emu/emu_p2p.cpp, -DDKT_EMU).  See DESIGN.md 5.1.
#ifndef DKT_P2P_CUH
#define DKT_P2P_CUH

#include <cstdint>
#include <string>
#include <vector>

#ifdef DKT_EMU
#include "cuda_emu.h"
struct cudaIpcMemHandle_t { char reserved[64]; };
#endif

namespace dkt
{
// Every rank owns one IPC-exported buffer  [flagR[64] flagW[64] .. 1 KiB | xr: ghost values, by owner | xw: partial
// sums coming back, by ghosting rank]  and maps the peers' buffers.  A "put" kernel gathers and stores straight into
// the peers' receive regions over NVLink (pack + send in one kernel), a one-block kernel then publishes this matvec's
// epoch in the peers' flag words, and the consumer waits for the epoch right before it needs the data - by then it
// has normally arrived behind the interior elements.  One stream, no NCCL kernel competing for SMs.
// Re-use is safe without double buffering: a rank's put of matvec e+1 into a peer follows its wait for that peer's
// write-back flag of matvec e, which the peer raised after it had consumed the data of matvec e.
constexpr size_t P2P_FLAG_BYTES = 1024;
constexpr int P2P_MAX_RANKS = 64;
static __global__ void k_p2p_put(const double *src, const uint32_t *idx, uint64_t n, const uint64_t *seg_off, double *const *peer_dst, int nranks)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int p = 0;
  while (p + 1 < nranks && i >= seg_off[p + 1]) p++;  // segment of peer p: [seg_off[p], seg_off[p+1])
  peer_dst[p][i - seg_off[p]] = idx ? src[idx[i]] : src[i];
  __threadfence_system();
}
static __global__ void k_p2p_signal(uint32_t *const *peer_flag, const uint64_t *seg_off, int nranks, uint32_t epoch)
{
  const int p = threadIdx.x;
  if (p >= nranks || seg_off[p + 1] == seg_off[p]) return;  // nothing was sent to p
  __threadfence_system();
  *(volatile uint32_t *)peer_flag[p] = epoch;
}
// every block waits for the epoch of all peers that send to this rank.  A peer that stays away for ~30 s (6e10 cycles) is an
// error: *err = 1 and the kernel traps, so the failure surfaces at the host's next synchronisation instead of stale data
// flowing into the result.
#ifdef DKT_EMU
#define DKT_P2P_TRAP() abort()
#else
#define DKT_P2P_TRAP() __trap()
#endif
__device__ __forceinline__ void p2p_wait(const volatile uint32_t *flags, const uint64_t *seg_off, int nranks, uint32_t epoch, int *err)
{
  const int p = threadIdx.x;
  if (p < nranks && seg_off[p + 1] != seg_off[p])
  {
    const long long t0 = clock64();
    while ((int32_t)(flags[p] - epoch) < 0)
    {
      __nanosleep(100);
      if (clock64() - t0 > 60000000000ll)
      {
        atomicExch(err, 1);
        __threadfence_system();
        DKT_P2P_TRAP();
      }
    }
    __threadfence_system();
  }
  __syncthreads();
}
// ONE block that waits: launched in front of the kernels below, so that their many blocks find the flags raised and never sit
// spinning on the SMs beside the element kernels
static __global__ void k_p2p_gate(const volatile uint32_t *flags, const uint64_t *seg_off, int nranks, uint32_t epoch, int *err)
{
  p2p_wait(flags, seg_off, nranks, epoch, err);
}
static __global__ void k_p2p_wait_copy(const volatile uint32_t *flags, const uint64_t *seg_off, int nranks, uint32_t epoch, const double *x,
                                double *dst, uint64_t n, int *err)
{
  p2p_wait(flags, seg_off, nranks, epoch, err);
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __ldcg(x + i);
}
// several peers may return contributions to the same owned node -> atomic
static __global__ void k_p2p_wait_add(const volatile uint32_t *flags, const uint64_t *seg_off, int nranks, uint32_t epoch, const double *x,
                               double *v, const uint32_t *idx, uint64_t n, int *err)
{
  p2p_wait(flags, seg_off, nranks, epoch, err);
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(v + idx[i], __ldcg(x + i));
}


struct P2PInfo
{
  cudaIpcMemHandle_t handle;
  uint64_t nGhost, totalSend;
  uint64_t recv_off[P2P_MAX_RANKS + 1], send_off[P2P_MAX_RANKS + 1];
};

// Where rank `me` stores into its peers: xr[p] / xw[p] = start of this rank's segment of peer p's ghost values / of the
// partial sums of p's owned nodes, fr[p] / fw[p] = this rank's epoch flags on p.  base[p]: rank p's exchange buffer as
// seen from this process, info[p]: its segment offsets.  Both sides of every list must agree.
inline bool p2p_tables(int me, int R, const uint64_t *send_off, const uint64_t *recv_off, const P2PInfo *info, void *const *base,
                       std::vector<double *> &xr, std::vector<double *> &xw, std::vector<uint32_t *> &fr, std::vector<uint32_t *> &fw,
                       std::string &why)
{
  xr.assign(R, nullptr); xw.assign(R, nullptr); fr.assign(R, nullptr); fw.assign(R, nullptr);
  for (int p = 0; p < R; p++)
  {
    if (p == me) continue;
    const uint64_t sc = send_off[p + 1] - send_off[p], rcv = recv_off[p + 1] - recv_off[p];
    if (info[p].recv_off[me + 1] - info[p].recv_off[me] != sc || info[p].send_off[me + 1] - info[p].send_off[me] != rcv)
    {
      why = "send/receive lists of ranks " + std::to_string(me) + " and " + std::to_string(p) + " disagree";
      return false;
    }
    if (!sc && !rcv) continue;
    if (!base[p]) { why = "no mapping of the exchange buffer of rank " + std::to_string(p); return false; }
    double *x = (double *)((char *)base[p] + P2P_FLAG_BYTES);
    xr[p] = x + info[p].recv_off[me];                    // peer p's ghost values owned by me
    xw[p] = x + info[p].nGhost + info[p].send_off[me];   // partial sums of p's owned nodes that I ghost
    fr[p] = (uint32_t *)base[p] + me;
    fw[p] = (uint32_t *)base[p] + P2P_MAX_RANKS + me;
  }
  return true;
}
}  // namespace dkt
#endif
