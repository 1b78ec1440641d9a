// Minimal CUDA-on-CPU emulation used ONLY by the CPU test-suite (tests/test_emu_chunks.py) to execute the
// table-building and matvec kernels of dendro-kt_b200/csrc/dkt_chunks.cu without a GPU.  It is test
// infrastructure: the product (libdkt.so) is never built with DKT_EMU and has no CPU path.
//
// Model: the threads of one block are fibers (ucontext) on one OS thread; __syncthreads() yields to a
// scheduler that resumes every live fiber once per barrier phase.  The order in which the fibers of a phase
// run is selectable (forward / reverse / shuffled, env DKT_EMU_ORDER=0/1/2) so that a missing barrier shows
// up as an order-dependent result.  Blocks run one after the other.  Only what dkt_chunks.cu uses exists.
#ifndef DKT_CUDA_EMU_H
#define DKT_CUDA_EMU_H

#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <random>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __constant__ static
#define __grid_constant__
#define __launch_bounds__(...)
#define __restrict__

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
struct uint2 { unsigned x, y; };
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
extern emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

typedef int cudaError_t;
#ifndef DKT_INTERNAL_H
typedef struct CUstream_st *cudaStream_t;
#endif
constexpr cudaError_t cudaSuccess = 0;
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaDevAttrMultiProcessorCount = 16 };

namespace emu
{
struct Fiber
{
  ucontext_t ctx;
  char *stack = nullptr;
  bool done = false;
  int wait = 0;                // 0 runnable, 1 at a __syncthreads(), 2 spinning (warp shuffle rendezvous)
  unsigned long barriers = 0;  // __syncthreads() executed by this thread in the current block
};
struct State
{
  ucontext_t sched;
  std::vector<Fiber> fibers;
  int current = -1;
  std::function<void()> body;
  char *dyn_smem = nullptr;
  size_t dyn_cap = 0;
  int order = 0;
  std::mt19937 rng{12345};
  uint64_t barriers = 0;
  // warp shuffles: per lane the value, the number of shuffles posted and the number read (reset per block)
  struct WarpX { uint64_t val[32]; unsigned posted[32], read[32]; };
  std::vector<WarpX> warps;
  unsigned nthreads = 0;
};
State &state();
void yield();
void spin_yield();  // give the other fibers a turn while waiting for one of them (not a barrier)
void run_block(unsigned nthreads, const std::function<void()> &body);

template <typename F>
struct Launch
{
  F f;
  unsigned grid, block;
  size_t smem;
  template <typename... A>
  void operator()(A... args)
  {
    State &s = state();
    // exactly the requested bytes + a canary behind them: a write past the dynamic shared memory aborts
    constexpr size_t CANARY = 256;
    free(s.dyn_smem);
    s.dyn_smem = (char *)aligned_alloc(128, (smem + CANARY + 127) & ~(size_t)127);
    s.dyn_cap = smem;
    gridDim.x = grid;
    blockDim.x = block;
    for (unsigned b = 0; b < grid; b++)
    {
      blockIdx.x = b;
      memset(s.dyn_smem, 0xFF, smem);  // poison (NaN as double): shared memory is not zero-initialised
      memset(s.dyn_smem + smem, 0x5A, CANARY);
      run_block(block, [&]() { f(args...); });
      for (size_t i = 0; i < CANARY; i++)
        if ((unsigned char)s.dyn_smem[smem + i] != 0x5A)
        {
          fprintf(stderr, "emu: write past the %zu bytes of dynamic shared memory (block %u)\n", smem, b);
          abort();
        }
    }
  }
};
template <typename F>
Launch<F *> make_launch(F *f, uint64_t grid, unsigned block, size_t smem) { return Launch<F *>{f, (unsigned)grid, block, smem}; }
// kernels WITHOUT barriers and shared memory (element-wise): the threads run as a plain loop, no fibers
template <typename F>
struct LaunchFlat
{
  F f;
  unsigned grid, block;
  template <typename... A>
  void operator()(A... args)
  {
    gridDim.x = grid;
    blockDim.x = block;
    state().current = -2;  // a __syncthreads() here is a bug
    for (unsigned b = 0; b < grid; b++)
    {
      blockIdx.x = b;
      for (unsigned t = 0; t < block; t++)
      {
        threadIdx.x = t;
        f(args...);
      }
    }
    state().current = -1;
  }
};
template <typename F>
LaunchFlat<F *> make_launch_flat(F *f, uint64_t grid, unsigned block) { return LaunchFlat<F *>{f, (unsigned)grid, block}; }
}  // namespace emu

#define DKT_LAUNCH(k, g, b, s, st) ::emu::make_launch(k, (g), (b), (s))
#define DKT_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(::emu::state().dyn_smem)

inline void __syncthreads() { emu::yield(); }

template <typename T> inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline long long __double_as_longlong(double v) { long long r; memcpy(&r, &v, 8); return r; }
inline double __hiloint2double(int hi, int lo) { const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double r; memcpy(&r, &b, 8); return r; }
// warp shuffle between the fibers of one warp: post, wait for the partner's post, read, wait until the partner has read
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int o)
{
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  emu::State &s = emu::state();
  const unsigned tid = threadIdx.x, lane = tid & 31u, partner = lane ^ (unsigned)o;
  if ((tid & ~31u) + partner >= s.nthreads) return v;
  emu::State::WarpX &x = s.warps[tid >> 5];
  const unsigned gen = x.posted[lane] + 1;
  x.val[lane] = 0;
  memcpy(&x.val[lane], &v, sizeof(T));
  x.posted[lane] = gen;
  while (x.posted[partner] < gen) emu::spin_yield();
  T r;
  memcpy(&r, &x.val[partner], sizeof(T));
  x.read[lane] = gen;
  while (x.read[partner] < gen) emu::spin_yield();
  return r;
}
// warp barrier: a butterfly of rendezvous (every lane has heard from every other lane after five rounds)
inline void __syncwarp(unsigned m = 0xffffffffu)
{
  int x = 0;
  for (int o = 1; o < 32; o <<= 1) x += __shfl_xor_sync(m, x, o);
}
inline void __threadfence_system() {}
inline void __nanosleep(unsigned) {}
inline long long clock64() { static long long c = 0; return c += 1000; }
template <typename T> inline T __ldcg(const T *p) { return *p; }
using std::max;
using std::min;

inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
// device allocations: [size header 256 B | payload (poisoned) | canary 256 B]; cudaFree aborts on a damaged canary
inline cudaError_t emu_malloc(void **p, size_t bytes)
{
  const size_t pay = (std::max<size_t>(bytes, 1) + 255) & ~(size_t)255;
  char *raw = (char *)aligned_alloc(256, pay + 512);
  if (!raw) return 2;
  *(size_t *)raw = bytes;
  memset(raw + 256, 0xCD, pay);  // device memory is not zero-initialised
  memset(raw + 256 + bytes, 0x5A, pay - bytes + 256);
  *p = raw + 256;
  return cudaSuccess;
}
template <typename T> inline cudaError_t cudaMalloc(T **p, size_t bytes) { return emu_malloc((void **)p, bytes); }
inline cudaError_t cudaFree(void *p)
{
  if (!p) return cudaSuccess;
  char *raw = (char *)p - 256;
  const size_t bytes = *(size_t *)raw;
  const size_t pay = (std::max<size_t>(bytes, 1) + 255) & ~(size_t)255;
  for (size_t i = bytes; i < pay + 256; i++)
    if ((unsigned char)raw[256 + i] != 0x5A)
    {
      fprintf(stderr, "emu: write past a device allocation of %zu bytes (offset %zu)\n", bytes, i);
      abort();
    }
  free(raw);
  return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, int) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaMemset(void *p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = 0; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = (cudaStream_t)(void *)0x1; return cudaSuccess; }
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
template <typename H> inline cudaError_t cudaIpcGetMemHandle(H *, void *) { return 2; }
template <typename H> inline cudaError_t cudaIpcOpenMemHandle(void **, H, unsigned) { return 2; }
inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
template <typename S> inline cudaError_t cudaMemcpyToSymbol(S &sym, const void *src, size_t n) { memcpy(&sym, src, n); return cudaSuccess; }
#ifndef DKT_INTERNAL_H
typedef struct CUevent_st *cudaEvent_t;
#endif
enum { cudaEventDisableTiming = 2, cudaStreamNonBlocking = 1 };
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (cudaEvent_t)(void *)0x1; return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t)(void *)0x1; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)(void *)0x1; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
template <typename K> inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
template <typename K> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, K, int, size_t) { *n = 2; return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int *v, int, int) { *v = 3; return cudaSuccess; }  // 3 "SMs": persistent loops iterate

// ---- the two cub block primitives dkt_chunks.cu uses (blocked arrangement, stable LSD radix semantics) ----
namespace cub
{
template <typename Key, int THREADS, int ITEMS, typename Value = char>
struct BlockRadixSort
{
  struct TempStorage
  {
    Key k[THREADS * ITEMS];
    Value v[THREADS * ITEMS];
  };
  TempStorage &t;
  explicit BlockRadixSort(TempStorage &ts) : t(ts) {}
  void sort_impl(Key *key, Value *val, int b0, int b1)
  {
    const int tid = threadIdx.x;
    for (int i = 0; i < ITEMS; i++)
    {
      t.k[tid * ITEMS + i] = key[i];
      if (val) t.v[tid * ITEMS + i] = val[i];
    }
    emu::yield();
    if (tid == 0)
    {
      std::vector<int> idx(THREADS * ITEMS);
      for (int i = 0; i < THREADS * ITEMS; i++) idx[i] = i;
      const uint64_t mask = (b1 - b0 >= 64) ? ~0ull : ((1ull << (b1 - b0)) - 1);
      std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
        return (((uint64_t)t.k[a] >> b0) & mask) < (((uint64_t)t.k[b] >> b0) & mask);
      });
      std::vector<Key> kk(THREADS * ITEMS);
      std::vector<Value> vv(THREADS * ITEMS);
      for (int i = 0; i < THREADS * ITEMS; i++) { kk[i] = t.k[idx[i]]; vv[i] = t.v[idx[i]]; }
      for (int i = 0; i < THREADS * ITEMS; i++) { t.k[i] = kk[i]; t.v[i] = vv[i]; }
    }
    emu::yield();
    for (int i = 0; i < ITEMS; i++)
    {
      key[i] = t.k[tid * ITEMS + i];
      if (val) val[i] = t.v[tid * ITEMS + i];
    }
  }
  void Sort(Key (&key)[ITEMS], Value (&val)[ITEMS], int b0 = 0, int b1 = sizeof(Key) * 8) { sort_impl(key, val, b0, b1); }
  void Sort(Key (&key)[ITEMS], int b0 = 0, int b1 = sizeof(Key) * 8) { sort_impl(key, nullptr, b0, b1); }
};
// device-wide primitives used by dkt_build.cu (two-phase calling convention: a null temp pointer only asks for its size)
struct DeviceRadixSort
{
  template <typename K, typename V>
  static cudaError_t SortPairs(void *tmp, size_t &bytes, const K *kin, K *kout, const V *vin, V *vout, int64_t n, int b0, int b1, cudaStream_t)
  {
    if (!tmp) { bytes = 16; return cudaSuccess; }
    std::vector<int64_t> idx(n);
    for (int64_t i = 0; i < n; i++) idx[i] = i;
    const int w = b1 - b0;
    const uint64_t mask = w >= 64 ? ~0ull : ((1ull << w) - 1);
    std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return (((uint64_t)kin[a] >> b0) & mask) < (((uint64_t)kin[b] >> b0) & mask); });
    std::vector<K> kk(n);
    std::vector<V> vv(n);
    for (int64_t i = 0; i < n; i++) { kk[i] = kin[idx[i]]; vv[i] = vin[idx[i]]; }
    for (int64_t i = 0; i < n; i++) { kout[i] = kk[i]; vout[i] = vv[i]; }
    return cudaSuccess;
  }
  template <typename K>
  static cudaError_t SortKeys(void *tmp, size_t &bytes, const K *kin, K *kout, int64_t n, int b0, int b1, cudaStream_t)
  {
    if (!tmp) { bytes = 16; return cudaSuccess; }
    const int w = b1 - b0;
    const uint64_t mask = w >= 64 ? ~0ull : ((1ull << w) - 1);
    std::vector<K> kk(kin, kin + n);
    std::stable_sort(kk.begin(), kk.end(), [&](K a, K b) { return (((uint64_t)a >> b0) & mask) < (((uint64_t)b >> b0) & mask); });
    for (int64_t i = 0; i < n; i++) kout[i] = kk[i];
    return cudaSuccess;
  }
};
struct DeviceScan
{
  template <typename I, typename O>
  static cudaError_t ExclusiveSum(void *tmp, size_t &bytes, const I *in, O *out, int64_t n, cudaStream_t)
  {
    if (!tmp) { bytes = 16; return cudaSuccess; }
    O acc = 0;
    for (int64_t i = 0; i < n; i++) { const O v = (O)in[i]; out[i] = acc; acc += v; }
    return cudaSuccess;
  }
};
struct DeviceRunLengthEncode
{
  template <typename K, typename C, typename N>
  static cudaError_t Encode(void *tmp, size_t &bytes, const K *in, K *uniq, C *counts, N *nruns, int64_t n, cudaStream_t)
  {
    if (!tmp) { bytes = 16; return cudaSuccess; }
    int64_t r = 0;
    for (int64_t i = 0; i < n;)
    {
      int64_t j = i;
      while (j < n && in[j] == in[i]) j++;
      uniq[r] = in[i];
      counts[r] = (C)(j - i);
      r++;
      i = j;
    }
    *nruns = (N)r;
    return cudaSuccess;
  }
};
template <typename T, int THREADS>
struct BlockScan
{
  struct TempStorage
  {
    T v[THREADS + 1];
  };
  TempStorage &t;
  explicit BlockScan(TempStorage &ts) : t(ts) {}
  void ExclusiveSum(T in, T &out, T &total)
  {
    const int tid = threadIdx.x;
    t.v[tid] = in;
    emu::yield();
    T acc = 0;
    for (int i = 0; i < tid; i++) acc += t.v[i];
    T tot = 0;
    for (int i = 0; i < THREADS; i++) tot += t.v[i];
    emu::yield();
    out = acc;
    total = tot;
  }
};
}  // namespace cub

#endif
