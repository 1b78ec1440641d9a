// Fiber scheduler of the CUDA-on-CPU emulation (see cuda_emu.h).  Test infrastructure only.
#include "cuda_emu.h"

#include <numeric>

emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace emu
{
static constexpr size_t STACK = 512 * 1024;

State &state()
{
  static State s;
  static bool init = false;
  if (!init)
  {
    init = true;
    if (const char *e = getenv("DKT_EMU_ORDER")) s.order = atoi(e);
  }
  return s;
}

static void leave(int why)
{
  State &s = state();
  if (s.current < 0)
  {
    fprintf(stderr, "emu: __syncthreads() or a warp shuffle in a kernel launched without fibers\n");
    abort();
  }
  Fiber &f = s.fibers[s.current];
  f.wait = why;
  swapcontext(&f.ctx, &s.sched);
}
void yield()
{
  State &s = state();
  s.barriers++;
  if (s.current >= 0) s.fibers[s.current].barriers++;
  leave(1);
}
void spin_yield() { leave(2); }

static void trampoline()
{
  State &s = state();
  s.body();
  s.fibers[s.current].done = true;
  swapcontext(&s.fibers[s.current].ctx, &s.sched);
}

// Fibers run until they reach a barrier, spin or finish.  Fibers at a barrier stay there until every live fiber of the
// block is at a barrier; spinning fibers get another turn every sweep.
void run_block(unsigned nthreads, const std::function<void()> &body)
{
  State &s = state();
  if (s.fibers.size() < nthreads) s.fibers.resize(nthreads);
  s.body = body;
  s.nthreads = nthreads;
  s.warps.assign((nthreads + 31) / 32, State::WarpX());
  for (auto &w : s.warps) memset(&w, 0, sizeof(w));
  for (unsigned t = 0; t < nthreads; t++)
  {
    Fiber &f = s.fibers[t];
    if (!f.stack) f.stack = (char *)malloc(STACK);
    f.done = false;
    f.wait = 0;
    f.barriers = 0;
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = STACK;
    f.ctx.uc_link = &s.sched;
    makecontext(&f.ctx, (void (*)())trampoline, 0);
  }
  std::vector<unsigned> order(nthreads);
  unsigned live = nthreads;
  unsigned long idle_sweeps = 0;
  while (live)
  {
    std::iota(order.begin(), order.end(), 0u);
    if (s.order == 1) std::reverse(order.begin(), order.end());
    else if (s.order == 2) std::shuffle(order.begin(), order.end(), s.rng);
    bool progress = false;
    for (unsigned t : order)
    {
      Fiber &f = s.fibers[t];
      if (f.done || f.wait == 1) continue;
      const int before = f.wait;
      s.current = (int)t;
      threadIdx.x = t;
      swapcontext(&s.sched, &f.ctx);
      if (f.done) live--;
      if (f.done || f.wait != 2 || before != 2) progress = true;
    }
    // release the barrier once every live fiber has arrived
    unsigned at = 0;
    for (unsigned t = 0; t < nthreads; t++) at += (!s.fibers[t].done && s.fibers[t].wait == 1);
    if (live && at == live)
    {
      for (unsigned t = 0; t < nthreads; t++) s.fibers[t].wait = 0;
      progress = true;
    }
    idle_sweeps = progress ? 0 : idle_sweeps + 1;
    if (idle_sweeps > 100000)
    {
      fprintf(stderr, "emu: block %u makes no progress (a warp shuffle whose partner never arrives, or threads stuck at different barriers)\n",
              blockIdx.x);
      abort();
    }
  }
  s.current = -1;
  // every thread of a block must pass the same barriers (a divergent __syncthreads() hangs a real GPU)
  for (unsigned t = 1; t < nthreads; t++)
    if (s.fibers[t].barriers != s.fibers[0].barriers)
    {
      fprintf(stderr, "emu: threads 0 and %u of block %u executed %lu and %lu barriers\n", t, blockIdx.x, s.fibers[0].barriers,
              s.fibers[t].barriers);
      abort();
    }
}
}  // namespace emu
