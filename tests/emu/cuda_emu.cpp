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

void yield()
{
  State &s = state();
  if (s.current < 0)
  {
    fprintf(stderr, "emu: __syncthreads() in a kernel launched without fibers\n");
    abort();
  }
  s.barriers++;
  Fiber &f = s.fibers[s.current];
  f.barriers++;
  swapcontext(&f.ctx, &s.sched);
}

static void trampoline()
{
  State &s = state();
  s.body();
  s.fibers[s.current].done = true;
  swapcontext(&s.fibers[s.current].ctx, &s.sched);
}

void run_block(unsigned nthreads, const std::function<void()> &body)
{
  State &s = state();
  if (s.fibers.size() < nthreads) s.fibers.resize(nthreads);
  s.body = body;
  for (unsigned t = 0; t < nthreads; t++)
  {
    Fiber &f = s.fibers[t];
    if (!f.stack) f.stack = (char *)malloc(STACK);
    f.done = false;
    f.barriers = 0;
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = STACK;
    f.ctx.uc_link = &s.sched;
    makecontext(&f.ctx, (void (*)())trampoline, 0);
  }
  std::vector<unsigned> order(nthreads);
  unsigned live = nthreads;
  while (live)
  {
    std::iota(order.begin(), order.end(), 0u);
    if (s.order == 1) std::reverse(order.begin(), order.end());
    else if (s.order == 2) std::shuffle(order.begin(), order.end(), s.rng);
    for (unsigned t : order)
    {
      Fiber &f = s.fibers[t];
      if (f.done) continue;
      s.current = (int)t;
      threadIdx.x = t;
      swapcontext(&s.sched, &f.ctx);
      if (f.done) live--;
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
