// Multi-process MPI stand-in (see mpi.h in this directory).  TEST INFRASTRUCTURE ONLY.
//
// Ranks are fork()ed children of one parent.  dktmp_world_create() maps an anonymous shared arena before the fork:
//   [World header | per-destination mailboxes | bump-allocated messages]
// A send copies its payload into a fresh message and appends it to the destination's mailbox (eager, never blocks; the
// arena is never recycled - test-sized runs only).  A receive scans its own mailbox in arrival order for the first message
// that matches (source, tag, communicator context): MPI's non-overtaking rule.  Collectives are sequences of such messages
// with reserved negative tags; reductions are combined in rank order at rank 0 of the communicator, so results do not
// depend on timing.  A wait that sees nothing for DKTMP_TIMEOUT seconds aborts every rank.
#include "mpi.h"

#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <vector>

namespace
{
constexpr int MAXR = 64;
struct Msg
{
  int src, tag;
  uint64_t ctx, bytes, next;
};
struct Box
{
  std::atomic<int> lock;
  uint64_t head, tail;
  char pad[40];
};
struct World
{
  int nranks;
  std::atomic<int> aborted;
  std::atomic<uint64_t> top;
  uint64_t cap;
  Box box[MAXR];
};
World *W = nullptr;
char *ARENA = nullptr;
int ME = 0;
long g_sent = 0;

struct Comm
{
  uint64_t ctx = 0;
  std::vector<int> ranks;  // world ranks
  int me = -1;             // my rank in the communicator
  uint64_t nsplit = 0;
};
std::vector<Comm> comms;                // handle = index
std::vector<std::vector<int>> groups;   // handle = index (world ranks)
struct Req
{
  bool active = false, recv = false;
  void *buf = nullptr;
  uint64_t bytes = 0, ctx = 0;
  int src = 0, tag = 0;  // src: world rank or MPI_ANY_SOURCE
  const Comm *comm = nullptr;
};
std::vector<Req> reqs;
std::vector<MPI_User_function *> userops;

double now()
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
[[noreturn]] void die(const char *what)
{
  fprintf(stderr, "[oracle mpi_mp] rank %d: %s\n", ME, what);
  if (W) W->aborted.store(1);
  _exit(86);
}
inline int tsize(MPI_Datatype t) { return t & 0xFFFFFF; }
inline int tkind(MPI_Datatype t) { return (t >> 24) & 0xFF; }
uint64_t mix(uint64_t a, uint64_t b)
{
  uint64_t x = a * 0x9E3779B97F4A7C15ull ^ (b + 0x7F4A7C15ull + (a << 6) + (a >> 2));
  x ^= x >> 31; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 29;
  return x | 1ull << 63;
}
void lock(Box &b)
{
  int z = 0;
  while (!b.lock.compare_exchange_weak(z, 1, std::memory_order_acquire)) { z = 0; sched_yield(); }
}
void unlock(Box &b) { b.lock.store(0, std::memory_order_release); }

Comm &C(MPI_Comm c)
{
  if (c <= 0 || c >= (int)comms.size() || comms[c].me < 0) die("invalid communicator");
  return comms[c];
}

void post(int dst_world, int tag, uint64_t ctx, const void *buf, uint64_t bytes)
{
  const uint64_t need = (sizeof(Msg) + bytes + 15) & ~15ull;
  const uint64_t off = W->top.fetch_add(need);
  if (off + need > W->cap) die("message arena exhausted (raise the arena size of dktmp_world_create)");
  Msg *m = (Msg *)(ARENA + off);
  m->src = ME; m->tag = tag; m->ctx = ctx; m->bytes = bytes; m->next = 0;
  if (bytes) memcpy(m + 1, buf, bytes);
  g_sent += (long)bytes;
  Box &b = W->box[dst_world];
  lock(b);
  if (b.tail) ((Msg *)(ARENA + b.tail))->next = off;
  else b.head = off;
  b.tail = off;
  unlock(b);
}
// blocking receive; src_world may be MPI_ANY_SOURCE, tag may be MPI_ANY_TAG
void take(int src_world, int tag, uint64_t ctx, void *buf, uint64_t cap_bytes, MPI_Status *st, const Comm *cm)
{
  Box &b = W->box[ME];
  const double t0 = now();
  static double timeout = getenv("DKTMP_TIMEOUT") ? atof(getenv("DKTMP_TIMEOUT")) : 120.0;
  for (unsigned spin = 0;; spin++)
  {
    lock(b);
    uint64_t prev = 0, cur = b.head;
    while (cur)
    {
      Msg *m = (Msg *)(ARENA + cur);
      if (m->ctx == ctx && (src_world == MPI_ANY_SOURCE || m->src == src_world) && (tag == MPI_ANY_TAG || m->tag == tag))
      {
        if (prev) ((Msg *)(ARENA + prev))->next = m->next;
        else b.head = m->next;
        if (b.tail == cur) b.tail = prev;
        unlock(b);
        if (m->bytes > cap_bytes) die("message longer than the receive buffer");
        if (m->bytes) memcpy(buf, m + 1, m->bytes);
        if (st)
        {
          int r = m->src;
          if (cm)
            for (size_t i = 0; i < cm->ranks.size(); i++)
              if (cm->ranks[i] == m->src) r = (int)i;
          st->MPI_SOURCE = r; st->MPI_TAG = m->tag; st->MPI_ERROR = 0; st->dkt_bytes = (long)m->bytes;
        }
        return;
      }
      prev = cur;
      cur = m->next;
    }
    unlock(b);
    if (W->aborted.load()) _exit(87);
    if ((spin & 255) == 255 && now() - t0 > timeout) die("timed out waiting for a message (deadlock?)");
    sched_yield();
  }
}
// collective helpers on communicator ranks, reserved tags
enum { T_BAR = -1001, T_BCAST = -1002, T_GATHER = -1003, T_RED = -1004, T_SCAN = -1005, T_A2A = -1006, T_SPLIT = -1007 };
void csend(const Comm &c, int dst, int tag, const void *buf, uint64_t bytes) { post(c.ranks[dst], tag, c.ctx, buf, bytes); }
void crecv(const Comm &c, int src, int tag, void *buf, uint64_t bytes) { take(c.ranks[src], tag, c.ctx, buf, bytes, nullptr, &c); }

template <typename T>
void combine_t(const T *in, T *io, int n, MPI_Op op)
{
  for (int i = 0; i < n; i++)
    switch (op)
    {
    case MPI_SUM: io[i] = (T)(in[i] + io[i]); break;
    case MPI_PROD: io[i] = (T)(in[i] * io[i]); break;
    case MPI_MIN: io[i] = std::min(in[i], io[i]); break;
    case MPI_MAX: io[i] = std::max(in[i], io[i]); break;
    case MPI_LAND: io[i] = (T)(in[i] && io[i]); break;
    case MPI_LOR: io[i] = (T)(in[i] || io[i]); break;
    default: die("unsupported reduction");
    }
}
void combine(const void *in, void *io, int n, MPI_Datatype t, MPI_Op op)
{
  if (op >= 100)
  {
    if (op - 100 >= (int)userops.size()) die("invalid user op");
    userops[op - 100]((void *)in, io, &n, &t);
    return;
  }
  const int k = tkind(t), s = tsize(t);
  if (k == 1 && s == 1) combine_t((const signed char *)in, (signed char *)io, n, op);
  else if (k == 1 && s == 2) combine_t((const short *)in, (short *)io, n, op);
  else if (k == 1 && s == 4) combine_t((const int *)in, (int *)io, n, op);
  else if (k == 1 && s == 8) combine_t((const long long *)in, (long long *)io, n, op);
  else if (k == 2 && s == 1) combine_t((const unsigned char *)in, (unsigned char *)io, n, op);
  else if (k == 2 && s == 2) combine_t((const unsigned short *)in, (unsigned short *)io, n, op);
  else if (k == 2 && s == 4) combine_t((const unsigned *)in, (unsigned *)io, n, op);
  else if (k == 2 && s == 8) combine_t((const unsigned long long *)in, (unsigned long long *)io, n, op);
  else if (k == 3 && s == 4) combine_t((const float *)in, (float *)io, n, op);
  else if (k == 3 && s == 8) combine_t((const double *)in, (double *)io, n, op);
  else if (k == 3 && s == 16) combine_t((const long double *)in, (long double *)io, n, op);
  else die("reduction on an opaque datatype without a user function");
}
int new_comm(uint64_t ctx, const std::vector<int> &ranks)
{
  Comm c;
  c.ctx = ctx;
  c.ranks = ranks;
  for (size_t i = 0; i < ranks.size(); i++)
    if (ranks[i] == ME) c.me = (int)i;
  comms.push_back(c);
  return (int)comms.size() - 1;
}
}  // namespace

extern "C"
{
int dktmp_world_create(int nranks, size_t arena_bytes)
{
  if (nranks < 1 || nranks > MAXR) return 1;
  const size_t total = sizeof(World) + arena_bytes;
  void *p = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (p == MAP_FAILED) return 2;
  W = new (p) World;
  W->nranks = nranks;
  W->aborted.store(0);
  W->cap = total;
  W->top.store((sizeof(World) + 63) & ~63ull);
  for (int i = 0; i < MAXR; i++) { W->box[i].lock.store(0); W->box[i].head = W->box[i].tail = 0; }
  ARENA = (char *)p;
  return 0;
}
int dktmp_set_rank(int rank)
{
  if (!W || rank < 0 || rank >= W->nranks) return 1;
  ME = rank;
  comms.clear(); groups.clear(); reqs.clear();
  comms.resize(3);
  std::vector<int> all(W->nranks);
  for (int i = 0; i < W->nranks; i++) all[i] = i;
  comms[MPI_COMM_WORLD].ctx = mix(1, 1); comms[MPI_COMM_WORLD].ranks = all; comms[MPI_COMM_WORLD].me = rank;
  comms[MPI_COMM_SELF].ctx = mix(2, (uint64_t)rank); comms[MPI_COMM_SELF].ranks = {rank}; comms[MPI_COMM_SELF].me = 0;
  groups.resize(1);
  reqs.resize(1);
  return 0;
}
long dktmp_bytes_sent(void) { return g_sent; }

int MPI_Init(int *, char ***) { return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm, int code)
{
  if (W) W->aborted.store(1);
  _exit(code ? code : 88);
}
double MPI_Wtime(void) { return now(); }
int MPI_Comm_rank(MPI_Comm c, int *r) { *r = C(c).me; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm c, int *s) { *s = (int)C(c).ranks.size(); return MPI_SUCCESS; }

int MPI_Barrier(MPI_Comm cc)
{
  const Comm &c = C(cc);
  const int n = (int)c.ranks.size();
  if (c.me == 0)
  {
    for (int r = 1; r < n; r++) crecv(c, r, T_BAR, nullptr, 0);
    for (int r = 1; r < n; r++) csend(c, r, T_BAR, nullptr, 0);
  }
  else
  {
    csend(c, 0, T_BAR, nullptr, 0);
    crecv(c, 0, T_BAR, nullptr, 0);
  }
  return MPI_SUCCESS;
}
int MPI_Bcast(void *buf, int n, MPI_Datatype t, int root, MPI_Comm cc)
{
  const Comm &c = C(cc);
  const uint64_t bytes = (uint64_t)n * tsize(t);
  if (c.me == root)
  {
    for (int r = 0; r < (int)c.ranks.size(); r++)
      if (r != root) csend(c, r, T_BCAST, buf, bytes);
  }
  else
    crecv(c, root, T_BCAST, buf, bytes);
  return MPI_SUCCESS;
}
int MPI_Gather(const void *s, int n, MPI_Datatype t, void *r, int rn, MPI_Datatype rt, int root, MPI_Comm cc)
{
  const Comm &c = C(cc);
  const uint64_t bytes = (uint64_t)n * tsize(t), rbytes = (uint64_t)rn * tsize(rt);
  if (c.me != root) { csend(c, root, T_GATHER, s, bytes); return MPI_SUCCESS; }
  for (int q = 0; q < (int)c.ranks.size(); q++)
  {
    char *dst = (char *)r + (uint64_t)q * rbytes;
    if (q == root) { if (s != MPI_IN_PLACE) memmove(dst, s, bytes); }
    else crecv(c, q, T_GATHER, dst, rbytes);
  }
  return MPI_SUCCESS;
}
int MPI_Allgather(const void *s, int n, MPI_Datatype t, void *r, int rn, MPI_Datatype rt, MPI_Comm cc)
{
  const Comm &c = C(cc);
  const uint64_t rbytes = (uint64_t)rn * tsize(rt);
  const void *src = s == MPI_IN_PLACE ? (const char *)r + (uint64_t)c.me * rbytes : s;
  MPI_Gather(src, n, t, r, rn, rt, 0, cc);
  return MPI_Bcast(r, (int)(rbytes * c.ranks.size()), MPI_BYTE, 0, cc);
}
int MPI_Allgatherv(const void *s, int n, MPI_Datatype t, void *r, const int *cnt, const int *dsp, MPI_Datatype rt, MPI_Comm cc)
{
  const Comm &c = C(cc);
  const int P = (int)c.ranks.size(), es = tsize(rt);
  // everybody sends its piece to everybody (eager), then collects in rank order
  for (int q = 0; q < P; q++)
    if (q != c.me) csend(c, q, T_GATHER, s, (uint64_t)n * tsize(t));
  for (int q = 0; q < P; q++)
  {
    char *dst = (char *)r + (uint64_t)dsp[q] * es;
    if (q == c.me) { if (s != MPI_IN_PLACE) memmove(dst, s, (uint64_t)n * tsize(t)); }
    else crecv(c, q, T_GATHER, dst, (uint64_t)cnt[q] * es);
  }
  return MPI_SUCCESS;
}
int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm cc)
{
  const Comm &c = C(cc);
  const uint64_t bytes = (uint64_t)n * tsize(t);
  const int P = (int)c.ranks.size();
  if (c.me != root) { csend(c, root, T_RED, s == MPI_IN_PLACE ? r : s, bytes); return MPI_SUCCESS; }
  // combine in rank order: acc = contribution of rank 0, then (in = rank q, inout = acc)
  std::vector<char> mine(bytes), tmp(bytes), acc(bytes);
  memcpy(mine.data(), s == MPI_IN_PLACE ? r : s, bytes);
  for (int q = 0; q < P; q++)
  {
    const char *contrib;
    if (q == root) contrib = mine.data();
    else { crecv(c, q, T_RED, tmp.data(), bytes); contrib = tmp.data(); }
    if (q == 0) memcpy(acc.data(), contrib, bytes);
    else
    {
      // MPI: inout = in op inout with `in` the earlier ranks' value; keep rank order for non-commutative user functions
      std::vector<char> io(contrib, contrib + bytes);
      combine(acc.data(), io.data(), n, t, op);
      acc.swap(io);
    }
  }
  memcpy(r, acc.data(), bytes);
  return MPI_SUCCESS;
}
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm cc)
{
  MPI_Reduce(s, r, n, t, op, 0, cc);
  return MPI_Bcast(r, n, t, 0, cc);
}
int MPI_Scan(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm cc)
{
  const Comm &c = C(cc);
  const uint64_t bytes = (uint64_t)n * tsize(t);
  std::vector<char> mine(bytes);
  memcpy(mine.data(), s == MPI_IN_PLACE ? r : s, bytes);
  if (c.me > 0)
  {
    std::vector<char> prev(bytes);
    crecv(c, c.me - 1, T_SCAN, prev.data(), bytes);
    combine(prev.data(), mine.data(), n, t, op);  // inclusive prefix: earlier ranks op mine
  }
  if (c.me + 1 < (int)c.ranks.size()) csend(c, c.me + 1, T_SCAN, mine.data(), bytes);
  memcpy(r, mine.data(), bytes);
  return MPI_SUCCESS;
}
int MPI_Alltoall(const void *s, int n, MPI_Datatype t, void *r, int rn, MPI_Datatype rt, MPI_Comm cc)
{
  const Comm &c = C(cc);
  const int P = (int)c.ranks.size();
  const uint64_t sb = (uint64_t)n * tsize(t), rb = (uint64_t)rn * tsize(rt);
  for (int q = 0; q < P; q++)
    if (q != c.me) csend(c, q, T_A2A, (const char *)s + q * sb, sb);
  for (int q = 0; q < P; q++)
  {
    if (q == c.me) memmove((char *)r + q * rb, (const char *)s + q * sb, sb);
    else crecv(c, q, T_A2A, (char *)r + q * rb, rb);
  }
  return MPI_SUCCESS;
}
int MPI_Alltoallv(const void *s, const int *sc, const int *sd, MPI_Datatype t, void *r, const int *rc, const int *rd, MPI_Datatype rt,
                  MPI_Comm cc)
{
  const Comm &c = C(cc);
  const int P = (int)c.ranks.size(), ss = tsize(t), rs = tsize(rt);
  for (int q = 0; q < P; q++)
    if (q != c.me) csend(c, q, T_A2A, (const char *)s + (uint64_t)sd[q] * ss, (uint64_t)sc[q] * ss);
  for (int q = 0; q < P; q++)
  {
    if (q == c.me) memmove((char *)r + (uint64_t)rd[q] * rs, (const char *)s + (uint64_t)sd[q] * ss, (uint64_t)sc[q] * ss);
    else crecv(c, q, T_A2A, (char *)r + (uint64_t)rd[q] * rs, (uint64_t)rc[q] * rs);
  }
  return MPI_SUCCESS;
}

// ---- point to point -----------------------------------------------------------------------------------------------------
int MPI_Send(const void *b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm cc)
{
  const Comm &c = C(cc);
  if (tag < 0) die("negative user tag");
  post(c.ranks[dst], tag, c.ctx, b, (uint64_t)n * tsize(t));
  return MPI_SUCCESS;
}
int MPI_Isend(const void *b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm cc, MPI_Request *rq)
{
  MPI_Send(b, n, t, dst, tag, cc);  // eager: complete at once
  Req q;
  q.active = true;
  reqs.push_back(q);
  *rq = (int)reqs.size() - 1;
  return MPI_SUCCESS;
}
int MPI_Issend(const void *b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm cc, MPI_Request *rq) { return MPI_Isend(b, n, t, dst, tag, cc, rq); }
int MPI_Recv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm cc, MPI_Status *st)
{
  const Comm &c = C(cc);
  take(src == MPI_ANY_SOURCE ? MPI_ANY_SOURCE : c.ranks[src], tag, c.ctx, b, (uint64_t)n * tsize(t), st, &c);
  return MPI_SUCCESS;
}
int MPI_Irecv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm cc, MPI_Request *rq)
{
  const Comm &c = C(cc);
  Req q;
  q.active = true; q.recv = true; q.buf = b; q.bytes = (uint64_t)n * tsize(t); q.ctx = c.ctx;
  q.src = src == MPI_ANY_SOURCE ? MPI_ANY_SOURCE : c.ranks[src]; q.tag = tag; q.comm = &c;
  reqs.push_back(q);
  *rq = (int)reqs.size() - 1;
  return MPI_SUCCESS;
}
int MPI_Wait(MPI_Request *rq, MPI_Status *st)
{
  if (!rq || *rq <= 0 || *rq >= (int)reqs.size()) return MPI_SUCCESS;
  Req &q = reqs[*rq];
  if (q.active && q.recv)
  {
    // comms may have been reallocated since the Irecv: look the communicator up again by context
    const Comm *cm = nullptr;
    for (const Comm &c : comms)
      if (c.ctx == q.ctx && c.me >= 0) cm = &c;
    take(q.src, q.tag, q.ctx, q.buf, q.bytes, st, cm);
  }
  q.active = false;
  *rq = MPI_REQUEST_NULL;
  return MPI_SUCCESS;
}
int MPI_Waitall(int n, MPI_Request *rq, MPI_Status *st)
{
  for (int i = 0; i < n; i++) MPI_Wait(rq + i, st ? st + i : nullptr);
  return MPI_SUCCESS;
}
int MPI_Sendrecv(const void *sb, int sn, MPI_Datatype stt, int dst, int stag, void *rb, int rn, MPI_Datatype rt, int src, int rtag, MPI_Comm cc,
                 MPI_Status *st)
{
  MPI_Send(sb, sn, stt, dst, stag, cc);
  return MPI_Recv(rb, rn, rt, src, rtag, cc, st);
}
int MPI_Get_count(const MPI_Status *st, MPI_Datatype t, int *c)
{
  *c = st ? (int)(st->dkt_bytes / std::max(1, tsize(t))) : 0;
  return MPI_SUCCESS;
}

// ---- datatypes, operations ------------------------------------------------------------------------------------------------
int MPI_Type_contiguous(int n, MPI_Datatype old, MPI_Datatype *nt) { *nt = DKT_MPI_T(0, n * tsize(old)); return MPI_SUCCESS; }
int MPI_Type_commit(MPI_Datatype *) { return MPI_SUCCESS; }
int MPI_Type_free(MPI_Datatype *) { return MPI_SUCCESS; }
int MPI_Op_create(MPI_User_function *f, int, MPI_Op *op)
{
  userops.push_back(f);
  *op = 100 + (int)userops.size() - 1;
  return MPI_SUCCESS;
}
int MPI_Op_free(MPI_Op *) { return MPI_SUCCESS; }

// ---- communicators and groups ----------------------------------------------------------------------------------------------
int MPI_Comm_split(MPI_Comm cc, int color, int key, MPI_Comm *out)
{
  Comm &c = C(cc);
  const int P = (int)c.ranks.size();
  const uint64_t seq = c.nsplit++;
  const uint64_t ctx = c.ctx;
  std::vector<int> mine = {color, key}, all(2 * P);
  MPI_Allgather(mine.data(), 2, MPI_INT, all.data(), 2, MPI_INT, cc);
  if (color == MPI_UNDEFINED) { *out = MPI_COMM_NULL; return MPI_SUCCESS; }
  std::vector<std::pair<std::pair<int, int>, int>> mem;  // ((key, old rank), world rank)
  const std::vector<int> ranks = comms[cc].ranks;
  for (int q = 0; q < P; q++)
    if (all[2 * q] == color) mem.push_back({{all[2 * q + 1], q}, ranks[q]});
  std::sort(mem.begin(), mem.end());
  std::vector<int> nr;
  for (auto &m : mem) nr.push_back(m.second);
  *out = new_comm(mix(mix(ctx, seq), (uint64_t)(unsigned)color), nr);
  return MPI_SUCCESS;
}
int MPI_Comm_dup(MPI_Comm cc, MPI_Comm *out)
{
  Comm &c = C(cc);
  const uint64_t seq = c.nsplit++;
  const std::vector<int> ranks = c.ranks;
  const uint64_t ctx = c.ctx;
  *out = new_comm(mix(mix(ctx, seq), 0xD0Dull), ranks);
  return MPI_SUCCESS;
}
int MPI_Comm_free(MPI_Comm *c) { if (c) *c = MPI_COMM_NULL; return MPI_SUCCESS; }
int MPI_Comm_group(MPI_Comm cc, MPI_Group *g)
{
  groups.push_back(C(cc).ranks);
  *g = (int)groups.size() - 1;
  return MPI_SUCCESS;
}
int MPI_Group_incl(MPI_Group g, int n, const int *ranks, MPI_Group *out)
{
  if (g <= 0 || g >= (int)groups.size()) die("invalid group");
  std::vector<int> r(n);
  for (int i = 0; i < n; i++) r[i] = groups[g][ranks[i]];
  groups.push_back(r);
  *out = (int)groups.size() - 1;
  return MPI_SUCCESS;
}
int MPI_Group_free(MPI_Group *g) { if (g) *g = MPI_GROUP_NULL; return MPI_SUCCESS; }
int MPI_Comm_create(MPI_Comm cc, MPI_Group g, MPI_Comm *out)
{
  Comm &c = C(cc);
  if (g <= 0 || g >= (int)groups.size()) die("invalid group");
  const uint64_t seq = c.nsplit++;  // collective over cc: every rank of cc counts it
  const uint64_t ctx = c.ctx;
  const std::vector<int> r = groups[g];
  const bool in = std::find(r.begin(), r.end(), ME) != r.end();
  if (!in || r.empty()) { *out = MPI_COMM_NULL; return MPI_SUCCESS; }
  *out = new_comm(mix(mix(ctx, seq), 0xC0000ull + (uint64_t)r[0]), r);
  return MPI_SUCCESS;
}
}  // extern "C"
