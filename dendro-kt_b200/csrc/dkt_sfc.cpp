// SFC rotation tables generated from first principles.
//
// Layout follows the reference (include/hcurvedata.h:49-59, src/KDhcurvedata.cpp:24-59):
//   rot_perm[r][sfc]      Morton child visited at SFC position `sfc` under rotation r
//   rot_inv [r][child_m]  SFC position of Morton child `child_m`
//   htab    [r][child_m]  rotation of that child
// Morton mode is the single identity rotation.  Hilbert mode is the k-D Hilbert curve whose
// base shape is the binary reflected Gray code; every orientation is a signed permutation
// (entry corner e, axis permutation P): position(w) = e ^ P(gray(w)).  Child orientations of
// the base curve follow a recursion on the dimension (axis k-1 is inserted into the (k-1)-D
// child permutation at position 0 for even w, popcount(w) for odd w, mirrored for the second
// half); all other rows follow by composing signed permutations.  No table data is taken
// from the reference; tests/test_sfc_tables.py checks that the resulting ORDER equals the
// reference's KDhcurvedata tables for dim 2,3,4 (state numbering differs, order does not).
#include "dkt_internal.h"

#include <map>
#include <vector>

namespace dkt
{
static inline int gray(int i) { return i ^ (i >> 1); }
static inline int popc(int x) { return __builtin_popcount((unsigned)x); }

// entry corner of sub-cell w of the base curve
static int entry_of(int w) { return w == 0 ? 0 : gray(2 * ((w - 1) / 2)); }

static std::vector<std::vector<int>> base_child_perms(int n)
{
  if (n == 1) return {{0}, {0}};
  std::vector<std::vector<int>> low = base_child_perms(n - 1);
  const int half = 1 << (n - 1);
  std::vector<std::vector<int>> out(1 << n);
  for (int w = 0; w < (1 << n); w++)
  {
    const int wl = w < half ? w : (1 << n) - 1 - w;
    std::vector<int> p = low[wl];
    const int pos = (wl % 2 == 0) ? 0 : popc(wl);
    p.insert(p.begin() + pos, n - 1);
    out[w] = p;
  }
  return out;
}

struct Orient
{
  int e;
  std::vector<int> p;
  bool operator<(const Orient &o) const { return e != o.e ? e < o.e : p < o.p; }
};

static int apply_perm(const std::vector<int> &p, int x)
{
  int r = 0;
  for (size_t j = 0; j < p.size(); j++) r |= ((x >> j) & 1) << p[j];
  return r;
}

void make_sfc_tables(int dim, int mode, SfcTables &t)
{
  const int nch = 1 << dim;
  t.dim = dim;
  t.nch = nch;
  if (mode == DKT_SFC_MORTON)
  {
    t.nrot = 1;
    t.rot_perm.assign(nch, 0);
    t.rot_inv.assign(nch, 0);
    t.htab.assign(nch, 0);
    for (int i = 0; i < nch; i++) t.rot_perm[i] = t.rot_inv[i] = (uint8_t)i;
    return;
  }
  std::vector<std::vector<int>> bp = base_child_perms(dim);
  std::vector<Orient> states;
  std::map<Orient, int> ids;
  Orient id0;
  id0.e = 0;
  for (int j = 0; j < dim; j++) id0.p.push_back(j);
  states.push_back(id0);
  ids[id0] = 0;
  std::vector<int> child_state; // [state][w]
  for (size_t k = 0; k < states.size(); k++)
  {
    const Orient s = states[k];
    for (int w = 0; w < nch; w++)
    {
      Orient c;
      c.e = s.e ^ apply_perm(s.p, entry_of(w));
      c.p.resize(dim);
      for (int j = 0; j < dim; j++) c.p[j] = s.p[bp[w][j]];
      auto it = ids.find(c);
      int cid;
      if (it == ids.end())
      {
        cid = (int)states.size();
        ids[c] = cid;
        states.push_back(c);
      }
      else
        cid = it->second;
      child_state.push_back(cid);
    }
  }
  t.nrot = (int)states.size();
  t.rot_perm.assign((size_t)t.nrot * nch, 0);
  t.rot_inv.assign((size_t)t.nrot * nch, 0);
  t.htab.assign((size_t)t.nrot * nch, 0);
  for (int r = 0; r < t.nrot; r++)
    for (int w = 0; w < nch; w++)
    {
      const int m = states[r].e ^ apply_perm(states[r].p, gray(w));
      t.rot_perm[(size_t)r * nch + w] = (uint8_t)m;
      t.rot_inv[(size_t)r * nch + m] = (uint8_t)w;
      t.htab[(size_t)r * nch + m] = (uint8_t)child_state[(size_t)r * nch + w];
    }
}
} // namespace dkt
