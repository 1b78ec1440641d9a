// ot::SFC_Tree<T,dim> over libdkt.so: tree construction and 2:1 balancing from points (include/tsort.h:109-336,
// src/tsort.cpp:566-877) on the GPU - dkt_tree_from_points.  Same static entry points, same arguments; the leaves come
// back in the reference's tree order.  The curve follows the reference's build switch: -DHILBERT_ORDERING selects the
// Hilbert curve, otherwise Morton (src/KDhcurvedata.cpp:47-57).
#ifndef DKT_HOST_TSORT_H
#define DKT_HOST_TSORT_H

#include <array>
#include <stdexcept>
#include <vector>

#include "treeNode.h"

namespace ot
{
typedef long long RankI;  // DendroIntL with USE_64BIT_INDICES (include/tsort.h:26)

template <typename T, unsigned int D>
struct SFC_Tree
{
  /** Complete linear tree with at most maxPtsPerRegion points per leaf (src/tsort.cpp:647-716).  `points`: TreeNodes of
   *  level m_uiMaxDepth, as ot::getPts() makes them (include/octUtils.h:29-66). */
  static void distTreeConstruction(std::vector<TreeNode<T, D>> &points, std::vector<TreeNode<T, D>> &tree, RankI maxPtsPerRegion,
                                   double /*loadFlexibility*/, MPI_Comm /*comm*/)
  {
    fromPoints(points, tree, maxPtsPerRegion, false);
  }
  /** The same followed by 2:1 balancing across faces, edges and corners (src/tsort.cpp:862-877). */
  static void distTreeBalancing(std::vector<TreeNode<T, D>> &points, std::vector<TreeNode<T, D>> &tree, RankI maxPtsPerRegion,
                                double /*loadFlexibility*/, MPI_Comm /*comm*/)
  {
    fromPoints(points, tree, maxPtsPerRegion, true);
  }
  /** src/tsort.cpp:829-856 */
  static void locTreeBalancing(std::vector<TreeNode<T, D>> &points, std::vector<TreeNode<T, D>> &tree, RankI maxPtsPerRegion)
  {
    fromPoints(points, tree, maxPtsPerRegion, true);
  }

private:
  static void fromPoints(const std::vector<TreeNode<T, D>> &points, std::vector<TreeNode<T, D>> &tree, RankI maxPts, bool balance)
  {
    tree.clear();
    if (points.empty()) return;  // the reference returns an empty tree
    std::vector<uint32_t> pts(points.size() * D);
    for (size_t i = 0; i < points.size(); i++)
    {
      if (points[i].getLevel() != m_uiMaxDepth)
        throw std::runtime_error("SFC_Tree: points must be TreeNodes of level m_uiMaxDepth (seeds of other levels are not supported)");
      for (unsigned d = 0; d < D; d++) pts[i * D + d] = points[i].getX(d);
    }
#ifdef HILBERT_ORDERING
    const int sfc = DKT_SFC_HILBERT;
#else
    const int sfc = DKT_SFC_MORTON;
#endif
    dkt_tree *h = nullptr;
    dkt_host::check(dkt_tree_from_points(D, m_uiMaxDepth, sfc, pts.data(), points.size(), (uint64_t)maxPts, balance ? 1 : 0, 0u, &h),
                    "SFC_Tree::fromPoints");
    uint64_t n = 0;
    dkt_tree_size(h, &n, nullptr);
    std::vector<uint32_t> xyz(n * D);
    std::vector<uint8_t> lev(n);
    const int rc = dkt_tree_export(h, xyz.data(), lev.data(), 0u);
    dkt_tree_destroy(h);
    dkt_host::check(rc, "SFC_Tree::fromPoints");
    tree.reserve(n);
    for (uint64_t i = 0; i < n; i++)
    {
      std::array<T, D> c;
      for (unsigned d = 0; d < D; d++) c[d] = xyz[i * D + d];
      tree.push_back(TreeNode<T, D>(1, c, lev[i]));
    }
  }
};
}  // namespace ot
#endif
