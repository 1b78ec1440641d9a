// ot::TreeNode<T,dim>: anchor + level, as include/treeNode.h:37-260 of the reference (the subset the
// matvec path and its callers use).
#ifndef DKT_HOST_TREENODE_H
#define DKT_HOST_TREENODE_H

#include <array>

#include "dendro.h"

namespace ot
{
template <typename T, unsigned int dim>
class TreeNode
{
protected:
  std::array<T, dim> m_uiCoords;
  unsigned int m_uiLevel;

public:
  using coordType = T;
  static constexpr unsigned int coordDim = dim;
  static constexpr char numChildren = (1u << dim);

  TreeNode() : m_uiLevel(0) { m_uiCoords.fill(0); }
  /** clips the coordinates to the anchor of the level-`level` cell (treeNode.tcc:30-45) */
  TreeNode(const std::array<T, dim> coords, unsigned int level) : m_uiLevel(level)
  {
    const T mask = ~((T(1) << (m_uiMaxDepth - level)) - 1);
    for (unsigned d = 0; d < dim; d++) m_uiCoords[d] = coords[d] & mask;
  }
  /** no clipping (treeNode.tcc:60-66) */
  TreeNode(const int, const std::array<T, dim> coords, unsigned int level) : m_uiCoords(coords), m_uiLevel(level) {}

  bool operator==(const TreeNode &o) const { return m_uiLevel == o.m_uiLevel && m_uiCoords == o.m_uiCoords; }
  bool operator!=(const TreeNode &o) const { return !(*this == o); }

  unsigned int getDim() const { return dim; }
  unsigned int getMaxDepth() const { return m_uiMaxDepth; }
  unsigned int getLevel() const { return m_uiLevel; }
  void setLevel(unsigned int l) { m_uiLevel = l; }
  T getX(int d) const { return m_uiCoords[d]; }
  void setX(int d, T c) { m_uiCoords[d] = c; }
  int getAnchor(std::array<T, dim> &xyz) const { xyz = m_uiCoords; return 1; }

  /** Morton child number at `level` (treeNode.tcc:245-258) */
  unsigned char getMortonIndex(T level) const
  {
    unsigned char c = 0;
    for (unsigned d = 0; d < dim; d++) c |= ((m_uiCoords[d] >> (m_uiMaxDepth - level)) & 1u) << d;
    return c;
  }
  unsigned char getMortonIndex() const { return getMortonIndex(m_uiLevel); }
  TreeNode getParent() const { return TreeNode(m_uiCoords, m_uiLevel - 1); }
  TreeNode getChildMorton(unsigned char child) const
  {
    TreeNode c(*this);
    c.m_uiLevel = m_uiLevel + 1;
    const T len = T(1) << (m_uiMaxDepth - c.m_uiLevel);
    for (unsigned d = 0; d < dim; d++)
      if ((child >> d) & 1u) c.m_uiCoords[d] += len;
    return c;
  }
  T minX(int d) const { return m_uiCoords[d]; }
  T maxX(int d) const { return m_uiCoords[d] + (T(1) << (m_uiMaxDepth - m_uiLevel)); }
  bool isRoot() const { return m_uiLevel == 0; }
  bool isAncestor(const TreeNode &o) const
  {
    if (!(m_uiLevel < o.m_uiLevel)) return false;
    for (unsigned d = 0; d < dim; d++)
      if (o.minX(d) < minX(d) || o.maxX(d) > maxX(d)) return false;
    return true;
  }
  /** treeNode.tcc:606-624 */
  bool isTouchingDomainBoundary() const
  {
    const T mask = (T(1) << m_uiMaxDepth) - 1, len = T(1) << (m_uiMaxDepth - m_uiLevel);
    for (unsigned d = 0; d < dim; d++)
      if (!(m_uiCoords[d] & mask) || !((m_uiCoords[d] + len) & mask)) return true;
    return false;
  }
  bool isOnDomainBoundary() const
  {
    const T mask = (T(1) << m_uiMaxDepth) - 1;
    for (unsigned d = 0; d < dim; d++)
      if (!(m_uiCoords[d] & mask)) return true;
    return false;
  }
};

/** Later Dendro-KT revisions wrap the distributed tree in ot::DistTree; provided so that callers
 *  written against either interface compile (the April-2019 reference passes std::vector). */
template <typename T, unsigned int dim>
class DistTree
{
  std::vector<TreeNode<T, dim>> m_tree;

public:
  DistTree() {}
  explicit DistTree(std::vector<TreeNode<T, dim>> &tree) { m_tree.swap(tree); }
  const std::vector<TreeNode<T, dim>> &getTreePartFiltered() const { return m_tree; }
  size_t size() const { return m_tree.size(); }
};
} // namespace ot
#endif
