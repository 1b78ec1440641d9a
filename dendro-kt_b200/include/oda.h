// ot::DA<dim> over libdkt.so: the distributed-array object of include/oda.h:41-525, for the matvec
// path.  Construction (src/oda.cpp:46-151: node generation, dedup, ordering, maps) happens on the
// GPU inside dkt_da_create; this class keeps the reference's getters and vector helpers.
#ifndef DKT_HOST_ODA_H
#define DKT_HOST_ODA_H

#include <cstring>
#include <vector>

#include "treeNode.h"

namespace ot
{
template <unsigned int dim>
class DA
{
  using C = unsigned int;

  dkt_da *m_handle = nullptr;
  unsigned int m_uiElementOrder = 1, m_uiNpE = 0;
  unsigned int m_uiTotalNodalSz = 0, m_uiLocalNodalSz = 0, m_uiLocalNodeBegin = 0, m_uiGlobalNodeSz = 0;
  unsigned int m_uiLocalElementSz = 0;
  MPI_Comm m_uiGlobalComm = MPI_COMM_WORLD;
  std::vector<TreeNode<C, dim>> m_tnCoords;
  TreeNode<C, dim> m_treePartFront, m_treePartBack;
  std::vector<unsigned int> m_uiBdyNodeIds;
  dkt_sizes m_sizes;

  DA(const DA &) = delete;
  DA &operator=(const DA &) = delete;

public:
  DA() { std::memset(&m_sizes, 0, sizeof(m_sizes)); }

  /** @param inTree 2:1 balanced, complete linear tree (sorted or not: the SFC sort runs on the GPU)
   *  @param sfcMode DKT_SFC_MORTON (the reference's default build) or DKT_SFC_HILBERT
   *  @param ip0,ip1 RefElement::getIMChild0/1() of the host's reference element, or null */
  DA(const TreeNode<C, dim> *inTree, unsigned int nEle, MPI_Comm comm, unsigned int order, unsigned int grainSz = 100,
     double sfc_tol = 0.3, int sfcMode = DKT_SFC_MORTON, const double *ip0 = nullptr, const double *ip1 = nullptr)
  {
    (void)grainSz; (void)sfc_tol;
    construct(inTree, nEle, comm, order, sfcMode, ip0, ip1);
  }
  DA(const DistTree<C, dim> &tree, MPI_Comm comm, unsigned int order, unsigned int grainSz = 100, double sfc_tol = 0.3)
  {
    (void)grainSz; (void)sfc_tol;
    construct(tree.getTreePartFiltered().data(), (unsigned)tree.size(), comm, order, DKT_SFC_MORTON, nullptr, nullptr);
  }
  ~DA() { if (m_handle) dkt_da_destroy(m_handle); }

  void construct(const TreeNode<C, dim> *inTree, unsigned int nEle, MPI_Comm comm, unsigned int order, int sfcMode = DKT_SFC_MORTON,
                 const double *ip0 = nullptr, const double *ip1 = nullptr)
  {
    m_uiGlobalComm = comm;
    m_uiElementOrder = order;
    std::vector<uint32_t> xyz((size_t)nEle * dim);
    std::vector<uint8_t> lev(nEle);
    for (unsigned i = 0; i < nEle; i++)
    {
      for (unsigned d = 0; d < dim; d++) xyz[(size_t)i * dim + d] = inTree[i].getX(d);
      lev[i] = (uint8_t)inTree[i].getLevel();
    }
    dkt_host::check(dkt_da_create(dim, order, m_uiMaxDepth, sfcMode, xyz.data(), lev.data(), nEle, ip0, ip1, 0u, &m_handle), "DA::construct");
    dkt_host::check(dkt_da_sizes(m_handle, &m_sizes), "dkt_da_sizes");
    m_uiNpE = m_sizes.nodes_per_elem;
    m_uiTotalNodalSz = m_uiLocalNodalSz = m_uiGlobalNodeSz = (unsigned)m_sizes.n_nodes;
    m_uiLocalNodeBegin = 0;
    m_uiLocalElementSz = nEle;
    // node coordinates in DA order (src/oda.cpp:131-140)
    std::vector<uint32_t> nx((size_t)m_sizes.n_nodes * dim);
    std::vector<uint8_t> nl(m_sizes.n_nodes);
    dkt_host::check(dkt_da_export_nodes(m_handle, nx.data(), nl.data()), "dkt_da_export_nodes");
    m_tnCoords.resize(m_sizes.n_nodes);
    for (size_t i = 0; i < m_tnCoords.size(); i++)
    {
      std::array<C, dim> c;
      for (unsigned d = 0; d < dim; d++) c[d] = nx[i * dim + d];
      m_tnCoords[i] = TreeNode<C, dim>(1, c, nl[i]);
    }
    // splitters (src/oda.cpp:85-86): first and last element of the sorted tree
    dkt_host::check(dkt_da_export_elements(m_handle, xyz.data(), lev.data()), "dkt_da_export_elements");
    auto elem = [&](size_t i) {
      std::array<C, dim> c;
      for (unsigned d = 0; d < dim; d++) c[d] = xyz[i * dim + d];
      return TreeNode<C, dim>(1, c, lev[i]);
    };
    m_treePartFront = elem(0);
    m_treePartBack = elem(nEle - 1);
    m_uiBdyNodeIds.resize(m_sizes.n_boundary);
    dkt_host::check(dkt_da_export_boundary(m_handle, m_uiBdyNodeIds.data()), "dkt_da_export_boundary");
  }

  // ---- getters of include/oda.h:180-264 ------------------------------------------------------------
  unsigned int getLocalNodalSz() const { return m_uiLocalNodalSz; }
  unsigned int getLocalNodeBegin() const { return m_uiLocalNodeBegin; }
  unsigned int getPreNodalSz() const { return 0; }
  unsigned int getPostNodalSz() const { return 0; }
  unsigned int getTotalNodalSz() const { return m_uiTotalNodalSz; }
  unsigned int getGlobalNodeSz() const { return m_uiGlobalNodeSz; }
  unsigned int getGlobalRankBegin() const { return 0; }
  unsigned int getLocalElemSz() const { return m_uiLocalElementSz; }
  bool isActive() { return true; }
  unsigned int getNumNodesPerElement() const { return m_uiNpE; }
  unsigned int getElementOrder() const { return m_uiElementOrder; }
  MPI_Comm getGlobalComm() const { return m_uiGlobalComm; }
  MPI_Comm getCommActive() const { return m_uiGlobalComm; }
  unsigned int getNpesAll() const { return 1; }
  unsigned int getNpesActive() const { return 1; }
  unsigned int getRankAll() const { return 0; }
  unsigned int getRankActive() const { return 0; }
  unsigned int getMaxDepth() const { return m_uiMaxDepth; }
  unsigned int getDimension() const { return dim; }
  const TreeNode<C, dim> *getTNCoords() const { return m_tnCoords.data(); }
  const TreeNode<C, dim> *getTreePartFront() const { return &m_treePartFront; }
  const TreeNode<C, dim> *getTreePartBack() const { return &m_treePartBack; }
  void getBoundaryNodeIndices(std::vector<unsigned int> &bdyIndex) const { bdyIndex = m_uiBdyNodeIds; }
  /** the device-side object and its tree class: not part of the reference API */
  dkt_da *handle() const { return m_handle; }
  const dkt_sizes &sizes() const { return m_sizes; }

  // ---- vector helpers of include/oda.h:273-413 (single rank: ghosted == local) ---------------------
  template <typename T>
  int createVector(T *&local, bool isElemental = false, bool isGhosted = false, unsigned int dof = 1) const
  {
    (void)isGhosted;
    if (isElemental) throw std::runtime_error("elemental vectors are not supported");
    local = new T[(size_t)dof * m_uiTotalNodalSz]();
    return 0;
  }
  template <typename T>
  int createVector(std::vector<T> &local, bool isElemental = false, bool isGhosted = false, unsigned int dof = 1) const
  {
    (void)isGhosted;
    if (isElemental) throw std::runtime_error("elemental vectors are not supported");
    local.assign((size_t)dof * m_uiTotalNodalSz, T());
    return 0;
  }
  template <typename T>
  void destroyVector(T *&local) const { delete[] local; local = nullptr; }
  template <typename T>
  void destroyVector(std::vector<T> &local) const { local.clear(); }
  template <typename T>
  void nodalVecToGhostedNodal(const T *in, T *&out, bool isAllocated = false, unsigned int dof = 1) const
  {
    if (!isAllocated) createVector(out, false, true, dof);
    std::memcpy(out + (size_t)dof * m_uiLocalNodeBegin, in, sizeof(T) * dof * m_uiLocalNodalSz);
  }
  template <typename T>
  void ghostedNodalToNodalVec(const T *gVec, T *&local, bool isAllocated = false, unsigned int dof = 1) const
  {
    if (!isAllocated) createVector(local, false, false, dof);
    std::memcpy(local, gVec + (size_t)dof * m_uiLocalNodeBegin, sizeof(T) * dof * m_uiLocalNodalSz);
  }
  // single rank: the exchanges are no-ops exactly as in the reference (include/oda.tcc:216,283,324,387)
  template <typename T> void readFromGhostBegin(T *, unsigned int = 1) {}
  template <typename T> void readFromGhostEnd(T *, unsigned int = 1) {}
  template <typename T> void writeToGhostsBegin(T *, unsigned int = 1) {}
  template <typename T> void writeToGhostsEnd(T *, unsigned int = 1) {}
};
} // namespace ot
#endif
