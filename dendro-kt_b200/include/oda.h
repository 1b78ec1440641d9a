// ot::DA<dim> over libdkt.so: the distributed-array object of include/oda.h:41-525, for the matvec
// path.  Construction (src/oda.cpp:46-151: node generation, dedup, ordering, maps) happens on the
// GPU inside dkt_da_create; this class keeps the reference's getters and vector helpers.
#ifndef DKT_HOST_ODA_H
#define DKT_HOST_ODA_H

#include <algorithm>
#include <cstring>
#include <functional>
#include <iostream>
#include <type_traits>
#include <vector>

#include "treeNode.h"
#include "mathUtils.h"
#include "refel.h"

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
  int m_rank = 0, m_nranks = 1;
  MPI_Comm m_uiGlobalComm = MPI_COMM_WORLD;
  std::vector<TreeNode<C, dim>> m_tnCoords;
  TreeNode<C, dim> m_treePartFront, m_treePartBack;
  std::vector<unsigned int> m_uiBdyNodeIds;
  dkt_sizes m_sizes;
  RefElement m_refel;  // include/oda.h:137

  DA(const DA &) = delete;
  DA &operator=(const DA &) = delete;

public:
  DA() : m_refel(dim, 1) { std::memset(&m_sizes, 0, sizeof(m_sizes)); }

  /** @param inTree 2:1 balanced, complete linear tree (sorted or not: the SFC sort runs on the GPU)
   *  @param sfcMode DKT_SFC_MORTON (the reference's default build) or DKT_SFC_HILBERT
   *  @param ip0,ip1 RefElement::getIMChild0/1() of the host's reference element, or null */
  DA(const TreeNode<C, dim> *inTree, unsigned int nEle, MPI_Comm comm, unsigned int order, unsigned int grainSz = 100,
     double sfc_tol = 0.3, int sfcMode = DKT_SFC_MORTON, const double *ip0 = nullptr, const double *ip1 = nullptr)
      : m_refel(dim, order)
  {
    (void)grainSz; (void)sfc_tol;
    construct(inTree, nEle, comm, order, sfcMode, ip0, ip1);
  }
  DA(const DistTree<C, dim> &tree, MPI_Comm comm, unsigned int order, unsigned int grainSz = 100, double sfc_tol = 0.3)
      : m_refel(dim, order)
  {
    (void)grainSz; (void)sfc_tol;
    construct(tree.getTreePartFiltered().data(), (unsigned)tree.size(), comm, order, DKT_SFC_MORTON, nullptr, nullptr);
  }
  /** Regular grid (include/oda.h:155, include/oda.tcc:15-62): the uniform tree of the smallest level with at least
   *  nProc * grainSz elements, 2^(dim * level) of them; one process drives one GPU here, so nProc = 1. */
  DA(MPI_Comm comm, unsigned int order, unsigned int grainSz = 100, double sfc_tol = 0.3) : m_refel(dim, order)
  {
    (void)sfc_tol;
    if (grainSz == 0) grainSz = 1;
    unsigned int bits = 0;
    for (unsigned int v = grainSz - 1; v; v >>= 1) bits++;        // binOp::binLength(nProc * grainSz - 1)
    const unsigned int endL = (bits + dim - 1) / dim;
    if (endL > m_uiMaxDepth) throw std::runtime_error("regular grid deeper than m_uiMaxDepth");
    const unsigned int n1 = 1u << endL, len = 1u << (m_uiMaxDepth - endL);
    size_t n = 1;
    for (unsigned d = 0; d < dim; d++) n *= n1;
    std::vector<TreeNode<C, dim>> tree(n);
    for (size_t i = 0; i < n; i++)
    {
      std::array<C, dim> c;
      size_t r = i;
      for (unsigned d = 0; d < dim; d++) { c[d] = (C)(r % n1) * len; r /= n1; }
      tree[i] = TreeNode<C, dim>(1, c, endL);
    }
    construct(tree.data(), (unsigned)n, comm, order);
  }
  /** The function-driven constructor of the reference (include/oda.h:167: refinement by interpolation error, then 2:1
   *  balancing) belongs to the tree pipeline, which is outside this path: build the tree with the reference's
   *  function2Octree / distTreeBalancing and pass it to the constructor above. */
  template <typename T>
  DA(std::function<void(const T *, T *)>, unsigned int, MPI_Comm, unsigned int, double, unsigned int = 100, double = 0.3) : m_refel(dim, 1)
  {
    throw std::runtime_error("ot::DA(function): tree generation is not part of the B200 matvec path; construct the DA from a tree");
  }
  ~DA() { if (m_handle) dkt_da_destroy(m_handle); }

  void construct(const TreeNode<C, dim> *inTree, unsigned int nEle, MPI_Comm comm, unsigned int order, int sfcMode = DKT_SFC_MORTON,
                 const double *ip0 = nullptr, const double *ip1 = nullptr)
  {
    m_uiGlobalComm = comm;
    m_uiElementOrder = order;
    if (!ip0 || !ip1) { ip0 = m_refel.getIMChild0(); ip1 = m_refel.getIMChild1(); }  // the reference uploads its own RefElement's
    std::vector<uint32_t> xyz((size_t)nEle * dim);
    std::vector<uint8_t> lev(nEle);
    for (unsigned i = 0; i < nEle; i++)
    {
      for (unsigned d = 0; d < dim; d++) xyz[(size_t)i * dim + d] = inTree[i].getX(d);
      lev[i] = (uint8_t)inTree[i].getLevel();
    }
    // Several ranks (one process per GPU): the DA is partitioned - SFC-contiguous element ranges, owned + ghost nodes, exchanges
    // over NCCL (dkt_da_create_dist).  The library partitions a tree that every rank holds in full: with MPI the ranks' pieces
    // (the reference's distributed tree) are gathered first; without MPI every rank must pass the whole tree.
    dkt_host::comm_rank_size(comm, m_rank, m_nranks);
    if (m_nranks > 1)
    {
#if defined(MPI_VERSION) || defined(DKT_HAVE_MPI)
      {
        int cnt = (int)nEle;
        std::vector<int> cnts(m_nranks), dx(m_nranks), dl(m_nranks), cx(m_nranks);
        MPI_Allgather(&cnt, 1, MPI_INT, cnts.data(), 1, MPI_INT, comm);
        size_t tot = 0;
        for (int r = 0; r < m_nranks; r++) { dl[r] = (int)tot; dx[r] = (int)(tot * dim); cx[r] = cnts[r] * (int)dim; tot += cnts[r]; }
        std::vector<uint32_t> ax(tot * dim);
        std::vector<uint8_t> al(tot);
        MPI_Allgatherv(xyz.data(), cnt * (int)dim, MPI_UNSIGNED, ax.data(), cx.data(), dx.data(), MPI_UNSIGNED, comm);
        MPI_Allgatherv(lev.data(), cnt, MPI_UNSIGNED_CHAR, al.data(), cnts.data(), dl.data(), MPI_UNSIGNED_CHAR, comm);
        xyz.swap(ax);
        lev.swap(al);
        nEle = (unsigned)tot;
      }
#endif
      char id[128];
      dkt_host::share_nccl_id(comm, m_rank, id);
      dkt_host::check(dkt_da_create_dist(dim, order, m_uiMaxDepth, sfcMode, xyz.data(), lev.data(), nEle, ip0, ip1, 0u, m_rank, m_nranks, id,
                                         &m_handle),
                      "DA::construct");
    }
    else
      dkt_host::check(dkt_da_create(dim, order, m_uiMaxDepth, sfcMode, xyz.data(), lev.data(), nEle, ip0, ip1, 0u, &m_handle), "DA::construct");
    dkt_host::check(dkt_da_sizes(m_handle, &m_sizes), "dkt_da_sizes");
    m_uiNpE = m_sizes.nodes_per_elem;
    // local vector layout [owned | ghosts]: getLocalNodeBegin() == 0, getPostNodalSz() == ghosts (the reference puts the ghosts owned by
    // lower ranks in front, include/oda.h:180-200; applications address the layout through these getters)
    m_uiLocalNodalSz = (unsigned)m_sizes.n_nodes;
    m_uiTotalNodalSz = (unsigned)(m_sizes.n_nodes + m_sizes.n_ghost_nodes);
    m_uiGlobalNodeSz = (unsigned)m_sizes.n_global_nodes;
    m_uiLocalNodeBegin = 0;
    m_uiLocalElementSz = nEle;
    // node coordinates in DA order (src/oda.cpp:131-140)
    std::vector<uint32_t> nx((size_t)m_uiTotalNodalSz * dim);
    std::vector<uint8_t> nl(m_uiTotalNodalSz);
    dkt_host::check(dkt_da_export_nodes(m_handle, nx.data(), nl.data()), "dkt_da_export_nodes");
    m_tnCoords.resize(m_uiTotalNodalSz);
    for (size_t i = 0; i < m_tnCoords.size(); i++)
    {
      std::array<C, dim> c;
      for (unsigned d = 0; d < dim; d++) c[d] = nx[i * dim + d];
      m_tnCoords[i] = TreeNode<C, dim>(1, c, nl[i]);
    }
    // splitters (src/oda.cpp:85-86): first and last element of the sorted tree (of the whole tree, also on a partitioned DA)
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
  unsigned int getPostNodalSz() const { return m_uiTotalNodalSz - m_uiLocalNodalSz; }
  unsigned int getTotalNodalSz() const { return m_uiTotalNodalSz; }
  unsigned int getGlobalNodeSz() const { return m_uiGlobalNodeSz; }
  unsigned int getGlobalRankBegin() const { return 0; }
  unsigned int getLocalElemSz() const { return m_uiLocalElementSz; }
  bool isActive() { return true; }
  unsigned int getNumNodesPerElement() const { return m_uiNpE; }
  unsigned int getElementOrder() const { return m_uiElementOrder; }
  MPI_Comm getGlobalComm() const { return m_uiGlobalComm; }
  MPI_Comm getCommActive() const { return m_uiGlobalComm; }
  unsigned int getNpesAll() const { return (unsigned)m_nranks; }
  unsigned int getNpesActive() const { return (unsigned)m_nranks; }
  unsigned int getRankAll() const { return (unsigned)m_rank; }
  unsigned int getRankActive() const { return (unsigned)m_rank; }
  /** partitioned DA: position of every owned node in the single-rank DA order (the order the oracle and the fixtures use) */
  std::vector<unsigned int> getOwnedGlobalIds() const
  {
    std::vector<unsigned int> ids(m_uiLocalNodalSz);
    if (m_nranks > 1) dkt_host::check(dkt_da_export_owned_ids(m_handle, ids.data()), "dkt_da_export_owned_ids");
    else
      for (unsigned i = 0; i < m_uiLocalNodalSz; i++) ids[i] = i;
    return ids;
  }
  unsigned int getMaxDepth() const { return m_uiMaxDepth; }
  unsigned int getDimension() const { return dim; }
  const TreeNode<C, dim> *getTNCoords() const { return m_tnCoords.data(); }
  const TreeNode<C, dim> *getTreePartFront() const { return &m_treePartFront; }
  const TreeNode<C, dim> *getTreePartBack() const { return &m_treePartBack; }
  void getBoundaryNodeIndices(std::vector<unsigned int> &bdyIndex) const { bdyIndex = m_uiBdyNodeIds; }
  const RefElement *getReferenceElement() const { return &m_refel; }  // include/oda.h:261
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
  // Ghost exchanges on HOST vectors of getTotalNodalSz() entries (include/oda.h:300-322).  One rank: no-ops exactly as in the
  // reference (include/oda.tcc:216,283,324,387).  Several ranks: Begin performs the exchange over NCCL (dkt_ghost_read_host /
  // dkt_ghost_write_host: upload, exchange, download of the changed segment), End has nothing left to wait for.  Resident device
  // vectors use dkt_ghost_read_begin/end, dkt_ghost_write_begin/end of the C ABI instead.
  template <typename T> void readFromGhostBegin(T *vec, unsigned int dof = 1) { ghostOp(vec, dof, 0, true); }
  template <typename T> void readFromGhostEnd(T *vec, unsigned int dof = 1) { ghostOp(vec, dof, 0, false); }
  template <typename T> void writeToGhostsBegin(T *vec, unsigned int dof = 1) { ghostOp(vec, dof, 1, true); }
  template <typename T> void writeToGhostsEnd(T *vec, unsigned int dof = 1) { ghostOp(vec, dof, 1, false); }

private:
  template <typename T>
  void ghostOp(T *vec, unsigned int dof, int which, bool begin)
  {
    if (m_sizes.n_ranks <= 1) return;
    if (dof != 1 || !std::is_same<T, double>::value) throw std::runtime_error("ghost exchange: dof == 1, double vectors");
    if (!begin) return;
    double *v = reinterpret_cast<double *>(vec);
    dkt_host::check(which == 0 ? dkt_ghost_read_host(m_handle, v) : dkt_ghost_write_host(m_handle, v), "ghost exchange");
  }
};
} // namespace ot
#endif
