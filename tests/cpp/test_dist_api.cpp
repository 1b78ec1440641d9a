// The reference's distributed API on several GPUs, one process per GPU: ot::DA<dim> built by every rank of the launch,
// feMatrix<LeafT,dim>::matVec with an elementalMatVec callback and Dirichlet hooks, readFromGhostBegin/End and
// writeToGhostsBegin/End on host vectors (include/oda.h:150,300-322; FEM/include/feMatrix.h:190-259).
// Launched by tests/test_gpu_dist.py once per rank with RANK / WORLD_SIZE / CUDA_VISIBLE_DEVICES / DKT_NCCL_ID_FILE set.
//   usage: test_dist_api <dim> <order> <maxDepth> <dirichlet 0|1> <dir>
//   reads  dir/elem_xyz.bin elem_lev.bin K.bin u.bin params.bin (u in the single-rank DA order)
//   writes dir/rank<r>_{ids.bin (position of each owned node in the single-rank order), v.bin, ghost.bin, nodes.bin}
#define DKT_DEFINE_GLOBALS
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "feMatrix.h"

template <typename T>
static std::vector<T> readAll(const std::string &path)
{
  std::ifstream f(path, std::ios::binary);
  if (!f) { std::cerr << "cannot open " << path << "\n"; std::exit(2); }
  f.seekg(0, std::ios::end);
  const size_t n = (size_t)f.tellg() / sizeof(T);
  f.seekg(0);
  std::vector<T> v(n);
  f.read((char *)v.data(), n * sizeof(T));
  return v;
}
template <typename T>
static void writeAll(const std::string &path, const std::vector<T> &v)
{
  std::ofstream f(path, std::ios::binary);
  f.write((const char *)v.data(), v.size() * sizeof(T));
}

template <unsigned dim>
class DenseMat : public feMatrix<DenseMat<dim>, dim>
{
public:
  std::vector<double> K;
  double alpha = 0.0;
  bool dirichlet = false;
  ot::DA<dim> *da;
  DenseMat(ot::DA<dim> *d) : feMatrix<DenseMat<dim>, dim>(d, 1), da(d) {}
  virtual void elementalMatVec(const VECType *in, VECType *out, double *coords, double scale)
  {
    const unsigned N = da->getNumNodesPerElement();
    const double h = coords[(size_t)(N - 1) * dim] - coords[0];
    const double s = scale * std::pow(h, alpha);
    for (unsigned i = 0; i < N; i++)
    {
      double acc = 0.0;
      for (unsigned j = 0; j < N; j++) acc += K[(size_t)i * N + j] * in[j];
      out[i] = s * acc;
    }
  }
  bool preMatVec(const VECType *, VECType *out, double) { zero(out); return true; }
  bool postMatVec(const VECType *, VECType *out, double) { zero(out); return true; }
  void zero(VECType *v)
  {
    if (!dirichlet) return;
    std::vector<unsigned int> b;
    da->getBoundaryNodeIndices(b);
    for (unsigned i : b) v[i] = 0.0;
  }
};

template <unsigned dim>
static int run(int order, int dirichlet, const std::string &dir)
{
  const auto exyz = readAll<unsigned int>(dir + "/elem_xyz.bin");
  const auto elev = readAll<unsigned char>(dir + "/elem_lev.bin");
  const auto K = readAll<double>(dir + "/K.bin");
  const auto u = readAll<double>(dir + "/u.bin");
  const auto prm = readAll<double>(dir + "/params.bin");
  std::vector<ot::TreeNode<unsigned int, dim>> tree(elev.size());
  for (size_t i = 0; i < elev.size(); i++)
  {
    std::array<unsigned int, dim> c;
    for (unsigned d = 0; d < dim; d++) c[d] = exyz[i * dim + d];
    tree[i] = ot::TreeNode<unsigned int, dim>(1, c, elev[i]);
  }
  ot::DA<dim> da(tree.data(), (unsigned)tree.size(), MPI_COMM_WORLD, order);
  const unsigned nLoc = da.getLocalNodalSz(), nTot = da.getTotalNodalSz(), rank = da.getRankAll();
  if (da.getLocalNodeBegin() + nLoc + da.getPostNodalSz() != nTot || da.getPreNodalSz() != da.getLocalNodeBegin()) return 4;
  const std::vector<unsigned int> ids = da.getOwnedGlobalIds();
  const std::string pre = dir + "/rank" + std::to_string(rank) + "_";
  writeAll(pre + "ids.bin", ids);

  // v = A u on the owned nodes, through the callback
  DenseMat<dim> mat(&da);
  mat.K = K;
  mat.alpha = prm[0];
  mat.dirichlet = dirichlet != 0;
  std::vector<double> uLoc(nLoc), vLoc(nLoc, 0.0);
  for (unsigned i = 0; i < nLoc; i++) uLoc[i] = u[ids[i]];
  mat.matVec(uLoc.data(), vLoc.data(), prm[1]);
  writeAll(pre + "v.bin", vLoc);

  // ghost read: owners' values (here: their single-rank position) appear in the ghost segment;
  // ghost write of ones: every owned node gains one per rank that ghosts it
  std::vector<double> g(nTot, -1.0);
  for (unsigned i = 0; i < nLoc; i++) g[da.getLocalNodeBegin() + i] = (double)ids[i];
  da.readFromGhostBegin(g.data(), 1);
  da.readFromGhostEnd(g.data(), 1);
  std::vector<double> w(nTot, 1.0);
  da.writeToGhostsBegin(w.data(), 1);
  da.writeToGhostsEnd(w.data(), 1);
  g.insert(g.end(), w.begin(), w.end());
  writeAll(pre + "ghost.bin", g);

  // node coordinates of the local vector [owned | ghosts]
  std::vector<unsigned int> nodes;
  const ot::TreeNode<unsigned int, dim> *tn = da.getTNCoords();
  for (unsigned i = 0; i < nTot; i++)
  {
    for (unsigned d = 0; d < dim; d++) nodes.push_back(tn[i].getX(d));
    nodes.push_back(tn[i].getLevel());
  }
  writeAll(pre + "nodes.bin", nodes);
  std::printf("rank %u of %u: %u owned, %u ghost nodes of %u\n", rank, da.getNpesAll(), nLoc, nTot - nLoc, da.getGlobalNodeSz());
  return 0;
}

int main(int argc, char **argv)
{
  if (argc < 6) { std::cerr << "usage: test_dist_api dim order maxDepth dirichlet dir\n"; return 2; }
  const int dim = std::atoi(argv[1]), order = std::atoi(argv[2]), diri = std::atoi(argv[4]);
  m_uiMaxDepth = (unsigned)std::atoi(argv[3]);
  try
  {
    if (dim == 2) return run<2>(order, diri, argv[5]);
    if (dim == 3) return run<3>(order, diri, argv[5]);
    if (dim == 4) return run<4>(order, diri, argv[5]);
  }
  catch (const std::exception &e)
  {
    std::cerr << e.what() << "\n";
    return 3;
  }
  return 2;
}
