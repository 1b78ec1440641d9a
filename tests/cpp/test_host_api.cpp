// Exercises the C++ host API (ot::DA, feMatrix<LeafT,dim>::matVec, feVector::computeVec) the way a
// user of the reference would: a leaf class with an elementalMatVec callback, Dirichlet pre/post hooks.
// Driven by tests/test_cpp_api.py (GPU): reads a tree, K and u from files, writes A u and DA data.
//   usage: test_host_api <dim> <order> <maxDepth> <mode> <dir>
//   mode 0: dense level-scaled operator; 1: same + Dirichlet hooks; 2: position-dependent callback
//   (must be refused); 3: feVector::computeVec with the same callback
#define DKT_DEFINE_GLOBALS
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "feMatrix.h"
#include "feVector.h"

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
  bool dirichlet = false, positionDependent = false;
  ot::DA<dim> *da;
  DenseMat(ot::DA<dim> *d) : feMatrix<DenseMat<dim>, dim>(d, 1), da(d) {}
  virtual void elementalMatVec(const VECType *in, VECType *out, double *coords, double scale)
  {
    const unsigned N = da->getNumNodesPerElement();
    const double h = coords[(size_t)(N - 1) * dim] - coords[0];
    double s = scale * std::pow(h, alpha);
    if (positionDependent) s *= 1.0 + coords[0];
    for (unsigned i = 0; i < N; i++)
    {
      double acc = 0.0;
      for (unsigned j = 0; j < N; j++) acc += K[(size_t)i * N + j] * in[j];
      out[i] = s * acc;
    }
  }
  // HeatMat-style boundary handling (FEM/examples/src/heatMat.cpp:120-139)
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
class DenseVec : public feVector<DenseVec<dim>, dim>
{
public:
  std::vector<double> K;
  double alpha = 0.0;
  ot::DA<dim> *da;
  DenseVec(ot::DA<dim> *d) : feVector<DenseVec<dim>, dim>(d, 1), da(d) {}
  virtual void elementalComputeVec(const VECType *in, VECType *out, double *coords, double scale)
  {
    const unsigned N = da->getNumNodesPerElement();
    const double s = scale * std::pow(coords[(size_t)(N - 1) * dim] - coords[0], alpha);
    for (unsigned i = 0; i < N; i++)
    {
      double acc = 0.0;
      for (unsigned j = 0; j < N; j++) acc += K[(size_t)i * N + j] * in[j];
      out[i] = s * acc;
    }
  }
};

template <unsigned dim>
static int run(unsigned order, int mode, const std::string &dir)
{
  const std::vector<uint32_t> xyz = readAll<uint32_t>(dir + "/elem_xyz.bin");
  const std::vector<uint8_t> lev = readAll<uint8_t>(dir + "/elem_lev.bin");
  const std::vector<double> K = readAll<double>(dir + "/K.bin");
  const std::vector<double> u = readAll<double>(dir + "/u.bin");
  const std::vector<double> prm = readAll<double>(dir + "/params.bin");  // alpha, scale
  std::vector<ot::TreeNode<unsigned, dim>> tree;
  for (size_t i = 0; i < lev.size(); i++)
  {
    std::array<unsigned, dim> c;
    for (unsigned d = 0; d < dim; d++) c[d] = xyz[i * dim + d];
    tree.push_back(ot::TreeNode<unsigned, dim>(1, c, lev[i]));
  }
  ot::DA<dim> da(tree.data(), (unsigned)tree.size(), MPI_COMM_WORLD, order);
  if (da.getTotalNodalSz() != u.size()) { std::cerr << "node count mismatch " << da.getTotalNodalSz() << " vs " << u.size() << "\n"; return 2; }
  // DA order of nodes, for the Python side to compare with the oracle
  std::vector<uint32_t> nodes;
  for (unsigned i = 0; i < da.getTotalNodalSz(); i++)
  {
    for (unsigned d = 0; d < dim; d++) nodes.push_back(da.getTNCoords()[i].getX(d));
    nodes.push_back(da.getTNCoords()[i].getLevel());
  }
  writeAll(dir + "/nodes.bin", nodes);
  std::vector<double> v(u.size(), 0.0);
  try
  {
    if (mode == 3)
    {
      DenseVec<dim> vec(&da);
      vec.K = K; vec.alpha = prm[0];
      vec.computeVec(u.data(), v.data(), prm[1]);
    }
    else
    {
      DenseMat<dim> mat(&da);
      mat.K = K; mat.alpha = prm[0];
      mat.dirichlet = (mode == 1);
      mat.positionDependent = (mode == 2);
      mat.matVec(u.data(), v.data(), prm[1]);
      mat.matVec(u.data(), v.data(), prm[1]);  // second call re-uses the recovered operator
    }
  }
  catch (const std::exception &e)
  {
    std::cerr << "refused: " << e.what() << "\n";
    return 3;
  }
  writeAll(dir + "/v.bin", v);
  return 0;
}

int main(int argc, char **argv)
{
  if (argc < 6) { std::cerr << "usage: test_host_api dim order maxDepth mode dir\n"; return 2; }
  const int dim = std::atoi(argv[1]), order = std::atoi(argv[2]), mode = std::atoi(argv[4]);
  m_uiMaxDepth = (unsigned)std::atoi(argv[3]);
  const std::string dir = argv[5];
  if (dim == 2) return run<2>(order, mode, dir);
  if (dim == 3) return run<3>(order, mode, dir);
  if (dim == 4) return run<4>(order, mode, dir);
  return 2;
}
