// The reference's tree pipeline driven through the drop-in headers: points -> SFC_Tree::distTreeConstruction /
// distTreeBalancing -> ot::DA (what bench/src/matvec_bench_adaptive.cpp:92-150 and test/testMovingBall.cpp do).
// usage: test_tsort dim maxDepth maxPts balance dir     reads dir/pts.bin, writes dir/tree.bin (dim anchors + level per leaf)
#define DKT_DEFINE_GLOBALS
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "oda.h"
#include "tsort.h"

template <unsigned int dim>
static int run(int maxPts, int balance, const std::string &dir)
{
  std::ifstream f(dir + "/pts.bin", std::ios::binary);
  f.seekg(0, std::ios::end);
  const size_t bytes = (size_t)f.tellg();
  f.seekg(0);
  std::vector<unsigned int> pts(bytes / sizeof(unsigned int));
  f.read((char *)pts.data(), bytes);
  std::vector<ot::TreeNode<unsigned int, dim>> points, tree;
  for (size_t i = 0; i < pts.size() / dim; i++)
  {
    std::array<unsigned int, dim> c;
    for (unsigned d = 0; d < dim; d++) c[d] = pts[i * dim + d];
    points.push_back(ot::TreeNode<unsigned int, dim>(c, m_uiMaxDepth));
  }
  if (balance) ot::SFC_Tree<unsigned int, dim>::distTreeBalancing(points, tree, maxPts, 0.3, MPI_COMM_WORLD);
  else ot::SFC_Tree<unsigned int, dim>::distTreeConstruction(points, tree, maxPts, 0.3, MPI_COMM_WORLD);
  std::vector<unsigned int> out;
  for (const auto &t : tree)
  {
    for (unsigned d = 0; d < dim; d++) out.push_back(t.getX(d));
    out.push_back(t.getLevel());
  }
  std::ofstream o(dir + "/tree.bin", std::ios::binary);
  o.write((const char *)out.data(), out.size() * sizeof(unsigned int));
  // a balanced tree feeds ot::DA like any other (an unbalanced one has no defined matvec; a balanced tree whose level jumps
  // touch the domain boundary is class U and is refused)
  if (balance)
  {
    try
    {
      ot::DA<dim> da(tree.data(), (unsigned)tree.size(), MPI_COMM_WORLD, 1);
      std::printf("%zu leaves, %u nodes\n", tree.size(), da.getLocalNodalSz());
    }
    catch (const std::exception &e)
    {
      std::printf("%zu leaves, DA refused: %s\n", tree.size(), e.what());
    }
  }
  return 0;
}

int main(int argc, char **argv)
{
  if (argc < 6) { std::cerr << "usage: test_tsort dim maxDepth maxPts balance dir\n"; return 2; }
  const int dim = std::atoi(argv[1]);
  m_uiMaxDepth = (unsigned)std::atoi(argv[2]);
  const int maxPts = std::atoi(argv[3]), balance = std::atoi(argv[4]);
  try
  {
    if (dim == 2) return run<2>(maxPts, balance, argv[5]);
    if (dim == 3) return run<3>(maxPts, balance, argv[5]);
    if (dim == 4) return run<4>(maxPts, balance, argv[5]);
  }
  catch (const std::exception &e)
  {
    std::cerr << e.what() << "\n";
    return 3;
  }
  return 2;
}
