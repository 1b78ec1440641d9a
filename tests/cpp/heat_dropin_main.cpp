// Driver of the drop-in check: the REFERENCE's own FEM/examples/src/heatMat.cpp and heatVec.cpp - compiled unmodified, where
// they lie, against dendro-kt_b200/include (oracle/build_dropin.sh) - run v = A u through feMatrix::matVec / b = M f through
// feVector::computeVec on the GPU.  usage: heat_dropin maxDepth dir   (dir: elem_xyz.bin, elem_lev.bin, u.bin -> v_mat.bin, v_vec.bin)
#define DKT_DEFINE_GLOBALS
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "heatMat.h"
#include "heatVec.h"

template <typename T>
static std::vector<T> readAll(const std::string &path)
{
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) { std::cerr << "cannot read " << path << "\n"; std::exit(2); }
  const size_t bytes = (size_t)f.tellg();
  std::vector<T> v(bytes / sizeof(T));
  f.seekg(0);
  f.read((char *)v.data(), bytes);
  return v;
}
static void writeAll(const std::string &path, const std::vector<double> &v)
{
  std::ofstream f(path, std::ios::binary);
  f.write((const char *)v.data(), v.size() * sizeof(double));
}

int main(int argc, char **argv)
{
  if (argc < 3) { std::cerr << "usage: heat_dropin maxDepth dir\n"; return 2; }
  m_uiMaxDepth = (unsigned)std::atoi(argv[1]);
  const std::string dir = argv[2];
  const std::vector<uint32_t> xyz = readAll<uint32_t>(dir + "/elem_xyz.bin");
  const std::vector<uint8_t> lev = readAll<uint8_t>(dir + "/elem_lev.bin");
  const std::vector<double> u = readAll<double>(dir + "/u.bin");
  std::vector<ot::TreeNode<unsigned, 3>> tree;
  for (size_t i = 0; i < lev.size(); i++)
    tree.push_back(ot::TreeNode<unsigned, 3>(1, std::array<unsigned, 3>{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, lev[i]));
  try
  {
    ot::DA<3> da(tree.data(), (unsigned)tree.size(), MPI_COMM_WORLD, 1);
    if (da.getTotalNodalSz() != u.size()) { std::cerr << "node count mismatch\n"; return 2; }
    std::vector<double> v(u.size(), 0.0);
    HeatEq::HeatMat<3> mat(&da, 1);
    mat.setProblemDimensions(Point<3>(-0.5, -0.5, -0.5), Point<3>(0.5, 0.5, 0.5));
    mat.matVec(u.data(), v.data(), 1.0);
    writeAll(dir + "/v_mat.bin", v);
    HeatEq::HeatVec<3> vec(&da, 1);
    vec.setProblemDimensions(Point<3>(-0.5, -0.5, -0.5), Point<3>(0.5, 0.5, 0.5));
    vec.computeVec(u.data(), v.data(), 1.0);
    writeAll(dir + "/v_vec.bin", v);
  }
  catch (const std::exception &e)
  {
    std::cerr << "failed: " << e.what() << "\n";
    return 3;
  }
  return 0;
}
