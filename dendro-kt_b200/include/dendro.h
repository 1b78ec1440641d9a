// Host C++ API of the B200 matvec path: common definitions.
// Mirrors include/dendro.h + include/dtypes.h of the reference (DendroScalar = double, dendro.h:36).
#ifndef DKT_HOST_DENDRO_H
#define DKT_HOST_DENDRO_H

#include <stdexcept>
#include <string>

#include "../../include/dkt.h"

typedef double DendroScalar;
typedef DendroScalar VECType;  // include/oda.h:35

// The reference's API carries MPI communicators.  One process drives one GPU here; when the host
// application is not an MPI program a stand-in type keeps the signatures compiling.
#if !defined(MPI_VERSION) && !defined(DKT_HAVE_MPI)
typedef int MPI_Comm;
#ifndef MPI_COMM_WORLD
#define MPI_COMM_WORLD 0
#endif
#ifndef MPI_COMM_NULL
#define MPI_COMM_NULL (-1)
#endif
#endif

extern unsigned int m_uiMaxDepth;  // include/treeNode.h:17 (defined once via DKT_DEFINE_GLOBALS)
#ifdef DKT_DEFINE_GLOBALS
unsigned int m_uiMaxDepth = 30;    // src/treeNode.cpp:4
#endif

namespace dkt_host
{
inline void check(int rc, const char *what)
{
  if (rc != DKT_OK) throw std::runtime_error(std::string(what) + ": " + dkt_last_error());
}
} // namespace dkt_host
#endif
