// Host C++ API of the B200 matvec path: common definitions.
// Mirrors include/dendro.h + include/dtypes.h of the reference (DendroScalar = double, dendro.h:36).
#ifndef DKT_HOST_DENDRO_H
#define DKT_HOST_DENDRO_H

#include <cstdio>
#include <cstdlib>
#include <ctime>
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

// ---- ranks -------------------------------------------------------------------------------------------------------------
// One process per GPU.  With MPI (the application includes <mpi.h> first, or -DDKT_HAVE_MPI) rank, size and the broadcast
// of the NCCL id come from the communicator.  Without MPI (this image has none) they come from the launcher's environment
// - DKT_RANK/DKT_NRANKS, else RANK/WORLD_SIZE (torchrun), OMPI_COMM_WORLD_*, PMI_* - and the id travels through a file:
// DKT_NCCL_ID_FILE=<path on a file system all ranks see>; every distributed DA of a process takes the next sequence number,
// so ranks must construct their DAs in the same order (they do in the reference as well: construction is collective).
inline int env_int(const char *const *names, int dflt)
{
  for (; *names; names++)
    if (const char *v = std::getenv(*names)) return std::atoi(v);
  return dflt;
}
inline void comm_rank_size(MPI_Comm comm, int &rank, int &size)
{
#if defined(MPI_VERSION) || defined(DKT_HAVE_MPI)
  MPI_Comm_rank(comm, &rank);
  MPI_Comm_size(comm, &size);
#else
  (void)comm;
  static const char *const r[] = {"DKT_RANK", "RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", nullptr};
  static const char *const n[] = {"DKT_NRANKS", "WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", nullptr};
  rank = env_int(r, 0);
  size = env_int(n, 1);
#endif
}
inline void share_nccl_id(MPI_Comm comm, int rank, char id[128])
{
#if defined(MPI_VERSION) || defined(DKT_HAVE_MPI)
  if (rank == 0) check(dkt_nccl_unique_id(id), "dkt_nccl_unique_id");
  MPI_Bcast(id, 128, MPI_BYTE, 0, comm);
#else
  (void)comm;
  static int seq = 0;
  const char *base = std::getenv("DKT_NCCL_ID_FILE");
  if (!base) throw std::runtime_error("distributed ot::DA without MPI: set DKT_NCCL_ID_FILE to a path all ranks can read and write");
  const std::string path = std::string(base) + "." + std::to_string(seq++);
  if (rank == 0)
  {
    check(dkt_nccl_unique_id(id), "dkt_nccl_unique_id");
    const std::string tmp = path + ".tmp";
    FILE *f = std::fopen(tmp.c_str(), "wb");
    if (!f || std::fwrite(id, 1, 128, f) != 128) throw std::runtime_error("cannot write " + tmp);
    std::fclose(f);
    if (std::rename(tmp.c_str(), path.c_str()) != 0) throw std::runtime_error("cannot publish " + path);
    return;
  }
  for (int tries = 0; tries < 6000; tries++)  // up to ~60 s
  {
    if (FILE *f = std::fopen(path.c_str(), "rb"))
    {
      const size_t got = std::fread(id, 1, 128, f);
      std::fclose(f);
      if (got == 128) return;
    }
    struct timespec ts = {0, 10 * 1000 * 1000};
    nanosleep(&ts, nullptr);
  }
  throw std::runtime_error("timed out waiting for the NCCL id in " + path);
#endif
}
} // namespace dkt_host
#endif
