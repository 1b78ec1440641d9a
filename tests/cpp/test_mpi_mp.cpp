// Unit test of the multi-process MPI stand-in oracle/shim_mp (test infrastructure of the multi-rank reference oracle): four fork()ed
// ranks check collectives, point-to-point, communicator splitting/creation against their closed-form results.
#include "mpi.h"
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>
#include <sys/wait.h>
#include <vector>
int main()
{
  const int P = 4;
  dktmp_world_create(P, 1 << 26);
  for (int r = 0; r < P; r++)
    if (fork() == 0)
    {
      dktmp_set_rank(r);
      int rank, size; MPI_Comm_rank(MPI_COMM_WORLD, &rank); MPI_Comm_size(MPI_COMM_WORLD, &size);
      long long v = rank + 1, s = 0; MPI_Allreduce(&v, &s, 1, MPI_LONG_LONG_INT, MPI_SUM, MPI_COMM_WORLD);
      long long sc = 0; MPI_Scan(&v, &sc, 1, MPI_LONG_LONG_INT, MPI_SUM, MPI_COMM_WORLD);
      double d = rank * 1.5, dm = 0; MPI_Allreduce(&d, &dm, 1, MPI_DOUBLE, MPI_MAX, MPI_COMM_WORLD);
      std::vector<int> sc_(P), sd(P), rc(P), rd(P); std::vector<int> sb, rb;
      for (int q = 0; q < P; q++) { sc_[q] = rank + q; sd[q] = (int)sb.size(); for (int k = 0; k < sc_[q]; k++) sb.push_back(rank * 100 + q); }
      MPI_Alltoall(sc_.data(), 1, MPI_INT, rc.data(), 1, MPI_INT, MPI_COMM_WORLD);
      int tot = 0; for (int q = 0; q < P; q++) { rd[q] = tot; tot += rc[q]; } rb.resize(tot);
      MPI_Alltoallv(sb.data(), sc_.data(), sd.data(), MPI_INT, rb.data(), rc.data(), rd.data(), MPI_INT, MPI_COMM_WORLD);
      bool ok = s == P * (P + 1) / 2 && sc == (rank + 1) * (rank + 2) / 2 && dm == (P - 1) * 1.5;
      for (int q = 0; q < P; q++) { ok &= rc[q] == q + rank; for (int k = 0; k < rc[q]; k++) ok &= rb[rd[q] + k] == q * 100 + rank; }
      MPI_Comm half; MPI_Comm_split(MPI_COMM_WORLD, rank % 2, rank, &half);
      int hr, hs; MPI_Comm_rank(half, &hr); MPI_Comm_size(half, &hs);
      long long hsum = 0; MPI_Allreduce(&v, &hsum, 1, MPI_LONG_LONG_INT, MPI_SUM, half);
      ok &= hs == P / 2 && hr == rank / 2 && hsum == (rank % 2 ? 2 + 4 : 1 + 3);
      int nb = (rank + 1) % P, pv = (rank + P - 1) % P, got = -1;
      MPI_Sendrecv(&rank, 1, MPI_INT, nb, 7, &got, 1, MPI_INT, pv, 7, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
      ok &= got == pv;
      MPI_Request rq[2]; int a = rank * 3, b = -1; MPI_Irecv(&b, 1, MPI_INT, pv, 9, MPI_COMM_WORLD, &rq[0]); MPI_Isend(&a, 1, MPI_INT, nb, 9, MPI_COMM_WORLD, &rq[1]);
      MPI_Waitall(2, rq, MPI_STATUSES_IGNORE); ok &= b == pv * 3;
      MPI_Group g, g2; MPI_Comm_group(MPI_COMM_WORLD, &g); int inc[2] = {1, 2}; MPI_Group_incl(g, 2, inc, &g2);
      MPI_Comm sub; MPI_Comm_create(MPI_COMM_WORLD, g2, &sub);
      if (rank == 1 || rank == 2) { int sr; MPI_Comm_rank(sub, &sr); ok &= sr == rank - 1; long long z = 0; MPI_Allreduce(&v, &z, 1, MPI_LONG_LONG_INT, MPI_SUM, sub); ok &= z == 5; } else ok &= sub == MPI_COMM_NULL;
      std::vector<int> ag(P); MPI_Allgather(&rank, 1, MPI_INT, ag.data(), 1, MPI_INT, MPI_COMM_WORLD); for (int q = 0; q < P; q++) ok &= ag[q] == q;
      MPI_Barrier(MPI_COMM_WORLD);
      printf("rank %d %s\n", rank, ok ? "OK" : "FAIL");
      fflush(stdout);
      _exit(ok ? 0 : 1);
    }
  int bad = 0; for (int r = 0; r < P; r++) { int st; wait(&st); bad |= st; }
  return bad != 0;
}
