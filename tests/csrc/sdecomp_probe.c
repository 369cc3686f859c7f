/*
 * sdecomp_probe.c -- prints what the seven SimpleDecomp getters the pipeline uses return, for every rank, every pencil
 * (X1 .. Z2) and every direction.  Test infrastructure: built twice by tests/test_host_logic.py -- against the genuine
 * SimpleDecomp inside oracle/_ref/libc3dref.so and against the stand-in contour3d_b200/libc3d_sdecomp_compat.so -- and
 * run under the fork-mode MPI stand-in; the outputs must agree line for line.
 *
 *   MINIMPI_NP=<ranks> sdecomp_probe dim_x dim_y dim_z n_x n_y n_z out_prefix
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdbool.h>
#include <mpi.h>
#include "sdecomp.h"

int main(int argc, char **argv) {
  if (argc < 8) return 2;
  MPI_Init(NULL, NULL);
  int rank = 0;
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  size_t dims[3] = {(size_t)atol(argv[1]), (size_t)atol(argv[2]), (size_t)atol(argv[3])};
  const size_t gl[3] = {(size_t)atol(argv[4]), (size_t)atol(argv[5]), (size_t)atol(argv[6])};
  sdecomp_info_t *info = NULL;
  if (0 != sdecomp.construct(MPI_COMM_WORLD, 3, dims, (bool[3]){true, false, true}, &info)) return 1;
  char name[512];
  snprintf(name, sizeof(name), "%s.%d", argv[7], rank);
  FILE *fp = fopen(name, "w");
  if (NULL == fp) return 1;
  int size = 0, me = 0;
  sdecomp.get_comm_size(info, &size);
  sdecomp.get_comm_rank(info, &me);
  fprintf(fp, "size %d rank %d\n", size, me);
  for (sdecomp_pencil_t pencil = 0; pencil < 6; pencil++) {
    for (sdecomp_dir_t dir = 0; dir < 3; dir++) {
      int nprocs = -1, myrank = -1, nb[2] = {-7, -7};
      size_t mysize = 0, offset = 0;
      const int a = sdecomp.get_nprocs(info, pencil, dir, &nprocs);
      const int b = sdecomp.get_myrank(info, pencil, dir, &myrank);
      const int c = sdecomp.get_neighbours(info, pencil, dir, nb);
      const int d = sdecomp.get_pencil_mysize(info, pencil, dir, gl[dir], &mysize);
      const int e = sdecomp.get_pencil_offset(info, pencil, dir, gl[dir], &offset);
      fprintf(fp, "pencil %d dir %d: rc %d%d%d%d%d nprocs %d myrank %d neighbours %d %d mysize %zu offset %zu\n",
              (int)pencil, (int)dir, a, b, c, d, e, nprocs, myrank, nb[0], nb[1], mysize, offset);
    }
  }
  fclose(fp);
  sdecomp.destruct(info);
  MPI_Finalize();
  return 0;
}
