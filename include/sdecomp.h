/*
 * sdecomp.h -- interface of SimpleDecomp (NaokiHori/SimpleDecomp,
 * include/sdecomp.h:35-154 at commit 7e10fd0) as consumed by contour3d.h.
 *
 * When this project is embedded in a solver that already ships SimpleDecomp,
 * put the solver's own sdecomp.h first on the include path and link its
 * implementation; this file is only the stand-in for hosts that have neither
 * SimpleDecomp nor MPI (contour3d_b200/csrc/c3d_sdecomp_compat.c implements
 * it over compat/minimpi).  Type names, enumerator values and the member
 * order of the `sdecomp` method table are those of the original, so the two
 * are link-compatible.
 */
#if !defined(SDECOMP_H)
#define SDECOMP_H

#include <stddef.h>
#include <stdbool.h>
#include <stdint.h>
#include <mpi.h>

#if defined(__cplusplus)
extern "C" {
#endif

typedef struct sdecomp_info_t_ sdecomp_info_t;
typedef struct sdecomp_transpose_plan_t_ sdecomp_transpose_plan_t;

typedef uint_fast8_t sdecomp_dir_t;
extern const sdecomp_dir_t SDECOMP_XDIR; /* 0 */
extern const sdecomp_dir_t SDECOMP_YDIR; /* 1 */
extern const sdecomp_dir_t SDECOMP_ZDIR; /* 2 */

typedef uint_fast8_t sdecomp_pencil_t;
extern const sdecomp_pencil_t SDECOMP_X1PENCIL; /* 0 */
extern const sdecomp_pencil_t SDECOMP_Y1PENCIL; /* 1 */
extern const sdecomp_pencil_t SDECOMP_Z1PENCIL; /* 2 */
extern const sdecomp_pencil_t SDECOMP_X2PENCIL; /* 3 */
extern const sdecomp_pencil_t SDECOMP_Y2PENCIL; /* 4 */
extern const sdecomp_pencil_t SDECOMP_Z2PENCIL; /* 5 */

typedef struct {
  int (* const construct)(const sdecomp_info_t * info, const sdecomp_pencil_t pencil_bef, const sdecomp_pencil_t pencil_aft,
                          const size_t * glsizes, const size_t size_of_element, sdecomp_transpose_plan_t ** plan);
  int (* const execute)(sdecomp_transpose_plan_t * restrict plan, const void * restrict sendbuf, void * restrict recvbuf);
  int (* const destruct)(sdecomp_transpose_plan_t * plan);
} sdecomp_transpose_t;

typedef struct {
  int (* const construct)(const MPI_Comm comm_default, const size_t ndims, const size_t * dims, const bool * periods, sdecomp_info_t ** info);
  int (* const destruct)(sdecomp_info_t * info);
  int (* const get_ndims)(const sdecomp_info_t * info, size_t * ndims);
  int (* const get_comm_size)(const sdecomp_info_t * info, int * nprocs);
  int (* const get_comm_rank)(const sdecomp_info_t * info, int * myrank);
  int (* const get_nprocs)(const sdecomp_info_t * info, const sdecomp_pencil_t pencil, const sdecomp_dir_t dir, int * nprocs);
  int (* const get_myrank)(const sdecomp_info_t * info, const sdecomp_pencil_t pencil, const sdecomp_dir_t dir, int * myrank);
  int (* const get_neighbours)(const sdecomp_info_t * info, const sdecomp_pencil_t pencil, const sdecomp_dir_t dir, int neighbours[2]);
  int (* const get_pencil_mysize)(const sdecomp_info_t * info, const sdecomp_pencil_t pencil, const sdecomp_dir_t dir, const size_t glsize, size_t * mysize);
  int (* const get_pencil_offset)(const sdecomp_info_t * info, const sdecomp_pencil_t pencil, const sdecomp_dir_t dir, const size_t glsize, size_t * offset);
  int (* const get_comm_cart)(const sdecomp_info_t * info, MPI_Comm * comm);
  const sdecomp_transpose_t transpose;
} sdecomp_t;

extern const sdecomp_t sdecomp;

#if defined(__cplusplus)
}
#endif

#endif /* SDECOMP_H */
