/*
 * c3d_sdecomp_compat.c -- minimal stand-in for SimpleDecomp (only what
 * Contour3D consumes: construct/destruct and the getters), for hosts that do
 * not ship the real library.  Built into libc3d_sdecomp_compat.so, NOT into
 * libcontour3d_b200.so: an embedding solver links its own SimpleDecomp.
 *
 * Behaviour follows NaokiHori/SimpleDecomp @ 7e10fd0:
 *   construct  -> MPI_Dims_create with x fixed to 1 + MPI_Cart_create  (sdecomp/main.c:42-162)
 *   pencil/dir -> Cartesian dimension                                  (sdecomp/get.c:30-57)
 *   mysize     =  (glsize + myrank) / nprocs, offset = prefix sum      (sdecomp/kernel.c:89,111-116)
 *   neighbours -> MPI_Cart_shift by one                                (sdecomp/get.c:163-181)
 * Pencil transposes are not provided (never used by Contour3D).
 */
#include <stdio.h>
#include <stdlib.h>
#include "sdecomp.h"

struct sdecomp_info_t_ {
  MPI_Comm comm_cart;
  size_t ndims;
};

const sdecomp_dir_t SDECOMP_XDIR = 0;
const sdecomp_dir_t SDECOMP_YDIR = 1;
const sdecomp_dir_t SDECOMP_ZDIR = 2;
const sdecomp_pencil_t SDECOMP_X1PENCIL = 0;
const sdecomp_pencil_t SDECOMP_Y1PENCIL = 1;
const sdecomp_pencil_t SDECOMP_Z1PENCIL = 2;
const sdecomp_pencil_t SDECOMP_X2PENCIL = 3;
const sdecomp_pencil_t SDECOMP_Y2PENCIL = 4;
const sdecomp_pencil_t SDECOMP_Z2PENCIL = 5;

#define COMPAT_ERROR(...) do { fprintf(stderr, "[SDECOMP ERROR (compat)] " __VA_ARGS__); fflush(stderr); } while (0)

/* which dimension of the x1-pencil process grid a (pencil, direction) pair refers to */
static int cart_dim(const size_t ndims, const sdecomp_pencil_t pencil, const sdecomp_dir_t dir, int *out) {
  static const int map2[2][2] = {{0, 1}, {1, 0}};
  static const int map3[6][3] = {{0, 1, 2}, {1, 0, 2}, {1, 2, 0}, {0, 2, 1}, {2, 0, 1}, {2, 1, 0}};
  if (2 == ndims) {
    if (pencil > 1 || dir > 1) return 1;
    *out = map2[pencil][dir];
    return 0;
  }
  if (pencil > 5 || dir > 2) return 1;
  *out = map3[pencil][dir];
  return 0;
}

static int grid_of(const sdecomp_info_t *info, int dims[3], int coords[3]) {
  int periods[3] = {0, 0, 0};
  return MPI_SUCCESS == MPI_Cart_get(info->comm_cart, (int)info->ndims, dims, periods, coords) ? 0 : 1;
}

static int construct(const MPI_Comm comm_default, const size_t ndims, const size_t *dims, const bool *periods, sdecomp_info_t **info) {
  if (NULL == info) return 1;
  *info = NULL;
  if (NULL == dims || NULL == periods) { COMPAT_ERROR("dims / periods is NULL\n"); return 1; }
  if (2 != ndims && 3 != ndims) { COMPAT_ERROR("ndims (%zu) should be 2 or 3\n", ndims); return 1; }
  int nprocs = 0;
  MPI_Comm_size(comm_default, &nprocs);
  int user = 0;
  for (size_t d = 0; d < ndims; d++) if (0 != dims[d]) user = 1;
  int idims[3] = {0, 0, 0}, iperiods[3] = {0, 0, 0};
  if (user) {
    if (1 != dims[0]) { COMPAT_ERROR("dims[0] should be 1 (%zu is given)\n", dims[0]); return 1; }
    size_t prod = 1;
    for (size_t d = 0; d < ndims; d++) { idims[d] = (int)dims[d]; prod *= dims[d]; }
    if ((size_t)nprocs != prod) { COMPAT_ERROR("number of processes (%d) and product of dims (%zu) do not match\n", nprocs, prod); return 1; }
  } else {
    idims[0] = 1;
    MPI_Dims_create(nprocs, (int)ndims, idims);
  }
  for (size_t d = 0; d < ndims; d++) iperiods[d] = periods[d] ? 1 : 0;
  MPI_Comm cart = MPI_COMM_NULL;
  if (MPI_SUCCESS != MPI_Cart_create(comm_default, (int)ndims, idims, iperiods, 1, &cart)) return 1;
  sdecomp_info_t *made = calloc(1, sizeof(*made));
  if (NULL == made) return 1;
  made->comm_cart = cart;
  made->ndims = ndims;
  *info = made;
  return 0;
}

static int destruct(sdecomp_info_t *info) {
  if (NULL == info) return 1;
  MPI_Comm_free(&info->comm_cart);
  free(info);
  return 0;
}

static int get_ndims(const sdecomp_info_t *info, size_t *ndims) {
  if (NULL == info || NULL == ndims) return 1;
  *ndims = info->ndims;
  return 0;
}

static int get_comm_size(const sdecomp_info_t *info, int *nprocs) {
  if (NULL == info || NULL == nprocs) return 1;
  return MPI_SUCCESS == MPI_Comm_size(info->comm_cart, nprocs) ? 0 : 1;
}

static int get_comm_rank(const sdecomp_info_t *info, int *myrank) {
  if (NULL == info || NULL == myrank) return 1;
  return MPI_SUCCESS == MPI_Comm_rank(info->comm_cart, myrank) ? 0 : 1;
}

static int get_comm_cart(const sdecomp_info_t *info, MPI_Comm *comm) {
  if (NULL == info || NULL == comm) return 1;
  *comm = info->comm_cart;
  return 0;
}

static int get_nprocs(const sdecomp_info_t *info, const sdecomp_pencil_t pencil, const sdecomp_dir_t dir, int *nprocs) {
  int dims[3], coords[3], dim = 0;
  if (NULL == info || NULL == nprocs || 0 != cart_dim(info->ndims, pencil, dir, &dim) || 0 != grid_of(info, dims, coords)) return 1;
  *nprocs = dims[dim];
  return 0;
}

static int get_myrank(const sdecomp_info_t *info, const sdecomp_pencil_t pencil, const sdecomp_dir_t dir, int *myrank) {
  int dims[3], coords[3], dim = 0;
  if (NULL == info || NULL == myrank || 0 != cart_dim(info->ndims, pencil, dir, &dim) || 0 != grid_of(info, dims, coords)) return 1;
  *myrank = coords[dim];
  return 0;
}

static int get_neighbours(const sdecomp_info_t *info, const sdecomp_pencil_t pencil, const sdecomp_dir_t dir, int neighbours[2]) {
  int dim = 0;
  if (NULL == info || NULL == neighbours || 0 != cart_dim(info->ndims, pencil, dir, &dim)) return 1;
  neighbours[0] = MPI_PROC_NULL;
  neighbours[1] = MPI_PROC_NULL;
  return MPI_SUCCESS == MPI_Cart_shift(info->comm_cart, dim, 1, neighbours + 0, neighbours + 1) ? 0 : 1;
}

static int get_pencil_mysize(const sdecomp_info_t *info, const sdecomp_pencil_t pencil, const sdecomp_dir_t dir, const size_t glsize, size_t *mysize) {
  int nprocs = 0, myrank = 0;
  if (NULL == mysize || 0 != get_nprocs(info, pencil, dir, &nprocs) || 0 != get_myrank(info, pencil, dir, &myrank)) return 1;
  if (glsize < (size_t)nprocs) { COMPAT_ERROR("glsize (%zu) should be equal to or more than nprocs (%d)\n", glsize, nprocs); return 1; }
  *mysize = (glsize + (size_t)myrank) / (size_t)nprocs;
  return 0;
}

static int get_pencil_offset(const sdecomp_info_t *info, const sdecomp_pencil_t pencil, const sdecomp_dir_t dir, const size_t glsize, size_t *offset) {
  int nprocs = 0, myrank = 0;
  if (NULL == offset || 0 != get_nprocs(info, pencil, dir, &nprocs) || 0 != get_myrank(info, pencil, dir, &myrank)) return 1;
  if (glsize < (size_t)nprocs) { COMPAT_ERROR("glsize (%zu) should be equal to or more than nprocs (%d)\n", glsize, nprocs); return 1; }
  size_t sum = 0;
  for (int r = 0; r < myrank; r++) sum += (glsize + (size_t)r) / (size_t)nprocs;
  *offset = sum;
  return 0;
}

static int transpose_construct(const sdecomp_info_t *info, const sdecomp_pencil_t bef, const sdecomp_pencil_t aft, const size_t *glsizes,
                               const size_t size_of_element, sdecomp_transpose_plan_t **plan) {
  (void)info; (void)bef; (void)aft; (void)glsizes; (void)size_of_element;
  if (plan) *plan = NULL;
  COMPAT_ERROR("pencil transposes are not part of this stand-in\n");
  return 1;
}
static int transpose_execute(sdecomp_transpose_plan_t *restrict plan, const void *restrict sendbuf, void *restrict recvbuf) {
  (void)plan; (void)sendbuf; (void)recvbuf;
  return 1;
}
static int transpose_destruct(sdecomp_transpose_plan_t *plan) {
  (void)plan;
  return 1;
}

const sdecomp_t sdecomp = {
  .construct = construct,
  .destruct = destruct,
  .get_ndims = get_ndims,
  .get_comm_size = get_comm_size,
  .get_comm_rank = get_comm_rank,
  .get_nprocs = get_nprocs,
  .get_myrank = get_myrank,
  .get_neighbours = get_neighbours,
  .get_pencil_mysize = get_pencil_mysize,
  .get_pencil_offset = get_pencil_offset,
  .get_comm_cart = get_comm_cart,
  .transpose = {
    .construct = transpose_construct,
    .execute = transpose_execute,
    .destruct = transpose_destruct,
  },
};
