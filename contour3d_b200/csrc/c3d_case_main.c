/*
 * c3d_case_main.c -- a small MPI program around the PUBLIC API only
 * (contour3d_execute + SimpleDecomp getters): renders a "case file".
 *
 *   [MINIMPI_NP=<ranks>] c3d_case <case.bin> <out.ppm> [dim_y dim_z] [repeat] [warmup] [flags]
 *
 *   dim_y dim_z = 0 0 leaves the process grid to SimpleDecomp (MPI_Dims_create, sdecomp/main.c:133-139);
 *   flags: 1 = construct the decomposition periodic in all directions, as the reference's demo does (src/main.c:119-125);
 *          2 = hand the arrays over as X2 pencils (y split by the third, z by the second grid dimension, sdecomp/get.c:44-57)
 *
 * Because it touches nothing but include/contour3d.h it is built twice from
 * this one source: against libcontour3d_b200.so (-> contour3d_b200/c3d_case,
 * the drop-in demonstration) and, by oracle/Makefile, against the UNMODIFIED
 * reference sources (-> oracle/_ref/ref_driver, the CPU baseline).
 *
 * The case file (contour3d_b200/casefile.py) holds the scene, the grids, the
 * contour / line objects, one or more global fields (fp32 or fp64, x
 * fastest) and the coordinate converter every object uses (0 identity,
 * 1 cylindrical, 2 stretched, 3 twisted: built-in functions below).  Every rank reads its own X1 pencil, widening fp32 to double
 * exactly, fills the public structs and calls contour3d_execute
 * (include/contour3d.h:58-86).  Rank 0 prints one JSON line with the wall
 * time of contour3d_execute (max over ranks, field I/O excluded).
 */
#if !defined(C3D_IMPL_NAME)
#define C3D_IMPL_NAME "contour3d_b200"
#endif
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <stdbool.h>
#include <time.h>
#include <mpi.h>
#include "sdecomp.h"
#include "contour3d.h"

#include <math.h>

/* coordinate converters a case file can ask for (header byte pad[0]); the same functions run inside the unmodified
   reference (ref_driver) and behind the drop-in (c3d_case), which tabulates them for the device */
static contour3d_vector_t identity(const contour3d_vector_t o) { return o; }
/* 1: annular sector, (r, theta, z) -> Cartesian: depends on two directions in two components */
static contour3d_vector_t cylindrical(const contour3d_vector_t o) {
  const double r = 0.75 + 0.5 * o.x, theta = 1.5 * o.y;
  const contour3d_vector_t c = {r * cos(theta) - 0.75, r * sin(theta), o.z};
  return c;
}
/* 2: stretched grid, separable */
static contour3d_vector_t stretched(const contour3d_vector_t o) {
  const contour3d_vector_t c = {o.x + 0.3 * o.x * o.x * o.x, 1.2 * o.y, o.z - 0.2 * o.z * o.z};
  return c;
}
/* 3: twisted column: every component depends on all three directions */
static contour3d_vector_t twisted(const contour3d_vector_t o) {
  const double a = 0.8 * o.z;
  const contour3d_vector_t c = {o.x * cos(a) - o.y * sin(a) + 0.02 * o.z * o.y, o.x * sin(a) + o.y * cos(a), o.z + 0.05 * sin(3. * o.x) * o.y};
  return c;
}
typedef contour3d_vector_t (*converter_t)(const contour3d_vector_t);
static converter_t converter_of(const int kind) {
  return 1 == kind ? cylindrical : (2 == kind ? stretched : (3 == kind ? twisted : identity));
}

typedef struct {
  char magic[8];
  uint64_t n[3], width, height, ncontours, nlines, elem_bytes, nfields;
  double camera[3], look_at[3], light[3], center[3], local_x[3], local_y[3];
  uint8_t bg[3], pad[5];
} case_header_t;

typedef struct { uint64_t field; double threshold; uint8_t color[3], pad[5]; } case_contour_t;
typedef struct { double edges[6]; uint64_t nitems; double width; uint8_t color[3], pad[5]; } case_line_t;

static double now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1.e-9 * (double)ts.tv_nsec;
}

int main(int argc, char **argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s case.bin out.ppm [dim_y dim_z] [repeat] [warmup] [flags]\n", argv[0]);
    return 2;
  }
  MPI_Init(NULL, NULL);
  int nranks = 1, rank = 0;
  MPI_Comm_size(MPI_COMM_WORLD, &nranks);
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  size_t dims[3] = {0, 0, 0};
  if (argc >= 5 && (0 != atol(argv[3]) || 0 != atol(argv[4]))) {
    dims[0] = 1;
    dims[1] = (size_t)atol(argv[3]);
    dims[2] = (size_t)atol(argv[4]);
  }
  const int repeat = argc >= 6 ? atoi(argv[5]) : 1;
  const int warmup = argc >= 7 ? atoi(argv[6]) : 0;      /* untimed calls before the `repeat` timed ones */
  const int flags = argc >= 8 ? atoi(argv[7]) : 0;
  const bool periodic = 0 != (flags & 1);
  const sdecomp_pencil_t pencil = 0 != (flags & 2) ? SDECOMP_X2PENCIL : SDECOMP_X1PENCIL;
  sdecomp_info_t *info = NULL;
  if (0 != sdecomp.construct(MPI_COMM_WORLD, 3, dims, (bool[3]){periodic, periodic, periodic}, &info)) return 1;
  FILE *fp = fopen(argv[1], "rb");
  if (NULL == fp) { perror(argv[1]); return 1; }
  case_header_t h;
  if (1 != fread(&h, sizeof(h), 1, fp) || 0 != memcmp(h.magic, "C3DCASE1", 8)) {
    fprintf(stderr, "bad case file\n");
    return 1;
  }
  double *grids[3];
  for (int d = 0; d < 3; d++) {
    grids[d] = malloc(h.n[d] * sizeof(double));
    if (h.n[d] != fread(grids[d], sizeof(double), h.n[d], fp)) return 1;
  }
  case_contour_t *cc = malloc((h.ncontours ? h.ncontours : 1) * sizeof(*cc));
  case_line_t *cl = malloc((h.nlines ? h.nlines : 1) * sizeof(*cl));
  if (h.ncontours != fread(cc, sizeof(*cc), h.ncontours, fp)) return 1;
  if (h.nlines != fread(cl, sizeof(*cl), h.nlines, fp)) return 1;
  const long fields_at = ftell(fp);
  /* my x pencil */
  size_t mysizes[3], offsets[3];
  for (sdecomp_dir_t d = 0; d < 3; d++) {
    sdecomp.get_pencil_mysize(info, pencil, d, h.n[d], mysizes + d);
    sdecomp.get_pencil_offset(info, pencil, d, h.n[d], offsets + d);
  }
  const size_t nlocal = mysizes[0] * mysizes[1] * mysizes[2];
  double **fields = malloc(h.nfields * sizeof(double *));
  void *row = malloc(mysizes[0] * h.elem_bytes);
  for (uint64_t f = 0; f < h.nfields; f++) {
    fields[f] = malloc(nlocal * sizeof(double));
    for (size_t k = 0; k < mysizes[2]; k++) {
      for (size_t j = 0; j < mysizes[1]; j++) {
        const uint64_t gidx = ((f * h.n[2] + k + offsets[2]) * h.n[1] + j + offsets[1]) * h.n[0] + offsets[0];
        fseek(fp, fields_at + (long)(gidx * h.elem_bytes), SEEK_SET);
        if (mysizes[0] != fread(row, h.elem_bytes, mysizes[0], fp)) return 1;
        double *dst = fields[f] + (k * mysizes[1] + j) * mysizes[0];
        for (size_t i = 0; i < mysizes[0]; i++) {
          dst[i] = 4 == h.elem_bytes ? (double)((float *)row)[i] : ((double *)row)[i];
        }
      }
    }
  }
  fclose(fp);
  contour3d_contour_obj_t *cobjs = calloc(h.ncontours ? h.ncontours : 1, sizeof(*cobjs));
  for (uint64_t n = 0; n < h.ncontours; n++) {
    cobjs[n].pencil = pencil;
    for (int d = 0; d < 3; d++) { cobjs[n].glsizes[d] = h.n[d]; cobjs[n].grids[d] = grids[d]; }
    cobjs[n].converter = converter_of(h.pad[0]);
    cobjs[n].threshold = cc[n].threshold;
    cobjs[n].color.r = cc[n].color[0]; cobjs[n].color.g = cc[n].color[1]; cobjs[n].color.b = cc[n].color[2];
    cobjs[n].array = fields[cc[n].field];
  }
  contour3d_line_obj_t *lobjs = calloc(h.nlines ? h.nlines : 1, sizeof(*lobjs));
  for (uint64_t n = 0; n < h.nlines; n++) {
    for (int e = 0; e < 2; e++) {
      lobjs[n].edges[e].x = cl[n].edges[3 * e + 0];
      lobjs[n].edges[e].y = cl[n].edges[3 * e + 1];
      lobjs[n].edges[e].z = cl[n].edges[3 * e + 2];
    }
    lobjs[n].nitems = cl[n].nitems;
    lobjs[n].converter = converter_of(h.pad[0]);
    lobjs[n].color.r = cl[n].color[0]; lobjs[n].color.g = cl[n].color[1]; lobjs[n].color.b = cl[n].color[2];
    lobjs[n].width = cl[n].width;
  }
  const contour3d_vector_t camera = {h.camera[0], h.camera[1], h.camera[2]};
  const contour3d_vector_t look_at = {h.look_at[0], h.look_at[1], h.look_at[2]};
  const contour3d_vector_t light = {h.light[0], h.light[1], h.light[2]};
  const contour3d_vector_t center = {h.center[0], h.center[1], h.center[2]};
  const contour3d_vector_t local[2] = {
    {h.local_x[0], h.local_x[1], h.local_x[2]},
    {h.local_y[0], h.local_y[1], h.local_y[2]},
  };
  const contour3d_color_t bg = {h.bg[0], h.bg[1], h.bg[2]};
  const size_t screen[2] = {h.width, h.height};
  double best = 1.e300, sum = 0.;
  int status = 0;
  for (int it = -warmup; it < repeat; it++) {
    MPI_Barrier(MPI_COMM_WORLD);
    const double t0 = now();
    status |= contour3d_execute(info, &camera, &look_at, &light, screen, &center, local, &bg,
                                h.ncontours, cobjs, h.nlines, lobjs, argv[2]);
    double dt = now() - t0;
    MPI_Allreduce(MPI_IN_PLACE, &dt, 1, MPI_DOUBLE, MPI_MAX, MPI_COMM_WORLD);
    if (it < 0) continue;
    if (dt < best) best = dt;
    sum += dt;
  }
  if (0 == rank) {
    printf("{\"impl\": \"" C3D_IMPL_NAME "\", \"ranks\": %d, \"status\": %d, \"repeat\": %d, \"seconds_best\": %.6f, \"seconds_mean\": %.6f, "
           "\"voxels\": %llu, \"contours\": %llu}\n",
           nranks, status, repeat, best, sum / repeat,
           (unsigned long long)(h.n[0] * h.n[1] * h.n[2]), (unsigned long long)h.ncontours);
    fflush(stdout);
  }
  sdecomp.destruct(info);
  MPI_Finalize();
  return status;
}
