/*
 * contour3d_execute.c -- the drop-in entry point (include/contour3d.h),
 * host C on top of the CUDA C-ABI (include/contour3d_b200.h).
 *
 * Replaces src/contour3d/main.c:68-146 of the reference together with the
 * per-object driver contour/main.c:11-119 and the host half of
 * contour/extend.c:150-282 (sizes, offsets, neighbours from the sdecomp
 * getters).  Same argument meaning, same 0/1 return convention, same
 * "[CONTOUR3D ERROR] ..." lines on stderr (logger.c:5-18).
 *
 * Differences a caller can observe (all documented in DESIGN.md):
 *  - `array` may be a host or a device pointer;
 *  - `converter` must be the identity map (NULL is accepted as identity);
 *    anything else is rejected, there is no host fallback;
 *  - arrays on Y/Z pencils (not x-fastest; the reference mis-indexes them,
 *    extend.c:199-208) and local extents too small to tessellate are rejected;
 *  - a process-wide context (stream, NCCL communicator, canvases) is created
 *    on the first call and released at exit.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#include <string.h>
#include <mpi.h>
#include "sdecomp.h"
#include "contour3d.h"
#include "contour3d_b200.h"

#define MAX_OBJECTS 16

static c3d_context_t *g_ctx = NULL;
static int g_rank = -1, g_size = -1;

/* Device arrays kept between calls.  A solver calls contour3d_execute() once per output step with arrays of
   the same shape; allocating and freeing gigabytes of device memory each time costs more than the frame.
   The contents are never reused (every call uploads afresh), so nothing the caller can observe is retained.
   CONTOUR3D_NO_CACHE=1 restores allocate-per-call. */
typedef struct {
  c3d_field_t *field;
  size_t glsizes[CONTOUR3D_NDIMS];
  c3d_decomp_t decomp;
  int in_use;
} cached_field_t;
static cached_field_t g_cache[MAX_OBJECTS];

static int log_error(const char *fmt, ...) {
  va_list args;
  va_start(args, fmt);
  fprintf(stderr, "[CONTOUR3D ERROR] ");
  vfprintf(stderr, fmt, args);
  fprintf(stderr, "\n");
  fflush(stderr);
  va_end(args);
  return 1;
}

static void release_context(void) {
  for (size_t n = 0; n < MAX_OBJECTS; n++) {
    if (NULL != g_cache[n].field) c3d_field_destroy(g_ctx, g_cache[n].field);
    memset(&g_cache[n], 0, sizeof(g_cache[n]));
  }
  if (NULL != g_ctx) {
    c3d_context_destroy(g_ctx);
    g_ctx = NULL;
  }
}

/* one communicator per process is assumed, like the reference's use of MPI_COMM_WORLD (output.c:82) */
static int acquire_context(const sdecomp_info_t *info) {
  int rank = 0, size = 1;
  if (0 != sdecomp.get_comm_rank(info, &rank)) return log_error("sdecomp.get_comm_rank failed");
  if (0 != sdecomp.get_comm_size(info, &size)) return log_error("sdecomp.get_comm_size failed");
  if (NULL != g_ctx && rank == g_rank && size == g_size) return 0;
  release_context();
  unsigned char id[C3D_NCCL_ID_BYTES];
  memset(id, 0, sizeof(id));
  if (size > 1) {
    MPI_Comm comm = MPI_COMM_NULL;
    if (0 != sdecomp.get_comm_cart(info, &comm)) return log_error("sdecomp.get_comm_cart failed");
    int failed = 0;
    if (0 == rank) failed = c3d_nccl_unique_id(id);
    MPI_Bcast(&failed, 1, MPI_INT, 0, comm);
    if (failed) return log_error("could not create an NCCL id");
    MPI_Bcast(id, C3D_NCCL_ID_BYTES, MPI_BYTE, 0, comm);
  }
  int device = -1;
  const char *env = getenv("CONTOUR3D_DEVICE");
  if (NULL != env && '\0' != *env) device = atoi(env);
  else if (NULL != (env = getenv("LOCAL_RANK")) && '\0' != *env) device = atoi(env);
  if (0 != c3d_context_create(&g_ctx, device, rank, size, id)) return log_error("device context creation failed");
  if (g_rank < 0) atexit(release_context);
  g_rank = rank;
  g_size = size;
  return 0;
}

static int cache_enabled(void) {
  const char *env = getenv("CONTOUR3D_NO_CACHE");
  return !(NULL != env && '\0' != *env && '0' != *env);
}

/* a device array of this shape: from the cache when one is free, else new (and cached when there is room) */
static int acquire_field(const c3d_decomp_t *decomp, const size_t glsizes[CONTOUR3D_NDIMS], c3d_field_t **field, int *cached) {
  *cached = 0;
  if (cache_enabled()) {
    for (size_t n = 0; n < MAX_OBJECTS; n++) {
      cached_field_t *c = &g_cache[n];
      if (NULL != c->field && !c->in_use && 0 == memcmp(c->glsizes, glsizes, sizeof(c->glsizes)) && 0 == memcmp(&c->decomp, decomp, sizeof(*decomp))) {
        c->in_use = 1;
        *field = c->field;
        *cached = 1;
        return 0;
      }
    }
  }
  if (0 != c3d_field_create(g_ctx, decomp, glsizes, (int)sizeof(double), field)) {
    /* out of device memory? drop what the cache holds and try once more */
    int dropped = 0;
    for (size_t n = 0; n < MAX_OBJECTS; n++) {
      if (NULL != g_cache[n].field && !g_cache[n].in_use) {
        c3d_field_destroy(g_ctx, g_cache[n].field);
        memset(&g_cache[n], 0, sizeof(g_cache[n]));
        dropped = 1;
      }
    }
    if (!dropped || 0 != c3d_field_create(g_ctx, decomp, glsizes, (int)sizeof(double), field)) return 1;
  }
  if (cache_enabled()) {
    for (size_t n = 0; n < MAX_OBJECTS; n++) {
      cached_field_t *c = &g_cache[n];
      if (NULL == c->field) {
        c->field = *field;
        memcpy(c->glsizes, glsizes, sizeof(c->glsizes));
        c->decomp = *decomp;
        c->in_use = 1;
        *cached = 1;
        break;
      }
    }
  }
  return 0;
}

static int is_identity(contour3d_vector_t (*converter)(const contour3d_vector_t), const double probes[][3], const size_t nprobes) {
  if (NULL == converter) return 1;
  for (size_t n = 0; n < nprobes; n++) {
    const contour3d_vector_t in = {probes[n][0], probes[n][1], probes[n][2]};
    const contour3d_vector_t out = converter(in);
    if (0 != memcmp(&in, &out, sizeof(in))) return 0;
  }
  return 1;
}

int contour3d_execute (
    const sdecomp_info_t * const sdecomp_info,
    const contour3d_vector_t * const camera_position,
    const contour3d_vector_t * const camera_look_at,
    const contour3d_vector_t * const light_direction,
    const size_t screen_sizes[2],
    const contour3d_vector_t * const screen_center,
    const contour3d_vector_t screen_local[2],
    const contour3d_color_t * const bg_color,
    const size_t num_contours,
    const contour3d_contour_obj_t * const contour3d_contour_objs,
    const size_t num_lines,
    const contour3d_line_obj_t * const contour3d_line_objs,
    const char fname[]
) {
  (void)camera_look_at; /* stored but never read by the reference either (src/contour3d/main.c:17-20) */
  c3d_field_t *fields[MAX_OBJECTS] = {NULL};
  int cached[MAX_OBJECTS] = {0};
  c3d_contour_t contours[MAX_OBJECTS];
  c3d_line_t *lines = NULL;
  int status = 1;
  if (NULL == sdecomp_info || NULL == camera_position || NULL == light_direction || NULL == screen_sizes
      || NULL == screen_center || NULL == screen_local || NULL == bg_color || NULL == fname) {
    return log_error("contour3d_execute: NULL argument");
  }
  if (num_contours > MAX_OBJECTS) return log_error("at most %d contour objects per call, received %zu", MAX_OBJECTS, num_contours);
  if (num_contours > 0 && NULL == contour3d_contour_objs) return log_error("contour objects are missing");
  if (num_lines > 0 && NULL == contour3d_line_objs) return log_error("line objects are missing");
  if (0 != acquire_context(sdecomp_info)) return 1;
  c3d_scene_t scene;
  memset(&scene, 0, sizeof(scene));
  scene.camera[0] = camera_position->x; scene.camera[1] = camera_position->y; scene.camera[2] = camera_position->z;
  scene.light[0] = light_direction->x; scene.light[1] = light_direction->y; scene.light[2] = light_direction->z;
  scene.center[0] = screen_center->x; scene.center[1] = screen_center->y; scene.center[2] = screen_center->z;
  scene.local_x[0] = screen_local[0].x; scene.local_x[1] = screen_local[0].y; scene.local_x[2] = screen_local[0].z;
  scene.local_y[0] = screen_local[1].x; scene.local_y[1] = screen_local[1].y; scene.local_y[2] = screen_local[1].z;
  scene.width = screen_sizes[0];
  scene.height = screen_sizes[1];
  scene.bg[0] = bg_color->r; scene.bg[1] = bg_color->g; scene.bg[2] = bg_color->b;
  /* contour objects -> device fields (one per distinct array) */
  for (size_t n = 0; n < num_contours; n++) {
    const contour3d_contour_obj_t *obj = contour3d_contour_objs + n;
    if (SDECOMP_X1PENCIL != obj->pencil && SDECOMP_X2PENCIL != obj->pencil) {
      log_error("contour object %zu: only x pencils (x-fastest arrays) are supported, received pencil %d", n, (int)obj->pencil);
      goto done;
    }
    if (NULL == obj->array || NULL == obj->grids[0] || NULL == obj->grids[1] || NULL == obj->grids[2]) {
      log_error("contour object %zu: array or grids missing", n);
      goto done;
    }
    const double probes[4][3] = {
      {obj->grids[0][0], obj->grids[1][0], obj->grids[2][0]},
      {obj->grids[0][obj->glsizes[0] - 1], obj->grids[1][obj->glsizes[1] - 1], obj->grids[2][obj->glsizes[2] - 1]},
      {obj->grids[0][obj->glsizes[0] / 2], obj->grids[1][obj->glsizes[1] / 3], obj->grids[2][obj->glsizes[2] / 5]},
      {0.3183098861837907, -2.718281828459045, 1.4142135623730951},
    };
    if (!is_identity(obj->converter, probes, 4)) {
      log_error("contour object %zu: a non-identity coordinate converter cannot be called from the device", n);
      goto done;
    }
    /* what contour/extend.c:160-181, 216-221 asks SimpleDecomp */
    c3d_decomp_t decomp;
    for (sdecomp_dir_t dir = 0; dir < CONTOUR3D_NDIMS; dir++) {
      if (0 != sdecomp.get_nprocs(sdecomp_info, obj->pencil, dir, &decomp.nprocs[dir])
          || 0 != sdecomp.get_myrank(sdecomp_info, obj->pencil, dir, &decomp.myrank[dir])
          || 0 != sdecomp.get_neighbours(sdecomp_info, obj->pencil, dir, decomp.neighbours[dir])) {
        log_error("sdecomp getters failed");
        goto done;
      }
      /* no wrap-around: boundary ranks clip instead (extend.c:222-252) */
      if (0 == decomp.myrank[dir]) decomp.neighbours[dir][0] = -1;
      if (decomp.nprocs[dir] - 1 == decomp.myrank[dir]) decomp.neighbours[dir][1] = -1;
    }
    size_t same = n;
    for (size_t m = 0; m < n; m++) {
      const contour3d_contour_obj_t *prev = contour3d_contour_objs + m;
      if (prev->array == obj->array && prev->pencil == obj->pencil
          && 0 == memcmp(prev->glsizes, obj->glsizes, sizeof(obj->glsizes))) {
        same = m;
        break;
      }
    }
    if (same < n) {
      contours[n].field = contours[same].field;
    } else {
      if (0 != acquire_field(&decomp, obj->glsizes, &fields[n], &cached[n])) {
        log_error("failed to extend domain");
        goto done;
      }
      if (0 != c3d_field_upload(g_ctx, fields[n], obj->array)) {
        log_error("failed to extend domain");
        goto done;
      }
      contours[n].field = fields[n];
    }
    for (int d = 0; d < 3; d++) contours[n].grids[d] = obj->grids[d];
    contours[n].threshold = obj->threshold;
    contours[n].color[0] = obj->color.r; contours[n].color[1] = obj->color.g; contours[n].color[2] = obj->color.b;
  }
  if (num_lines > 0) {
    lines = calloc(num_lines, sizeof(c3d_line_t));
    if (NULL == lines) { log_error("out of host memory"); goto done; }
  }
  for (size_t n = 0; n < num_lines; n++) {
    const contour3d_line_obj_t *obj = contour3d_line_objs + n;
    const double probes[3][3] = {
      {obj->edges[0].x, obj->edges[0].y, obj->edges[0].z},
      {obj->edges[1].x, obj->edges[1].y, obj->edges[1].z},
      {0.3183098861837907, -2.718281828459045, 1.4142135623730951},
    };
    if (!is_identity(obj->converter, probes, 3)) {
      log_error("line object %zu: a non-identity coordinate converter cannot be called from the device", n);
      goto done;
    }
    lines[n].edges[0][0] = obj->edges[0].x; lines[n].edges[0][1] = obj->edges[0].y; lines[n].edges[0][2] = obj->edges[0].z;
    lines[n].edges[1][0] = obj->edges[1].x; lines[n].edges[1][1] = obj->edges[1].y; lines[n].edges[1][2] = obj->edges[1].z;
    lines[n].nitems = obj->nitems;
    lines[n].width = obj->width;
    lines[n].color[0] = obj->color.r; lines[n].color[1] = obj->color.g; lines[n].color[2] = obj->color.b;
  }
  if (0 != c3d_frame_render(g_ctx, &scene, num_contours, contours, num_lines, lines, NULL, fname)) {
    log_error("frame rendering failed");
    goto done;
  }
  status = 0;
done:
  /* nothing is retained after return, like the reference (contour/main.c:114-117) */
  for (size_t n = 0; n < num_contours && n < MAX_OBJECTS; n++) {
    if (NULL != fields[n] && !cached[n]) c3d_field_destroy(g_ctx, fields[n]);
  }
  for (size_t n = 0; n < MAX_OBJECTS; n++) g_cache[n].in_use = 0;
  {
    const char *env = getenv("CONTOUR3D_VERBOSE");
    c3d_timings_t t;
    if (0 == status && NULL != env && '\0' != *env && '0' != *env && 0 == c3d_frame_timings(g_ctx, &t)) {
      fprintf(stderr, "[CONTOUR3D] upload %.2f sweep %.2f expand %.2f raster %.2f shade %.2f reduce %.2f download %.2f | device total %.2f ms\n",
              t.upload, t.sweep, t.expand, t.raster, t.shade, t.reduce, t.download, t.upload + t.total);
    }
  }
  free(lines);
  return status;
}
