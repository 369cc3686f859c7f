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
 *  - `converter` is evaluated on the host, once per grid point at most, and
 *    tabulated for the kernels (c3d_positions.c): any deterministic function
 *    works as long as its tables fit CONTOUR3D_CONVERTER_MAX_BYTES (default
 *    4 GiB; a function of all three coordinates in every component needs
 *    24 bytes per point).  NULL is accepted as the identity;
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

/* the callback of include/contour3d.h:31-33 behind the plain-double signature of c3d_positions_probe */
typedef struct { contour3d_vector_t (*fn)(const contour3d_vector_t); } converter_box_t;

static void call_converter(const double in[3], double out[3], void *user) {
  const converter_box_t *box = user;
  const contour3d_vector_t arg = {in[0], in[1], in[2]};
  const contour3d_vector_t res = box->fn(arg);
  out[0] = res.x; out[1] = res.y; out[2] = res.z;
}

static size_t converter_max_bytes(void) {
  const char *env = getenv("CONTOUR3D_CONVERTER_MAX_BYTES");
  if (NULL != env && '\0' != *env) return (size_t)strtoull(env, NULL, 10);
  return (size_t)4 << 30;
}

/* the points of a poly-line, converted: s_n = head + n * deltas (line.c:138-140,157-158; e_n == s_(n+1) bit for bit).
   *points stays NULL when the converter leaves every point unchanged. */
static int convert_line(const contour3d_line_obj_t *obj, double **points) {
  *points = NULL;
  if (NULL == obj->converter || obj->nitems < 2) return 0;
  double *p = malloc(3 * obj->nitems * sizeof(double));
  if (NULL == p) return log_error("out of host memory");
  const double w = 1. / (obj->nitems - 1);
  const contour3d_vector_t head = obj->edges[0];
  const contour3d_vector_t deltas = {w * (obj->edges[1].x - head.x), w * (obj->edges[1].y - head.y), w * (obj->edges[1].z - head.z)};
  int identity = 1;
  for (size_t n = 0; n < obj->nitems; n++) {
    const contour3d_vector_t s = {head.x + n * deltas.x, head.y + n * deltas.y, head.z + n * deltas.z};
    const contour3d_vector_t c = obj->converter(s);
    if (0 != memcmp(&s, &c, sizeof(s))) identity = 0;
    p[3 * n + 0] = c.x; p[3 * n + 1] = c.y; p[3 * n + 2] = c.z;
  }
  if (identity) free(p);
  else *points = p;
  return 0;
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
  c3d_positions_t positions[MAX_OBJECTS];
  int has_positions[MAX_OBJECTS] = {0};
  converter_box_t boxes[MAX_OBJECTS];
  c3d_line_t *lines = NULL;
  double **line_points = NULL;
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
  /* Everything up to the frame can fail on ONE rank only (a converter table over the memory bound for that rank's block, an
     allocation, a halo that does not fit): such a rank must not leave the others alone in the frame's collectives.
     `failed` is agreed on (maximum over the ranks, MPI_Allreduce on the Cartesian communicator) before the frame starts. */
  int failed = 1;
  /* contour objects -> device fields (one per distinct array) */
  for (size_t n = 0; n < num_contours; n++) {
    const contour3d_contour_obj_t *obj = contour3d_contour_objs + n;
    if (SDECOMP_X1PENCIL != obj->pencil && SDECOMP_X2PENCIL != obj->pencil) {
      log_error("contour object %zu: only x pencils (x-fastest arrays) are supported, received pencil %d", n, (int)obj->pencil);
      goto agree;
    }
    if (NULL == obj->array || NULL == obj->grids[0] || NULL == obj->grids[1] || NULL == obj->grids[2]) {
      log_error("contour object %zu: array or grids missing", n);
      goto agree;
    }
    /* what contour/extend.c:160-181, 216-221 asks SimpleDecomp */
    c3d_decomp_t decomp;
    for (sdecomp_dir_t dir = 0; dir < CONTOUR3D_NDIMS; dir++) {
      if (0 != sdecomp.get_nprocs(sdecomp_info, obj->pencil, dir, &decomp.nprocs[dir])
          || 0 != sdecomp.get_myrank(sdecomp_info, obj->pencil, dir, &decomp.myrank[dir])
          || 0 != sdecomp.get_neighbours(sdecomp_info, obj->pencil, dir, decomp.neighbours[dir])) {
        log_error("sdecomp getters failed");
        goto agree;
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
      /* a device array that needs no halo room (one rank) is swept where it lies; everything else is copied into an
         array with halo room (the padding copy of contour/extend.c:199-208) */
      const int wrapped = c3d_field_wrap(g_ctx, &decomp, obj->glsizes, (int)sizeof(double), obj->array, &fields[n]);
      if (1 == wrapped) {
        log_error("failed to extend domain");
        goto agree;
      }
      if (0 != wrapped) {
        if (0 != acquire_field(&decomp, obj->glsizes, &fields[n], &cached[n])) {
          log_error("failed to extend domain");
          goto agree;
        }
        if (0 != c3d_field_upload(g_ctx, fields[n], obj->array)) {
          log_error("failed to extend domain");
          goto agree;
        }
      }
      contours[n].field = fields[n];
    }
    for (int d = 0; d < 3; d++) contours[n].grids[d] = obj->grids[d];
    /* the coordinate converter, tabulated over the points of the extended local array (c3d_positions.c) */
    contours[n].positions = NULL;
    if (NULL != obj->converter) {
      size_t reuse = n;
      for (size_t m = 0; m < n; m++) {
        if (contours[m].field == contours[n].field && contour3d_contour_objs[m].converter == obj->converter
            && 0 == memcmp(contour3d_contour_objs[m].grids, obj->grids, sizeof(obj->grids))) {
          reuse = m;
          break;
        }
      }
      if (reuse < n) {
        contours[n].positions = contours[reuse].positions;
      } else {
        size_t ext[3], ext_off[3], mysizes[3], offsets[3];
        int identity = 1;
        if (0 != c3d_field_geometry(contours[n].field, ext, ext_off, mysizes, offsets)) goto agree;
        boxes[n].fn = obj->converter;
        if (0 != c3d_positions_probe(call_converter, &boxes[n], obj->grids, ext_off, ext, converter_max_bytes(), &positions[n], &identity)) {
          log_error("contour object %zu: the coordinate converter could not be tabulated", n);
          goto agree;
        }
        if (!identity) {
          has_positions[n] = 1;
          contours[n].positions = &positions[n];
        }
      }
    }
    contours[n].threshold = obj->threshold;
    contours[n].color[0] = obj->color.r; contours[n].color[1] = obj->color.g; contours[n].color[2] = obj->color.b;
  }
  if (num_lines > 0) {
    lines = calloc(num_lines, sizeof(c3d_line_t));
    line_points = calloc(num_lines, sizeof(double *));
    if (NULL == lines || NULL == line_points) { log_error("out of host memory"); goto agree; }
  }
  for (size_t n = 0; n < num_lines; n++) {
    const contour3d_line_obj_t *obj = contour3d_line_objs + n;
    if (0 != convert_line(obj, &line_points[n])) goto agree;
    lines[n].points = line_points[n];
    lines[n].edges[0][0] = obj->edges[0].x; lines[n].edges[0][1] = obj->edges[0].y; lines[n].edges[0][2] = obj->edges[0].z;
    lines[n].edges[1][0] = obj->edges[1].x; lines[n].edges[1][1] = obj->edges[1].y; lines[n].edges[1][2] = obj->edges[1].z;
    lines[n].nitems = obj->nitems;
    lines[n].width = obj->width;
    lines[n].color[0] = obj->color.r; lines[n].color[1] = obj->color.g; lines[n].color[2] = obj->color.b;
  }
  failed = 0;
agree:
  if (g_size > 1) {
    MPI_Comm comm = MPI_COMM_NULL;
    /* which objects carry converter tables must be the same everywhere too: it selects how the ranks exchange data */
    int word[1 + 2 * MAX_OBJECTS], agreed[1 + 2 * MAX_OBJECTS];
    word[0] = failed;
    for (size_t n = 0; n < MAX_OBJECTS; n++) {
      const int counts = n < num_contours && !failed;
      word[1 + n] = (counts && NULL != contours[n].positions) ? 1 : 0;                  /* "I have tables" ... */
      word[1 + MAX_OBJECTS + n] = (counts && NULL == contours[n].positions) ? 1 : 0;    /* ... "I have none": both seen = mismatch, seen by all */
    }
    if (0 != sdecomp.get_comm_cart(sdecomp_info, &comm) || MPI_SUCCESS != MPI_Allreduce(word, agreed, 1 + 2 * MAX_OBJECTS, MPI_INT, MPI_MAX, comm)) {
      log_error("the ranks could not agree on the state of the call");
      goto done;
    }
    if (0 != agreed[0]) {
      if (!failed) log_error("another rank could not prepare the frame");
      goto done;
    }
    for (size_t n = 0; n < num_contours; n++) {
      if (agreed[1 + n] && agreed[1 + MAX_OBJECTS + n]) {
        log_error("contour object %zu: the coordinate converter is the identity on some ranks' blocks and not on others'", n);
        goto done;
      }
    }
  }
  if (failed) goto done;
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
  for (size_t n = 0; n < MAX_OBJECTS; n++) {
    if (has_positions[n]) c3d_positions_free(&positions[n]);
  }
  for (size_t n = 0; NULL != line_points && n < num_lines; n++) free(line_points[n]);
  free(line_points);
  free(lines);
  return status;
}
