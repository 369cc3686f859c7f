/*
 * c3d_oracle.c -- CPU restatement of Contour3D's per-frame pipeline.
 * TEST INFRASTRUCTURE ONLY (see c3d_oracle.h).
 *
 * Parity status: PINNED.  Checked bit-for-bit against the unmodified
 * reference built by oracle/Makefile (tests/test_oracle_vs_reference.py:
 * per-cube triangle dumps through the reference's extern stage functions and
 * whole images incl. the reference's own demo scene), and against golden
 * vectors generated from that build (tests/golden/, tools/make_golden.py).
 *
 * The restatement is NOT a copy: cubes are stored sparsely (the reference
 * keeps a 2408-byte lattice_t for every cube), the cube-edge table and the
 * smoothing neighbourhoods are derived from geometry instead of being spelled
 * out, and threads emulate ranks.  What is kept exactly is the arithmetic:
 * every expression below is evaluated in the reference's order, in fp64,
 * without FMA contraction (build with -std=c99).
 *
 * File:line references are to /root/reference (commit c66854e).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <errno.h>
#include <pthread.h>
#include "c3d_oracle.h"

typedef oracle_vec_t vec;

/* ---- 3-vector helpers, operation order of src/contour3d/vector.c:5-76 ---- */
static inline vec v_add(const vec a, const vec b) { const vec r = {a.x + b.x, a.y + b.y, a.z + b.z}; return r; }
static inline vec v_sub(const vec a, const vec b) { const vec r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }
static inline vec v_mul(const double w, const vec a) { const vec r = {w * a.x, w * a.y, w * a.z}; return r; }
static inline double v_dot(const vec a, const vec b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline vec v_cross(const vec a, const vec b) {
  const vec r = {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
  return r;
}
static inline vec v_unit(const vec a) {
  /* multiply by the reciprocal root, vector.c:66-76 */
  const double inv = 1. / sqrt(v_dot(a, a));
  const vec r = {a.x * inv, a.y * inv, a.z * inv};
  return r;
}
static inline vec v_from(const double p[3]) { const vec r = {p[0], p[1], p[2]}; return r; }

/* ---- marching-tetrahedra tables (contour/triangulate.c) ---- */

/* six tetrahedra around the 0-7 diagonal, triangulate.c:145-152 */
static const uint8_t TETRA[6][4] = {
  {0, 7, 3, 1}, {0, 7, 1, 5}, {0, 7, 5, 4}, {0, 7, 4, 6}, {0, 7, 6, 2}, {0, 7, 2, 3},
};

/* folded mask 1..7 -> one or two triangles given as (tail, head) tetra-vertex
   pairs per triangle vertex, triangulate.c:204-270 */
typedef struct { uint8_t n; uint8_t tail[2][3]; uint8_t head[2][3]; } case_t;
static const case_t CASES[8] = {
  {0, {{0, 0, 0}, {0, 0, 0}}, {{0, 0, 0}, {0, 0, 0}}},
  {1, {{0, 0, 0}, {0, 0, 0}}, {{1, 3, 2}, {0, 0, 0}}},
  {1, {{1, 1, 1}, {0, 0, 0}}, {{0, 2, 3}, {0, 0, 0}}},
  {2, {{0, 1, 0}, {0, 1, 1}}, {{2, 3, 3}, {2, 2, 3}}},
  {1, {{2, 2, 2}, {0, 0, 0}}, {{0, 3, 1}, {0, 0, 0}}},
  {2, {{2, 0, 2}, {2, 0, 0}}, {{1, 3, 3}, {1, 1, 3}}},
  {2, {{0, 2, 1}, {0, 1, 0}}, {{2, 3, 3}, {2, 3, 1}}},
  {1, {{3, 3, 3}, {0, 0, 0}}, {{0, 2, 1}, {0, 0, 0}}},
};

/* cube-edge id of the edge between cube corners a and b (corner bit0 = +x,
   bit1 = +y, bit2 = +z).  Same numbering as the 8x8 table of
   triangulate.c:23-32, derived from which axes differ:
     0-3 x edges (dj + 2 dk), 4-7 y edges (di + 2 dk), 8-11 z edges (di + 2 dj),
     12/13 yz-face diagonal (di), 14/15 xz-face diagonal (dj),
     16/17 xy-face diagonal (dk), 18 body diagonal */
static inline uint32_t cube_edge(const uint32_t a, const uint32_t b) {
  const uint32_t lo = a & b;          /* the offsets shared by both ends */
  const uint32_t di = lo & 1u, dj = (lo >> 1) & 1u, dk = (lo >> 2) & 1u;
  switch (a ^ b) {
    case 1u: return 0u + dj + 2u * dk;
    case 2u: return 4u + di + 2u * dk;
    case 4u: return 8u + di + 2u * dj;
    case 6u: return 12u + di;
    case 5u: return 14u + dj;
    case 3u: return 16u + dk;
    default: return 18u;
  }
}

/* internal view of a contour object: a window of a strided array */
typedef struct {
  size_t n[3];                /* points of the window */
  size_t sy, sz;              /* strides (in elements) between rows / planes */
  const double *field;        /* first element of the window */
  const double *gx, *gy, *gz; /* coordinates of the window's points */
  double threshold;
  uint8_t color[3];
  oracle_converter_t converter;
} view_t;

static view_t view_of(const oracle_contour_t *c) {
  view_t v;
  v.n[0] = c->dims[0]; v.n[1] = c->dims[1]; v.n[2] = c->dims[2];
  v.sy = c->dims[0];
  v.sz = c->dims[0] * c->dims[1];
  v.field = c->field;
  v.gx = c->grid[0]; v.gy = c->grid[1]; v.gz = c->grid[2];
  v.threshold = c->threshold;
  memcpy(v.color, c->color, 3);
  v.converter = c->converter;
  return v;
}

typedef struct {
  vec v[3];
  vec fn;
  vec vn[3];
  double area;
  uint8_t edge[3];
} tri_t;

/* the eight corner scalars and positions of cube (i, j, k), triangulate.c:156-173, 302-317 */
typedef struct { double val[8]; vec pos[8]; } cube_t;

static void load_cube(const view_t *w, const size_t i, const size_t j, const size_t k, cube_t *cube) {
  for (uint32_t c = 0; c < 8; c++) {
    const size_t di = c & 1u, dj = (c >> 1) & 1u, dk = (c >> 2) & 1u;
    cube->val[c] = w->field[(k + dk) * w->sz + (j + dj) * w->sy + (i + di)];
    const vec ortho = {w->gx[i + di], w->gy[j + dj], w->gz[k + dk]};
    cube->pos[c] = w->converter ? w->converter(ortho) : ortho;
  }
}

/* pre-fold 4-bit mask of tetrahedron n: bit m set when threshold < value of
   tetra vertex m (strict), triangulate.c:176-179 */
static inline uint32_t tetra_mask(const cube_t *cube, const double threshold, const int n) {
  uint32_t mask = 0;
  for (int m = 0; m < 4; m++) {
    mask |= (uint32_t)(threshold < cube->val[TETRA[n][m]]) << m;
  }
  return mask;
}

/* triangle vertex on the tetra edge tail -> head, triangulate.c:40-73 */
static inline vec cut_edge(const double threshold, const cube_t *cube, const uint32_t tail, const uint32_t head) {
  const double v0 = cube->val[tail], v1 = cube->val[head];
  if (fabs(v1 - v0) < 1.e-8) {
    return cube->pos[tail];
  }
  const double factor = (threshold - v0) / (v1 - v0);
  return v_add(cube->pos[tail], v_mul(factor, v_sub(cube->pos[head], cube->pos[tail])));
}

/* all triangles of one cube, in the reference's order (tetra 0..5, A before B);
   returns the count (<= 12) */
static int tessellate_cube(const cube_t *cube, const double threshold, tri_t out[12], uint32_t *code) {
  int count = 0;
  uint32_t packed = 0;
  for (int n = 0; n < 6; n++) {
    uint32_t mask = tetra_mask(cube, threshold, n);
    packed |= mask << (4 * n);
    /* fold the upper half onto the lower with reversed winding, triangulate.c:189-192 */
    const int reverse = mask >= 8u;
    if (reverse) mask = 15u - mask;
    const case_t *cs = &CASES[mask];
    for (int t = 0; t < cs->n; t++) {
      tri_t *tri = &out[count++];
      for (int m = 0; m < 3; m++) {
        const int pick = reverse ? 2 - m : m;       /* triangulate.c:108-112 */
        const uint32_t tail = TETRA[n][cs->tail[t][pick]];
        const uint32_t head = TETRA[n][cs->head[t][pick]];
        tri->v[m] = cut_edge(threshold, cube, tail, head);
        tri->edge[m] = (uint8_t)cube_edge(tail, head);
      }
      /* face normal and |cross|, triangulate.c:75-93 */
      const vec c = v_cross(v_sub(tri->v[1], tri->v[0]), v_sub(tri->v[2], tri->v[0]));
      tri->area = sqrt(v_dot(c, c));
      tri->fn = v_mul(1. / tri->area, c);
      const vec zero = {0., 0., 0.};
      tri->vn[0] = zero; tri->vn[1] = zero; tri->vn[2] = zero;
    }
  }
  if (code) *code = packed;
  return count;
}

/* ---- sparse slice of cubes ---- */
typedef struct {
  size_t ncubes;
  uint32_t *start;            /* ncubes + 1 offsets into tris */
  tri_t *tris;
  size_t cap;
} slice_t;

static int slice_init(slice_t *s, const size_t ncubes) {
  s->ncubes = ncubes;
  s->start = calloc(ncubes + 1, sizeof(uint32_t));
  s->cap = 1024;
  s->tris = malloc(s->cap * sizeof(tri_t));
  return (NULL == s->start || NULL == s->tris) ? 1 : 0;
}

static void slice_free(slice_t *s) {
  free(s->start);
  free(s->tris);
  s->start = NULL;
  s->tris = NULL;
}

/* reference: contour3d_contour_triangulate_slice, triangulate.c:287-342 */
static int slice_fill(slice_t *s, const view_t *w, const size_t k, oracle_stats_t *stats) {
  const size_t cx = w->n[0] - 1, cy = w->n[1] - 1;
  size_t used = 0;
  for (size_t j = 0; j < cy; j++) {
    for (size_t i = 0; i < cx; i++) {
      if (used + 12 > s->cap) {
        s->cap *= 2;
        tri_t *grown = realloc(s->tris, s->cap * sizeof(tri_t));
        if (NULL == grown) return 1;
        s->tris = grown;
      }
      cube_t cube;
      load_cube(w, i, j, k, &cube);
      const int n = tessellate_cube(&cube, w->threshold, s->tris + used, NULL);
      s->start[j * cx + i] = (uint32_t)used;
      used += (size_t)n;
      if (stats) {
        stats->cubes++;
        stats->triangles += (uint64_t)n;
        if (n > 0) stats->active_cubes++;
      }
    }
  }
  s->start[cx * cy] = (uint32_t)used;
  return 0;
}

/* add area * face_normal of every triangle of `cube` in `s` that has a vertex
   on cube edge `edge`, in triangle order.  reference: average(), normal.c:6-43 */
static inline void gather(const slice_t *s, const size_t cube, const uint32_t edge, vec *sum) {
  for (uint32_t n = s->start[cube]; n < s->start[cube + 1]; n++) {
    const tri_t *t = &s->tris[n];
    if (edge == t->edge[0] || edge == t->edge[1] || edge == t->edge[2]) {
      *sum = v_add(*sum, v_mul(t->area, t->fn));
    }
  }
}

/* vertex normal of the triangle vertex sitting on edge `e` of cube (i, j) of
   the middle slice.  The cubes sharing that edge are visited in global
   (k, j, i) order and the edge is renamed to each cube's local id.
   reference: the 19-way branch of normal.c:78-158 */
static vec vertex_normal(const slice_t *const ring[3], const size_t cx, const size_t i, const size_t j, const uint32_t e) {
  vec sum = {0., 0., 0.};
  if (e < 12u) {
    /* axis edge: four cubes around it.  (a, b) = the two in-plane offsets of
       the edge inside the cube; neighbours run over the two transverse axes,
       slow axis first */
    const uint32_t axis = e >> 2;           /* 0: x edge, 1: y edge, 2: z edge */
    const uint32_t a = e & 1u, b = (e >> 1) & 1u;
    for (uint32_t q = 0; q < 2; q++) {      /* slow transverse axis */
      for (uint32_t p = 0; p < 2; p++) {    /* fast transverse axis */
        /* neighbour offset (-1 or 0) + own offset; local edge id mirrors */
        const int da = (int)a - 1 + (int)p, db = (int)b - 1 + (int)q;
        const uint32_t local = 4u * axis + (1u - p) + 2u * (1u - q);
        size_t ii = i, jj = j;
        int kk = 1;
        if (0u == axis) { jj = (size_t)((long)j + da); kk += db; }        /* x edge: (j, k) */
        else if (1u == axis) { ii = (size_t)((long)i + da); kk += db; }   /* y edge: (i, k) */
        else { ii = (size_t)((long)i + da); jj = (size_t)((long)j + db); }/* z edge: (i, j) */
        gather(ring[kk], jj * cx + ii, local, &sum);
      }
    }
  } else if (e < 18u) {
    /* face diagonal: two cubes on either side of the face */
    const uint32_t kind = (e - 12u) >> 1;   /* 0: yz face (x normal), 1: xz face, 2: xy face */
    const uint32_t a = e & 1u;
    for (uint32_t p = 0; p < 2; p++) {
      const int da = (int)a - 1 + (int)p;
      const uint32_t local = 12u + 2u * kind + (1u - p);
      size_t ii = i, jj = j;
      int kk = 1;
      if (0u == kind) ii = (size_t)((long)i + da);
      else if (1u == kind) jj = (size_t)((long)j + da);
      else kk += da;
      gather(ring[kk], jj * cx + ii, local, &sum);
    }
  } else {
    gather(ring[1], j * cx + i, 18u, &sum);
  }
  return v_unit(sum);                       /* normal.c:159 */
}

/* reference: contour3d_contour_compute_vertex_normals, normal.c:46-165 */
static void slice_normals(slice_t *const ring[3], const size_t cx, const size_t cy) {
  if (cx < 3 || cy < 3) return;
  const slice_t *const cring[3] = {ring[0], ring[1], ring[2]};
  for (size_t j = 1; j < cy - 1; j++) {
    for (size_t i = 1; i < cx - 1; i++) {
      const size_t cube = j * cx + i;
      for (uint32_t n = ring[1]->start[cube]; n < ring[1]->start[cube + 1]; n++) {
        tri_t *t = &ring[1]->tris[n];
        for (int m = 0; m < 3; m++) {
          t->vn[m] = vertex_normal(cring, cx, i, j, t->edge[m]);
        }
      }
    }
  }
}

/* ---- camera / screen ---- */
typedef struct {
  vec camera, light, center, lx, ly, normal;
  size_t width, height;
} cam_t;

/* reference: init_camera/light/screen, src/contour3d/main.c:13-63 */
static cam_t cam_of(const oracle_scene_t *s) {
  cam_t c;
  c.camera = v_from(s->camera);
  c.light = v_unit(v_from(s->light));
  c.center = v_from(s->center);
  c.lx = v_from(s->local_x);
  c.ly = v_from(s->local_y);
  c.normal = v_unit(v_cross(c.lx, c.ly));
  c.width = s->width;
  c.height = s->height;
  return c;
}

/* reference: contour3d_project, project.c:9-67.  returns 0 on success */
static int project(const cam_t *c, const vec p, vec *out) {
  const vec ray = v_sub(p, c->camera);
  const double nc = v_dot(c->normal, c->camera);
  const double nr = v_dot(c->normal, ray);
  const double d = v_dot(c->normal, c->center);
  if (-1.e-8 < nr && nr < 1.e-8) return 1;
  const double t = (d - nc) / nr;
  if (t <= 0. || 1. <= t) return 1;
  const vec delta = {
    c->camera.x + t * ray.x - c->center.x,
    c->camera.y + t * ray.y - c->center.y,
    c->camera.z + t * ray.z - c->center.z,
  };
  out->x = v_dot(delta, c->lx) / v_dot(c->lx, c->lx);
  out->y = v_dot(delta, c->ly) / v_dot(c->ly, c->ly);
  out->z = nr;
  return 0;
}

static inline long clampl(const long lo, const long hi, const long v) {
  /* min(hi, max(lo, v)) as in render.c:86-89 */
  const long m = v < lo ? lo : v;
  return hi < m ? hi : m;
}

static inline double edge_fn(const vec *a, const vec *b, const vec *c) {
  /* render.c:38-46 */
  return (c->x - a->x) * (b->y - a->y) - (c->y - a->y) * (b->x - a->x);
}

/* reference: contour3d_contour_render_triangle, render.c:48-163 */
static void raster_triangle(const cam_t *c, const uint8_t color[3], const tri_t *tri, double *depth, uint8_t *rgb) {
  vec q[3];
  for (int m = 0; m < 3; m++) {
    if (0 != project(c, tri->v[m], &q[m])) return;
  }
  const double W = (double)c->width, H = (double)c->height;
  const long xmin = (long)((0.5 + fmin(q[0].x, fmin(q[1].x, q[2].x))) * W - 1);
  const long xmax = (long)((0.5 + fmax(q[0].x, fmax(q[1].x, q[2].x))) * W + 1);
  const long ymin = (long)((0.5 + fmin(q[0].y, fmin(q[1].y, q[2].y))) * H - 1);
  const long ymax = (long)((0.5 + fmax(q[0].y, fmax(q[1].y, q[2].y))) * H + 1);
  const size_t imin = (size_t)clampl(0, (long)c->width - 1, xmin);
  const size_t imax = (size_t)clampl(0, (long)c->width - 1, xmax);
  const size_t jmin = (size_t)clampl(0, (long)c->height - 1, ymin);
  const size_t jmax = (size_t)clampl(0, (long)c->height - 1, ymax);
  const double area = edge_fn(&q[0], &q[1], &q[2]);
  const double area_inv = 1. / (area < 0. ? fmin(area, -1.e-16) : fmax(area, +1.e-16));
  for (size_t j = jmin; j <= jmax; j++) {
    for (size_t i = imin; i <= imax; i++) {
      const vec p = {1. * ((double)i + 0.5) / W - 0.5, 1. * ((double)j + 0.5) / H - 0.5, 0.};
      const double w0 = area_inv * edge_fn(&q[1], &q[2], &p);
      const double w1 = area_inv * edge_fn(&q[2], &q[0], &p);
      const double w2 = area_inv * edge_fn(&q[0], &q[1], &p);
      const double slack = -1.e-8;
      if (w0 < slack || w1 < slack || w2 < slack) continue;
      const size_t pix = j * c->width + i;
      const double dist = 1. / (w0 / q[0].z + w1 / q[1].z + w2 / q[2].z);
      if (dist < depth[pix]) continue;
      depth[pix] = dist;
      const vec n = v_unit(v_add(v_mul(w0, tri->vn[0]), v_add(v_mul(w1, tri->vn[1]), v_mul(w2, tri->vn[2]))));
      const double factor = fmax(0.15, fabs(v_dot(n, c->light)));
      rgb[3 * pix + 0] = (uint8_t)(factor * color[0]);
      rgb[3 * pix + 1] = (uint8_t)(factor * color[1]);
      rgb[3 * pix + 2] = (uint8_t)(factor * color[2]);
    }
  }
}

/* the three-slice march of contour/main.c:49-113.  `sink` receives every
   triangle once its slice is final (vertex normals present when has_vn) */
typedef void (*sink_t)(void *ctx, size_t i, size_t j, size_t k, uint32_t t, const tri_t *tri, int has_vn);

static int march(const view_t *w, oracle_stats_t *stats, sink_t sink, void *ctx, const int interior_only) {
  if (w->n[0] < 2 || w->n[1] < 2 || w->n[2] < 2) return 1;
  const size_t cx = w->n[0] - 1, cy = w->n[1] - 1, cz = w->n[2] - 1;
  slice_t store[3];
  for (int n = 0; n < 3; n++) {
    if (0 != slice_init(&store[n], cx * cy)) return 1;
  }
  int status = 0;
  for (size_t k = 0; k < cz && 0 == status; k++) {
    status = slice_fill(&store[k % 3], w, k, stats);
    if (0 != status) break;
    /* boundary slices in z never get normals: flush the first one as it is */
    if (!interior_only && 0 == k) {
      const slice_t *s = &store[0];
      for (size_t c = 0; c < cx * cy; c++)
        for (uint32_t n = s->start[c]; n < s->start[c + 1]; n++)
          sink(ctx, c % cx, c / cx, 0, n - s->start[c], &s->tris[n], 0);
    }
    if (k < 2) continue;
    slice_t *const ring[3] = {&store[(k - 2) % 3], &store[(k - 1) % 3], &store[k % 3]};
    slice_normals(ring, cx, cy);
    const slice_t *mid = ring[1];
    for (size_t j = 0; j < cy; j++) {
      for (size_t i = 0; i < cx; i++) {
        const int inner = (i >= 1 && i + 1 < cx && j >= 1 && j + 1 < cy);
        if (interior_only && !inner) continue;
        const size_t c = j * cx + i;
        for (uint32_t n = mid->start[c]; n < mid->start[c + 1]; n++) {
          sink(ctx, i, j, k - 1, n - mid->start[c], &mid->tris[n], inner);
        }
      }
    }
  }
  if (0 == status && !interior_only && cz > 1) {
    /* last slice (k = cz - 1) */
    const slice_t *s = &store[(cz - 1) % 3];
    for (size_t c = 0; c < cx * cy; c++)
      for (uint32_t n = s->start[c]; n < s->start[c + 1]; n++)
        sink(ctx, c % cx, c / cx, cz - 1, n - s->start[c], &s->tris[n], 0);
  }
  for (int n = 0; n < 3; n++) slice_free(&store[n]);
  return status;
}

/* ---- public: case codes ---- */
size_t oracle_case_codes(const oracle_contour_t *c, uint32_t *codes, uint8_t *counts) {
  const view_t w = view_of(c);
  if (w.n[0] < 2 || w.n[1] < 2 || w.n[2] < 2) return 0;
  const size_t cx = w.n[0] - 1, cy = w.n[1] - 1, cz = w.n[2] - 1;
  size_t idx = 0;
  for (size_t k = 0; k < cz; k++) {
    for (size_t j = 0; j < cy; j++) {
      for (size_t i = 0; i < cx; i++, idx++) {
        cube_t cube;
        load_cube(&w, i, j, k, &cube);
        uint32_t packed = 0;
        uint32_t count = 0;
        for (int n = 0; n < 6; n++) {
          uint32_t mask = tetra_mask(&cube, w.threshold, n);
          packed |= mask << (4 * n);
          if (mask >= 8u) mask = 15u - mask;
          count += CASES[mask].n;
        }
        if (codes) codes[idx] = packed;
        if (counts) counts[idx] = (uint8_t)count;
      }
    }
  }
  return idx;
}

/* ---- public: triangle dump ---- */
typedef struct {
  size_t cx, cy;
  size_t capacity, count;
  oracle_triangle_t *out;
} dump_t;

static void dump_sink(void *ctx, size_t i, size_t j, size_t k, uint32_t t, const tri_t *tri, int has_vn) {
  dump_t *d = ctx;
  if (d->count < d->capacity) {
    oracle_triangle_t *o = &d->out[d->count];
    memset(o, 0, sizeof(*o));
    o->cube = (uint64_t)((k * d->cy + j) * d->cx + i);
    o->tri = t;
    o->has_vn = (uint32_t)has_vn;
    o->area = tri->area;
    for (int m = 0; m < 3; m++) {
      o->edge[m] = tri->edge[m];
      o->v[3 * m + 0] = tri->v[m].x; o->v[3 * m + 1] = tri->v[m].y; o->v[3 * m + 2] = tri->v[m].z;
      o->vn[3 * m + 0] = tri->vn[m].x; o->vn[3 * m + 1] = tri->vn[m].y; o->vn[3 * m + 2] = tri->vn[m].z;
    }
    o->fn[0] = tri->fn.x; o->fn[1] = tri->fn.y; o->fn[2] = tri->fn.z;
  }
  d->count++;
}

size_t oracle_tessellate(const oracle_contour_t *c, size_t capacity, oracle_triangle_t *out, oracle_stats_t *stats) {
  const view_t w = view_of(c);
  dump_t d = {w.n[0] - 1, w.n[1] - 1, capacity, 0, out};
  oracle_stats_t local = {0, 0, 0, 0};
  if (0 != march(&w, &local, dump_sink, &d, 0)) return 0;
  if (stats) *stats = local;
  return d.count;
}

/* ---- public: canvas ---- */
int oracle_canvas_clear(const oracle_scene_t *s, double *depth, uint8_t *rgb) {
  /* src/contour3d/main.c:90-103 */
  const size_t n = s->width * s->height;
  for (size_t p = 0; p < n; p++) {
    depth[p] = -1. * DBL_MAX;
    rgb[3 * p + 0] = s->bg[0];
    rgb[3 * p + 1] = s->bg[1];
    rgb[3 * p + 2] = s->bg[2];
  }
  return 0;
}

typedef struct {
  const cam_t *cam;
  const uint8_t *color;
  double *depth;
  uint8_t *rgb;
  uint64_t rendered;
} paint_t;

static void paint_sink(void *ctx, size_t i, size_t j, size_t k, uint32_t t, const tri_t *tri, int has_vn) {
  (void)i; (void)j; (void)k; (void)t; (void)has_vn;
  paint_t *p = ctx;
  raster_triangle(p->cam, p->color, tri, p->depth, p->rgb);
  p->rendered++;
}

static int render_view(const cam_t *cam, const view_t *w, double *depth, uint8_t *rgb, oracle_stats_t *stats) {
  paint_t p = {cam, w->color, depth, rgb, 0};
  oracle_stats_t local = {0, 0, 0, 0};
  if (0 != march(w, &local, paint_sink, &p, 1)) return 1;
  if (stats) {
    stats->cubes += local.cubes;
    stats->active_cubes += local.active_cubes;
    stats->triangles += local.triangles;
    stats->rendered += p.rendered;
  }
  return 0;
}

int oracle_render_contour(const oracle_scene_t *s, const oracle_contour_t *c, double *depth, uint8_t *rgb, oracle_stats_t *stats) {
  const cam_t cam = cam_of(s);
  const view_t w = view_of(c);
  return render_view(&cam, &w, depth, rgb, stats);
}

/* reference: draw_line_between_two_points, line.c:26-118 */
static void raster_segment(const cam_t *c, const oracle_line_t *l, const vec s_ortho, const vec e_ortho, double *depth, uint8_t *rgb) {
  const vec s_cart = l->converter ? l->converter(s_ortho) : s_ortho;
  const vec e_cart = l->converter ? l->converter(e_ortho) : e_ortho;
  vec p0 = {0., 0., 0.}, p1 = {0., 0., 0.};
  if (0 != project(c, s_cart, &p0)) return;
  if (0 != project(c, e_cart, &p1)) return;
  const double W = (double)c->width, H = (double)c->height;
  p0.x = (0.5 + p0.x) * W; p0.y = (0.5 + p0.y) * H;
  p1.x = (0.5 + p1.x) * W; p1.y = (0.5 + p1.y) * H;
  const long xmin = (long)(fmin(p0.x, p1.x) - 1);
  const long xmax = (long)(fmax(p0.x, p1.x) + 1);
  const long ymin = (long)(fmin(p0.y, p1.y) - 1);
  const long ymax = (long)(fmax(p0.y, p1.y) + 1);
  const size_t imin = (size_t)clampl(0, (long)c->width - 1, xmin);
  const size_t imax = (size_t)clampl(0, (long)c->width - 1, xmax);
  const size_t jmin = (size_t)clampl(0, (long)c->height - 1, ymin);
  const size_t jmax = (size_t)clampl(0, (long)c->height - 1, ymax);
  const double a = +p1.y - p0.y;
  const double b = -p1.x + p0.x;
  const double cc = (p1.x - p0.x) * p0.y - (p1.y - p0.y) * p0.x;
  for (size_t j = jmin; j <= jmax; j++) {
    for (size_t i = imin; i <= imax; i++) {
      const double fi = (double)i, fj = (double)j;
      const double limit = pow(0.5 * l->width, 2.);
      const double dist2 = pow(a * fi + b * fj + cc, 2.) / (a * a + b * b);
      if (limit < dist2) continue;
      const double param = -1. * ((p1.x - p0.x) * (p0.x - fi) + (p1.y - p0.y) * (p0.y - fj))
                         / (pow(p1.x - p0.x, 2.) + pow(p1.y - p0.y, 2.));
      const double dist = 1. / (param / p1.z + (1. - param) / p0.z);
      const size_t pix = j * c->width + i;
      if (dist < depth[pix]) continue;
      rgb[3 * pix + 0] = l->color[0];
      rgb[3 * pix + 1] = l->color[1];
      rgb[3 * pix + 2] = l->color[2];
      depth[pix] = dist;
    }
  }
}

/* reference: contour3d_process_line_obj, line.c:120-170 */
int oracle_render_line(const oracle_scene_t *s, const oracle_line_t *l, double *depth, uint8_t *rgb) {
  const cam_t cam = cam_of(s);
  if (l->nitems < 2) {
    fprintf(stderr, "[ORACLE] a line needs at least two points, received %zu\n", l->nitems);
    return 1;
  }
  const vec head = v_from(l->edges[0]);
  const vec tail = v_from(l->edges[1]);
  const vec delta = v_mul(1. / (double)(l->nitems - 1), v_sub(tail, head));
  for (size_t n = 0; n < l->nitems - 1; n++) {
    const vec s0 = v_add(head, v_mul((double)(n + 0), delta));
    const vec s1 = v_add(head, v_mul((double)(n + 1), delta));
    raster_segment(&cam, l, s0, s1, depth, rgb);
  }
  return 0;
}

int oracle_canvas_merge(size_t npixels, double *dst_depth, uint8_t *dst_rgb, const double *src_depth, const uint8_t *src_rgb) {
  for (size_t p = 0; p < npixels; p++) {
    if (dst_depth[p] < src_depth[p]) {
      dst_depth[p] = src_depth[p];
      memcpy(dst_rgb + 3 * p, src_rgb + 3 * p, 3);
    }
  }
  return 0;
}

int oracle_write_ppm(const char *fname, size_t width, size_t height, const uint8_t *rgb) {
  FILE *fp = fopen(fname, "w");
  if (NULL == fp) {
    fprintf(stderr, "[ORACLE] %s: %s\n", fname, strerror(errno));
    return 1;
  }
  fprintf(fp, "P6\n%zu %zu\n255\n", width, height);
  for (size_t j = 0; j < height; j++) {
    fwrite(rgb + 3 * (height - j - 1) * width, 3, width, fp);
  }
  fclose(fp);
  return 0;
}

/* ---- public: whole frame, threads emulating ranks of a y decomposition ---- */

/* window of rank r of p along an axis of n points, after extension by two
   points per interior side (contour/extend.c:216-252; sizes and offsets as
   SimpleDecomp/src/sdecomp/kernel.c:89,111-116) */
static void window(const size_t n, const int p, const int r, size_t *start, size_t *count) {
  size_t off = 0;
  for (int q = 0; q < r; q++) off += (n + (size_t)q) / (size_t)p;
  const size_t mine = (n + (size_t)r) / (size_t)p;
  const size_t lo = r > 0 ? 2 : 0, hi = r < p - 1 ? 2 : 0;
  *start = off - lo;
  *count = mine + lo + hi;
}

typedef struct {
  const oracle_scene_t *scene;
  const cam_t *cam;
  size_t ncontours, nlines;
  const oracle_contour_t *contours;
  const oracle_line_t *lines;
  int rank, nranks;
  double *depth;
  uint8_t *rgb;
  oracle_stats_t stats;
  int failed;
} rank_job_t;

static void *rank_main(void *arg) {
  rank_job_t *job = arg;
  oracle_canvas_clear(job->scene, job->depth, job->rgb);
  for (size_t n = 0; n < job->ncontours; n++) {
    view_t w = view_of(&job->contours[n]);
    size_t start = 0, count = 0;
    window(w.n[1], job->nranks, job->rank, &start, &count);
    w.field += start * w.sy;
    w.gy += start;
    w.n[1] = count;
    if (0 != render_view(job->cam, &w, job->depth, job->rgb, &job->stats)) job->failed = 1;
  }
  /* every rank draws every line, src/contour3d/main.c:119-129 */
  for (size_t n = 0; n < job->nlines; n++) {
    if (0 != oracle_render_line(job->scene, &job->lines[n], job->depth, job->rgb)) job->failed = 1;
  }
  return NULL;
}

int oracle_frame(const oracle_scene_t *s, size_t ncontours, const oracle_contour_t *contours,
                 size_t nlines, const oracle_line_t *lines, int nthreads,
                 double *depth, uint8_t *rgb, oracle_stats_t *stats) {
  const cam_t cam = cam_of(s);
  const size_t npix = s->width * s->height;
  if (nthreads < 1) nthreads = 1;
  for (size_t n = 0; n < ncontours; n++) {
    if (contours[n].dims[1] < (size_t)(4 * nthreads)) nthreads = 1;
  }
  int failed = 0;
  rank_job_t *jobs = calloc((size_t)nthreads, sizeof(rank_job_t));
  pthread_t *tids = calloc((size_t)nthreads, sizeof(pthread_t));
  for (int r = 0; r < nthreads; r++) {
    jobs[r].scene = s; jobs[r].cam = &cam;
    jobs[r].ncontours = ncontours; jobs[r].contours = contours;
    jobs[r].nlines = nlines; jobs[r].lines = lines;
    jobs[r].rank = r; jobs[r].nranks = nthreads;
    jobs[r].depth = 0 == r ? depth : malloc(npix * sizeof(double));
    jobs[r].rgb = 0 == r ? rgb : malloc(3 * npix);
    if (NULL == jobs[r].depth || NULL == jobs[r].rgb) failed = 1;
  }
  if (!failed) {
    for (int r = 1; r < nthreads; r++) pthread_create(&tids[r], NULL, rank_main, &jobs[r]);
    rank_main(&jobs[0]);
    for (int r = 1; r < nthreads; r++) pthread_join(tids[r], NULL);
  }
  oracle_stats_t total = {0, 0, 0, 0};
  for (int r = 0; r < nthreads; r++) {
    failed |= jobs[r].failed;
    total.cubes += jobs[r].stats.cubes;
    total.active_cubes += jobs[r].stats.active_cubes;
    total.triangles += jobs[r].stats.triangles;
    total.rendered += jobs[r].stats.rendered;
    if (r > 0) {
      /* reference: MPI_Reduce with "keep the larger depth", output.c:10-30, 82 */
      if (!failed) oracle_canvas_merge(npix, depth, rgb, jobs[r].depth, jobs[r].rgb);
      free(jobs[r].depth);
      free(jobs[r].rgb);
    }
  }
  free(jobs);
  free(tids);
  if (stats) *stats = total;
  return failed;
}

/* ---- coordinate converters for the tests ----
 * The `converter` callback belongs to the caller (include/contour3d.h:31-33 of the reference); the oracle only calls it
 * where the reference does (triangulate.c:165-170, line.c:39-40).  These three are the ones the case-file driver
 * offers (contour3d_b200/csrc/c3d_case_main.c: cylindrical, stretched, twisted), restated here so that tests can
 * hand the SAME functions to the oracle through its Python binding. */
#include <math.h>

static oracle_vec_t converter_cylindrical(const oracle_vec_t o) {
  const double r = 0.75 + 0.5 * o.x, theta = 1.5 * o.y;
  const oracle_vec_t c = {r * cos(theta) - 0.75, r * sin(theta), o.z};
  return c;
}
static oracle_vec_t converter_stretched(const oracle_vec_t o) {
  const oracle_vec_t c = {o.x + 0.3 * o.x * o.x * o.x, 1.2 * o.y, o.z - 0.2 * o.z * o.z};
  return c;
}
static oracle_vec_t converter_twisted(const oracle_vec_t o) {
  const double a = 0.8 * o.z;
  const oracle_vec_t c = {o.x * cos(a) - o.y * sin(a) + 0.02 * o.z * o.y, o.x * sin(a) + o.y * cos(a), o.z + 0.05 * sin(3. * o.x) * o.y};
  return c;
}

/* kind: 0 identity (NULL), 1 cylindrical, 2 stretched, 3 twisted -- the numbering of the case files */
oracle_converter_t oracle_converter_of(int kind) {
  return 1 == kind ? converter_cylindrical : (2 == kind ? converter_stretched : (3 == kind ? converter_twisted : NULL));
}

/* out[3n..] = converter(in[3n..]) for n points: lets a test tabulate the function without leaving C arithmetic */
void oracle_convert_points(int kind, size_t n, const double *in, double *out) {
  const oracle_converter_t fn = oracle_converter_of(kind);
  for (size_t q = 0; q < n; q++) {
    const oracle_vec_t o = {in[3 * q], in[3 * q + 1], in[3 * q + 2]};
    const oracle_vec_t c = fn ? fn(o) : o;
    out[3 * q] = c.x; out[3 * q + 1] = c.y; out[3 * q + 2] = c.z;
  }
}
