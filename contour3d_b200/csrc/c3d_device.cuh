/*
 * c3d_device.cuh -- device-side geometry shared by all kernels.
 *
 * Numerical contract: every floating-point expression that the reference
 * evaluates after classification (interpolation, cross products, projection,
 * barycentric weights, depth, shading) is evaluated here in fp64 in the SAME
 * operation order with explicitly rounded intrinsics (__dadd_rn, __dmul_rn,
 * __ddiv_rn, __dsqrt_rn), which the compiler never contracts into FMAs.  The
 * reference is built with -std=c99 (no contraction, Makefile:2), so the two
 * agree bit for bit, not just within a tolerance.
 *
 * File:line references are to NaokiHori/Contour3D @ c66854e.
 */
#pragma once

#include "c3d_internal.h"

#define C3D_FULL_MASK 0xffffffffu

struct V3 { double x, y, z; };

__device__ __forceinline__ double r_add(const double a, const double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double r_sub(const double a, const double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double r_mul(const double a, const double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double r_div(const double a, const double b) { return __ddiv_rn(a, b); }

/* vector.c:5-76 */
__device__ __forceinline__ V3 v_add(const V3 a, const V3 b) { return V3{r_add(a.x, b.x), r_add(a.y, b.y), r_add(a.z, b.z)}; }
__device__ __forceinline__ V3 v_sub(const V3 a, const V3 b) { return V3{r_sub(a.x, b.x), r_sub(a.y, b.y), r_sub(a.z, b.z)}; }
__device__ __forceinline__ V3 v_mul(const double w, const V3 a) { return V3{r_mul(w, a.x), r_mul(w, a.y), r_mul(w, a.z)}; }
__device__ __forceinline__ double v_dot(const V3 a, const V3 b) {
  return r_add(r_add(r_mul(a.x, b.x), r_mul(a.y, b.y)), r_mul(a.z, b.z));
}
__device__ __forceinline__ V3 v_cross(const V3 a, const V3 b) {
  return V3{r_sub(r_mul(a.y, b.z), r_mul(a.z, b.y)), r_sub(r_mul(a.z, b.x), r_mul(a.x, b.z)), r_sub(r_mul(a.x, b.y), r_mul(a.y, b.x))};
}
__device__ __forceinline__ V3 v_unit(const V3 a) {
  const double inv = r_div(1., __dsqrt_rn(v_dot(a, a)));
  return V3{r_mul(a.x, inv), r_mul(a.y, inv), r_mul(a.z, inv)};
}
__device__ __forceinline__ V3 v_of(const double p[3]) { return V3{p[0], p[1], p[2]}; }

/* total order on doubles as unsigned integers (larger double <-> larger key) */
__device__ __forceinline__ unsigned long long ord_of(const double d) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(d);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord_to_double(const unsigned long long o) {
  const unsigned long long b = (o >> 63) ? (o & 0x7fffffffffffffffull) : ~o;
  return __longlong_as_double((long long)b);
}

__device__ __forceinline__ double ld_value(const FieldView &f, const long long idx) {
  return 4 == f.elem_bytes ? (double)__ldg((const float *)f.data + idx) : __ldg((const double *)f.data + idx);
}

/* Cartesian position of grid point (i, j, k) under a tabulated coordinate converter (PosMap, c3d_internal.h) */
__device__ __forceinline__ V3 mapped_pos(const PosMap *__restrict__ m, const long long i, const long long j, const long long k) {
  return V3{__ldg(m->pos[0] + (i * m->ps[0][0] + j * m->ps[0][1] + k * m->ps[0][2])),
            __ldg(m->pos[1] + (i * m->ps[1][0] + j * m->ps[1][1] + k * m->ps[1][2])),
            __ldg(m->pos[2] + (i * m->ps[2][0] + j * m->ps[2][1] + k * m->ps[2][2]))};
}

/* eight corner scalars + the two coordinates per axis (identity converter) or the
   cube's place in the position tables (any other converter);
   triangulate.c:156-173, 302-317 */
struct Cube {
  double val[8];
  double px[2], py[2], pz[2];
  const PosMap *map;
  long long i, j, k;
};

__device__ __forceinline__ void load_cube(const FieldView &f, const long long i, const long long j, const long long k, Cube &c) {
  const long long base = k * f.sz + j * f.sy + i;
#pragma unroll
  for (int n = 0; n < 8; n++) {
    c.val[n] = ld_value(f, base + (n & 1) + ((n >> 1) & 1) * f.sy + (n >> 2) * f.sz);
  }
  c.map = f.map;
  c.i = i; c.j = j; c.k = k;
  c.px[0] = c.px[1] = c.py[0] = c.py[1] = c.pz[0] = c.pz[1] = 0.;
  if (nullptr == f.map) {
    c.px[0] = __ldg(f.g[0] + i); c.px[1] = __ldg(f.g[0] + i + 1);
    c.py[0] = __ldg(f.g[1] + j); c.py[1] = __ldg(f.g[1] + j + 1);
    c.pz[0] = __ldg(f.g[2] + k); c.pz[1] = __ldg(f.g[2] + k + 1);
  }
}

/* corner bits: bit n set when thr < value[n] (strict), triangulate.c:176-179 */
__device__ __forceinline__ unsigned int corner_code(const Cube &c, const double thr) {
  unsigned int code = 0;
#pragma unroll
  for (int n = 0; n < 8; n++) code |= (unsigned int)(thr < c.val[n]) << n;
  return code;
}

__device__ __forceinline__ V3 corner_pos(const Cube &c, const unsigned int n) {
  if (nullptr != c.map) return mapped_pos(c.map, c.i + (n & 1u), c.j + ((n >> 1) & 1u), c.k + (n >> 2));
  return V3{c.px[n & 1u], c.py[(n >> 1) & 1u], c.pz[(n >> 2) & 1u]};
}

/* triangle vertex on the edge tail -> head, triangulate.c:40-73 */
__device__ __forceinline__ V3 cut_edge(const Cube &c, const double thr, const unsigned int tail, const unsigned int head) {
  const double v0 = c.val[tail], v1 = c.val[head];
  const V3 p0 = corner_pos(c, tail);
  const double dv = r_sub(v1, v0);
  if (fabs(dv) < 1.e-8) return p0;
  const double factor = r_div(r_sub(thr, v0), dv);
  const V3 p1 = corner_pos(c, head);
  return v_add(p0, v_mul(factor, v_sub(p1, p0)));
}

/* the same vertex straight from global memory (no register-indexed cube copy):
   used where a thread needs one triangle of a cube, not all of them */
/* MAPPED = false compiles the converter-table branch out (the raster's hot path when no object of the frame has one) */
template <bool MAPPED>
__device__ __forceinline__ V3 cut_edge_global_t(const FieldView &f, const long long i, const long long j, const long long k,
                                                const double thr, const unsigned int tail, const unsigned int head) {
  const long long base = k * f.sz + j * f.sy + i;
  const double v0 = ld_value(f, base + (tail & 1u) + ((tail >> 1) & 1u) * f.sy + (tail >> 2) * f.sz);
  const double v1 = ld_value(f, base + (head & 1u) + ((head >> 1) & 1u) * f.sy + (head >> 2) * f.sz);
  const double dv = r_sub(v1, v0);
  if (MAPPED && nullptr != f.map) {      /* converted coordinates: interpolation happens in Cartesian space (triangulate.c:165-170) */
    const V3 m0 = mapped_pos(f.map, i + (tail & 1u), j + ((tail >> 1) & 1u), k + (tail >> 2));
    if (fabs(dv) < 1.e-8) return m0;
    const double mfactor = r_div(r_sub(thr, v0), dv);
    const V3 m1 = mapped_pos(f.map, i + (head & 1u), j + ((head >> 1) & 1u), k + (head >> 2));
    return v_add(m0, v_mul(mfactor, v_sub(m1, m0)));
  }
  const V3 p0 = V3{__ldg(f.g[0] + i + (tail & 1u)), __ldg(f.g[1] + j + ((tail >> 1) & 1u)), __ldg(f.g[2] + k + (tail >> 2))};
  if (fabs(dv) < 1.e-8) return p0;
  const double factor = r_div(r_sub(thr, v0), dv);
  const V3 p1 = V3{__ldg(f.g[0] + i + (head & 1u)), __ldg(f.g[1] + j + ((head >> 1) & 1u)), __ldg(f.g[2] + k + (head >> 2))};
  return v_add(p0, v_mul(factor, v_sub(p1, p0)));
}

__device__ __forceinline__ V3 cut_edge_global(const FieldView &f, const long long i, const long long j, const long long k,
                                              const double thr, const unsigned int tail, const unsigned int head) {
  return cut_edge_global_t<true>(f, i, j, k, thr, tail, head);
}

__device__ __forceinline__ void triangle_vertices(const Cube &c, const double thr, const unsigned int pairs, V3 v[3]) {
#pragma unroll
  for (int m = 0; m < 3; m++) {
    v[m] = cut_edge(c, thr, (pairs >> (6 * m)) & 7u, (pairs >> (6 * m + 3)) & 7u);
  }
}

/* face normal and |cross|, triangulate.c:75-93 */
__device__ __forceinline__ void face_normal(const V3 v[3], V3 &fn, double &area) {
  const V3 c = v_cross(v_sub(v[1], v[0]), v_sub(v[2], v[0]));
  area = __dsqrt_rn(v_dot(c, c));
  fn = v_mul(r_div(1., area), c);
}

/* add area * face_normal of the triangles of cube (i, j, k) that touch its
   local edge `edge`, in triangle order; average(), normal.c:6-43 */
static __device__ __noinline__ void gather_cube(const FieldView &f, const DevTables &t, const long long i, const long long j, const long long k,
                                         const unsigned int edge, V3 &sum) {
  Cube c;
  load_cube(f, i, j, k, c);
  const unsigned int code = corner_code(c, f.thr);
  const unsigned int n = __ldg(t.count + code);
  for (unsigned int s = 0; s < n; s++) {
    if ((__ldg(t.emask + code * C3D_MAX_TRIS + s) >> edge) & 1u) {
      V3 v[3], fn;
      double area;
      triangle_vertices(c, f.thr, __ldg(t.pairs + code * C3D_MAX_TRIS + s), v);
      face_normal(v, fn, area);
      sum = v_add(sum, v_mul(area, fn));
    }
  }
}

/* vertex normal of the vertex on edge e of cube (i, j, k): the cubes sharing
   the edge are visited in global (k, j, i) order with the edge renamed to each
   cube's local id; the 19-way branch of normal.c:78-158, then normalise (159) */
static __device__ __noinline__ V3 vertex_normal(const FieldView &f, const DevTables &t, const long long i, const long long j, const long long k,
                                         const unsigned int e) {
  V3 sum = V3{0., 0., 0.};
  if (e < 12u) {
    const unsigned int axis = e >> 2, a = e & 1u, b = (e >> 1) & 1u;
    for (unsigned int q = 0; q < 2; q++) {
      for (unsigned int p = 0; p < 2; p++) {
        const long long da = (long long)a - 1 + p, db = (long long)b - 1 + q;
        const unsigned int local = 4u * axis + (1u - p) + 2u * (1u - q);
        long long ii = i, jj = j, kk = k;
        if (0u == axis) { jj += da; kk += db; }
        else if (1u == axis) { ii += da; kk += db; }
        else { ii += da; jj += db; }
        gather_cube(f, t, ii, jj, kk, local, sum);
      }
    }
  } else if (e < 18u) {
    const unsigned int kind = (e - 12u) >> 1, a = e & 1u;
    for (unsigned int p = 0; p < 2; p++) {
      const long long da = (long long)a - 1 + p;
      const unsigned int local = 12u + 2u * kind + (1u - p);
      long long ii = i, jj = j, kk = k;
      if (0u == kind) ii += da;
      else if (1u == kind) jj += da;
      else kk += da;
      gather_cube(f, t, ii, jj, kk, local, sum);
    }
  } else {
    gather_cube(f, t, i, j, k, 18u, sum);
  }
  return v_unit(sum);
}

/* project.c:9-67; returns false when the point is rejected */
__device__ __forceinline__ bool project(const Cam &c, const V3 p, V3 &out) {
  const V3 cam = v_of(c.cam), nrm = v_of(c.normal);
  const V3 ray = v_sub(p, cam);
  const double nr = v_dot(nrm, ray);
  if (-1.e-8 < nr && nr < 1.e-8) return false;
  const double tt = r_div(r_sub(c.d, c.nc), nr);
  if (tt <= 0. || 1. <= tt) return false;
  const V3 delta = V3{
    r_sub(r_add(c.cam[0], r_mul(tt, ray.x)), c.center[0]),
    r_sub(r_add(c.cam[1], r_mul(tt, ray.y)), c.center[1]),
    r_sub(r_add(c.cam[2], r_mul(tt, ray.z)), c.center[2]),
  };
  out.x = r_div(v_dot(delta, v_of(c.lx)), c.lxlx);
  out.y = r_div(v_dot(delta, v_of(c.ly)), c.lyly);
  out.z = nr;
  return true;
}

/* a projected triangle prepared for per-pixel tests (render.c:74-115).
   The differences (b - a) of the three edge functions are per-triangle
   constants of render.c:38-46; hoisting them changes no rounding. */
struct Tri2D {
  V3 q[3];
  double ey[3], ex[3];     /* edge m: a = q[(m+1)%3], b = q[(m+2)%3]; ey = b.y - a.y, ex = b.x - a.x */
  double area_inv;
  long long imin, imax, jmin, jmax;
};

__device__ __forceinline__ long long clamp_ll(const long long lo, const long long hi, const long long v) {
  const long long m = v < lo ? lo : v;
  return hi < m ? hi : m;
}

/* bounding box, edge differences and inverse area of a triangle whose projected vertices t.q[] are set */
__device__ __forceinline__ void setup_projected(const Cam &c, Tri2D &t) {
  const double xlo = fmin(t.q[0].x, fmin(t.q[1].x, t.q[2].x));
  const double xhi = fmax(t.q[0].x, fmax(t.q[1].x, t.q[2].x));
  const double ylo = fmin(t.q[0].y, fmin(t.q[1].y, t.q[2].y));
  const double yhi = fmax(t.q[0].y, fmax(t.q[1].y, t.q[2].y));
  /* truncation toward zero BEFORE clamping, render.c:81-89 */
  const long long xmin = __double2ll_rz(r_sub(r_mul(r_add(0.5, xlo), c.W), 1.));
  const long long xmax = __double2ll_rz(r_add(r_mul(r_add(0.5, xhi), c.W), 1.));
  const long long ymin = __double2ll_rz(r_sub(r_mul(r_add(0.5, ylo), c.H), 1.));
  const long long ymax = __double2ll_rz(r_add(r_mul(r_add(0.5, yhi), c.H), 1.));
  t.imin = clamp_ll(0, c.w - 1, xmin);
  t.imax = clamp_ll(0, c.w - 1, xmax);
  t.jmin = clamp_ll(0, c.h - 1, ymin);
  t.jmax = clamp_ll(0, c.h - 1, ymax);
#pragma unroll
  for (int m = 0; m < 3; m++) {
    const V3 &a = t.q[(m + 1) % 3], &b = t.q[(m + 2) % 3];
    t.ey[m] = r_sub(b.y, a.y);
    t.ex[m] = r_sub(b.x, a.x);
  }
  /* area = edge_function(v0, v1, v2), render.c:102: a = q0, b = q1 (edge 2), evaluated at q2 */
  const double area = r_sub(r_mul(r_sub(t.q[2].x, t.q[0].x), t.ey[2]), r_mul(r_sub(t.q[2].y, t.q[0].y), t.ex[2]));
  t.area_inv = r_div(1., area < 0. ? fmin(area, -1.e-16) : fmax(area, +1.e-16));
}

__device__ __forceinline__ bool setup_triangle(const Cam &c, const V3 v[3], Tri2D &t) {
#pragma unroll
  for (int m = 0; m < 3; m++) {
    if (!project(c, v[m], t.q[m])) return false;    /* render.c:62-72 */
  }
  setup_projected(c, t);
  return true;
}

/* pixel-centre coordinate, render.c:94-98 (tabulated per screen size by k_pixel_tables) */
__device__ __forceinline__ double pixel_centre(const long long i, const double n) {
  return r_sub(r_div(r_add((double)i, 0.5), n), 0.5);
}

/* barycentric weights at screen position (px, py); false when outside (render.c:94-115) */
__device__ __forceinline__ bool pixel_weights(const Tri2D &t, const double px, const double py, double w[3]) {
#pragma unroll
  for (int m = 0; m < 3; m++) {
    const V3 &a = t.q[(m + 1) % 3];
    /* edge_function(a, b, p), render.c:38-46 */
    w[m] = r_mul(t.area_inv, r_sub(r_mul(r_sub(px, a.x), t.ey[m]), r_mul(r_sub(py, a.y), t.ex[m])));
  }
  const double slack = -1.e-8;
  return !(w[0] < slack || w[1] < slack || w[2] < slack);
}

/* harmonic depth, render.c:120-124 */
__device__ __forceinline__ double pixel_depth(const Tri2D &t, const double w[3]) {
  return r_div(1., r_add(r_add(r_div(w[0], t.q[0].z), r_div(w[1], t.q[1].z)), r_div(w[2], t.q[2].z)));
}

/* 128-bit compare-and-swap on a z-buffer entry (ATOMG.E.CAS.128) */
__device__ __forceinline__ Pixel cas_pixel(Pixel *addr, const Pixel expect, const Pixel desired) {
  Pixel old;
  asm volatile(
      "{\n\t.reg .b128 c, n, d;\n\t"
      "mov.b128 c, {%2, %3};\n\t"
      "mov.b128 n, {%4, %5};\n\t"
      "atom.global.cas.b128 d, [%6], c, n;\n\t"
      "mov.b128 {%0, %1}, d;\n\t}"
      : "=l"(old.depth), "=l"(old.key)
      : "l"(expect.depth), "l"(expect.key), "l"(desired.depth), "l"(desired.key), "l"(addr)
      : "memory");
  return old;
}

/* canvas[pix] = lexicographic max(canvas[pix], (o, key)): the reference's
   sequential "if (dist < stored) skip; else overwrite" (render.c:125-132)
   made independent of the order in which fragments arrive */
__device__ __forceinline__ void submit_fragment(Pixel *slot, const unsigned long long o, const unsigned long long key) {
  const ulonglong2 seen = __ldcg(reinterpret_cast<const ulonglong2 *>(slot));   /* a hint (L2); the CAS decides */
  Pixel cur{seen.x, seen.y};
  while (o > cur.depth || (o == cur.depth && key > cur.key)) {
    const Pixel prev = cas_pixel(slot, cur, Pixel{o, key});
    if (prev.depth == cur.depth && prev.key == cur.key) break;
    cur = prev;
  }
}

/* Out-of-line on purpose: the raster and shading kernels are large, and their warps sit in different phases;
   one shared copy of "cut the edge and project" keeps the hot code within reach of the instruction cache. */
/* triangle vertex on the cube edge tail -> head (triangulate.c:40-73), projected (project.c:9-67);
   returned by value (registers), .w = 1 when the projection accepted the point */
struct ProjectedVertex { double x, y, z, w; };
template <bool MAPPED>
static __device__ __noinline__ ProjectedVertex exact_vertex_t(const Cam &cam, const FieldView &f, const long long i, const long long j,
                                                              const long long k, const unsigned int tail, const unsigned int head) {
  V3 q = V3{0., 0., 0.};
  const bool ok = project(cam, cut_edge_global_t<MAPPED>(f, i, j, k, f.thr, tail, head), q);
  return ProjectedVertex{q.x, q.y, q.z, ok ? 1. : 0.};
}
/* the same from the two corner values and positions already in registers: callers that need several vertices of a cube
   issue all their loads first (one exposure to the memory latency instead of one per vertex -- a peer read over NVLink
   takes microseconds) and then come here */
static __device__ __noinline__ ProjectedVertex exact_vertex_from(const Cam &cam, const double thr, const double v0, const double v1,
                                                                 const double x0, const double y0, const double z0,
                                                                 const double x1, const double y1, const double z1) {
  const V3 p0 = V3{x0, y0, z0};
  const double dv = r_sub(v1, v0);
  V3 c = p0;
  if (!(fabs(dv) < 1.e-8)) {                       /* triangulate.c:62-71 */
    const double factor = r_div(r_sub(thr, v0), dv);
    c = v_add(p0, v_mul(factor, v_sub(V3{x1, y1, z1}, p0)));
  }
  V3 q = V3{0., 0., 0.};
  const bool ok = project(cam, c, q);
  return ProjectedVertex{q.x, q.y, q.z, ok ? 1. : 0.};
}
__device__ __forceinline__ ProjectedVertex exact_vertex(const FrameParams &p, const unsigned int obj, const long long i, const long long j,
                                                        const long long k, const unsigned int tail, const unsigned int head) {
  return exact_vertex_t<true>(p.cam, p.view[obj], i, j, k, tail, head);
}

/* the view a list entry refers to: this rank's own array, or (screen-partitioned frames) the array of the rank that listed it */
/* staged: a shared-memory copy of the views in use, indexed src * ncontours + obj (nullptr: the device table p.pviews) */
template <bool PEER>
__device__ __forceinline__ const FieldView &entry_view(const FrameParams &p, const FieldView *staged, const unsigned long long entry) {
  if (PEER) {
    if (nullptr != staged) return staged[C3D_ENTRY_SRC(entry) * p.ncontours + C3D_ENTRY_OBJ(entry)];
    return p.pviews[C3D_PEER_VIEW(C3D_ENTRY_SRC(entry), C3D_ENTRY_OBJ(entry))];
  }
  return p.view[C3D_ENTRY_OBJ(entry)];
}

/* winner key of triangle s of cube (i, j, k) of view f (c3d_internal.h): global cube indices */
__device__ __forceinline__ unsigned long long winner_key(const FieldView &f, const unsigned int obj, const long long i, const long long j,
                                                         const long long k, const unsigned int s) {
  return ((unsigned long long)(obj + 1u) << C3D_KEY_OBJ_SHIFT) | ((unsigned long long)(k + f.off[2]) << 36)
       | ((unsigned long long)(j + f.off[1]) << 20) | ((unsigned long long)(i + f.off[0]) << 4) | s;
}
