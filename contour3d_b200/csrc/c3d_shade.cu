/*
 * c3d_shade.cu -- deferred shading of the winning triangle of every pixel.
 *
 * Replaces contour3d_contour_compute_vertex_normals + average (contour/normal.c:6-165)
 * and the colour half of contour3d_contour_render_triangle (contour/render.c:137-159).
 * The reference smooths EVERY triangle of every slice; here a vertex normal is
 * computed only for vertices of triangles that own a pixel, and only ONCE per
 * vertex: a vertex lives on a cube edge (triangulate.c:15-32), its normal is the
 * normalised sum of `area * face_normal` over the triangles that touch that edge
 * in the 1, 2 or 4 cubes sharing it, visited in a fixed global order
 * (normal.c:78-158) -- a function of the GLOBAL edge alone.  At 1024^3 on a 4K
 * screen ~20 pixels reference the same vertex, so the pixels first agree on the
 * set of distinct vertices through a hash table keyed by the global edge.
 * Sums are evaluated in the reference's order: bit-identical, and independent of
 * any atomic ordering (the table only decides WHERE a normal is stored).
 *
 *   k_shade_select   one thread per pixel: background, lines and pixels lost to
 *                    another rank are written at once; pixels won by a contour
 *                    triangle are compacted into a work list
 *   k_shade_setup    one thread per listed pixel: rebuilds the winning triangle,
 *                    the barycentric weights of the pixel centre (render.c:102-110)
 *                    and enters its three vertices into the table
 *   k_shade_normals  eight lanes per distinct vertex: the face terms of the
 *                    touching triangles in parallel, added up in order, normalised
 *   k_shade_blend    one thread per listed pixel: blend, light, truncate
 *                    (render.c:137-159)
 *
 * File:line references are to NaokiHori/Contour3D @ c66854e.
 */
#include <float.h>
#include "c3d_device.cuh"


/* ------------------------------------------------------------------ select */

__global__ void __launch_bounds__(256) k_shade_select(const __grid_constant__ FrameParams p, const LineDev *__restrict__ lines,
                                                      const Pixel *__restrict__ canvas, const unsigned long long *__restrict__ depth_global,
                                                      unsigned int *__restrict__ rgba, const size_t npix, unsigned int *__restrict__ work,
                                                      unsigned int *__restrict__ nwork) {
  const unsigned int lane = threadIdx.x & 31u;
  for (size_t base = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) & ~(size_t)31; base < npix; base += (size_t)gridDim.x * blockDim.x) {
    const size_t pix = base + lane;
    bool mine = false;
    if (pix < npix) {
      const ulonglong2 px = *reinterpret_cast<const ulonglong2 *>(canvas + pix);
      unsigned int out = 0u;     /* 0 = nothing from this rank: background, or another rank is nearer */
      if (0ull != px.y && (nullptr == depth_global || px.x == depth_global[pix])) {
        const unsigned int obj = (unsigned int)(px.y >> C3D_KEY_OBJ_SHIFT) - 1u;
        if (obj >= p.ncontours) {
          const LineDev &ln = lines[obj - p.ncontours];
          out = 0xFF000000u | ((unsigned int)ln.color[2] << 16) | ((unsigned int)ln.color[1] << 8) | ln.color[0];
        } else {
          mine = true;
        }
      }
      if (!mine) rgba[pix] = out;
    }
    const unsigned int who = __ballot_sync(C3D_FULL_MASK, mine);
    if (0u != who) {
      unsigned int at = 0u;
      if (0u == lane) at = atomicAdd(nwork, (unsigned int)__popc(who));
      at = __shfl_sync(C3D_FULL_MASK, at, 0);
      if (mine) work[at + __popc(who & ((1u << lane) - 1u))] = (unsigned int)pix;
    }
  }
}

/* ----------------------------------------------------------------- geometry */

/* the c-th cube (reference visiting order) sharing edge e of cube (i, j, k), and the edge's id in that cube;
   the 19-way branch of normal.c:78-158 */
__device__ __forceinline__ bool edge_neighbour(const unsigned int e, const unsigned int c, int &di, int &dj, int &dk, unsigned int &local) {
  di = dj = dk = 0;
  if (e < 12u) {
    const unsigned int axis = e >> 2, a = e & 1u, b = (e >> 1) & 1u, pp = c & 1u, qq = c >> 1;
    const int da = (int)a - 1 + (int)pp, db = (int)b - 1 + (int)qq;
    local = 4u * axis + (1u - pp) + 2u * (1u - qq);
    if (0u == axis) { dj = da; dk = db; }
    else if (1u == axis) { di = da; dk = db; }
    else { di = da; dj = db; }
    return true;
  }
  if (e < 18u) {
    if (c >= 2u) return false;
    const unsigned int kind = (e - 12u) >> 1, a = e & 1u;
    const int da = (int)a - 1 + (int)c;
    local = 12u + 2u * kind + (1u - c);
    if (0u == kind) di = da;
    else if (1u == kind) dj = da;
    else dk = da;
    return true;
  }
  local = 18u;
  return 0u == c;
}

__device__ __forceinline__ unsigned int cube_code(const FieldView &f, const long long i, const long long j, const long long k) {
  const long long base = k * f.sz + j * f.sy + i;
  unsigned int code = 0u;
#pragma unroll
  for (int n = 0; n < 8; n++) {
    code |= (unsigned int)(f.thr < ld_value(f, base + (n & 1) + ((n >> 1) & 1) * f.sy + (n >> 2) * f.sz)) << n;
  }
  return code;
}

/* area * face_normal of triangle s of a cube (triangulate.c:75-93, normal.c:39-40) */
__device__ __forceinline__ V3 face_term(const FrameParams &p, const unsigned int obj, const long long i, const long long j, const long long k,
                                        const unsigned int code, const unsigned int s) {
  const FieldView &f = p.view[obj];
  const unsigned int pairs = __ldg(p.tab.pairs + code * C3D_MAX_TRIS + s);
  V3 v[3], fn;
  double area;
#pragma unroll
  for (int m = 0; m < 3; m++) v[m] = cut_edge_global(f, i, j, k, f.thr, (pairs >> (6 * m)) & 7u, (pairs >> (6 * m + 3)) & 7u);
  face_normal(v, fn, area);
  return v_mul(area, fn);
}


/* ------------------------------------------------------------------- setup */

struct ShadePix {              /* 56 bytes per listed pixel */
  double w[3];                 /* barycentric weights of the pixel centre (render.c:102-110) */
  unsigned int vslot[3];       /* table slots of the three vertices; C3D_NO_SLOT: table full, normal computed on the spot */
  unsigned int pix;
  unsigned short i, j, k;      /* cube of the winning triangle */
  unsigned char obj, edges[3]; /* cube-edge ids of its three vertices */
};
static_assert(sizeof(ShadePix) == C3D_SHADE_PIX_BYTES, "ShadePix layout");

#define C3D_NO_SLOT 0xffffffffu
#define SH_PROBES 32

/* the global edge a cube-local edge id names: the cube whose edge 0 / 4 / 8 / 12 / 14 / 16 / 18 it is (the numbering of
   triangulate.c:15-32: 0-3 x edges at (dj, dk), 4-7 y edges at (di, dk), 8-11 z edges at (di, dj), 12/13 yz-face
   diagonals at di, 14/15 xz at dj, 16/17 xy at dk, 18 the body diagonal), packed as
   1:1 | object:4 | type:3 | k:16 | j:16 | i:16 (never 0 = empty slot) */
__device__ __forceinline__ unsigned long long vertex_key(const unsigned int obj, const unsigned int i, const unsigned int j, const unsigned int k,
                                                         const unsigned int e) {
  unsigned int ci = i, cj = j, ck = k, type;
  if (e < 12u) {
    const unsigned int axis = e >> 2, a = e & 1u, b = (e >> 1) & 1u;
    type = axis;
    if (0u == axis) { cj += a; ck += b; }
    else if (1u == axis) { ci += a; ck += b; }
    else { ci += a; cj += b; }
  } else if (e < 18u) {
    const unsigned int kind = (e - 12u) >> 1, a = e & 1u;
    type = 3u + kind;
    if (0u == kind) ci += a;
    else if (1u == kind) cj += a;
    else ck += a;
  } else {
    type = 6u;
  }
  return (1ull << 63) | ((unsigned long long)obj << 51) | ((unsigned long long)type << 48) | ((unsigned long long)ck << 32)
       | ((unsigned long long)cj << 16) | (unsigned long long)ci;
}

__global__ void __launch_bounds__(128) k_shade_setup(const __grid_constant__ FrameParams p, const Pixel *__restrict__ canvas,
                                                     const unsigned int *__restrict__ work, unsigned int *__restrict__ counters,
                                                     ShadePix *__restrict__ out, unsigned long long *__restrict__ table, const unsigned int mask,
                                                     unsigned int *__restrict__ vlist, unsigned long long *__restrict__ overflowed) {
  const unsigned int lane = threadIdx.x & 31u;
  const unsigned int n = counters[0];
  /* whole warps iterate together (ballots inside) */
  for (unsigned int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < n; base += gridDim.x * blockDim.x) {
    const unsigned int idx = base + lane;
    const bool valid = idx < n;
    ShadePix me;
    me.w[0] = me.w[1] = me.w[2] = 0.;
    me.vslot[0] = me.vslot[1] = me.vslot[2] = C3D_NO_SLOT;
    me.pix = 0u; me.i = me.j = me.k = 0; me.obj = 0; me.edges[0] = me.edges[1] = me.edges[2] = 0;
    unsigned long long vkey[3] = {0ull, 0ull, 0ull};
    if (valid) {
      const unsigned int pix = __ldg(work + idx);
      const unsigned long long key = canvas[pix].key;
      const unsigned int obj = (unsigned int)(key >> C3D_KEY_OBJ_SHIFT) - 1u;
      const unsigned long long prim = key & C3D_KEY_PRIM_MASK;
      const unsigned int s = (unsigned int)(prim & 15ull);
      const long long i = (long long)((prim >> 4) & 65535ull), j = (long long)((prim >> 20) & 65535ull), k = (long long)((prim >> 36) & 65535ull);
      const FieldView &f = p.view[obj];
      const unsigned int code = cube_code(f, i, j, k);
      const unsigned int pairs = __ldg(p.tab.pairs + code * C3D_MAX_TRIS + s);
      const unsigned int edges = __ldg(p.tab.edges + code * C3D_MAX_TRIS + s);
      Tri2D t;
      bool ok = true;
#pragma unroll
      for (int m = 0; m < 3; m++) {
        const ProjectedVertex v = exact_vertex(p, obj, i, j, k, (pairs >> (6 * m)) & 7u, (pairs >> (6 * m + 3)) & 7u);
        ok = ok && (0. != v.w);
        t.q[m] = V3{v.x, v.y, v.z};
      }
      if (ok) {
        setup_projected(p.cam, t);
        pixel_weights(t, __ldg(p.px + (pix % (unsigned int)p.cam.w)), __ldg(p.py + (pix / (unsigned int)p.cam.w)), me.w);
      }
      me.pix = pix;
      me.i = (unsigned short)i; me.j = (unsigned short)j; me.k = (unsigned short)k;
      me.obj = (unsigned char)obj;
#pragma unroll
      for (int m = 0; m < 3; m++) {
        me.edges[m] = (unsigned char)((edges >> (5 * m)) & 31u);
        vkey[m] = vertex_key(obj, (unsigned int)i, (unsigned int)j, (unsigned int)k, me.edges[m]);
      }
    }
#pragma unroll
    for (int m = 0; m < 3; m++) {
      unsigned int slot = C3D_NO_SLOT;
      bool inserted = false;
      if (valid) {
        unsigned int h = (unsigned int)((vkey[m] * 0x9E3779B97F4A7C15ull) >> 32) & mask;
        for (int probe = 0; probe < SH_PROBES; probe++) {
          const unsigned long long old = atomicCAS(table + h, 0ull, vkey[m]);
          if (0ull == old) { inserted = true; slot = h; break; }
          if (old == vkey[m]) { slot = h; break; }
          h = (h + 1u) & mask;
        }
        if (C3D_NO_SLOT == slot) atomicAdd(overflowed, 1ull);
      }
      /* the first pixel to name a vertex lists it for k_shade_normals */
      const unsigned int who = __ballot_sync(C3D_FULL_MASK, inserted);
      if (0u != who) {
        unsigned int at = 0u;
        if (0u == lane) at = atomicAdd(counters + 1, (unsigned int)__popc(who));
        at = __shfl_sync(C3D_FULL_MASK, at, 0);
        if (inserted) vlist[at + __popc(who & ((1u << lane) - 1u))] = slot;
      }
      me.vslot[m] = slot;
    }
    if (valid) out[idx] = me;
  }
}

/* ----------------------------------------------------------------- normals */

/* the n-th (0-based) set bit of x */
__device__ __forceinline__ unsigned int nth_set_bit(unsigned int x, unsigned int n) {
  for (; n > 0u; n--) x &= x - 1u;
  return (unsigned int)__ffs((int)x) - 1u;
}

__global__ void __launch_bounds__(128) k_shade_normals(const __grid_constant__ FrameParams p, const unsigned long long *__restrict__ table,
                                                       const unsigned int *__restrict__ vlist, const unsigned int *__restrict__ counters,
                                                       double *__restrict__ vn) {
  const unsigned int lane = threadIdx.x & 31u, group = lane >> 3, sub = lane & 7u, lead = group << 3;
  const unsigned int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const unsigned int nv = counters[1];
  for (unsigned int base = warp * 4u; base < nv; base += nwarps * 4u) {      /* warp-uniform */
    const unsigned int v = base + group;
    const bool valid = v < nv;
    unsigned int slot = 0u, obj = 0u, e = 18u;
    long long ci = 0, cj = 0, ck = 0;
    if (valid) {
      slot = __ldg(vlist + v);
      const unsigned long long key = table[slot];
      const unsigned int type = (unsigned int)(key >> 48) & 7u;
      obj = (unsigned int)(key >> 51) & 15u;
      e = type < 3u ? 4u * type : (type < 6u ? 6u + 2u * type : 18u);      /* edge 0 / 4 / 8 / 12 / 14 / 16 / 18 of cube (ci, cj, ck) */
      ck = (long long)((key >> 32) & 65535ull); cj = (long long)((key >> 16) & 65535ull); ci = (long long)(key & 65535ull);
    }
    /* lanes 0..3 of the group: the cubes sharing the edge, in the reference's visiting order */
    long long ii = 0, jj = 0, kk = 0;
    unsigned int code = 0u, touch = 0u;
    {
      int di, dj, dk;
      unsigned int local;
      if (valid && sub < 4u && edge_neighbour(e, sub, di, dj, dk, local)) {
        ii = ci + di; jj = cj + dj; kk = ck + dk;
        code = cube_code(p.view[obj], ii, jj, kk);
        touch = __ldg(p.tab.touch + code * 19u + local);
      }
    }
    const unsigned int cnt = __popc(touch);
    const unsigned int c0 = __shfl_sync(C3D_FULL_MASK, cnt, lead), c1 = __shfl_sync(C3D_FULL_MASK, cnt, lead + 1u);
    const unsigned int c2 = __shfl_sync(C3D_FULL_MASK, cnt, lead + 2u), c3 = __shfl_sync(C3D_FULL_MASK, cnt, lead + 3u);
    const unsigned int total = c0 + c1 + c2 + c3;
    const unsigned int most = __reduce_max_sync(C3D_FULL_MASK, total);
    V3 sum = V3{0., 0., 0.};
    for (unsigned int it0 = 0u; it0 < most; it0 += 8u) {
      /* item it0 + sub = the within-th touching triangle of sharing cube c */
      const unsigned int item = it0 + sub;
      const bool have = valid && item < total;
      const unsigned int c = item < c0 ? 0u : (item < c0 + c1 ? 1u : (item < c0 + c1 + c2 ? 2u : 3u));
      const unsigned int within = item - (0u == c ? 0u : (1u == c ? c0 : (2u == c ? c0 + c1 : c0 + c1 + c2)));
      const long long sii = __shfl_sync(C3D_FULL_MASK, ii, lead + c), sjj = __shfl_sync(C3D_FULL_MASK, jj, lead + c);
      const long long skk = __shfl_sync(C3D_FULL_MASK, kk, lead + c);
      const unsigned int scode = __shfl_sync(C3D_FULL_MASK, code, lead + c), stouch = __shfl_sync(C3D_FULL_MASK, touch, lead + c);
      V3 term = V3{0., 0., 0.};
      if (have) term = face_term(p, obj, sii, sjj, skk, scode, nth_set_bit(stouch, within));
      /* add in (cube, triangle) order, average() of normal.c:6-43 */
#pragma unroll
      for (unsigned int t = 0u; t < 8u; t++) {
        const V3 other = V3{__shfl_sync(C3D_FULL_MASK, term.x, lead + t), __shfl_sync(C3D_FULL_MASK, term.y, lead + t),
                            __shfl_sync(C3D_FULL_MASK, term.z, lead + t)};
        if (it0 + t < total) sum = v_add(sum, other);
      }
    }
    if (valid && 0u == sub) {
      const V3 unit = v_unit(sum);                                  /* normal.c:159 */
      vn[3ull * slot + 0ull] = unit.x; vn[3ull * slot + 1ull] = unit.y; vn[3ull * slot + 2ull] = unit.z;
    }
  }
}

/* ------------------------------------------------------------------- blend */

__global__ void __launch_bounds__(128) k_shade_blend(const __grid_constant__ FrameParams p, const ShadePix *__restrict__ in,
                                                     const unsigned int *__restrict__ counters, const double *__restrict__ vn,
                                                     unsigned int *__restrict__ rgba) {
  const unsigned int n = counters[0];
  for (unsigned int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
    const ShadePix sp = in[idx];
    const FieldView &f = p.view[sp.obj];
    V3 nv[3];
#pragma unroll
    for (int m = 0; m < 3; m++) {
      const unsigned int slot = sp.vslot[m];
      if (C3D_NO_SLOT != slot) nv[m] = V3{vn[3ull * slot + 0ull], vn[3ull * slot + 1ull], vn[3ull * slot + 2ull]};
      else nv[m] = vertex_normal(f, p.tab, (long long)sp.i, (long long)sp.j, (long long)sp.k, (unsigned int)sp.edges[m]);
    }
    /* w0*n0 + (w1*n1 + w2*n2), render.c:137-146 */
    const V3 nrm = v_unit(v_add(v_mul(sp.w[0], nv[0]), v_add(v_mul(sp.w[1], nv[1]), v_mul(sp.w[2], nv[2]))));
    const double factor = fmax(0.15, fabs(v_dot(nrm, v_of(p.cam.light))));
    const unsigned int r = (unsigned int)__double2int_rz(r_mul(factor, (double)f.color[0])) & 255u;
    const unsigned int g = (unsigned int)__double2int_rz(r_mul(factor, (double)f.color[1])) & 255u;
    const unsigned int b = (unsigned int)__double2int_rz(r_mul(factor, (double)f.color[2])) & 255u;
    rgba[sp.pix] = 0xFF000000u | (b << 16) | (g << 8) | r;
  }
}

extern "C" int c3d_k_shade(const FrameParams *p, const LineDev *lines_dev, const Pixel *canvas, const unsigned long long *depth_global,
                           unsigned int *rgba, size_t npix, const ShadeScratch *sc, unsigned long long *overflowed,
                           int sm_count, cudaStream_t s) {
  cudaError_t err = cudaMemsetAsync(sc->counters, 0, 2 * sizeof(unsigned int), s);
  if (cudaSuccess != err) return (int)err;
  err = cudaMemsetAsync(sc->table, 0, ((size_t)sc->table_mask + 1u) * sizeof(unsigned long long), s);
  if (cudaSuccess != err) return (int)err;
  k_shade_select<<<sm_count * 8, 256, 0, s>>>(*p, lines_dev, canvas, depth_global, rgba, npix, sc->work, sc->counters);
  k_shade_setup<<<sm_count * 12, 128, 0, s>>>(*p, canvas, sc->work, sc->counters, static_cast<ShadePix *>(sc->pix), sc->table, sc->table_mask,
                                              sc->vlist, overflowed);
  k_shade_normals<<<sm_count * 12, 128, 0, s>>>(*p, sc->table, sc->vlist, sc->counters, sc->vn);
  k_shade_blend<<<sm_count * 12, 128, 0, s>>>(*p, static_cast<const ShadePix *>(sc->pix), sc->counters, sc->vn, rgba);
  return (int)cudaGetLastError();
}
