/*
 * c3d_kernels.cu -- the sm_100a kernels of the per-frame pipeline and their
 * launchers.  Compiled with -gencode arch=compute_100a,code=sm_100a.
 *
 *   k_sweep      streaming classification of the scalar field + compaction of
 *                active cubes for arrays the TMA path cannot take (rows not
 *                16-byte aligned); the main sweep is c3d_sweep.cu
 *   k_raster     per-cube occlusion test, per-triangle pre-test, exact fp64
 *                tessellation + projection + coverage of the survivors into a
 *                z-buffer of (depth, winner key) pairs updated by ONE 128-bit
 *                compare-and-swap per fragment (single pass)
 *   k_hiz_build  coarse z-buffer between parts of the raster (optional)
 *   k_lines      poly-line segments into the same z-buffer
 *   (deferred shading lives in c3d_shade.cu, the tiled sweep in c3d_sweep.cu)
 *   k_pack       y-flip + RGB byte packing (PPM payload)
 *   + halo pack/unpack, debug/inspection kernels, synthetic fields
 *
 * File:line references are to NaokiHori/Contour3D @ c66854e.
 */
#include <float.h>
#include <stdlib.h>
#include "c3d_device.cuh"
#include "c3d_synth.h"

/* ------------------------------------------------------------------ clear */

__global__ void k_clear(Pixel *__restrict__ canvas, const size_t npix) {
  /* every pixel starts at depth -DBL_MAX with no winner (src/contour3d/main.c:95-103) */
  const ulonglong2 empty = make_ulonglong2(ord_of(-DBL_MAX), 0ull);
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (size_t)gridDim.x * blockDim.x) {
    reinterpret_cast<ulonglong2 *>(canvas)[p] = empty;
  }
}

extern "C" int c3d_k_clear(Pixel *canvas, size_t npix, cudaStream_t s) {
  const int threads = 256;
  const int blocks = (int)((npix + threads - 1) / threads < 148 * 16 ? (npix + threads - 1) / threads : 148 * 16);
  k_clear<<<blocks > 0 ? blocks : 1, threads, 0, s>>>(canvas, npix);
  return (int)cudaGetLastError();
}

/* screen-partitioned frames: only this rank's tiles are ever read, so only they are cleared */
__global__ void __launch_bounds__(256) k_clear_part(Pixel *__restrict__ canvas, const Partition part, const unsigned int w, const unsigned int h) {
  const ulonglong2 empty = make_ulonglong2(ord_of(-DBL_MAX), 0ull);
  const unsigned int tile = 1u << part.tile_shift, ntiles = (unsigned int)(part.tiles_x * part.tiles_y);
  for (unsigned int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const unsigned int tx = t % (unsigned int)part.tiles_x, ty = t / (unsigned int)part.tiles_x;
    if (tile_owner(part, tx, ty) != (unsigned int)part.me) continue;
    for (unsigned int q = threadIdx.x; q < tile * tile; q += blockDim.x) {
      const unsigned int pi = tx * tile + (q & (tile - 1u)), pj = ty * tile + (q >> part.tile_shift);
      if (pi < w && pj < h) reinterpret_cast<ulonglong2 *>(canvas)[(size_t)pj * w + pi] = empty;
    }
  }
}

extern "C" int c3d_k_clear_part(Pixel *canvas, const Partition *part, long long w, long long h, int sm_count, cudaStream_t s) {
  k_clear_part<<<sm_count * 4, 256, 0, s>>>(canvas, *part, (unsigned int)w, (unsigned int)h);
  return (int)cudaGetLastError();
}

/* pixel-centre coordinates of every column and row (render.c:94-98), once per screen size; the float copies
   feed the conservative fp32 coverage pre-test of the raster */
__global__ void k_pixel_tables(double *__restrict__ px, double *__restrict__ py, float *__restrict__ pxf, float *__restrict__ pyf,
                               const long long w, const long long h) {
  const double W = (double)w, H = (double)h;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < w + h; q += (long long)gridDim.x * blockDim.x) {
    if (q < w) { px[q] = pixel_centre(q, W); pxf[q] = (float)px[q]; }
    else { py[q - w] = pixel_centre(q - w, H); pyf[q - w] = (float)py[q - w]; }
  }
}

extern "C" int c3d_k_pixel_tables(double *px, double *py, float *pxf, float *pyf, long long w, long long h, cudaStream_t s) {
  k_pixel_tables<<<64, 256, 0, s>>>(px, py, pxf, pyf, w, h);
  return (int)cudaGetLastError();
}

/* ------------------------------------------------------------------ sweep */

struct SweepParams {
  const void *data;
  long long mx, my, mz, sy, sz;
  float thr_ge;
  double thr;
  long long ilo, ihi, jlo, jhi, klo, khi;   /* inclusive cube ranges to emit */
  unsigned int obj;                         /* contour object the entries are tagged with */
  int nxseg, nytile, nzchunk, zc;
  unsigned long long ntiles;
  unsigned long long *list, *count, capacity;
};

template <typename T> struct Above;
template <> struct Above<float> {
  float t;
  __device__ __forceinline__ Above(const SweepParams &p) : t(p.thr_ge) {}
  /* thr < (double)v  <=>  v >= (smallest float above thr) */
  __device__ __forceinline__ unsigned int operator()(const float v) const { return v >= t ? 1u : 0u; }
};
template <> struct Above<double> {
  double t;
  __device__ __forceinline__ Above(const SweepParams &p) : t(p.thr) {}
  __device__ __forceinline__ unsigned int operator()(const double v) const { return t < v ? 1u : 0u; }
};

template <typename T> struct Quad { T a, b, c, d; };

template <typename T> __device__ __forceinline__ Quad<T> ld_quad(const T *p);
template <> __device__ __forceinline__ Quad<float> ld_quad<float>(const float *p) {
  const float4 v = __ldcs(reinterpret_cast<const float4 *>(p));     /* 128-bit streaming load */
  return Quad<float>{v.x, v.y, v.z, v.w};
}
template <> __device__ __forceinline__ Quad<double> ld_quad<double>(const double *p) {
  const double2 lo = __ldcs(reinterpret_cast<const double2 *>(p));
  const double2 hi = __ldcs(reinterpret_cast<const double2 *>(p) + 1);
  return Quad<double>{lo.x, lo.y, hi.x, hi.y};
}

/*
 * One warp owns a tile of 128 points in x (lane l: points 4l..4l+3) by TY cube
 * rows in y and marches over zc cube planes in z.  For every (row, plane) the
 * lane keeps 5 bits: "above threshold" of its four points plus the first point
 * of the lane to its right, so the eight corner bits of all its cubes are
 * register-resident; nothing but the field is read and nothing but the compact
 * list of active cubes is written.
 *
 * Each plane step loads TY+1 rows with independent 128-bit loads (the next
 * plane is requested before the current one is classified).
 */
template <typename T, int TY, bool VEC>
__global__ void __launch_bounds__(256) k_sweep(const SweepParams p) {
  const int lane = threadIdx.x & 31;
  const unsigned long long tile = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tile >= p.ntiles) return;
  const int xseg = (int)(tile % (unsigned long long)p.nxseg);
  const int ytile = (int)((tile / (unsigned long long)p.nxseg) % (unsigned long long)p.nytile);
  const int zchunk = (int)(tile / ((unsigned long long)p.nxseg * (unsigned long long)p.nytile));
  const long long x0 = (long long)xseg * 128 + lane * 4;
  const long long j0 = p.jlo + (long long)ytile * TY;
  const long long k0 = p.klo + (long long)zchunk * p.zc;
  const long long k1 = (k0 + p.zc - 1 < p.khi) ? k0 + p.zc - 1 : p.khi;
  const T *const base = static_cast<const T *>(p.data);
  const Above<T> above(p);
  constexpr int ROWS = TY + 1;
  constexpr unsigned long long ALL = (1ull << (5 * ROWS)) - 1ull;

  Quad<T> raw[ROWS];
  T extra[ROWS];

  auto request = [&](const long long k) {
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
      const long long j = j0 + r;
      raw[r] = Quad<T>{T(0), T(0), T(0), T(0)};
      extra[r] = T(0);
      if (j < p.my && k < p.mz) {
        const T *row = base + k * p.sz + j * p.sy;
        if (VEC) {
          if (x0 < p.mx) raw[r] = ld_quad<T>(row + x0);
        } else {
          if (x0 + 0 < p.mx) raw[r].a = __ldcs(row + x0 + 0);
          if (x0 + 1 < p.mx) raw[r].b = __ldcs(row + x0 + 1);
          if (x0 + 2 < p.mx) raw[r].c = __ldcs(row + x0 + 2);
          if (x0 + 3 < p.mx) raw[r].d = __ldcs(row + x0 + 3);
        }
        /* the point right of the warp's segment belongs to another warp: lane 31 fetches it */
        if (31 == lane && x0 + 4 < p.mx) extra[r] = __ldg(row + x0 + 4);
      }
    }
  };
  /* out-of-range points were loaded as 0 and must classify as "not above"
     regardless of the threshold: mask them */
  auto classify = [&](const long long k) -> unsigned long long {
    unsigned long long bits = 0ull;
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
      const long long j = j0 + r;
      unsigned int b = 0u;
      if (j < p.my && k < p.mz) {
        b = above(raw[r].a) | (above(raw[r].b) << 1) | (above(raw[r].c) << 2) | (above(raw[r].d) << 3);
        if (!VEC) {
          const long long left = p.mx - x0;
          b &= left >= 4 ? 15u : (left > 0 ? (1u << left) - 1u : 0u);
        } else if (x0 >= p.mx) {
          b = 0u;
        }
        unsigned int nb = __shfl_down_sync(C3D_FULL_MASK, b & 1u, 1);
        if (31 == lane) nb = (x0 + 4 < p.mx) ? above(extra[r]) : 0u;
        b |= nb << 4;
      }
      bits |= (unsigned long long)b << (5 * r);
    }
    return bits;
  };

  request(k0);
  unsigned long long lower = classify(k0);
  request(k0 + 1);
  for (long long k = k0; k <= k1; k++) {
    const unsigned long long upper = classify(k + 1);
    if (k < k1) request(k + 2);
    /* quick reject: every point of both planes on the same side */
    const bool candidate = ((lower | upper) != 0ull) && ((lower & upper) != ALL);
    if (__any_sync(C3D_FULL_MASK, candidate)) {
      unsigned int active = 0u;
      if (candidate) {
#pragma unroll
        for (int r = 0; r < TY; r++) {
          const long long j = j0 + r;
          const unsigned int l0 = (unsigned int)(lower >> (5 * r)) & 31u, l1 = (unsigned int)(lower >> (5 * r + 5)) & 31u;
          const unsigned int u0 = (unsigned int)(upper >> (5 * r)) & 31u, u1 = (unsigned int)(upper >> (5 * r + 5)) & 31u;
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const unsigned int code = ((l0 >> c) & 3u) | (((l1 >> c) & 3u) << 2) | (((u0 >> c) & 3u) << 4) | (((u1 >> c) & 3u) << 6);
            const long long i = x0 + c;
            const bool ok = code != 0u && code != 255u && i >= p.ilo && i <= p.ihi && j <= p.jhi;
            active |= (ok ? 1u : 0u) << (r * 4 + c);
          }
        }
      }
      /* warp-aggregated append: one atomic per warp and plane */
      const unsigned int n = __popc(active);
      unsigned int incl = n;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned int up = __shfl_up_sync(C3D_FULL_MASK, incl, d);
        if (lane >= d) incl += up;
      }
      const unsigned int total = __shfl_sync(C3D_FULL_MASK, incl, 31);
      if (total > 0u) {
        unsigned long long start = 0ull;
        if (0 == lane) start = atomicAdd(p.count, (unsigned long long)total);
        start = __shfl_sync(C3D_FULL_MASK, start, 0);
        unsigned long long at = start + incl - n;
        while (active) {
          const int bit = __ffs((int)active) - 1;
          active &= active - 1u;
          const int r = bit >> 2, c = bit & 3;
          const unsigned int l0 = (unsigned int)(lower >> (5 * r)) & 31u, l1 = (unsigned int)(lower >> (5 * r + 5)) & 31u;
          const unsigned int u0 = (unsigned int)(upper >> (5 * r)) & 31u, u1 = (unsigned int)(upper >> (5 * r + 5)) & 31u;
          const unsigned int code = ((l0 >> c) & 3u) | (((l1 >> c) & 3u) << 2) | (((u0 >> c) & 3u) << 4) | (((u1 >> c) & 3u) << 6);
          /* no patch analysis on this path: marked irregular, i.e. never skipped by the raster's visibility tests */
          if (at < p.capacity) p.list[at] = C3D_ENTRY(p.obj, x0 + c, j0 + r, k, code) | C3D_ENTRY_IRREGULAR_BIT;
          at++;
        }
      }
    }
    lower = upper;
  }
}

extern "C" int c3d_k_sweep(const FieldView *f, const SweepRange *r, unsigned int obj, unsigned long long *list, unsigned long long *count,
                           unsigned long long capacity, int sm_count, cudaStream_t s) {
  SweepParams p;
  p.data = f->data;
  p.mx = f->n[0]; p.my = f->n[1]; p.mz = f->n[2];
  p.sy = f->sy; p.sz = f->sz;
  p.thr_ge = f->thr_ge; p.thr = f->thr;
  p.ilo = r->lo[0]; p.ihi = r->hi[0];
  p.jlo = r->lo[1]; p.jhi = r->hi[1];
  p.klo = r->lo[2]; p.khi = r->hi[2];
  p.obj = obj;
  p.list = list; p.count = count; p.capacity = capacity;
  if (p.ihi < p.ilo || p.jhi < p.jlo || p.khi < p.klo) return 0;
  const int ty = 4 == f->elem_bytes ? 8 : 4;
  p.nxseg = (int)((p.ihi + 1 + 127) / 128);              /* segments start at x = 0 to stay 512-byte aligned */
  p.nytile = (int)((p.jhi - p.jlo + 1 + ty - 1) / ty);
  const long long nk = p.khi - p.klo + 1;
  const long long per_chunk = (long long)p.nxseg * p.nytile;
  const long long target = 8ll * sm_count * 4;           /* a few waves of 8-warp CTAs */
  long long want = (target + per_chunk - 1) / per_chunk;
  if (want < 1) want = 1;
  long long zc = nk / want;
  if (zc > 64) zc = 64;
  if (zc < 4) zc = 4;
  if (zc > nk) zc = nk;
  p.zc = (int)zc;
  p.nzchunk = (int)((nk + zc - 1) / zc);
  p.ntiles = (unsigned long long)per_chunk * (unsigned long long)p.nzchunk;
  const int warps = 8;
  const unsigned int blocks = (unsigned int)((p.ntiles + warps - 1) / warps);
  const bool vec = (0 == p.mx % 4) && (0 == p.sy % 4) && (0 == p.sz % 4) && (0 == ((uintptr_t)p.data % 16));
  if (4 == f->elem_bytes) {
    if (vec) k_sweep<float, 8, true><<<blocks, warps * 32, 0, s>>>(p);
    else k_sweep<float, 8, false><<<blocks, warps * 32, 0, s>>>(p);
  } else {
    if (vec) k_sweep<double, 4, true><<<blocks, warps * 32, 0, s>>>(p);
    else k_sweep<double, 4, false><<<blocks, warps * 32, 0, s>>>(p);
  }
  return (int)cudaGetLastError();
}

/* ----------------------------------------------------------------- raster */

/*
 * One pass over the frame's active-cube list (all contour objects mixed).
 *
 * Load balance: a warp takes 32 cubes, and spreads their 1..12 triangles over
 * its lanes (prefix sum + one REDUX per round), so every lane rasterises ONE
 * triangle at a time whatever the case mix.  A triangle is rebuilt from the
 * field (positions are never stored), projected, and its +-1 pixel bounding
 * box (render.c:81-89) is walked in 8x8 pixel tiles in two phases: phase A
 * evaluates only the three edge functions and collects the covered pixels in
 * a 64-bit mask, phase B computes depth and submits the few that are covered.
 * That keeps the divergent depth/atomic path off the common "outside" case.
 *
 * Z-buffer: one 128-bit CAS per surviving fragment on (ord(depth), key); see
 * submit_fragment().
 */
/* pixel box imin | imax << 16 | jmin << 32 | jmax << 48 that certainly holds every pixel centre a triangle can cover
   (triangle_hidden), or C3D_NO_BOX when the pre-test did not get that far */
#define C3D_NO_BOX 0xffffffffffffffffull
#define C3D_QUEUE 64              /* triangles a warp can hold back until 32 of them are ready for the exact path */

template <bool MAPPED, bool PEER>
__device__ __forceinline__ void raster_triangle(const FrameParams &p, const FieldView *staged, const unsigned long long entry, const unsigned int s,
                                                const unsigned long long box, const double *qval, Pixel *__restrict__ canvas) {
  const unsigned int obj = C3D_ENTRY_OBJ(entry), code = C3D_ENTRY_CODE(entry);
  const FieldView &f = entry_view<PEER>(p, staged, entry);
  const long long i = C3D_ENTRY_I(entry), j = C3D_ENTRY_J(entry), k = C3D_ENTRY_K(entry);
  const unsigned int pairs = __ldg(p.tab.pairs + code * C3D_MAX_TRIS + s);
  Tri2D t;
  /* one out-of-line copy of "cut the edge and project" serves the three vertices: the kernel's code must stay
     within reach of the instruction cache while its warps sit in different phases */
  if (MAPPED && nullptr != f.map) {
#pragma unroll
    for (int m = 0; m < 3; m++) {
      const ProjectedVertex v = exact_vertex_t<MAPPED>(p.cam, f, i, j, k, (pairs >> (6 * m)) & 7u, (pairs >> (6 * m + 3)) & 7u);
      if (0. == v.w) return;                          /* render.c:62-72 */
      t.q[m] = V3{v.x, v.y, v.z};
    }
  } else {
    /* the six corner values came with the triangle (read once per cube, k_raster); the six coordinates are loaded together */
    double val[6];
#pragma unroll
    for (int v = 0; v < 6; v++) val[v] = qval[v * C3D_QUEUE];
    const double gx0 = __ldg(f.g[0] + i), gx1 = __ldg(f.g[0] + i + 1), gy0 = __ldg(f.g[1] + j), gy1 = __ldg(f.g[1] + j + 1);
    const double gz0 = __ldg(f.g[2] + k), gz1 = __ldg(f.g[2] + k + 1);
#pragma unroll
    for (int m = 0; m < 3; m++) {
      const unsigned int tail = (pairs >> (6 * m)) & 7u, head = (pairs >> (6 * m + 3)) & 7u;
      const ProjectedVertex v = exact_vertex_from(p.cam, f.thr, val[2 * m], val[2 * m + 1],
                                                  (tail & 1u) ? gx1 : gx0, (tail & 2u) ? gy1 : gy0, (tail & 4u) ? gz1 : gz0,
                                                  (head & 1u) ? gx1 : gx0, (head & 2u) ? gy1 : gy0, (head & 4u) ? gz1 : gz0);
      if (0. == v.w) return;                          /* render.c:62-72 */
      t.q[m] = V3{v.x, v.y, v.z};
    }
  }
  setup_projected(p.cam, t);
  const unsigned long long mykey = winner_key(f, obj, i, j, k, s);
  /*
   * Coverage in two steps.  Most pixels of the +-1 pixel box are far outside the triangle, so each is first
   * tested with the three edge functions in fp32 and dropped when one of them is below -tol, tol being a
   * rigorous bound of the fp32 evaluation error (inputs rounded to 2^-24, four rounded operations:
   * <= 5.01 * 2^-24 * [(|px| + |ax|) |ey| + (|py| + |ay|) |ex|], taken as 4e-7 * [...] with |px|, |py| <= 0.5)
   * plus the reference's own slack: w < -1e-8 <=> sign(area) E < -1e-8 |area| (render.c:102-115).  A dropped
   * pixel is therefore outside for the exact test too; everything else is decided by the exact fp64 test below.
   */
  float fax[3], fay[3], fex[3], fey[3], ntol[3];
  {
    const double sgn = t.area_inv < 0. ? -1. : 1.;
    const double slack = 1.0001e-8 / fabs(t.area_inv);
#pragma unroll
    for (int m = 0; m < 3; m++) {
      const V3 &a = t.q[(m + 1) % 3];
      fax[m] = (float)a.x; fay[m] = (float)a.y;
      fex[m] = (float)(sgn * t.ex[m]); fey[m] = (float)(sgn * t.ey[m]);
      const double tol = 4.e-7 * ((0.5 + fabs(a.x)) * fabs(t.ey[m]) + (0.5 + fabs(a.y)) * fabs(t.ex[m])) + slack;
      ntol[m] = -((float)tol * 1.000001f + 1.e-37f);       /* NaN / inf propagate: then nothing is dropped */
    }
  }
  int jmin = (int)t.jmin, jmax = (int)t.jmax, imin = (int)t.imin, imax = (int)t.imax;
  if (C3D_NO_BOX != box) {
    /* the reference searches its +-1 pixel box (render.c:81-89), 9-16 pixels for a sub-pixel triangle; the pre-test has
       already bounded where a pixel centre can lie inside the triangle, usually one or two pixels */
    imin = max(imin, (int)(box & 65535ull)); imax = min(imax, (int)((box >> 16) & 65535ull));
    jmin = max(jmin, (int)((box >> 32) & 65535ull)); jmax = min(jmax, (int)(box >> 48));
  }
  for (int j0 = jmin; j0 <= jmax; j0 += 4) {
    const int nj = min(jmax - j0, 3);
    for (int i0 = imin; i0 <= imax; i0 += 8) {
      const int ni = min(imax - i0, 7);
      unsigned int maybe = 0u;
      for (int dj = 0; dj <= nj; dj++) {
        const float py = __ldg(p.pyf + j0 + dj);
        /* row terms of the three edge functions; the fused form below has three roundings instead of four, within the same bound */
        const float r0 = (py - fay[0]) * fex[0], r1 = (py - fay[1]) * fex[1], r2 = (py - fay[2]) * fex[2];
        unsigned int bit = 1u << (8 * dj);
        for (int di = 0; di <= ni; di++, bit <<= 1) {
          const float px = __ldg(p.pxf + i0 + di);
          const float e0 = __fmaf_rn(px - fax[0], fey[0], -r0);
          const float e1 = __fmaf_rn(px - fax[1], fey[1], -r1);
          const float e2 = __fmaf_rn(px - fax[2], fey[2], -r2);
          if (!(e0 < ntol[0] || e1 < ntol[1] || e2 < ntol[2])) maybe |= bit;
        }
      }
      while (maybe) {
        const int b = __ffs((int)maybe) - 1;
        maybe &= maybe - 1u;
        const long long pi = i0 + (b & 7), pj = j0 + (b >> 3);
        if (PEER && !pixel_owned(p.part, (int)pi, (int)pj)) continue;     /* another rank draws that tile */
        double w[3];
        if (pixel_weights(t, __ldg(p.px + pi), __ldg(p.py + pj), w)) {
          submit_fragment(canvas + pj * p.cam.w + pi, ord_of(pixel_depth(t, w)), mykey);
        }
      }
    }
  }
}

/* every pixel of [ceil(x0), floor(x1)] x [ceil(y0), floor(y1)] (clipped to the screen) holds a depth above `limit`;
   an empty range is hidden only when `empty_is_hidden` (no pixel centre can be covered); more than 64 pixels: not decided.
   PEER (screen-partitioned frame): only this rank's pixels count, the others are not drawn here at all. */
template <bool PEER>
__device__ __forceinline__ bool region_hidden(const Pixel *__restrict__ canvas, const CullCam &q, const Partition &part, const float x0, const float x1,
                                              const float y0, const float y1, const bool empty_is_hidden, const unsigned long long limit) {
  const int imin = max(0, (int)ceilf(x0)), imax = min(q.w - 1, (int)floorf(x1));
  const int jmin = max(0, (int)ceilf(y0)), jmax = min(q.h - 1, (int)floorf(y1));
  if (imax < imin || jmax < jmin) return empty_is_hidden;
  if ((imax - imin + 1) * (jmax - jmin + 1) > 64) return false;
  for (int pj = jmin; pj <= jmax; pj++) {
    const unsigned long long *row = reinterpret_cast<const unsigned long long *>(canvas + (long long)pj * q.w);
    for (int pi = imin; pi <= imax; pi++) {
      if (PEER && !pixel_owned(part, pi, pj)) continue;
      if (__ldcg(row + 2 * pi) <= limit) return false;
    }
  }
  return true;
}

/* Coarse z-buffer ("Hi-Z"): one word per C3D_HIZ_TILE x C3D_HIZ_TILE pixel tile = the FARTHEST depth stored in the tile when
   the table was built (k_hiz_build), across ranks the maximum of that (ncclMax).  Stored depths only grow
   (render.c:125-132), so every pixel of the tile holds -- on the rank that will win it -- at least this depth for the
   rest of the frame: a region whose tiles all exceed `limit` is hidden, by the same argument as region_hidden(),
   without touching the pixels and including what OTHER ranks have drawn in front.  Undecided (false) otherwise. */
struct HiZ {
  const unsigned long long *tile;    /* nullptr: no table */
  int tw;                            /* tiles per row */
};

__device__ __forceinline__ bool hiz_hidden(const HiZ &z, const CullCam &q, const float x0, const float x1, const float y0, const float y1,
                                           const unsigned long long limit) {
  const int imin = max(0, (int)ceilf(x0)), imax = min(q.w - 1, (int)floorf(x1));
  const int jmin = max(0, (int)ceilf(y0)), jmax = min(q.h - 1, (int)floorf(y1));
  if (imax < imin || jmax < jmin) return true;              /* no pixel centre inside: nothing can be drawn */
  const int ti0 = imin / C3D_HIZ_TILE, ti1 = imax / C3D_HIZ_TILE, tj0 = jmin / C3D_HIZ_TILE, tj1 = jmax / C3D_HIZ_TILE;
  if ((ti1 - ti0 + 1) * (tj1 - tj0 + 1) > 16) return false;
  for (int tj = tj0; tj <= tj1; tj++) {
    for (int ti = ti0; ti <= ti1; ti++) {
      if (__ldg(z.tile + (long long)tj * z.tw + ti) <= limit) return false;
    }
  }
  return true;
}

/*
 * Conservative occlusion test of a whole cube against the z-buffer as it
 * stands: true only when every pixel any of the cube's triangles could touch
 * already holds a depth strictly nearer than anything inside the cube.  Such
 * a cube cannot change the image (depths only grow, render.c:125-132), so its
 * triangles are never built.  The test runs in fp32 with margins far above the
 * fp32 error (0.25 pixel, 1e-4 relative depth); anything doubtful is kept.
 * With a depth complexity of ~20 most cubes of a turbulent field are hidden.
 */
template <bool MAPPED, bool PEER>
__device__ __forceinline__ bool cube_hidden(const FrameParams &p, const FieldView *staged, const unsigned long long entry, const Pixel *__restrict__ canvas, const HiZ &hiz) {
  /* A triangle whose vertices have collapsed onto one cube corner is drawn by the reference as a blob: the
     clamped area (render.c:103-106) makes all three weights +-0, every pixel of the +-1 pixel box passes the
     inside test and the depth is 1 / +-0 = +-inf -- it beats any finite depth.  Collapse needs an interpolation
     factor within rounding of 0 or 1, i.e. a corner value (almost) equal to the iso-value; k_expand marks the
     cubes of every 4 x 8 patch where that cannot be excluded (C3D_ENTRY_IRREGULAR) and those are never culled. */
  if (C3D_ENTRY_IRREGULAR(entry)) return false;
  const CullCam &q = p.cull;
  const FieldView &f = entry_view<PEER>(p, staged, entry);
  const long long i = C3D_ENTRY_I(entry), j = C3D_ENTRY_J(entry), k = C3D_ENTRY_K(entry);
  float xlo = 3.e38f, xhi = -3.e38f, ylo = 3.e38f, yhi = -3.e38f, znear = -3.e38f, tlo = 3.e38f, thi = -3.e38f;
  if (MAPPED && nullptr != f.map) {
    /* tabulated converter: the eight corners are arbitrary points; every triangle vertex still lies on a segment between
       two of them (triangulate.c:40-73), i.e. inside their hull */
#pragma unroll 1
    for (int c = 0; c < 8; c++) {
      float rx, ry, rz;
      mapped_posf(f.map, i + (c & 1), j + ((c >> 1) & 1), k + (c >> 2), rx, ry, rz);
      rx -= q.cam[0]; ry -= q.cam[1]; rz -= q.cam[2];
      const float nr = q.n[0] * rx + q.n[1] * ry + q.n[2] * rz;
      const float t = __fdividef(q.c, nr);
      const float sx = q.ax + t * (q.lxs[0] * rx + q.lxs[1] * ry + q.lxs[2] * rz);
      const float sy = q.ay + t * (q.lys[0] * rx + q.lys[1] * ry + q.lys[2] * rz);
      xlo = fminf(xlo, sx); xhi = fmaxf(xhi, sx);
      ylo = fminf(ylo, sy); yhi = fmaxf(yhi, sy);
      znear = fmaxf(znear, nr);
      tlo = fminf(tlo, t); thi = fmaxf(thi, t);
    }
  } else {
    float an[2], al[2], am[2], bn[2], bl[2], bm[2], cn[2], cl[2], cm[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const float rx = __ldg(f.gf[0] + i + e) - q.cam[0];
      const float ry = __ldg(f.gf[1] + j + e) - q.cam[1];
      const float rz = __ldg(f.gf[2] + k + e) - q.cam[2];
      an[e] = q.n[0] * rx; al[e] = q.lxs[0] * rx; am[e] = q.lys[0] * rx;
      bn[e] = q.n[1] * ry; bl[e] = q.lxs[1] * ry; bm[e] = q.lys[1] * ry;
      cn[e] = q.n[2] * rz; cl[e] = q.lxs[2] * rz; cm[e] = q.lys[2] * rz;
    }
#pragma unroll
    for (int c = 0; c < 8; c++) {
      const int a = c & 1, b = (c >> 1) & 1, d = c >> 2;
      const float nr = an[a] + bn[b] + cn[d];
      const float t = __fdividef(q.c, nr);
      const float sx = q.ax + t * (al[a] + bl[b] + cl[d]);
      const float sy = q.ay + t * (am[a] + bm[b] + cm[d]);
      xlo = fminf(xlo, sx); xhi = fmaxf(xhi, sx);
      ylo = fminf(ylo, sy); yhi = fmaxf(yhi, sy);
      znear = fmaxf(znear, nr);
      tlo = fminf(tlo, t); thi = fmaxf(thi, t);
    }
  }
  /* every corner strictly beyond the screen plane as seen from the camera (project.c:44-46 accepts 0 < t < 1;
     then all n.ray share the sign of c and the cube projects inside the hull of its corners), else the
     bounds mean nothing and the cube is kept */
  if (!(tlo > 1.e-4f && thi < 0.9999f)) return false;
  if (!(xlo > -1.e6f && xhi < 1.e6f && ylo > -1.e6f && yhi < 1.e6f)) return false;
  const double zcap = (double)znear + (1.e-4 * fabs((double)znear) + 1.e-6);
  const unsigned long long limit = ord_of(zcap);
  /* a fragment exists only where a PIXEL CENTRE lies inside a triangle (render.c:94-115; the +-1 pixel box of
     render.c:81-89 is just the search range), and every triangle lies inside the hull of the corners:
     the pixels to look at are those whose centre (i + 0.5) / W - 0.5 falls within the bounds, +- 0.05 pixel */
  const float X0 = (0.5f + xlo) * q.W - 0.55f, X1 = (0.5f + xhi) * q.W - 0.45f, Y0 = (0.5f + ylo) * q.H - 0.55f, Y1 = (0.5f + yhi) * q.H - 0.45f;
  if (!PEER && nullptr != hiz.tile && hiz_hidden(hiz, q, X0, X1, Y0, Y1, limit)) return true;
  return region_hidden<PEER>(canvas, q, p.part, X0, X1, Y0, Y1, true, limit);
}

/*
 * The same question for ONE triangle, asked before its exact (and expensive) fp64 set-up: an fp32
 * estimate of the three vertices with explicit error bounds gives a box that certainly contains the
 * exact projected triangle.  No pixel centre inside the box (true for about half of the sub-pixel
 * triangles of a 1024^3 field on a 4K screen), or every such pixel already nearer in the z-buffer
 * than the triangle can be -> it cannot produce a surviving fragment and is dropped.
 *
 * Error budget per vertex: the cut position p0 + f (p1 - p0) has f = (thr - v0) / (v1 - v0) with an
 * absolute error below ferr = 2e-7 (|thr| + |v0| + |v1|) / |v1 - v0| (inputs rounded to fp32, three
 * rounded operations); the vertex moves by at most ferr |p1 - p0| in space, i.e. by at most
 * G ferr |p1 - p0|_1 on the screen with G = |t| N (|l| + |ray.l| |t| / |c|) the local magnification of
 * the projection (x1.5 for its curvature over that distance); fp32 evaluation of the projection itself
 * stays below 0.01 pixel on screens up to 2^20 pixels wide; 0.05 pixel is added on top.  Anything that
 * does not fit these assumptions (ferr > 0.02, vertices near the screen plane, NaN) keeps the triangle.
 */
template <bool MAPPED, bool PEER>
__device__ __forceinline__ bool triangle_hidden(const FrameParams &p, const FieldView *staged, const unsigned long long entry, const unsigned int s,
                                                const double *corner, const Pixel *__restrict__ canvas, unsigned long long &box) {
  box = C3D_NO_BOX;
  const CullCam &q = p.cull;
  const unsigned int code = C3D_ENTRY_CODE(entry);
  const FieldView &f = entry_view<PEER>(p, staged, entry);
  const long long i = C3D_ENTRY_I(entry), j = C3D_ENTRY_J(entry), k = C3D_ENTRY_K(entry);
  const unsigned int pairs = __ldg(p.tab.pairs + code * C3D_MAX_TRIS + s);
  const float thr = (float)f.thr;
  /* the cube's six grid coordinates once, not per vertex (identity converter) */
  float gx0 = 0.f, gx1 = 0.f, gy0 = 0.f, gy1 = 0.f, gz0 = 0.f, gz1 = 0.f;
  if (!(MAPPED && nullptr != f.map)) {
    gx0 = __ldg(f.gf[0] + i); gx1 = __ldg(f.gf[0] + i + 1);
    gy0 = __ldg(f.gf[1] + j); gy1 = __ldg(f.gf[1] + j + 1);
    gz0 = __ldg(f.gf[2] + k); gz1 = __ldg(f.gf[2] + k + 1);
  }
  float xlo = 3.e38f, xhi = -3.e38f, ylo = 3.e38f, yhi = -3.e38f, znear = -3.e38f, zslack = 0.f;
  /* fp32 with explicit fused multiply-adds (the translation unit is compiled with -fmad=false for the exact fp64 path;
     here fewer roundings only help) and ONE reciprocal per vertex: rcp is within 1 ulp and the product adds half an ulp,
     so the factor is off by < 2e-7 on top of the bound below -- covered by the extra 3e-7 in ferr */
  /* the cube's eight corner values sit in shared memory (corner[n * 32]: read from the array once per cube by k_raster) */
#pragma unroll 1
  for (int m = 0; m < 3; m++) {
    const unsigned int tail = (pairs >> (6 * m)) & 7u, head = (pairs >> (6 * m + 3)) & 7u;
    const float v0 = (float)corner[tail * 32u];
    const float v1 = (float)corner[head * 32u];
    const float dv = v1 - v0;
    const float inv = __frcp_rn(dv);
    const float ferr = 2.00002e-7f * (fabsf(thr) + fabsf(v0) + fabsf(v1)) * fabsf(inv) + 3.e-7f;
    if (!(ferr < 0.02f) || !(fabsf(dv) > 1.e-7f)) return false;    /* incl. the |dv| < 1e-8 branch of triangulate.c:62-71 */
    const float fr = fminf(fmaxf((thr - v0) * inv, 0.f), 1.f);
    /* a vertex that may sit within 1e-5 of a cube corner: the triangle may have collapsed to a point, which the
       reference draws as its whole +-1 pixel box (see cube_hidden); leave those to the exact path */
    if (!(fr - ferr > 1.e-5f && fr + ferr < 0.99999f)) return false;
    float x0, y0, z0, ex, ey, ez;
    if (MAPPED && nullptr != f.map) {
      mapped_posf(f.map, i + (tail & 1u), j + ((tail >> 1) & 1u), k + (tail >> 2), x0, y0, z0);
      mapped_posf(f.map, i + (head & 1u), j + ((head >> 1) & 1u), k + (head >> 2), ex, ey, ez);
      ex -= x0; ey -= y0; ez -= z0;
    } else {
      x0 = (tail & 1u) ? gx1 : gx0; y0 = (tail & 2u) ? gy1 : gy0; z0 = (tail & 4u) ? gz1 : gz0;
      ex = ((head & 1u) ? gx1 : gx0) - x0; ey = ((head & 2u) ? gy1 : gy0) - y0; ez = ((head & 4u) ? gz1 : gz0) - z0;
    }
    const float rx = __fmaf_rn(fr, ex, x0) - q.cam[0], ry = __fmaf_rn(fr, ey, y0) - q.cam[1], rz = __fmaf_rn(fr, ez, z0) - q.cam[2];
    const float nr = __fmaf_rn(q.n[2], rz, __fmaf_rn(q.n[1], ry, q.n[0] * rx));
    const float t = __fdividef(q.c, nr);
    if (!(t > 1.e-4f && t < 0.9999f)) return false;
    const float rl = __fmaf_rn(rz, q.lxs[2], __fmaf_rn(ry, q.lxs[1], rx * q.lxs[0]));
    const float rm = __fmaf_rn(rz, q.lys[2], __fmaf_rn(ry, q.lys[1], rx * q.lys[0]));
    /* pixel coordinates: pixel i has its centre at X = i */
    const float X = __fmaf_rn(__fmaf_rn(t, rl, 0.5f + q.ax), q.W, -0.5f), Y = __fmaf_rn(__fmaf_rn(t, rm, 0.5f + q.ay), q.H, -0.5f);
    const float span = ferr * (fabsf(ex) + fabsf(ey) + fabsf(ez));
    const float tc = t * q.inv_abs_c;
    const float st = 1.5f * span * t;
    const float mx = __fmaf_rn(st * q.W, __fmaf_rn(fabsf(rl), tc, q.lx_norm), 0.05f);
    const float my = __fmaf_rn(st * q.H, __fmaf_rn(fabsf(rm), tc, q.ly_norm), 0.05f);
    xlo = fminf(xlo, X - mx); xhi = fmaxf(xhi, X + mx);
    ylo = fminf(ylo, Y - my); yhi = fmaxf(yhi, Y + my);
    znear = fmaxf(znear, nr);
    zslack = fmaxf(zslack, span);
  }
  if (!(xlo > -1.e6f && xhi < 1.e6f && ylo > -1.e6f && yhi < 1.e6f)) return false;
  /* the pixels whose centre lies inside the box (screens are below 2^16 pixels per side here: make_cull) */
  {
    const int imin = max(0, (int)ceilf(xlo)), imax = min(q.w - 1, (int)floorf(xhi));
    const int jmin = max(0, (int)ceilf(ylo)), jmax = min(q.h - 1, (int)floorf(yhi));
    if (imax < imin || jmax < jmin) return true;              /* no pixel centre inside: nothing can be drawn */
    box = (unsigned long long)imin | ((unsigned long long)imax << 16) | ((unsigned long long)jmin << 32) | ((unsigned long long)jmax << 48);
  }
  const double zcap = (double)znear + (1.e-4 * fabs((double)znear) + 1.e-6 + 2. * (double)zslack);
  return region_hidden<PEER>(canvas, q, p.part, xlo, xhi, ylo, yhi, true, ord_of(zcap));
}

/* list entry + triangle index in its top four bits (the irregular flag has been used by then; the source rank of a
   screen-partitioned frame, which lives in three of those bits, travels in a side array of the queue) */
#define C3D_TRI_PACK(entry, s) (((entry) & 0x0fffffffffffffffull) | ((unsigned long long)(s) << 60))
#define C3D_TRI_ENTRY(q) ((q) & 0x0fffffffffffffffull)
#define C3D_TRI_INDEX(q) ((unsigned int)((q) >> 60))

template <int MINB, bool MAPPED, bool PEER>
__global__ void __launch_bounds__(128, MINB) k_raster(const __grid_constant__ FrameParams p, const unsigned long long *__restrict__ list,
                                                     const unsigned long long *__restrict__ count, const unsigned long long capacity,
                                                     Pixel *__restrict__ canvas, unsigned long long *__restrict__ culled,
                                                     const HiZ hiz, const unsigned int part_begin, const unsigned int part_end) {
  __shared__ unsigned long long s_entry[4][32];
  __shared__ unsigned int s_first[4][32];
  __shared__ unsigned long long s_queue[4][C3D_QUEUE];       /* triangles that passed the pre-test, waiting for a full warp */
  __shared__ unsigned long long s_qbox[4][C3D_QUEUE];        /* ... and the pixel boxes the pre-test left them */
  __shared__ unsigned char s_qsrc[4][PEER ? C3D_QUEUE : 1];  /* ... and the rank their cube indices refer to */
  /* The corner values of a surviving cube are read from the array ONCE (8 loads in flight per lane, one exposure to the
     memory latency -- in a partitioned frame most are peer reads over NVLink) and kept here for the cube's 1..12 triangles;
     a queued triangle takes the values of its three cut edges along (tail and head of each). */
  __shared__ double s_corner[4][8][32];
  __shared__ double s_qval[4][6][C3D_QUEUE];
  /* the views of a partitioned frame live in device memory (ranks x objects of them); the few in use are staged here, so that
     the per-cube tests read them like the single-rank kernel reads its constant bank */
  constexpr unsigned int STAGED = PEER ? 32u : 1u;
  __shared__ FieldView s_views[STAGED];
  const FieldView *staged = nullptr;
  if (PEER && (unsigned int)p.part.nranks * p.ncontours <= STAGED) {
    const unsigned int words = (unsigned int)(sizeof(FieldView) / 4u);
    for (unsigned int q = threadIdx.x; q < (unsigned int)p.part.nranks * p.ncontours * words; q += blockDim.x) {
      const unsigned int v = q / words, w = q % words;
      reinterpret_cast<unsigned int *>(s_views)[q] = reinterpret_cast<const unsigned int *>(p.pviews + C3D_PEER_VIEW(v / p.ncontours, v % p.ncontours))[w];
    }
    __syncthreads();
    staged = s_views;
  }
  unsigned long long n = *count;
  if (n > capacity) n = capacity;
  const unsigned int lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
  const unsigned long long warp = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + wid;
  const unsigned long long nwarps = (unsigned long long)gridDim.x * (blockDim.x >> 5);
  unsigned int hidden_total = 0u, hidden_tris = 0u, queued = 0u;
  /* this launch draws the entries [n * part_begin / 64, n * part_end / 64) of the nearest-first list (whole warps of 32) */
  const unsigned long long first = (n * part_begin / C3D_RASTER_PARTS) & ~31ull;
  if (part_end < C3D_RASTER_PARTS) n = (n * part_end / C3D_RASTER_PARTS) & ~31ull;
  for (unsigned long long base = first + warp * 32ull; base < n; base += nwarps * 32ull) {
    unsigned long long entry = 0ull;
    unsigned int ntri = 0u;
    if (base + lane < n) {
      entry = __ldg(list + base + lane);
      ntri = __ldg(p.tab.count + C3D_ENTRY_CODE(entry));
      if (p.cull.enabled && cube_hidden<MAPPED, PEER>(p, staged, entry, canvas, hiz)) ntri = 0u;
    }
    const unsigned int alive = __ballot_sync(C3D_FULL_MASK, ntri > 0u);
    hidden_total += __popc(__ballot_sync(C3D_FULL_MASK, base + lane < n)) - __popc(alive);
    unsigned int incl = ntri;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned int up = __shfl_up_sync(C3D_FULL_MASK, incl, d);
      if (lane >= (unsigned int)d) incl += up;
    }
    const unsigned int total = __shfl_sync(C3D_FULL_MASK, incl, 31);
    const unsigned int excl = incl - ntri;
    /* surviving cubes, compacted: slot r = r-th surviving lane */
    __syncwarp();
    if (ntri > 0u) {
      const unsigned int slot = __popc(alive & ((1u << lane) - 1u));
      s_entry[wid][slot] = entry;
      s_first[wid][slot] = excl;
      const FieldView &f = entry_view<PEER>(p, staged, entry);
      const long long at = C3D_ENTRY_K(entry) * f.sz + C3D_ENTRY_J(entry) * f.sy + C3D_ENTRY_I(entry);
      double c[8];
#pragma unroll
      for (int nn = 0; nn < 8; nn++) c[nn] = ld_value(f, at + (nn & 1) + ((nn >> 1) & 1) * f.sy + (nn >> 2) * f.sz);
#pragma unroll
      for (int nn = 0; nn < 8; nn++) s_corner[wid][nn][slot] = c[nn];
    }
    __syncwarp();
    for (unsigned int t0 = 0; t0 < total; t0 += 32u) {
      /* source cube of triangle t0 + lane: the last surviving cube whose first triangle is <= it */
      const bool starts_here = ntri > 0u && excl >= t0 && excl < t0 + 32u;
      const unsigned int starts = __reduce_or_sync(C3D_FULL_MASK, starts_here ? 1u << (excl - t0) : 0u);
      const unsigned int before = __popc(__ballot_sync(C3D_FULL_MASK, ntri > 0u && excl < t0));
      const unsigned int src = (before + __popc(starts & (0xffffffffu >> (31u - lane))) - 1u) & 31u;
      bool keep = false;
      unsigned long long mine = 0ull, mybox = C3D_NO_BOX;
      unsigned int mysrc = 0u;
      if (t0 + lane < total) {
        const unsigned long long e = s_entry[wid][src];
        const unsigned int s = t0 + lane - s_first[wid][src];
        mine = C3D_TRI_PACK(e, s);
        mysrc = C3D_ENTRY_SRC(e);
        keep = !(p.cull.enabled && !C3D_ENTRY_IRREGULAR(e) && triangle_hidden<MAPPED, PEER>(p, staged, e, s, &s_corner[wid][0][src], canvas, mybox));
      }
      const unsigned int kept = __ballot_sync(C3D_FULL_MASK, keep);
      hidden_tris += __popc(__ballot_sync(C3D_FULL_MASK, t0 + lane < total)) - __popc(kept);
      if (keep) {
        const unsigned int at = queued + __popc(kept & ((1u << lane) - 1u));
        s_queue[wid][at] = mine;
        s_qbox[wid][at] = mybox;
        if (PEER) s_qsrc[wid][at] = (unsigned char)mysrc;
        const unsigned int pairs = __ldg(p.tab.pairs + C3D_ENTRY_CODE(mine) * C3D_MAX_TRIS + C3D_TRI_INDEX(mine));
#pragma unroll
        for (int m = 0; m < 3; m++) {
          s_qval[wid][2 * m][at] = s_corner[wid][(pairs >> (6 * m)) & 7u][src];
          s_qval[wid][2 * m + 1][at] = s_corner[wid][(pairs >> (6 * m + 3)) & 7u][src];
        }
      }
      queued += __popc(kept);
      __syncwarp();
      if (queued >= 32u) {
        /* a full warp of triangles that need the exact treatment */
        const unsigned long long q = s_queue[wid][lane], qb = s_qbox[wid][lane];
        const unsigned long long qs = PEER ? (unsigned long long)s_qsrc[wid][lane] << C3D_ENTRY_SRC_SHIFT : 0ull;
        /* the exact path reads its values from the first 32 queue slots while the rest moves down: keep them in registers */
        double qv[6];
#pragma unroll
        for (int v = 0; v < 6; v++) qv[v] = s_qval[wid][v][lane];
        const unsigned long long rest = (lane + 32u < queued) ? s_queue[wid][lane + 32u] : 0ull;
        const unsigned long long restb = (lane + 32u < queued) ? s_qbox[wid][lane + 32u] : 0ull;
        const unsigned char rests = (PEER && lane + 32u < queued) ? s_qsrc[wid][lane + 32u] : (unsigned char)0;
        double restv[6];
#pragma unroll
        for (int v = 0; v < 6; v++) restv[v] = (lane + 32u < queued) ? s_qval[wid][v][lane + 32u] : 0.;
        __syncwarp();
        /* slots 32..63 take the drawn triangles' values (read back below through the same pointer arithmetic as the final flush) */
#pragma unroll
        for (int v = 0; v < 6; v++) s_qval[wid][v][lane + 32u] = qv[v];
        if (lane + 32u < queued) {
          s_queue[wid][lane] = rest; s_qbox[wid][lane] = restb; if (PEER) s_qsrc[wid][lane] = rests;
#pragma unroll
          for (int v = 0; v < 6; v++) s_qval[wid][v][lane] = restv[v];
        }
        queued -= 32u;
        __syncwarp();
        raster_triangle<MAPPED, PEER>(p, staged, C3D_TRI_ENTRY(q) | qs, C3D_TRI_INDEX(q), qb, &s_qval[wid][0][lane + 32u], canvas);
        __syncwarp();
      }
    }
  }
  __syncwarp();
  if (lane < queued) {
    const unsigned long long q = s_queue[wid][lane];
    const unsigned long long qs = PEER ? (unsigned long long)s_qsrc[wid][lane] << C3D_ENTRY_SRC_SHIFT : 0ull;
    raster_triangle<MAPPED, PEER>(p, staged, C3D_TRI_ENTRY(q) | qs, C3D_TRI_INDEX(q), s_qbox[wid][lane], &s_qval[wid][0][lane], canvas);
  }
  if (0u == lane && nullptr != culled) {
    if (hidden_total > 0u) atomicAdd(culled, (unsigned long long)hidden_total);
    if (hidden_tris > 0u) atomicAdd(culled + 1, (unsigned long long)hidden_tris);
  }
}

extern "C" int c3d_k_raster(const FrameParams *p, const unsigned long long *list, const unsigned long long *count, unsigned long long capacity,
                            Pixel *canvas, unsigned long long *culled, const unsigned long long *hiz_tiles, unsigned int part_begin,
                            unsigned int part_end, int sm_count, cudaStream_t s) {
  /* resident blocks per SM: the kernel is bound by dependent fp64 chains, so more warps help until spills hurt (tuned by measurement) */
  static int occ = -1;
  if (occ < 0) {
    const char *env = getenv("CONTOUR3D_RASTER_OCC");
    occ = (NULL != env && '\0' != *env) ? atoi(env) : 6;
  }
  if (part_end > C3D_RASTER_PARTS || part_begin >= part_end) return (int)cudaErrorInvalidValue;
  HiZ hiz;
  hiz.tile = hiz_tiles;
  hiz.tw = (int)((p->cam.w + C3D_HIZ_TILE - 1) / C3D_HIZ_TILE);
  /* the converter-table branches are compiled out unless an object of this frame has a tabulated converter */
  bool mapped = false;
  for (unsigned int n = 0; n < p->ncontours; n++) mapped = mapped || (nullptr != p->view[n].map);
  const bool peer = p->part.nranks > 1;      /* screen-partitioned frame: entries name their source rank, identity converters only */
  if (peer && mapped) return (int)cudaErrorInvalidValue;
#define C3D_RASTER_LAUNCH(MINB, PER_SM) \
  do { \
    if (peer) k_raster<MINB, false, true><<<sm_count * PER_SM, 128, 0, s>>>(*p, list, count, capacity, canvas, culled, hiz, part_begin, part_end); \
    else if (mapped) k_raster<MINB, true, false><<<sm_count * PER_SM, 128, 0, s>>>(*p, list, count, capacity, canvas, culled, hiz, part_begin, part_end); \
    else k_raster<MINB, false, false><<<sm_count * PER_SM, 128, 0, s>>>(*p, list, count, capacity, canvas, culled, hiz, part_begin, part_end); \
  } while (0)
  if (occ >= 8) C3D_RASTER_LAUNCH(8, 16);
  else if (occ >= 6) C3D_RASTER_LAUNCH(6, 12);
  else if (occ >= 5) C3D_RASTER_LAUNCH(5, 10);
  else C3D_RASTER_LAUNCH(4, 8);
#undef C3D_RASTER_LAUNCH
  return (int)cudaGetLastError();
}

/* farthest stored depth of every C3D_HIZ_TILE^2 pixel tile (pixels outside the screen do not count); one warp per tile */
__global__ void __launch_bounds__(256) k_hiz_build(const Pixel *__restrict__ canvas, const int w, const int h, const int tw, const int th,
                                                   unsigned long long *__restrict__ tiles) {
  const unsigned int lane = threadIdx.x & 31u;
  const int ntiles = tw * th;
  for (int t = (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)); t < ntiles; t += (int)(gridDim.x * (blockDim.x >> 5))) {
    const int ti = t % tw, tj = t / tw;
    unsigned long long lo = ~0ull;
    for (int q = (int)lane; q < C3D_HIZ_TILE * C3D_HIZ_TILE; q += 32) {
      const int pi = ti * C3D_HIZ_TILE + q % C3D_HIZ_TILE, pj = tj * C3D_HIZ_TILE + q / C3D_HIZ_TILE;
      if (pi < w && pj < h) {
        const unsigned long long d = __ldcg(reinterpret_cast<const unsigned long long *>(canvas + (long long)pj * w + pi));
        lo = d < lo ? d : lo;
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const unsigned long long other = __shfl_xor_sync(C3D_FULL_MASK, lo, d);
      lo = other < lo ? other : lo;
    }
    if (0u == lane) tiles[t] = lo;
  }
}

extern "C" int c3d_k_hiz_build(const Pixel *canvas, long long w, long long h, unsigned long long *tiles, int sm_count, cudaStream_t s) {
  const int tw = (int)((w + C3D_HIZ_TILE - 1) / C3D_HIZ_TILE), th = (int)((h + C3D_HIZ_TILE - 1) / C3D_HIZ_TILE);
  k_hiz_build<<<sm_count * 8, 256, 0, s>>>(canvas, (int)w, (int)h, tw, th, tiles);
  return (int)cudaGetLastError();
}

/* flag[0] = 1 when a work buffer of this frame overflowed (agreed across ranks by a maximum before anybody redoes the frame) */
__global__ void k_overflow_flag(const unsigned long long *__restrict__ nrec, const unsigned long long rec_cap,
                                const unsigned long long *__restrict__ nlist, const unsigned long long list_cap,
                                const unsigned long long *__restrict__ nout, const unsigned long long out_cap,
                                const unsigned long long *__restrict__ nover, unsigned long long *__restrict__ flag) {
  bool over = *nrec > rec_cap || *nlist > list_cap;
  if (nullptr != nout) over = over || *nout > out_cap;
  if (nullptr != nover) over = over || 0ull != *nover;
  *flag = over ? 1ull : 0ull;
}

extern "C" int c3d_k_overflow_flag(const unsigned long long *nrec, unsigned long long rec_cap, const unsigned long long *nlist,
                                   unsigned long long list_cap, const unsigned long long *nout, unsigned long long out_cap,
                                   const unsigned long long *nover, unsigned long long *flag, cudaStream_t s) {
  k_overflow_flag<<<1, 1, 0, s>>>(nrec, rec_cap, nlist, list_cap, nout, out_cap, nover, flag);
  return (int)cudaGetLastError();
}

/* ------------------------------------------------------------------ lines */

/* one CTA per segment; draw_line_between_two_points, line.c:26-118 */
__global__ void __launch_bounds__(128) k_lines(const Cam cam, const LineDev *__restrict__ lines, const unsigned int nlines,
                                               const unsigned long long first_obj, Pixel *__restrict__ canvas) {
  const unsigned long long seg_global = blockIdx.x;
  unsigned int l = 0;
  while (l + 1 < nlines && seg_global >= lines[l + 1].first_segment) l++;
  const LineDev &ln = lines[l];
  const unsigned long long seg = seg_global - ln.first_segment;
  if (seg >= ln.nsegments) return;
  /* s = head + n * delta, e = head + (n + 1) * delta (line.c:157-158) */
  const V3 e0 = v_of(ln.e0), dl = v_of(ln.delta);
  V3 ps, pe;
  if (nullptr != ln.points) {       /* converted on the host with the caller's function (line.c:39-40) */
    ps = V3{__ldg(ln.points + 3ull * seg), __ldg(ln.points + 3ull * seg + 1ull), __ldg(ln.points + 3ull * seg + 2ull)};
    pe = V3{__ldg(ln.points + 3ull * seg + 3ull), __ldg(ln.points + 3ull * seg + 4ull), __ldg(ln.points + 3ull * seg + 5ull)};
  } else {
    ps = v_add(e0, v_mul((double)seg, dl));
    pe = v_add(e0, v_mul((double)(seg + 1ull), dl));
  }
  V3 p0, p1;
  if (!project(cam, ps, p0)) return;
  if (!project(cam, pe, p1)) return;
  p0.x = r_mul(r_add(0.5, p0.x), cam.W); p0.y = r_mul(r_add(0.5, p0.y), cam.H);
  p1.x = r_mul(r_add(0.5, p1.x), cam.W); p1.y = r_mul(r_add(0.5, p1.y), cam.H);
  const long long imin = clamp_ll(0, cam.w - 1, __double2ll_rz(r_sub(fmin(p0.x, p1.x), 1.)));
  const long long imax = clamp_ll(0, cam.w - 1, __double2ll_rz(r_add(fmax(p0.x, p1.x), 1.)));
  const long long jmin = clamp_ll(0, cam.h - 1, __double2ll_rz(r_sub(fmin(p0.y, p1.y), 1.)));
  const long long jmax = clamp_ll(0, cam.h - 1, __double2ll_rz(r_add(fmax(p0.y, p1.y), 1.)));
  const double a = r_sub(p1.y, p0.y);
  const double b = r_add(-p1.x, p0.x);
  const double dx = r_sub(p1.x, p0.x), dy = r_sub(p1.y, p0.y);
  const double cc = r_sub(r_mul(dx, p0.y), r_mul(dy, p0.x));
  const double ab2 = r_add(r_mul(a, a), r_mul(b, b));
  const double len2 = r_add(r_mul(dx, dx), r_mul(dy, dy));
  const unsigned long long mykey = ((first_obj + l + 1ull) << C3D_KEY_OBJ_SHIFT) | seg;
  const long long nx = imax - imin + 1, ny = jmax - jmin + 1;
  for (long long q = threadIdx.x; q < nx * ny; q += blockDim.x) {
    const long long i = imin + q % nx, j = jmin + q / nx;
    const double fi = (double)i, fj = (double)j;
    const double lin = r_add(r_add(r_mul(a, fi), r_mul(b, fj)), cc);
    const double dist2 = r_div(r_mul(lin, lin), ab2);
    if (ln.half_width_sq < dist2) continue;
    const double param = r_div(r_mul(-1., r_add(r_mul(dx, r_sub(p0.x, fi)), r_mul(dy, r_sub(p0.y, fj)))), len2);
    const double dist = r_div(1., r_add(r_div(param, p1.z), r_div(r_sub(1., param), p0.z)));
    submit_fragment(canvas + j * cam.w + i, ord_of(dist), mykey);
  }
}

extern "C" int c3d_k_lines(const Cam *c, const LineDev *lines_dev, const LineDev *lines_host, size_t nlines, unsigned long long first_obj,
                           Pixel *canvas, cudaStream_t s) {
  if (0 == nlines) return 0;
  const unsigned long long total = lines_host[nlines - 1].first_segment + lines_host[nlines - 1].nsegments;
  if (0 == total) return 0;
  k_lines<<<(unsigned int)total, 128, 0, s>>>(*c, lines_dev, (unsigned int)nlines, first_obj, canvas);
  return (int)cudaGetLastError();
}

/* ------------------------------------------------------------------ shade */

/* depth plane of the canvas, contiguous, for the cross-rank ncclMax (output.c:10-30) */
__global__ void k_depth_extract(const Pixel *__restrict__ canvas, unsigned long long *__restrict__ depth, const size_t npix) {
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (size_t)gridDim.x * blockDim.x) {
    depth[p] = canvas[p].depth;
  }
}

extern "C" int c3d_k_depth_extract(const Pixel *canvas, unsigned long long *depth, size_t npix, cudaStream_t s) {
  k_depth_extract<<<148 * 8, 256, 0, s>>>(canvas, depth, npix);
  return (int)cudaGetLastError();
}

/* fp32 copies of the grid coordinates (visibility pre-tests) */
__global__ void k_to_float(const double *__restrict__ in, float *__restrict__ out, const size_t n) {
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) out[q] = (float)in[q];
}

extern "C" int c3d_k_to_float(const double *in, float *out, size_t n, cudaStream_t s) {
  if (0 == n) return 0;
  k_to_float<<<(unsigned int)((n + 255) / 256 < 256 ? (n + 255) / 256 : 256), 256, 0, s>>>(in, out, n);
  return (int)cudaGetLastError();
}

/* ------------------------------------------------------------------- pack */

/* RGB bytes in file order: row j of the file is canvas row h-1-j (output.c:113-123) */
__global__ void k_pack(const unsigned int *__restrict__ rgba, const unsigned int bg, const long long w, const long long h,
                       unsigned char *__restrict__ out) {
  const long long npix = w * h;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < npix; q += (long long)gridDim.x * blockDim.x) {
    const long long j = q / w, i = q % w;
    unsigned int px = rgba[(h - 1 - j) * w + i];
    if (0u == (px >> 24)) px = bg;
    out[3 * q + 0] = (unsigned char)(px & 255u);
    out[3 * q + 1] = (unsigned char)((px >> 8) & 255u);
    out[3 * q + 2] = (unsigned char)((px >> 16) & 255u);
  }
}

extern "C" int c3d_k_pack(const unsigned int *rgba, const unsigned char bg[3], long long w, long long h, unsigned char *out, cudaStream_t s) {
  const unsigned int bgw = ((unsigned int)bg[2] << 16) | ((unsigned int)bg[1] << 8) | bg[0];
  const long long npix = w * h;
  long long blocks = (npix + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_pack<<<(unsigned int)(blocks > 0 ? blocks : 1), 256, 0, s>>>(rgba, bgw, w, h, out);
  return (int)cudaGetLastError();
}

__global__ void k_canvas_split(const Pixel *__restrict__ canvas, double *__restrict__ depth, unsigned long long *__restrict__ key, const size_t npix) {
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (size_t)gridDim.x * blockDim.x) {
    if (depth) depth[p] = ord_to_double(canvas[p].depth);
    if (key) key[p] = canvas[p].key;
  }
}

extern "C" int c3d_k_canvas_split(const Pixel *canvas, double *depth, unsigned long long *key, size_t npix, cudaStream_t s) {
  k_canvas_split<<<148 * 8, 256, 0, s>>>(canvas, depth, key, npix);
  return (int)cudaGetLastError();
}

/* ------------------------------------------------- inspection (parity tests) */

__global__ void k_case_codes(const FieldView f, const DevTables tab, unsigned int *__restrict__ codes, unsigned char *__restrict__ counts) {
  const long long cx = f.n[0] - 1, cy = f.n[1] - 1, cz = f.n[2] - 1;
  const long long ncubes = cx * cy * cz;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < ncubes; q += (long long)gridDim.x * blockDim.x) {
    const long long i = q % cx, j = (q / cx) % cy, k = q / (cx * cy);
    const long long base = k * f.sz + j * f.sy + i;
    unsigned int code = 0;
#pragma unroll
    for (int n = 0; n < 8; n++) {
      const double v = ld_value(f, base + (n & 1) + ((n >> 1) & 1) * f.sy + (n >> 2) * f.sz);
      code |= (unsigned int)(f.thr < v) << n;
    }
    codes[q] = __ldg(tab.code24 + code);
    counts[q] = __ldg(tab.count + code);
  }
}

extern "C" int c3d_k_case_codes(const FieldView *f, const DevTables *t, unsigned int *codes, unsigned char *counts, cudaStream_t s) {
  k_case_codes<<<148 * 8, 256, 0, s>>>(*f, *t, codes, counts);
  return (int)cudaGetLastError();
}

struct TriangleRecord {     /* == c3d_triangle_t, 208 bytes */
  unsigned long long cube;
  unsigned int tri, has_vn, edge[3], pad;
  double area, v[9], fn[3], vn[9];
};

/* full triangle records for every listed cube (mesh export / parity dumps) */
__global__ void __launch_bounds__(128) k_emit(const FieldView f, const DevTables tab, const unsigned long long *__restrict__ list,
                                              const unsigned long long nlist, unsigned long long *cursor,
                                              const unsigned long long capacity, TriangleRecord *__restrict__ out) {
  const long long cx = f.n[0] - 1, cy = f.n[1] - 1, cz = f.n[2] - 1;
  for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < nlist;
       idx += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long entry = list[idx];
    const unsigned int code = C3D_ENTRY_CODE(entry);
    const long long i = C3D_ENTRY_I(entry), j = C3D_ENTRY_J(entry), k = C3D_ENTRY_K(entry);
    const unsigned long long cube = (unsigned long long)((k * cy + j) * cx + i);
    const unsigned int ntri = __ldg(tab.count + code);
    const unsigned long long at = atomicAdd(cursor, (unsigned long long)ntri);
    /* the reference smooths cubes 1..m-3 of every direction only (normal.c:54-55, contour/main.c:71,92-93) */
    const bool inner = i >= 1 && i + 1 < cx && j >= 1 && j + 1 < cy && k >= 1 && k + 1 < cz;
    Cube c;
    load_cube(f, i, j, k, c);
    for (unsigned int s = 0; s < ntri; s++) {
      if (at + s >= capacity) break;
      TriangleRecord rec;
      V3 v[3], fn;
      triangle_vertices(c, f.thr, __ldg(tab.pairs + code * C3D_MAX_TRIS + s), v);
      face_normal(v, fn, rec.area);
      const unsigned int edges = __ldg(tab.edges + code * C3D_MAX_TRIS + s);
      rec.cube = cube; rec.tri = s; rec.has_vn = inner ? 1u : 0u; rec.pad = 0u;
      for (int m = 0; m < 3; m++) {
        const unsigned int e = (edges >> (5 * m)) & 31u;
        rec.edge[m] = e;
        rec.v[3 * m + 0] = v[m].x; rec.v[3 * m + 1] = v[m].y; rec.v[3 * m + 2] = v[m].z;
        V3 n = V3{0., 0., 0.};
        if (inner) n = vertex_normal(f, tab, i, j, k, e);
        rec.vn[3 * m + 0] = n.x; rec.vn[3 * m + 1] = n.y; rec.vn[3 * m + 2] = n.z;
      }
      rec.fn[0] = fn.x; rec.fn[1] = fn.y; rec.fn[2] = fn.z;
      out[at + s] = rec;
    }
  }
}

extern "C" int c3d_k_emit(const FieldView *f, const DevTables *t, const unsigned long long *list, unsigned long long nlist,
                          unsigned long long *cursor, unsigned long long capacity, void *records, cudaStream_t s) {
  if (0 == nlist) return 0;
  k_emit<<<148 * 4, 128, 0, s>>>(*f, *t, list, nlist, cursor, capacity, static_cast<TriangleRecord *>(records));
  return (int)cudaGetLastError();
}

/* ------------------------------------------------------------------- halo */

/* rows [j0, j0+2) of planes [k0, k0+nk) <-> contiguous buffer (the y halo is
   strided: communicate_in_y's hvector, contour/extend.c:71-77) */
template <typename T, bool PACK>
__global__ void k_halo_y(T *__restrict__ data, const long long nx, const long long sy, const long long sz, const long long k0,
                         const long long nk, const long long j0, T *__restrict__ buf) {
  const long long total = nk * 2 * nx;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const long long i = q % nx, r = (q / nx) % 2, k = q / (2 * nx);
    T *cell = data + (k0 + k) * sz + (j0 + r) * sy + i;
    if (PACK) buf[q] = *cell; else *cell = buf[q];
  }
}

extern "C" int c3d_k_halo_pack_y(const void *data, int elem_bytes, long long nx, long long sy, long long sz, long long k0, long long nk,
                                 long long j0, void *buf, cudaStream_t s) {
  if (4 == elem_bytes) k_halo_y<float, true><<<148 * 4, 256, 0, s>>>((float *)data, nx, sy, sz, k0, nk, j0, (float *)buf);
  else k_halo_y<double, true><<<148 * 4, 256, 0, s>>>((double *)data, nx, sy, sz, k0, nk, j0, (double *)buf);
  return (int)cudaGetLastError();
}

extern "C" int c3d_k_halo_unpack_y(void *data, int elem_bytes, long long nx, long long sy, long long sz, long long k0, long long nk,
                                   long long j0, const void *buf, cudaStream_t s) {
  if (4 == elem_bytes) k_halo_y<float, false><<<148 * 4, 256, 0, s>>>((float *)data, nx, sy, sz, k0, nk, j0, (float *)buf);
  else k_halo_y<double, false><<<148 * 4, 256, 0, s>>>((double *)data, nx, sy, sz, k0, nk, j0, (double *)buf);
  return (int)cudaGetLastError();
}

/* -------------------------------------------------------- synthetic fields */

__device__ __forceinline__ unsigned int lowbias32(unsigned int x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

__device__ __forceinline__ float lattice_value(const int x, const int y, const int z, const unsigned int seed) {
  const unsigned int h = lowbias32((unsigned int)x * 0x9E3779B1u ^ lowbias32((unsigned int)y * 0x85EBCA77u ^ lowbias32((unsigned int)z * 0xC2B2AE3Du ^ seed)));
  return (float)(h >> 8) * (1.0f / 16777216.0f);
}

__device__ float value_noise(const float x, const float y, const float z, const unsigned int seed) {
  const float fx = floorf(x), fy = floorf(y), fz = floorf(z);
  const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
  float tx = x - fx, ty = y - fy, tz = z - fz;
  tx = tx * tx * (3.f - 2.f * tx); ty = ty * ty * (3.f - 2.f * ty); tz = tz * tz * (3.f - 2.f * tz);
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 8; c++) {
    const int dx = c & 1, dy = (c >> 1) & 1, dz = c >> 2;
    const float w = (dx ? tx : 1.f - tx) * (dy ? ty : 1.f - ty) * (dz ? tz : 1.f - tz);
    acc += w * lattice_value(ix + dx, iy + dy, iz + dz, seed);
  }
  return acc;
}

/* sum_o gain^o (2 valuenoise(2^(o+2) (p + 0.5), seed + o) - 1) / sum_o gain^o, in [-1, 1]; the host twin is c3d_synth_host.c */
__device__ float fractal_noise(const float x, const float y, const float z, const unsigned int seed, const int octaves, const float gain) {
  float amp = 1.f, sum = 0.f, norm = 0.f, freq = 4.f;
  for (int o = 0; o < octaves; o++) {
    sum += amp * (2.f * value_noise((x + 0.5f) * freq, (y + 0.5f) * freq, (z + 0.5f) * freq, seed + (unsigned int)o) - 1.f);
    norm += amp;
    amp *= gain;
    freq *= 2.f;
  }
  return sum / norm;
}

/* |curl F| of F = (noise(seed), noise(seed + 101), noise(seed + 202)), central differences of half a cell, scaled and shifted */
__device__ float curl_magnitude(const float x, const float y, const float z, const float h, const unsigned int seed, const int octaves,
                                const float gain, const float scale) {
  const float inv = 1.f / (2.f * h);
  const float dFz_dy = (fractal_noise(x, y + h, z, seed + 202u, octaves, gain) - fractal_noise(x, y - h, z, seed + 202u, octaves, gain)) * inv;
  const float dFy_dz = (fractal_noise(x, y, z + h, seed + 101u, octaves, gain) - fractal_noise(x, y, z - h, seed + 101u, octaves, gain)) * inv;
  const float dFx_dz = (fractal_noise(x, y, z + h, seed, octaves, gain) - fractal_noise(x, y, z - h, seed, octaves, gain)) * inv;
  const float dFz_dx = (fractal_noise(x + h, y, z, seed + 202u, octaves, gain) - fractal_noise(x - h, y, z, seed + 202u, octaves, gain)) * inv;
  const float dFy_dx = (fractal_noise(x + h, y, z, seed + 101u, octaves, gain) - fractal_noise(x - h, y, z, seed + 101u, octaves, gain)) * inv;
  const float dFx_dy = (fractal_noise(x, y + h, z, seed, octaves, gain) - fractal_noise(x, y - h, z, seed, octaves, gain)) * inv;
  const float cx = dFz_dy - dFy_dz, cy = dFx_dz - dFz_dx, cz = dFy_dx - dFx_dy;
  return scale * sqrtf(cx * cx + cy * cy + cz * cz) - 1.f;
}

/* kind 0: gyroid sin X cos Y + sin Y cos Z + sin Z cos X, X = 2 pi 4 x;
   kind 1: sum_o gain^o * valuenoise(2^(o+2) p, seed + o), mapped to [-1, 1];
   kind 2: "vorticity magnitude": scaled |curl| of a three-component noise field (config 4 of BASELINE.json).
   Coordinates are the cell centres x_n = -0.5 + (n + 0.5) / N of the GLOBAL grid. */
template <typename T>
__global__ void k_synth(T *__restrict__ data, const long long nx, const long long ny, const long long nz, const long long sy,
                        const long long sz, const long long ox, const long long oy, const long long oz, const long long gx,
                        const long long gy, const long long gz, const int kind, const unsigned int seed, const int octaves,
                        const float gain) {
  const long long total = nx * ny * nz;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const long long i = q % nx, j = (q / nx) % ny, k = q / (nx * ny);
    const float x = -0.5f + ((float)(i + ox) + 0.5f) / (float)gx;
    const float y = -0.5f + ((float)(j + oy) + 0.5f) / (float)gy;
    const float z = -0.5f + ((float)(k + oz) + 0.5f) / (float)gz;
    float v;
    if (0 == kind) {
      const float s = 6.2831853071795864769f * 4.f;
      v = sinf(s * x) * cosf(s * y) + sinf(s * y) * cosf(s * z) + sinf(s * z) * cosf(s * x);
    } else if (1 == kind) {
      v = fractal_noise(x, y, z, seed, octaves, gain);
    } else {
      v = curl_magnitude(x, y, z, 0.5f / (float)gx, seed, octaves, gain, C3D_CURL_SCALE);
    }
    data[k * sz + j * sy + i] = (T)v;
  }
}

extern "C" int c3d_k_synth(void *data, int elem_bytes, const long long n[3], long long sy, long long sz, const long long goff[3],
                           const long long gl[3], int kind, unsigned int seed, int octaves, double gain, cudaStream_t s) {
  if (4 == elem_bytes)
    k_synth<float><<<148 * 16, 256, 0, s>>>((float *)data, n[0], n[1], n[2], sy, sz, goff[0], goff[1], goff[2], gl[0], gl[1], gl[2], kind, seed, octaves, (float)gain);
  else
    k_synth<double><<<148 * 16, 256, 0, s>>>((double *)data, n[0], n[1], n[2], sy, sz, goff[0], goff[1], goff[2], gl[0], gl[1], gl[2], kind, seed, octaves, (float)gain);
  return (int)cudaGetLastError();
}
