/*
 * c3d_internal.h -- structures shared by the host runtime (c3d_runtime.cu)
 * and the kernels (c3d_kernels.cu).  Not part of the C-ABI.
 */
#if !defined(C3D_INTERNAL_H)
#define C3D_INTERNAL_H

#include <stdint.h>
#include <cuda_runtime.h>
#include "c3d_tables.h"

struct DevTables {
  const uint8_t *count;
  const uint32_t *code24;
  const uint32_t *pairs;
  const uint32_t *emask;
  const uint32_t *edges;
  const uint16_t *touch;
};

/* Cartesian positions of the grid points of a contour object whose coordinate converter is not the identity
   (include/contour3d.h:31-33; the reference calls it for every cube corner, triangulate.c:156-173): component c of point
   (i, j, k) of the extended local array is pos[c][i * ps[c][0] + j * ps[c][1] + k * ps[c][2]].  The tables are filled on the
   host with the caller's own function; a stride of 0 marks a direction the component does not depend on. */
struct PosMap {
  const double *pos[3];
  const float *posf[3];    /* the same rounded to fp32 (conservative visibility pre-tests only) */
  long long ps[3][3];
};

#if defined(__CUDACC__)
__device__ __forceinline__ void mapped_posf(const PosMap *__restrict__ m, const long long i, const long long j, const long long k,
                                            float &x, float &y, float &z) {
  x = __ldg(m->posf[0] + (i * m->ps[0][0] + j * m->ps[0][1] + k * m->ps[0][2]));
  y = __ldg(m->posf[1] + (i * m->ps[1][0] + j * m->ps[1][1] + k * m->ps[1][2]));
  z = __ldg(m->posf[2] + (i * m->ps[2][0] + j * m->ps[2][1] + k * m->ps[2][2]));
}
#endif

/* one contour object as the kernels see it: the extended local array */
struct FieldView {
  const void *data;        /* first point of the extended array */
  int elem_bytes;          /* 4 or 8 */
  long long n[3];          /* points in x, y, z */
  long long sy, sz;        /* row / plane strides in elements */
  const double *g[3];      /* device coordinates of exactly those points */
  const float *gf[3];      /* the same rounded to fp32 (conservative visibility pre-tests only) */
  const PosMap *map;       /* device pointer; nullptr: identity converter, positions are (g[0][i], g[1][j], g[2][k]) */
  long long off[3];        /* global index of the array's first point: winner keys carry GLOBAL cube indices, so that the
                              drawing order that decides depth ties (render.c:125-132) is the one-rank order on any decomposition */
  double thr;              /* iso value; classification is thr < value (strict) */
  float thr_ge;            /* fp32 fields: smallest float f with (double)f > thr, so thr < v <=> v >= thr_ge */
  unsigned char color[3];
};

/* camera + screen + light, with the projection invariants of project.c:34-36 hoisted */
struct Cam {
  double cam[3], center[3], lx[3], ly[3], normal[3], light[3];
  double lxlx, lyly;       /* lx.lx, ly.ly */
  double nc, d;            /* normal.cam, normal.center */
  double W, H;             /* (double)width, (double)height */
  long long w, h;
};

struct LineDev {
  double e0[3], delta[3];  /* first point and per-segment step (line.c:138-140) */
  double half_width_sq;    /* pow(0.5 * width, 2), line.c:87 */
  const double *points;    /* device; nullptr: identity converter (points are e0 + n * delta), else nitems x 3 converted points */
  unsigned long long first_segment; /* index of this line's first segment in the frame-wide numbering */
  unsigned long long nsegments;
  unsigned char color[3];
};

/* z-buffer entry: order-preserving integer image of the fp64 depth + winner key.
   One 128-bit compare-and-swap keeps the pair consistent, so the raster needs a
   single pass: the lexicographic maximum of (depth, key) is exactly "nearest
   wins, later in drawing order wins ties" (render.c:125-132). */
struct __align__(16) Pixel { unsigned long long depth, key; };

/* 64-bit winner key: object index + 1 in the top 12 bits, primitive below.
   contours: primitive = k:16 | j:16 | i:16 | triangle:4 (cube indices on the
   extended local array); lines: primitive = segment.
   Larger key = later in the reference's drawing order (contour, k, j, i, t;
   then lines), which is what wins a depth tie (render.c:125-132). */
#define C3D_KEY_OBJ_SHIFT 52
#define C3D_KEY_PRIM_MASK ((1ull << C3D_KEY_OBJ_SHIFT) - 1ull)
#define C3D_MAX_AXIS 65535ll      /* cubes per direction that fit the 16-bit fields */

/* active-cube list entry: code:8 | i:16 | j:16 | k:16 | object:4 | irregular:1 (bit 60: a corner value may be
   within rounding of the iso-value, so a triangle may have collapsed to a point; see cube_hidden) */
#define C3D_ENTRY(obj, i, j, k, code) \
  (((unsigned long long)(obj) << 56) | ((unsigned long long)(k) << 40) | ((unsigned long long)(j) << 24) | ((unsigned long long)(i) << 8) | (unsigned long long)(code))
#define C3D_ENTRY_CODE(e) ((unsigned int)((e) & 255ull))
#define C3D_ENTRY_I(e) ((long long)(((e) >> 8) & 65535ull))
#define C3D_ENTRY_J(e) ((long long)(((e) >> 24) & 65535ull))
#define C3D_ENTRY_K(e) ((long long)(((e) >> 40) & 65535ull))
#define C3D_ENTRY_OBJ(e) ((unsigned int)((e) >> 56) & 15u)
#define C3D_ENTRY_IRREGULAR_BIT (1ull << 60)
#define C3D_ENTRY_IRREGULAR(e) (0ull != ((e) & C3D_ENTRY_IRREGULAR_BIT))
/* screen-partitioned multi-rank frames: bits 61-63 name the rank whose array the cube indices refer to */
#define C3D_ENTRY_SRC_SHIFT 61
#define C3D_ENTRY_SRC(e) ((unsigned int)((e) >> C3D_ENTRY_SRC_SHIFT) & 7u)
#define C3D_MAX_RANKS 8               /* ranks of a screen-partitioned frame (3 bits of the entry) */
#define C3D_PEER_VIEW(src, obj) (((src) << 4) | (obj))

#define C3D_MAX_OBJECTS 16
#define C3D_DEPTH_BUCKETS 2048        /* view-depth slabs of the front-to-back ordering (multiple of 1024) */
#define C3D_PART_BUCKETS 1024         /* the same in screen-partitioned frames, where every rank keeps one set per destination rank */
#define C3D_HIZ_TILE 8                /* pixels per side of a coarse z-buffer tile */
#define C3D_RASTER_PARTS 64u          /* the raster's list is cut at multiples of 1/64 of its length */
#define C3D_SWEEP_MAX_THR 4        /* iso-values of one array classified by a single sweep launch */

struct SweepRange { long long lo[3], hi[3]; }; /* inclusive cube ranges to emit */

/* fp32 copy of the projection (project.c:28-65) for the conservative occlusion test of whole cubes:
   sx = ax + t * (ray . lxs), sy = ay + t * (ray . lys), t = c / (n . ray), ray = p - cam */
struct CullCam {
  float cam[3], n[3], lxs[3], lys[3];
  float ax, ay, c;
  float inv_abs_c, lx_norm, ly_norm;   /* 1 / |c|, |lxs|, |lys|: screen magnification bound of the triangle pre-test */
  float W, H;
  int w, h;
  int enabled;
  int valid;                           /* the constants are finite and the screen fits: cube_owners() may use them even when culling is off */
};

/*
 * Screen partition of a multi-rank frame ("sort-first"): the screen is cut into square tiles dealt out to the ranks in a
 * fixed pattern; a rank rasterises EVERY cube (its own and the other ranks', read through peer memory) that can touch
 * its tiles, nearest first, so the occlusion tests see the whole scene exactly as on one GPU and the raster work divides by
 * the number of ranks.  Shading goes back to the rank that owns the winning cube's array (it needs the neighbouring cubes).
 */
struct Partition {
  int nranks, me;                      /* nranks == 0: not partitioned */
  int tile_shift;                      /* log2 of the tile edge in pixels */
  int tiles_x, tiles_y;
  unsigned int recip;                  /* ceil(2^32 / nranks): v % nranks for small v without a division */
  int py, pz;                          /* process grid (x is never split) */
  unsigned short yoff[C3D_MAX_RANKS + 1], zoff[C3D_MAX_RANKS + 1];   /* global index of the first point of block c; [p] = N */
  unsigned char home[C3D_MAX_RANKS];   /* communicator rank of block (cy, cz): home[cy * pz + cz] */
  const unsigned char *tile_map;       /* device: owner of tile ty * tiles_x + tx; nullptr: the fixed pattern (tx + 3 ty) mod nranks.
                                          The same table on every rank; rebuilt when the ranks' loads drift apart (frame_once) */
};

#if defined(__CUDACC__)
__device__ __forceinline__ unsigned int tile_owner(const Partition &q, const unsigned int tx, const unsigned int ty) {
  if (nullptr != q.tile_map) return (unsigned int)__ldg(q.tile_map + ty * (unsigned int)q.tiles_x + tx);
  const unsigned int v = tx + 3u * ty;
  return v - __umulhi(v, q.recip) * (unsigned int)q.nranks;
}
__device__ __forceinline__ bool pixel_owned(const Partition &q, const int pi, const int pj) {
  return tile_owner(q, (unsigned int)pi >> q.tile_shift, (unsigned int)pj >> q.tile_shift) == (unsigned int)q.me;
}
/* the rank whose INTERIOR holds point (gj, gk): it lists (and can shade) the cube whose low corner that is */
__device__ __forceinline__ unsigned int cube_home(const Partition &q, const unsigned int gj, const unsigned int gk) {
  unsigned int cy = 0u, cz = 0u;
  for (int c = 1; c < q.py; c++) cy += (gj >= (unsigned int)q.yoff[c]) ? 1u : 0u;
  for (int c = 1; c < q.pz; c++) cz += (gk >= (unsigned int)q.zoff[c]) ? 1u : 0u;
  return (unsigned int)q.home[cy * (unsigned int)q.pz + cz];
}
#endif

/* everything the raster / shading kernels need, passed by value (constant bank) */
struct FrameParams {
  Cam cam;
  CullCam cull;
  DevTables tab;
  const double *px, *py;     /* pixel-centre coordinates, render.c:94-98, tabulated once per screen size */
  const float *pxf, *pyf;    /* the same rounded to fp32 (conservative coverage pre-test) */
  unsigned int ncontours;
  FieldView view[C3D_MAX_OBJECTS];
  /* screen-partitioned frames only */
  Partition part;
  const FieldView *pviews;   /* device table: view C3D_PEER_VIEW(src, obj) = object `obj` as rank `src` holds it (data through peer memory) */
};

/* a pixel won by a contour triangle, on its way to the rank that shades it */
struct __align__(16) WorkItem { unsigned int pix, pad; unsigned long long key; };

/* device scratch of the deferred shading (c3d_shade.cu), sized by the screen */
#define C3D_SHADE_PIX_BYTES 56
struct ShadeScratch {
  WorkItem *work;                /* pixels won by a contour triangle, compacted: npix items */
  unsigned int *counters;        /* [0] listed pixels, [1] distinct vertices, [2] pixels of the first pass, [3] vertices of the first pass */
  void *pix;                     /* per listed pixel: weights, cube, vertex slots: npix x C3D_SHADE_PIX_BYTES */
  unsigned long long *table;     /* open-addressing set of the vertices (global cube edges) the listed pixels reference */
  unsigned int table_mask;       /* slots in use this frame - 1 (a power of two, at most table_slots) */
  unsigned int table_slots;      /* slots allocated */
  unsigned int *vlist;           /* occupied slots in order of first insertion: min(slots, 3 npix) words */
  double *vn;                    /* unit normal of listed vertex v: 3 doubles each, vn_cap vertices */
  unsigned int vn_cap;
};

#if defined(__cplusplus)
extern "C" {
#endif

/* launchers (c3d_kernels.cu); all asynchronous on `s`, return cudaError_t as int */
int c3d_k_clear(Pixel *canvas, size_t npix, cudaStream_t s);
int c3d_k_pixel_tables(double *px, double *py, float *pxf, float *pyf, long long w, long long h, cudaStream_t s);
int c3d_k_sweep(const FieldView *f, const SweepRange *r, unsigned int obj, unsigned long long *list, unsigned long long *count,
                unsigned long long capacity, int sm_count, cudaStream_t s);
/* tiled bulk-copy sweep (c3d_sweep.cu): needs 16-byte aligned rows; up to C3D_SWEEP_MAX_THR iso-values per launch */
int c3d_k_sweep_tiles_ok(const FieldView *f);
int c3d_k_sweep_tiles(const FieldView *f, const SweepRange *r, int nthr, const double *thr, const float *thr_ge, const unsigned int *obj,
                      const int flip[3], unsigned int *tile_cursor, void *records, unsigned long long *count, unsigned long long capacity,
                      int sm_count, cudaStream_t s);
/* sorted != 0: counting sort of the entries by view depth, nearest first (needs fp->cull and the range of n.(p - cam)) */
int c3d_k_expand(const FrameParams *fp, int sorted, float near_depth, float far_depth,
                 const void *records, const unsigned long long *nrecords, unsigned long long rec_capacity, unsigned long long *list,
                 unsigned long long *count, unsigned long long capacity, unsigned int *histogram, unsigned long long *cursors,
                 unsigned long long *per_object, unsigned int nobjects, int sm_count, cudaStream_t s);
/* draws the entries [n * part_begin / C3D_RASTER_PARTS, n * part_end / C3D_RASTER_PARTS) of the list; hiz_tiles may be NULL */
int c3d_k_raster(const FrameParams *p, const unsigned long long *list, const unsigned long long *count, unsigned long long capacity,
                 Pixel *canvas, unsigned long long *culled, const unsigned long long *hiz_tiles, unsigned int part_begin,
                 unsigned int part_end, int sm_count, cudaStream_t s);
int c3d_k_hiz_build(const Pixel *canvas, long long w, long long h, unsigned long long *tiles, int sm_count, cudaStream_t s);
/* nout / nover may be NULL */
int c3d_k_overflow_flag(const unsigned long long *nrec, unsigned long long rec_cap, const unsigned long long *nlist,
                        unsigned long long list_cap, const unsigned long long *nout, unsigned long long out_cap,
                        const unsigned long long *nover, unsigned long long *flag, cudaStream_t s);
int c3d_k_lines(const Cam *c, const LineDev *lines_dev, const LineDev *lines_host, size_t nlines, unsigned long long first_obj,
                Pixel *canvas, cudaStream_t s);
int c3d_k_depth_extract(const Pixel *canvas, unsigned long long *depth, size_t npix, cudaStream_t s);
int c3d_k_shade(const FrameParams *p, const LineDev *lines_dev, const Pixel *canvas, const unsigned long long *depth_global,
                unsigned int *rgba, size_t npix, const ShadeScratch *scratch, unsigned long long *overflowed,
                int sm_count, cudaStream_t s);
int c3d_k_shade_begin(const ShadeScratch *scratch, cudaStream_t s);
int c3d_k_shade_finish(const FrameParams *p, unsigned int *rgba, const ShadeScratch *scratch, unsigned long long *overflowed, int helpers, int sm_count, cudaStream_t s);
int c3d_k_shade_select_part(const FrameParams *p, const LineDev *lines_dev, const Pixel *canvas, unsigned int *rgba_out,
                            WorkItem *seg, unsigned int seg_cap, unsigned int *nseg, unsigned long long *overflowed, int sm_count, cudaStream_t s);
/* screen-partitioned frames (c3d_sweep.cu, c3d_kernels.cu) */
int c3d_k_expand_part(const FrameParams *fp, float near_depth, float far_depth, const void *records, const unsigned long long *nrecords,
                      unsigned long long rec_capacity, unsigned long long *outbox, unsigned long long *count, unsigned long long capacity,
                      unsigned int *histogram, unsigned long long *cursors, unsigned long long *seg_start, unsigned char *masks,
                      unsigned long long *per_object, unsigned int nobjects, int sm_count, cudaStream_t s);
int c3d_k_clear_part(Pixel *canvas, const Partition *part, long long w, long long h, int sm_count, cudaStream_t s);
int c3d_k_to_float(const double *in, float *out, size_t n, cudaStream_t s);
int c3d_k_pack(const unsigned int *rgba, const unsigned char bg[3], long long w, long long h, unsigned char *out, cudaStream_t s);
int c3d_k_canvas_split(const Pixel *canvas, double *depth, unsigned long long *key, size_t npix, cudaStream_t s);
int c3d_k_case_codes(const FieldView *f, const DevTables *t, unsigned int *codes, unsigned char *counts, cudaStream_t s);
int c3d_k_emit(const FieldView *f, const DevTables *t, const unsigned long long *list, unsigned long long nlist,
               unsigned long long *cursor, unsigned long long capacity, void *records, cudaStream_t s);
int c3d_k_halo_pack_y(const void *data, int elem_bytes, long long nx, long long sy, long long sz, long long k0, long long nk,
                      long long j0, void *buf, cudaStream_t s);
int c3d_k_halo_unpack_y(void *data, int elem_bytes, long long nx, long long sy, long long sz, long long k0, long long nk,
                        long long j0, const void *buf, cudaStream_t s);
int c3d_k_synth(void *data, int elem_bytes, const long long n[3], long long sy, long long sz, const long long goff[3],
                const long long gl[3], int kind, unsigned int seed, int octaves, double gain, cudaStream_t s);

#if defined(__cplusplus)
}
#endif

#endif /* C3D_INTERNAL_H */
