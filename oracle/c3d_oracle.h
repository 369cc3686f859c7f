/*
 * c3d_oracle.h -- CPU restatement of Contour3D's per-frame pipeline.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product library
 * (contour3d_b200/) may include, link or call this.  It exists so that the
 * CUDA path can be checked on machines where the reference tree is absent
 * (the GPU boxes).  It is pinned against the unmodified reference
 * (oracle/_ref, built from /root/reference by oracle/Makefile): see
 * tests/test_oracle_vs_reference.py and tests/golden/.
 *
 * All arithmetic is fp64, evaluated in the reference's operation order and
 * compiled without FMA contraction (-std=c99), so results are expected to be
 * bit-identical to the reference's, not merely close.
 */
#if !defined(C3D_ORACLE_H)
#define C3D_ORACLE_H

#include <stddef.h>
#include <stdint.h>

typedef struct { double x, y, z; } oracle_vec_t;
typedef oracle_vec_t (*oracle_converter_t)(const oracle_vec_t orthogonal);

/* one contour object, described on the EXTENDED LOCAL array of one rank
   (for a single rank: the global array).  reference: contour3d_contour_obj_t,
   include/contour3d.h:22-40, after contour/extend.c has run */
typedef struct {
  size_t dims[3];             /* points in x, y, z of the (extended) array */
  const double *grid[3];      /* coordinates of exactly those points */
  const double *field;        /* x fastest: field[(k*dims[1]+j)*dims[0]+i] */
  double threshold;
  uint8_t color[3];
  oracle_converter_t converter; /* NULL = identity */
} oracle_contour_t;

/* reference: contour3d_line_obj_t, include/contour3d.h:42-56 */
typedef struct {
  double edges[2][3];
  size_t nitems;
  double width;
  uint8_t color[3];
  oracle_converter_t converter; /* NULL = identity */
} oracle_line_t;

/* reference: arguments of contour3d_execute, include/contour3d.h:58-86 */
typedef struct {
  double camera[3];
  double light[3];            /* not normalised (the oracle normalises, src/contour3d/main.c:24-28) */
  double center[3];
  double local_x[3];
  double local_y[3];
  size_t width, height;
  uint8_t bg[3];
} oracle_scene_t;

/* flat triangle record for dumps (208 bytes, no implicit padding) */
typedef struct {
  uint64_t cube;              /* (k*(dims[1]-1)+j)*(dims[0]-1)+i */
  uint32_t tri;               /* index within the cube, 0..11 */
  uint32_t has_vn;            /* 1 when vertex normals were computed (interior cube) */
  uint32_t edge[3];           /* cube-edge id 0..18 of each vertex */
  uint32_t pad;
  double area;                /* |cross| (twice the geometric area) */
  double v[9];                /* three vertices */
  double fn[3];               /* unit face normal */
  double vn[9];               /* three vertex normals */
} oracle_triangle_t;

typedef struct {
  uint64_t cubes;             /* cubes visited */
  uint64_t active_cubes;      /* cubes with at least one triangle */
  uint64_t triangles;         /* triangles generated (all cubes) */
  uint64_t rendered;          /* triangles handed to the rasteriser (interior cubes) */
} oracle_stats_t;

/* six pre-fold 4-bit tetrahedron masks per cube (24 bits, tetra n in bits
   4n..4n+3) and the triangle count per cube.  returns number of cubes */
size_t oracle_case_codes(const oracle_contour_t *c, uint32_t *codes, uint8_t *counts);

/* all triangles of all cubes in (k, j, i, tri) order, vertex normals where the
   reference computes them.  returns the total count; writes min(total, capacity) */
size_t oracle_tessellate(const oracle_contour_t *c, size_t capacity, oracle_triangle_t *out, oracle_stats_t *stats);

/* canvas = depth[W*H] + rgb[3*W*H] in canvas order (row j, column i; NOT y-flipped) */
int oracle_canvas_clear(const oracle_scene_t *s, double *depth, uint8_t *rgb);
int oracle_render_contour(const oracle_scene_t *s, const oracle_contour_t *c, double *depth, uint8_t *rgb, oracle_stats_t *stats);
int oracle_render_line(const oracle_scene_t *s, const oracle_line_t *l, double *depth, uint8_t *rgb);
/* "larger depth wins" merge of src into dst (reference: output.c:10-30) */
int oracle_canvas_merge(size_t npixels, double *dst_depth, uint8_t *dst_rgb, const double *src_depth, const uint8_t *src_rgb);
/* y-flipped binary P6 (reference: output.c:105-149) */
int oracle_write_ppm(const char *fname, size_t width, size_t height, const uint8_t *rgb);

/* whole frame on one "rank"; with nthreads > 1 the domain of every contour is
   split along y into nthreads windows that are rendered into private canvases
   and merged, exactly like the reference's ranks (decomposition invariant) */
int oracle_frame(const oracle_scene_t *s, size_t ncontours, const oracle_contour_t *contours,
                 size_t nlines, const oracle_line_t *lines, int nthreads,
                 double *depth, uint8_t *rgb, oracle_stats_t *stats);

/* the converters of the case-file driver (c3d_case_main.c), for tests: 0 identity (NULL), 1 cylindrical, 2 stretched, 3 twisted */
oracle_converter_t oracle_converter_of(int kind);
void oracle_convert_points(int kind, size_t n, const double *in, double *out);

#endif /* C3D_ORACLE_H */
