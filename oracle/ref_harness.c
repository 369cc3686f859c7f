/*
 * ref_harness.c -- stage-level access to the UNMODIFIED reference.
 * TEST INFRASTRUCTURE ONLY; compiled by oracle/Makefile into
 * oracle/_ref/libc3dref.so together with the reference's own sources.
 *
 * It calls the reference's extern stage functions
 * (src/contour3d/contour/internal.h:17-48) exactly the way the reference's
 * per-object driver does (contour/main.c:37-113: three-slice ring, normals of
 * the middle slice once three slices exist) and copies every lattice_t into
 * the flat 208-byte record that the oracle restatement and the CUDA library
 * also produce, so the three can be compared field by field.
 */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "contour3d.h"
#include "struct.h"
#include "contour/internal.h"

typedef struct {
  uint64_t cube;
  uint32_t tri, has_vn, edge[3], pad;
  double area, v[9], fn[3], vn[9];
} record_t;

static contour3d_vector_t identity(const contour3d_vector_t orthogonal) { return orthogonal; }

static void flush(const lattice_t *slice, const size_t cx, const size_t cy, const size_t k, const int with_normals,
                  const size_t capacity, record_t *out, size_t *count) {
  for (size_t j = 0; j < cy; j++) {
    for (size_t i = 0; i < cx; i++) {
      const lattice_t *lat = slice + j * cx + i;
      const int inner = with_normals && i >= 1 && i + 1 < cx && j >= 1 && j + 1 < cy;
      for (size_t t = 0; t < lat->num_triangles; t++) {
        if (*count < capacity) {
          const triangle_t *tri = lat->triangles + t;
          record_t *r = out + *count;
          memset(r, 0, sizeof(*r));
          r->cube = (uint64_t)((k * cy + j) * cx + i);
          r->tri = (uint32_t)t;
          r->has_vn = (uint32_t)inner;
          r->area = tri->area;
          for (int m = 0; m < 3; m++) {
            r->edge[m] = (uint32_t)tri->cube_indices[m];
            r->v[3 * m + 0] = tri->vertices[m].x; r->v[3 * m + 1] = tri->vertices[m].y; r->v[3 * m + 2] = tri->vertices[m].z;
            if (inner) {
              r->vn[3 * m + 0] = tri->vertex_normals[m].x; r->vn[3 * m + 1] = tri->vertex_normals[m].y; r->vn[3 * m + 2] = tri->vertex_normals[m].z;
            }
          }
          r->fn[0] = tri->face_normal.x; r->fn[1] = tri->face_normal.y; r->fn[2] = tri->face_normal.z;
        }
        (*count)++;
      }
    }
  }
}

/* all triangles of a (dims[0] x dims[1] x dims[2], x fastest) array in (k, j, i, t) order */
size_t ref_tessellate(const size_t dims[3], double *const grids[3], const double *field, const double threshold,
                      const size_t capacity, record_t *out) {
  const size_t cx = dims[0] - 1, cy = dims[1] - 1, cz = dims[2] - 1;
  lattice_t *slices[3];
  for (int n = 0; n < 3; n++) slices[n] = calloc(cx * cy, sizeof(lattice_t));
  size_t count = 0;
  for (size_t k = 0; k < cz; k++) {
    contour3d_contour_triangulate_slice(
        (size_t[2]){dims[0], dims[1]},
        (double *const[3]){grids[0], grids[1], grids[2] + k},
        identity, field + k * dims[0] * dims[1], threshold, slices[k % 3]);
    if (0 == k) flush(slices[0], cx, cy, 0, 0, capacity, out, &count);
    if (k < 2) continue;
    contour3d_contour_compute_vertex_normals(
        (size_t[2]){dims[0], dims[1]},
        (lattice_t *[3]){slices[(k - 2) % 3], slices[(k - 1) % 3], slices[k % 3]});
    flush(slices[(k - 1) % 3], cx, cy, k - 1, 1, capacity, out, &count);
  }
  if (cz > 1) flush(slices[(cz - 1) % 3], cx, cy, cz - 1, 0, capacity, out, &count);
  for (int n = 0; n < 3; n++) free(slices[n]);
  return count;
}
