/*
 * c3d_tables.h -- marching-tetrahedra case tables of the product library.
 *
 * The reference decides per tetrahedron at run time (six 4-bit masks, fold,
 * 7-way branch; contour/triangulate.c:145-276).  Here everything that depends
 * only on the eight corner bits of a cube is expanded ONCE on the host into
 * 256-entry tables that the kernels index with the 8-bit corner code:
 *
 *   count[code]          number of triangles (0..12)
 *   code24[code]         the six pre-fold tetra masks, tetra n in bits 4n..4n+3
 *   pairs[code][t]       triangle t: for vertex m the (tail, head) cube corners,
 *                        3+3 bits each at bit 6m (winding reversal applied)
 *   emask[code][t]       bit e set when a vertex of triangle t sits on cube edge e (0..18)
 *   edges[code][t]       the three cube-edge ids, 5 bits each at bit 5m
 *   touch[code][e]       bit t set when triangle t has a vertex on cube edge e (emask transposed)
 *
 * Corner numbering: bit0 = +x, bit1 = +y, bit2 = +z (triangulate.c:15-19).
 */
#if !defined(C3D_TABLES_H)
#define C3D_TABLES_H

#include <stdint.h>

#define C3D_MAX_TRIS 12

typedef struct {
  uint8_t count[256];
  uint32_t code24[256];
  uint32_t pairs[256 * C3D_MAX_TRIS];
  uint32_t emask[256 * C3D_MAX_TRIS];
  uint32_t edges[256 * C3D_MAX_TRIS];
  uint16_t touch[256 * 19];
} c3d_tables_t;

/* cube-edge id between corners a and b: 0-3 x edges (dj + 2dk), 4-7 y edges
   (di + 2dk), 8-11 z edges (di + 2dj), 12/13 yz-face diagonals (di), 14/15 xz
   (dj), 16/17 xy (dk), 18 body diagonal; equals the reference's 8x8 lookup
   (triangulate.c:23-32) on every pair the six tetrahedra can produce */
static inline uint32_t c3d_cube_edge(const uint32_t a, const uint32_t b) {
  const uint32_t lo = a & b;
  const uint32_t di = lo & 1u, dj = (lo >> 1) & 1u, dk = (lo >> 2) & 1u;
  switch (a ^ b) {
    case 1u: return dj + 2u * dk;
    case 2u: return 4u + di + 2u * dk;
    case 4u: return 8u + di + 2u * dj;
    case 6u: return 12u + di;
    case 5u: return 14u + dj;
    case 3u: return 16u + dk;
    default: return 18u;
  }
}

static inline void c3d_tables_build(c3d_tables_t *t) {
  /* tetrahedra around the 0-7 diagonal (triangulate.c:145-152) */
  static const uint8_t tetra[6][4] = {{0, 7, 3, 1}, {0, 7, 1, 5}, {0, 7, 5, 4}, {0, 7, 4, 6}, {0, 7, 6, 2}, {0, 7, 2, 3}};
  /* folded mask -> triangles as tetra-vertex (tail, head) pairs (triangulate.c:204-270) */
  static const uint8_t ntri[8] = {0, 1, 1, 2, 1, 2, 2, 1};
  static const uint8_t tails[8][2][3] = {
    {{0, 0, 0}, {0, 0, 0}}, {{0, 0, 0}, {0, 0, 0}}, {{1, 1, 1}, {0, 0, 0}}, {{0, 1, 0}, {0, 1, 1}},
    {{2, 2, 2}, {0, 0, 0}}, {{2, 0, 2}, {2, 0, 0}}, {{0, 2, 1}, {0, 1, 0}}, {{3, 3, 3}, {0, 0, 0}},
  };
  static const uint8_t heads[8][2][3] = {
    {{0, 0, 0}, {0, 0, 0}}, {{1, 3, 2}, {0, 0, 0}}, {{0, 2, 3}, {0, 0, 0}}, {{2, 3, 3}, {2, 2, 3}},
    {{0, 3, 1}, {0, 0, 0}}, {{1, 3, 3}, {1, 1, 3}}, {{2, 3, 3}, {2, 3, 1}}, {{0, 2, 1}, {0, 0, 0}},
  };
  for (uint32_t code = 0; code < 256; code++) {
    uint32_t count = 0, packed = 0;
    for (int n = 0; n < 6; n++) {
      uint32_t mask = 0;
      for (int m = 0; m < 4; m++) mask |= ((code >> tetra[n][m]) & 1u) << m;
      packed |= mask << (4 * n);
      const int reverse = mask >= 8u;               /* triangulate.c:189-192 */
      if (reverse) mask = 15u - mask;
      for (int s = 0; s < ntri[mask]; s++) {
        uint32_t pairs = 0, emask = 0, edges = 0;
        for (int m = 0; m < 3; m++) {
          const int pick = reverse ? 2 - m : m;     /* triangulate.c:108-112 */
          const uint32_t tail = tetra[n][tails[mask][s][pick]];
          const uint32_t head = tetra[n][heads[mask][s][pick]];
          const uint32_t e = c3d_cube_edge(tail, head);
          pairs |= (tail | (head << 3)) << (6 * m);
          emask |= 1u << e;
          edges |= e << (5 * m);
        }
        t->pairs[code * C3D_MAX_TRIS + count] = pairs;
        t->emask[code * C3D_MAX_TRIS + count] = emask;
        t->edges[code * C3D_MAX_TRIS + count] = edges;
        count++;
      }
    }
    for (uint32_t s = count; s < C3D_MAX_TRIS; s++) {
      t->pairs[code * C3D_MAX_TRIS + s] = 0;
      t->emask[code * C3D_MAX_TRIS + s] = 0;
      t->edges[code * C3D_MAX_TRIS + s] = 0;
    }
    t->count[code] = (uint8_t)count;
    t->code24[code] = packed;
    for (uint32_t e = 0; e < 19; e++) {
      uint16_t mask = 0;
      for (uint32_t s = 0; s < count; s++) mask |= (uint16_t)(((t->emask[code * C3D_MAX_TRIS + s] >> e) & 1u) << s);
      t->touch[code * 19 + e] = mask;
    }
  }
}

#endif /* C3D_TABLES_H */
