/*
 * c3d_positions.c -- host-side tabulation of a coordinate converter.
 *
 * The reference calls the user's `converter` (include/contour3d.h:31-33) on
 * every cube corner inside its innermost loop (contour/triangulate.c:156-173:
 * 24 calls per cube).  A device kernel cannot call a host function, so the
 * drop-in evaluates the SAME function on the host, once per grid point at
 * most, and hands the kernels tables (c3d_positions_t, contour3d_b200.h).  The
 * values are the callback's own doubles: whatever it computes (sin, cos,
 * table look-ups, ...), the device geometry starts from bit-identical corner
 * positions.
 *
 * A converter rarely depends on all three directions in every component
 * (identity: 1 + 1 + 1; cylindrical (r, theta, z): 2 + 2 + 1), so the
 * dependence of each output component on each direction is probed first and a
 * component gets a table only over the directions it depends on.  The probe
 * compares bit patterns along sample lines through the block; the tables are
 * then verified on a scattered sample of full (i, j, k) points, and a
 * component that fails is promoted to a full 3-D table.
 *
 * Plain C99, no device needed (tests/test_host_logic.py runs it on CPU).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "contour3d_b200.h"

#define N_BASES 6
#define LINE_SAMPLES 48
#define VERIFY_SAMPLES 4096

static int same_bits(const double a, const double b) { return 0 == memcmp(&a, &b, sizeof(double)); }

static uint64_t next_random(uint64_t *state) {
  *state = *state * 6364136223846793005ull + 1442695040888963407ull;
  return *state >> 33;
}

static void evaluate(c3d_converter_fn convert, void *user, const double *const g[3], const size_t idx[3], double out[3]) {
  const double in[3] = {g[0][idx[0]], g[1][idx[1]], g[2][idx[2]]};
  convert(in, out, user);
}

void c3d_positions_free(c3d_positions_t *p) {
  if (NULL == p) return;
  for (int c = 0; c < 3; c++) {
    free((void *)p->table[c]);
    p->table[c] = NULL;
  }
}

static int fail(const char *msg) {
  fprintf(stderr, "[CONTOUR3D ERROR] %s\n", msg);
  fflush(stderr);
  return 1;
}

/* table of component c over the directions dep[c][.] marks, x fastest */
static int fill_table(c3d_converter_fn convert, void *user, const double *const g[3], const size_t n[3], const int dep[3],
                      const int c, size_t stride[3], double **table, size_t *bytes_left) {
  size_t len = 1;
  for (int a = 0; a < 3; a++) {
    stride[a] = dep[a] ? len : 0;
    if (dep[a]) len *= n[a];
  }
  if (len > *bytes_left / sizeof(double)) return fail("the coordinate converter needs tables larger than CONTOUR3D_CONVERTER_MAX_BYTES allows");
  double *t = malloc(len * sizeof(double));
  if (NULL == t) return fail("out of host memory for the converter tables");
  *bytes_left -= len * sizeof(double);
  const size_t m[3] = {dep[0] ? n[0] : 1, dep[1] ? n[1] : 1, dep[2] ? n[2] : 1};
  size_t idx[3];
  for (idx[2] = 0; idx[2] < m[2]; idx[2]++) {
    for (idx[1] = 0; idx[1] < m[1]; idx[1]++) {
      for (idx[0] = 0; idx[0] < m[0]; idx[0]++) {
        double out[3];
        evaluate(convert, user, g, idx, out);
        t[idx[0] * stride[0] + idx[1] * stride[1] + idx[2] * stride[2]] = out[c];
      }
    }
  }
  *table = t;
  return 0;
}

int c3d_positions_probe(c3d_converter_fn convert, void *user, const double *const grids[3], const size_t offset[3],
                        const size_t count[3], size_t max_bytes, c3d_positions_t *out, int *identity) {
  if (NULL == convert || NULL == grids || NULL == offset || NULL == count || NULL == out || NULL == identity) {
    return fail("c3d_positions_probe: NULL argument");
  }
  memset(out, 0, sizeof(*out));
  *identity = 1;
  const double *g[3];
  size_t n[3];
  for (int d = 0; d < 3; d++) {
    if (NULL == grids[d] || 0 == count[d]) return fail("c3d_positions_probe: empty grid");
    g[d] = grids[d] + offset[d];
    n[d] = count[d];
  }
  /* 1. dependence of every component on every direction, along lines through a few base points */
  int dep[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  uint64_t rng = 0x9E3779B97F4A7C15ull ^ (uint64_t)(n[0] * 73856093u) ^ (uint64_t)(n[1] * 19349663u) ^ (uint64_t)(n[2] * 83492791u);
  for (int b = 0; b < N_BASES; b++) {
    size_t base[3];
    for (int d = 0; d < 3; d++) {
      base[d] = 0 == b ? 0 : (1 == b ? n[d] - 1 : (2 == b ? n[d] / 2 : (size_t)(next_random(&rng) % n[d])));
    }
    double ref[3];
    evaluate(convert, user, g, base, ref);
    {
      const double in[3] = {g[0][base[0]], g[1][base[1]], g[2][base[2]]};
      for (int c = 0; c < 3; c++) if (!same_bits(in[c], ref[c])) *identity = 0;
    }
    for (int a = 0; a < 3; a++) {
      const size_t samples = n[a] < LINE_SAMPLES ? n[a] : LINE_SAMPLES;
      for (size_t q = 0; q < samples; q++) {
        size_t idx[3] = {base[0], base[1], base[2]};
        idx[a] = samples > 1 ? q * (n[a] - 1) / (samples - 1) : 0;
        double val[3];
        evaluate(convert, user, g, idx, val);
        for (int c = 0; c < 3; c++) {
          if (!same_bits(val[c], ref[c])) dep[c][a] = 1;
          if (!same_bits(val[c], g[c][idx[c]])) *identity = 0;
        }
      }
    }
  }
  /* The lines above see the whole range of every direction but only a few base points.  A function that is the identity
     (or lower-dimensional) except in some region away from those lines would be taken for what it is not: the identity
     verdict is therefore checked on scattered points as well -- at least one in every z plane -- and a component caught
     differing is tabulated in full.  This is sampling, not proof; CONTOUR3D_CONVERTER_FULL_TABLES=1 skips all guessing
     and tabulates every component over all three directions (24 bytes per grid point). */
  {
    const char *env = getenv("CONTOUR3D_CONVERTER_FULL_TABLES");
    if (NULL != env && '\0' != *env && '0' != *env) {
      *identity = 0;
      for (int c = 0; c < 3; c++) dep[c][0] = dep[c][1] = dep[c][2] = 1;
    }
  }
  if (*identity) {
    uint64_t vr = rng ^ 0x5DEECE66Dull;
    const size_t nscatter = (size_t)VERIFY_SAMPLES + n[2];
    for (size_t q = 0; q < nscatter; q++) {
      const size_t idx[3] = {(size_t)(next_random(&vr) % n[0]), (size_t)(next_random(&vr) % n[1]), q < n[2] ? q : (size_t)(next_random(&vr) % n[2])};
      double val[3];
      evaluate(convert, user, g, idx, val);
      for (int c = 0; c < 3; c++) {
        if (!same_bits(val[c], g[c][idx[c]])) {
          *identity = 0;
          dep[c][0] = dep[c][1] = dep[c][2] = 1;
        }
      }
    }
  }
  if (*identity) return 0;
  /* 2. tables over the dependent directions, 3. verification on scattered points; a component that fails becomes 3-D */
  size_t bytes_left = max_bytes;
  for (int c = 0; c < 3; c++) {
    for (int attempt = 0; attempt < 2; attempt++) {
      double *table = NULL;
      if (0 != fill_table(convert, user, g, n, dep[c], c, out->stride[c], &table, &bytes_left)) {
        c3d_positions_free(out);
        return 1;
      }
      int ok = 1;
      uint64_t vr = rng + (uint64_t)c;
      for (int q = 0; q < VERIFY_SAMPLES && ok; q++) {
        const size_t idx[3] = {(size_t)(next_random(&vr) % n[0]), (size_t)(next_random(&vr) % n[1]), (size_t)(next_random(&vr) % n[2])};
        double val[3];
        evaluate(convert, user, g, idx, val);
        ok = same_bits(val[c], table[idx[0] * out->stride[c][0] + idx[1] * out->stride[c][1] + idx[2] * out->stride[c][2]]);
      }
      if (ok || (dep[c][0] && dep[c][1] && dep[c][2])) {
        if (!ok) {
          free(table);
          c3d_positions_free(out);
          return fail("the coordinate converter does not return the same value for the same argument");
        }
        out->table[c] = table;
        break;
      }
      size_t len = 1;
      for (int a = 0; a < 3; a++) if (dep[c][a]) len *= n[a];
      bytes_left += len * sizeof(double);
      free(table);
      dep[c][0] = dep[c][1] = dep[c][2] = 1;
    }
  }
  return 0;
}
