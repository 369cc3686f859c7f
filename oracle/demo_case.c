/*
 * demo_case.c -- writes the reference's built-in demo (config 1 of
 * BASELINE.json) as a case file, so that the unmodified reference
 * (ref_driver), the oracle restatement and the CUDA path all consume
 * IDENTICAL input bits.  TEST INFRASTRUCTURE ONLY.
 *
 * It restates the field/scene construction of /root/reference/src/main.c
 * (init_field 23-111: Chebyshev x grid, uniform y/z grids, 64 sine modes with
 * libc rand() phases, min/max normalisation; scene 126-306) for a single rank.
 * It is pinned by tests/test_oracle_vs_reference.py: the image the reference
 * renders from this file has the same md5 as the reference's own ./a.out.
 *
 *   demo_case <out.bin> [n=128] [width=1440] [height=1080]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#include <float.h>

typedef struct {
  char magic[8];
  uint64_t n[3], width, height, ncontours, nlines, elem_bytes, nfields;
  double camera[3], look_at[3], light[3], center[3], local_x[3], local_y[3];
  uint8_t bg[3], pad[5];
} case_header_t;
typedef struct { uint64_t field; double threshold; uint8_t color[3], pad[5]; } case_contour_t;
typedef struct { double edges[6]; uint64_t nitems; double width; uint8_t color[3], pad[5]; } case_line_t;

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s out.bin [n] [width] [height]\n", argv[0]); return 2; }
  const size_t n = argc > 2 ? (size_t)atol(argv[2]) : 128;
  const size_t width = argc > 3 ? (size_t)atol(argv[3]) : 1440;
  const size_t height = argc > 4 ? (size_t)atol(argv[4]) : 1080;
  const double pi = 3.1415926535897932;
  const double len = 1.;
  double *gx = malloc(n * sizeof(double)), *gy = malloc(n * sizeof(double)), *gz = malloc(n * sizeof(double));
  for (size_t m = 0; m < n; m++) {
    gx[m] = -0.5 * len + len * 0.5 * (1. + cos(pi - pi * (m + 0.5) / n));
    gy[m] = -0.5 * len + len / n * 0.5 * (2 * m + 1);
    gz[m] = -0.5 * len + len / n * 0.5 * (2 * m + 1);
  }
  double *f = calloc(n * n * n, sizeof(double));
  for (size_t kz = 1; kz < 5; kz++) {
    for (size_t ky = 1; ky < 5; ky++) {
      for (size_t kx = 1; kx < 5; kx++) {
        const double xphase = 0.;
        const double yphase = 2. * pi * rand() / RAND_MAX;
        const double zphase = 2. * pi * rand() / RAND_MAX;
        for (size_t k = 0; k < n; k++) {
          const double z = gz[k];
          for (size_t j = 0; j < n; j++) {
            const double y = gy[j];
            for (size_t i = 0; i < n; i++) {
              const double x = gx[i];
              f[(k * n + j) * n + i] += 1.
                * sin(2. * pi / len * kx * x + xphase)
                * sin(2. * pi / len * ky * y + yphase)
                * sin(2. * pi / len * kz * z + zphase);
            }
          }
        }
      }
    }
  }
  double lo = +1. * DBL_MAX, hi = -1. * DBL_MAX;
  for (size_t m = 0; m < n * n * n; m++) { lo = fmin(lo, f[m]); hi = fmax(hi, f[m]); }
  for (size_t m = 0; m < n * n * n; m++) f[m] = (f[m] - lo) / (hi - lo) * 2. - 1.;
  case_header_t h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, "C3DCASE1", 8);
  h.n[0] = h.n[1] = h.n[2] = n;
  h.width = width; h.height = height;
  h.ncontours = 2; h.nlines = 8; h.elem_bytes = 8; h.nfields = 1;
  const double camera[3] = {+8.660254037844386e-1, -1.5, +1.0};
  const double center[3] = {+4.330127018922194e-1, -7.500000000000003e-1, +5.000000000000003e-1};
  const double local_x[3] = {+1.154700538379251e+0, +6.666666666666665e-1, 0.};
  const double local_y[3] = {-2.5e-1, +4.330127018922195e-1, +8.660254037844386e-1};
  const double light[3] = {0., 2., -1.};
  memcpy(h.camera, camera, sizeof(camera));
  memcpy(h.center, center, sizeof(center));
  memcpy(h.local_x, local_x, sizeof(local_x));
  memcpy(h.local_y, local_y, sizeof(local_y));
  memcpy(h.light, light, sizeof(light));
  case_contour_t cs[2];
  memset(cs, 0, sizeof(cs));
  cs[0].threshold = -0.25; cs[0].color[0] = 0x00; cs[0].color[1] = 0xFF; cs[0].color[2] = 0xFF;
  cs[1].threshold = +0.25; cs[1].color[0] = 0xFF; cs[1].color[1] = 0xFF; cs[1].color[2] = 0x00;
  case_line_t ls[8];
  memset(ls, 0, sizeof(ls));
  /* four edges along y, then four along z (src/main.c:223-306) */
  const double corners[4][2] = {{-0.5, -0.5}, {+0.5, -0.5}, {-0.5, +0.5}, {+0.5, +0.5}};
  for (int m = 0; m < 4; m++) {
    /* y lines: (x, z) corners in the order (-,-) (+,-) (-,+) (+,+) */
    const double e[6] = {corners[m][0], -0.5, corners[m][1], corners[m][0], +0.5, corners[m][1]};
    memcpy(ls[m].edges, e, sizeof(e));
  }
  const double zc[4][2] = {{-0.5, -0.5}, {+0.5, -0.5}, {-0.5, +0.5}, {+0.5, +0.5}};
  for (int m = 0; m < 4; m++) {
    /* z lines: (x, y) corners */
    const double e[6] = {zc[m][0], zc[m][1], -0.5, zc[m][0], zc[m][1], +0.5};
    memcpy(ls[4 + m].edges, e, sizeof(e));
  }
  for (int m = 0; m < 8; m++) {
    ls[m].nitems = n;
    ls[m].width = 4.;
    ls[m].color[0] = ls[m].color[1] = ls[m].color[2] = 0xFF;
  }
  FILE *fp = fopen(argv[1], "wb");
  if (NULL == fp) { perror(argv[1]); return 1; }
  fwrite(&h, sizeof(h), 1, fp);
  fwrite(gx, sizeof(double), n, fp);
  fwrite(gy, sizeof(double), n, fp);
  fwrite(gz, sizeof(double), n, fp);
  fwrite(cs, sizeof(cs[0]), 2, fp);
  fwrite(ls, sizeof(ls[0]), 8, fp);
  fwrite(f, sizeof(double), n * n * n, fp);
  fclose(fp);
  return 0;
}
