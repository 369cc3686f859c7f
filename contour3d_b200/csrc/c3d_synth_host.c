/*
 * c3d_synth_host.c -- the benchmark's synthetic fields on the HOST, bit for bit the values the device generator
 * (k_synth, c3d_kernels.cu) produces: multi-octave value noise from an integer hash, and the magnitude of the curl of
 * a three-component noise field ("vorticity magnitude", config 4 of BASELINE.json).
 *
 * Purpose: the reference arm of bench.py (the unmodified CPU reference) must see the same field as the GPU arm without
 * touching the CUDA library.  A function of the GLOBAL grid index, so any window of the field can be produced on its own.
 * Plain C, single precision, no fused multiply-adds (build with -ffp-contract=off; the CUDA side is built with
 * -fmad=false), same operation order as the kernel.  The gyroid (kind 0) is NOT here: sinf / cosf differ between
 * libm and CUDA in the last bit; configurations that use it upload a numpy array (scenes.gyroid) on both sides.
 *
 * Built into its own small library (libc3d_synth_host.so); not part of libcontour3d_b200.so.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include "c3d_synth.h"

static inline uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

static float value_noise(const float x, const float y, const float z, const uint32_t seed) {
  const float fx = floorf(x), fy = floorf(y), fz = floorf(z);
  const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
  float tx = x - fx, ty = y - fy, tz = z - fz;
  tx = tx * tx * (3.f - 2.f * tx); ty = ty * ty * (3.f - 2.f * ty); tz = tz * tz * (3.f - 2.f * tz);
  /* the nested hash of lattice_value() with its inner levels shared between corners (integer arithmetic: same values) */
  uint32_t hz[2], hyz[4];
  for (int dz = 0; dz < 2; dz++) hz[dz] = lowbias32((uint32_t)(iz + dz) * 0xC2B2AE3Du ^ seed);
  for (int c = 0; c < 4; c++) hyz[c] = lowbias32((uint32_t)(iy + (c & 1)) * 0x85EBCA77u ^ hz[c >> 1]);
  float acc = 0.f;
  for (int c = 0; c < 8; c++) {
    const int dx = c & 1, dy = (c >> 1) & 1, dz = c >> 2;
    const float w = (dx ? tx : 1.f - tx) * (dy ? ty : 1.f - ty) * (dz ? tz : 1.f - tz);
    const uint32_t h = lowbias32((uint32_t)(ix + dx) * 0x9E3779B1u ^ hyz[dy + 2 * dz]);
    acc += w * ((float)(h >> 8) * (1.0f / 16777216.0f));
  }
  return acc;
}

/* sum_o gain^o (2 valuenoise(2^(o+2) (p + 0.5), seed + o) - 1) / sum_o gain^o, in [-1, 1] */
static float fractal(const float x, const float y, const float z, const uint32_t seed, const int octaves, const float gain) {
  float amp = 1.f, sum = 0.f, norm = 0.f, freq = 4.f;
  for (int o = 0; o < octaves; o++) {
    sum += amp * (2.f * value_noise((x + 0.5f) * freq, (y + 0.5f) * freq, (z + 0.5f) * freq, seed + (uint32_t)o) - 1.f);
    norm += amp;
    amp *= gain;
    freq *= 2.f;
  }
  return sum / norm;
}

/* |curl F| of F = (fractal(seed), fractal(seed + 101), fractal(seed + 202)) by central differences of half a cell,
   scaled by `scale` and shifted to be roughly centred on zero: scale * |curl F| - 1 */
static float curl_magnitude(const float x, const float y, const float z, const float h, const uint32_t seed, const int octaves, const float gain,
                            const float scale) {
  const float inv = 1.f / (2.f * h);
  const float dFz_dy = (fractal(x, y + h, z, seed + 202u, octaves, gain) - fractal(x, y - h, z, seed + 202u, octaves, gain)) * inv;
  const float dFy_dz = (fractal(x, y, z + h, seed + 101u, octaves, gain) - fractal(x, y, z - h, seed + 101u, octaves, gain)) * inv;
  const float dFx_dz = (fractal(x, y, z + h, seed, octaves, gain) - fractal(x, y, z - h, seed, octaves, gain)) * inv;
  const float dFz_dx = (fractal(x + h, y, z, seed + 202u, octaves, gain) - fractal(x - h, y, z, seed + 202u, octaves, gain)) * inv;
  const float dFy_dx = (fractal(x + h, y, z, seed + 101u, octaves, gain) - fractal(x - h, y, z, seed + 101u, octaves, gain)) * inv;
  const float dFx_dy = (fractal(x, y + h, z, seed, octaves, gain) - fractal(x, y - h, z, seed, octaves, gain)) * inv;
  const float cx = dFz_dy - dFy_dz, cy = dFx_dz - dFz_dx, cz = dFy_dx - dFx_dy;
  return scale * sqrtf(cx * cx + cy * cy + cz * cz) - 1.f;
}

typedef struct {
  float *out;
  long long n[3], off[3], gl[3];
  int kind, octaves;
  uint32_t seed;
  float gain;
  long long k0, k1;
} job_t;

static void *worker(void *arg) {
  const job_t *j = arg;
  for (long long k = j->k0; k < j->k1; k++) {
    const float z = -0.5f + ((float)(k + j->off[2]) + 0.5f) / (float)j->gl[2];
    for (long long jj = 0; jj < j->n[1]; jj++) {
      const float y = -0.5f + ((float)(jj + j->off[1]) + 0.5f) / (float)j->gl[1];
      float *row = j->out + (k * j->n[1] + jj) * j->n[0];
      for (long long i = 0; i < j->n[0]; i++) {
        const float x = -0.5f + ((float)(i + j->off[0]) + 0.5f) / (float)j->gl[0];
        row[i] = 1 == j->kind ? fractal(x, y, z, j->seed, j->octaves, j->gain)
                              : curl_magnitude(x, y, z, 0.5f / (float)j->gl[0], j->seed, j->octaves, j->gain, C3D_CURL_SCALE);
      }
    }
  }
  return NULL;
}

/* window n[3] (x, y, z points) at global offset off[3] of the gl[3] field, x fastest; kind 1 noise, 2 curl magnitude.
   Returns 0, or 1 on a bad argument. */
int c3d_synth_host(float *out, const long long n[3], const long long off[3], const long long gl[3], int kind, uint32_t seed, int octaves,
                   double gain, int nthreads) {
  if (NULL == out || (1 != kind && 2 != kind) || nthreads < 1) return 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t tid[256];
  job_t jobs[256];
  int started = 0;
  for (int t = 0; t < nthreads; t++) {
    job_t *j = &jobs[t];
    j->out = out;
    for (int d = 0; d < 3; d++) { j->n[d] = n[d]; j->off[d] = off[d]; j->gl[d] = gl[d]; }
    j->kind = kind; j->octaves = octaves; j->seed = seed; j->gain = (float)gain;
    j->k0 = n[2] * t / nthreads;
    j->k1 = n[2] * (t + 1) / nthreads;
    if (0 != pthread_create(&tid[t], NULL, worker, j)) { worker(j); tid[t] = 0; continue; }
    started |= 1;
  }
  for (int t = 0; t < nthreads; t++) {
    if (0 != tid[t]) pthread_join(tid[t], NULL);
  }
  (void)started;
  return 0;
}
