/*
 * c3d_sweep.cu -- the HBM-bound kernel of the pipeline: one streaming pass
 * over the scalar field that classifies every grid point against up to four
 * iso-values at once and writes a compact description of the cut cubes.
 *
 * Replaces the classification half of contour3d_contour_triangulate_slice
 * (contour/triangulate.c:287-342: gather 8 corners, `threshold < value`,
 * 176-179) for ALL cubes of the extended array; nothing else of the
 * tessellation runs on the >90 % of cubes that the surface does not cut.
 *
 * Design (sm_100a):
 *  - every warp owns tiles of 124 x 31 cubes (128 x 32 points) and marches
 *    through `zc` planes; tiles are handed out by an atomic cursor;
 *  - rows travel HBM -> shared memory as 1-D bulk async copies
 *    (cp.async.bulk ... mbarrier::complete_tx::bytes, SASS UBLKCP) into a
 *    warp-private ring of SW_SLOTS slots of 8 rows, so SW_SLOTS x 4 KiB per
 *    warp are always in flight and no register is tied up by a load;
 *  - a lane reads its 4 points of a row with one LDS.128 and folds
 *    `thr < v` of the 4 points into bit 4r+c of a 32-bit word per iso-value
 *    (8 rows x 4 points; FSET + LOP3 each), so one word describes 32 points;
 *  - "is the cube cut?" is evaluated for 32 cubes per lane at once with ~12
 *    word-wide logic instructions (OR/AND over the 8 corners, built from a
 *    1-lane shuffle for x+1, a 4-bit funnel shift for y+1 and the previous
 *    plane for z+1);
 *  - lanes that own cut cubes append one 32-byte record (the corner bits of
 *    their 4x8 patch) to a warp-private staging buffer that is flushed with
 *    one atomicAdd; k_expand turns records into per-cube list entries.
 *
 * Algorithmic bytes: sizeof(T) per point of the extended array per launch,
 * whatever the number of iso-values fused into the launch.
 *
 * File:line references are to NaokiHori/Contour3D @ c66854e.
 */
#include <float.h>
#include <stdint.h>
#include "c3d_internal.h"

#define SW_SLOTS 3          /* ring depth per warp */
#define SW_ROWS 8           /* rows per slot = rows per 32-bit word */
#define SW_GROUPS 4         /* row groups per tile: 32 point rows, 31 cube rows */
#define SW_XCUBES 124       /* cubes per tile row (lane 31 only supplies the x+1 neighbour) */
#define SW_YCUBES 31
#define SW_STAGE 32         /* records staged per warp before a flush */
#define SW_FULL 0xffffffffu

struct SweepRecord {        /* 32 bytes: the corner bits of a 4 (x) x 8 (y) patch of cubes, two planes */
  unsigned int pl, ql;      /* lower plane: points x..x+3 / x+1..x+4, bit 4r+c */
  unsigned int pu, qu;      /* upper plane */
  unsigned int active;      /* bit 4r+c: cube (x+c, j+r, k) is cut and inside the requested range */
  unsigned int row8;        /* row j+8 (needed by r = 7): pl, ql, pu, qu nibbles at bits 0, 4, 8, 12 */
  unsigned int xj;          /* x | j << 16 */
  unsigned int kobj;        /* k | object << 16 */
};

struct TileSweepParams {
  const void *data;
  long long mx, my, mz, sy, sz;
  int nthr;
  float thr_ge[C3D_SWEEP_MAX_THR];   /* fp32 fields: thr < (double)v  <=>  v >= thr_ge */
  double thr[C3D_SWEEP_MAX_THR];     /* fp64 fields: thr < v */
  unsigned int obj[C3D_SWEEP_MAX_THR];
  long long ilo, ihi, jlo, jhi, klo, khi;   /* inclusive cube ranges to report */
  int nxt, nyt, nzc, zc;
  unsigned int ntiles;
  unsigned int *tile_cursor;
  SweepRecord *records;
  unsigned long long *count, capacity;
};

/* ---- PTX helpers ---- */

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(const uint32_t bar, const uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(const uint32_t bar, const uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(const uint32_t bar, const uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity)
      : "memory");
}
/* 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (TMA engine, no tensor map needed) */
__device__ __forceinline__ void bulk_g2s(const uint32_t dst, const void *src, const uint32_t bytes, const uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

/* `threshold < value` (triangulate.c:176-179): fp32 fields compare against the smallest float above the
   threshold, which is the same predicate exactly; NaN is "not above" in both forms */
__device__ __forceinline__ bool above(const float v, const float thr_ge) { return v >= thr_ge; }
__device__ __forceinline__ bool above(const double v, const double thr) { return thr < v; }

template <typename T> struct Thr;
template <> struct Thr<float> {
  __device__ static __forceinline__ float get(const TileSweepParams &p, const int t) { return p.thr_ge[t]; }
};
template <> struct Thr<double> {
  __device__ static __forceinline__ double get(const TileSweepParams &p, const int t) { return p.thr[t]; }
};

template <typename T> struct Quad4 { T a, b, c, d; };
__device__ __forceinline__ Quad4<float> lds_quad(const float *row, const unsigned int lane) {
  const float4 v = *reinterpret_cast<const float4 *>(row + 4u * lane);
  return Quad4<float>{v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ Quad4<double> lds_quad(const double *row, const unsigned int lane) {
  const double2 lo = *reinterpret_cast<const double2 *>(row + 4u * lane);
  const double2 hi = *reinterpret_cast<const double2 *>(row + 4u * lane + 2u);
  return Quad4<double>{lo.x, lo.y, hi.x, hi.y};
}

/*
 * One warp = one tile at a time.  Shared memory per warp:
 *   ring   SW_SLOTS x SW_ROWS rows x 128 points
 *   stage  SW_STAGE records
 *   bars   SW_SLOTS mbarriers
 */
template <typename T, int NT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k_sweep_tiles(const __grid_constant__ TileSweepParams p) {
  constexpr unsigned int ROWB = 128u * sizeof(T);
  constexpr unsigned int SLOTB = SW_ROWS * ROWB;
  extern __shared__ __align__(128) unsigned char smem[];
  const unsigned int lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
  unsigned char *ring = smem + (size_t)wid * (SW_SLOTS * SLOTB);
  SweepRecord *stage = reinterpret_cast<SweepRecord *>(smem + (size_t)WARPS * (SW_SLOTS * SLOTB)) + (size_t)wid * SW_STAGE;
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + (size_t)WARPS * (SW_SLOTS * SLOTB + SW_STAGE * sizeof(SweepRecord))) + wid * SW_SLOTS;
  const uint32_t ring_s = smem_u32(ring), bars_s = smem_u32(bars);

  if (0u == lane) {
#pragma unroll
    for (int s = 0; s < SW_SLOTS; s++) mbar_init(bars_s + 8u * s, 1u);
  }
  fence_proxy_async();
  __syncwarp();

  unsigned int cslot = 0u, cphase = 0u;     /* consume cursor of the ring (continues across tiles) */
  unsigned int fill = 0u;                   /* records waiting in the staging buffer */
  const T *const base = static_cast<const T *>(p.data);

  auto flush = [&]() {
    __syncwarp();
    unsigned long long at = 0ull;
    if (0u == lane) at = atomicAdd(p.count, (unsigned long long)fill);
    at = __shfl_sync(SW_FULL, at, 0);
    const uint4 *src = reinterpret_cast<const uint4 *>(stage);
    uint4 *dst = reinterpret_cast<uint4 *>(p.records);
    for (unsigned int w = lane; w < 2u * fill; w += 32u) {
      if (at + (w >> 1) < p.capacity) dst[2ull * at + w] = src[w];
    }
    __syncwarp();
    fill = 0u;
  };

  for (;;) {
    unsigned int tile = 0u;
    if (0u == lane) tile = atomicAdd(p.tile_cursor, 1u);
    tile = __shfl_sync(SW_FULL, tile, 0);
    if (tile >= p.ntiles) break;
    const int xt = (int)(tile % (unsigned int)p.nxt);
    const int yt = (int)((tile / (unsigned int)p.nxt) % (unsigned int)p.nyt);
    const int zt = (int)(tile / ((unsigned int)p.nxt * (unsigned int)p.nyt));
    const long long x0 = (long long)xt * SW_XCUBES;
    const long long j0 = p.jlo + (long long)yt * SW_YCUBES;
    const long long k0 = p.klo + (long long)zt * p.zc;
    const long long k1 = (k0 + p.zc - 1 < p.khi) ? k0 + p.zc - 1 : p.khi;
    const int nplanes = (int)(k1 - k0) + 2;                 /* point planes k0 .. k1 + 1 */
    const long long xleft = p.mx - x0;
    const uint32_t xbytes = (uint32_t)((xleft < 128 ? xleft : 128) * (long long)sizeof(T));
    /* rows available per group; valid groups form a prefix */
    int nrows[SW_GROUPS], ng = 0;
#pragma unroll
    for (int g = 0; g < SW_GROUPS; g++) {
      const long long left = p.my - (j0 + SW_ROWS * g);
      nrows[g] = (int)(left < 0 ? 0 : (left > SW_ROWS ? SW_ROWS : left));
      if (nrows[g] > 0) ng = g + 1;
    }
    /* which of my 32 cubes per group may be reported */
    unsigned int valid[SW_GROUPS];
    {
      unsigned int xm = 0u;
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const long long i = x0 + 4 * (long long)lane + c;
        if (lane < 31u && i >= p.ilo && i <= p.ihi) xm |= 1u << c;
      }
#pragma unroll
      for (int g = 0; g < SW_GROUPS; g++) {
        unsigned int m = 0u;
#pragma unroll
        for (int r = 0; r < SW_ROWS; r++) {
          const int row = SW_ROWS * g + r;
          if (row < SW_YCUBES && j0 + row <= p.jhi) m |= xm << (4 * r);
        }
        valid[g] = m;
      }
    }

    /* producer side: one step = one row group of one plane */
    const int nsteps = nplanes * ng;
    int ipl = 0, ig = 0, issued = 0;
    unsigned int islot = cslot;
    auto issue = [&]() {
      const int rows = (0 == ig) ? nrows[0] : (1 == ig) ? nrows[1] : (2 == ig) ? nrows[2] : nrows[3];
      const uint32_t bar = bars_s + 8u * islot;
      if (0u == lane) mbar_expect_tx(bar, (uint32_t)rows * xbytes);
      __syncwarp();
      if ((int)lane < rows) {
        const T *src = base + (k0 + ipl) * p.sz + (j0 + SW_ROWS * ig + (long long)lane) * p.sy + x0;
        bulk_g2s(ring_s + islot * SLOTB + lane * ROWB, src, xbytes, bar);
      }
      issued++;
      if (++ig == ng) { ig = 0; ipl++; }
      if (++islot == SW_SLOTS) islot = 0u;
    };
    for (int s = 0; s < SW_SLOTS && s < nsteps; s++) issue();

    unsigned int LP[NT][SW_GROUPS], LQ[NT][SW_GROUPS];     /* lower plane: points x..x+3 / x+1..x+4 */
#pragma unroll
    for (int t = 0; t < NT; t++) {
#pragma unroll
      for (int g = 0; g < SW_GROUPS; g++) { LP[t][g] = 0u; LQ[t][g] = 0u; }
    }

    for (int pl = 0; pl < nplanes; pl++) {
      unsigned int UP[NT][SW_GROUPS], UQ[NT][SW_GROUPS];
#pragma unroll
      for (int g = 0; g < SW_GROUPS; g++) {
        unsigned int acc[NT];
#pragma unroll
        for (int t = 0; t < NT; t++) acc[t] = 0u;
        if (g < ng) {
          mbar_wait(bars_s + 8u * cslot, cphase);
          const T *rows = reinterpret_cast<const T *>(ring + cslot * SLOTB);
#pragma unroll
          for (int r = 0; r < SW_ROWS; r++) {
            const Quad4<T> v = lds_quad(rows + r * 128, lane);
#pragma unroll
            for (int t = 0; t < NT; t++) {
              const auto thr = Thr<T>::get(p, t);
              /* FSETP + predicated LOP3 per point */
              if (above(v.a, thr)) acc[t] |= 1u << (4 * r + 0);
              if (above(v.b, thr)) acc[t] |= 1u << (4 * r + 1);
              if (above(v.c, thr)) acc[t] |= 1u << (4 * r + 2);
              if (above(v.d, thr)) acc[t] |= 1u << (4 * r + 3);
            }
          }
          /* the slot is free again: refill it SW_SLOTS steps ahead */
          fence_proxy_async();
          __syncwarp();
          if (issued < nsteps) issue();
          if (++cslot == SW_SLOTS) { cslot = 0u; cphase ^= 1u; }
        }
#pragma unroll
        for (int t = 0; t < NT; t++) {
          const unsigned int right = __shfl_down_sync(SW_FULL, acc[t], 1);
          UP[t][g] = acc[t];
          UQ[t][g] = ((acc[t] >> 1) & 0x77777777u) | ((right << 3) & 0x88888888u);
        }
      }
      if (pl > 0) {
        const long long k = k0 + pl - 1;
#pragma unroll
        for (int t = 0; t < NT; t++) {
#pragma unroll
          for (int g = 0; g < SW_GROUPS; g++) {
            /* OR / AND over the 8 corners of each of my 32 cubes */
            const unsigned int lo_or = LP[t][g] | LQ[t][g], lo_and = LP[t][g] & LQ[t][g];
            const unsigned int up_or = UP[t][g] | UQ[t][g], up_and = UP[t][g] & UQ[t][g];
            const unsigned int z_or = lo_or | up_or, z_and = lo_and & up_and;
            unsigned int nz_or = 0u, nz_and = 0u;             /* first row of the next group */
            if (g + 1 < SW_GROUPS) {
              nz_or = LP[t][g + 1] | LQ[t][g + 1] | UP[t][g + 1] | UQ[t][g + 1];
              nz_and = LP[t][g + 1] & LQ[t][g + 1] & UP[t][g + 1] & UQ[t][g + 1];
            }
            const unsigned int any = z_or | __funnelshift_r(z_or, nz_or, 4);
            const unsigned int all = z_and & __funnelshift_r(z_and, nz_and, 4);
            const unsigned int active = any & ~all & valid[g];
            const unsigned int who = __ballot_sync(SW_FULL, 0u != active);
            if (0u != who) {
              const unsigned int cnt = __popc(who);
              if (fill + cnt > SW_STAGE) flush();
              if (0u != active) {
                unsigned int row8 = 0u;
                if (g + 1 < SW_GROUPS) {
                  row8 = (LP[t][g + 1] & 15u) | ((LQ[t][g + 1] & 15u) << 4) | ((UP[t][g + 1] & 15u) << 8) | ((UQ[t][g + 1] & 15u) << 12);
                }
                const unsigned int at = fill + __popc(who & ((1u << lane) - 1u));
                uint4 *dst = reinterpret_cast<uint4 *>(stage + at);
                dst[0] = make_uint4(LP[t][g], LQ[t][g], UP[t][g], UQ[t][g]);
                dst[1] = make_uint4(active, row8, (unsigned int)(x0 + 4 * (long long)lane) | ((unsigned int)(j0 + SW_ROWS * g) << 16),
                                    (unsigned int)k | (p.obj[t] << 16));
              }
              fill += cnt;
            }
          }
        }
      }
#pragma unroll
      for (int t = 0; t < NT; t++) {
#pragma unroll
        for (int g = 0; g < SW_GROUPS; g++) { LP[t][g] = UP[t][g]; LQ[t][g] = UQ[t][g]; }
      }
    }
  }
  if (fill > 0u) flush();
}

/* ------------------------------------------------------------------ expand */

/*
 * records -> one list entry per cut cube: code:8 | i:16 | j:16 | k:16 | object:4
 * (corner numbering bit0 = +x, bit1 = +y, bit2 = +z, triangulate.c:15-19).
 */
__global__ void __launch_bounds__(256) k_expand(const SweepRecord *__restrict__ records, const unsigned long long *__restrict__ nrecords,
                                                const unsigned long long rec_capacity, unsigned long long *__restrict__ list,
                                                unsigned long long *__restrict__ count, const unsigned long long capacity,
                                                unsigned long long *__restrict__ per_object, const unsigned int nobjects) {
  unsigned long long n = *nrecords;
  if (n > rec_capacity) n = rec_capacity;
  const unsigned int lane = threadIdx.x & 31u;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  /* whole warps iterate together (shuffles inside) */
  for (unsigned long long q0 = (unsigned long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); q0 < n; q0 += stride) {
    const unsigned long long q = q0 + lane;
    uint4 a = make_uint4(0u, 0u, 0u, 0u), b = make_uint4(0u, 0u, 0u, 0u);
    if (q < n) {
      a = __ldg(reinterpret_cast<const uint4 *>(records) + 2ull * q);
      b = __ldg(reinterpret_cast<const uint4 *>(records) + 2ull * q + 1ull);
    }
    unsigned int active = b.x;
    const unsigned int mine = __popc(active);
    unsigned int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned int up = __shfl_up_sync(SW_FULL, incl, d);
      if (lane >= (unsigned int)d) incl += up;
    }
    const unsigned int total = __shfl_sync(SW_FULL, incl, 31);
    if (0u == total) continue;
    unsigned long long at = 0ull;
    if (0u == lane) at = atomicAdd(count, (unsigned long long)total);
    at = __shfl_sync(SW_FULL, at, 0) + (incl - mine);
    /* rows 0..8 of the four bit planes */
    const unsigned long long pl = a.x | ((unsigned long long)(b.y & 15u) << 32), ql = a.y | ((unsigned long long)((b.y >> 4) & 15u) << 32);
    const unsigned long long pu = a.z | ((unsigned long long)((b.y >> 8) & 15u) << 32), qu = a.w | ((unsigned long long)((b.y >> 12) & 15u) << 32);
    const unsigned int x = b.z & 65535u, j = b.z >> 16, k = b.w & 65535u, obj = b.w >> 16;
    /* statistics: cut cubes per contour object */
    for (unsigned int o = 0; o < nobjects; o++) {
      const unsigned int sum = __reduce_add_sync(SW_FULL, obj == o ? mine : 0u);
      if (0u == lane && 0u != sum) atomicAdd(per_object + o, (unsigned long long)sum);
    }
    while (active) {
      const int bit = __ffs((int)active) - 1;
      active &= active - 1u;
      const unsigned int code = (unsigned int)((pl >> bit) & 1ull) | ((unsigned int)((ql >> bit) & 1ull) << 1)
                              | ((unsigned int)((pl >> (bit + 4)) & 1ull) << 2) | ((unsigned int)((ql >> (bit + 4)) & 1ull) << 3)
                              | ((unsigned int)((pu >> bit) & 1ull) << 4) | ((unsigned int)((qu >> bit) & 1ull) << 5)
                              | ((unsigned int)((pu >> (bit + 4)) & 1ull) << 6) | ((unsigned int)((qu >> (bit + 4)) & 1ull) << 7);
      if (at < capacity) list[at] = C3D_ENTRY(obj, x + (unsigned int)(bit & 3), j + (unsigned int)(bit >> 2), k, code);
      at++;
    }
  }
}

/* ---------------------------------------------------------------- launchers */

template <typename T, int NT, int WARPS>
static int launch_tiles(const TileSweepParams &p, const int sm_count, cudaStream_t s) {
  const size_t smem = (size_t)WARPS * (SW_SLOTS * SW_ROWS * 128 * sizeof(T) + SW_STAGE * sizeof(SweepRecord) + SW_SLOTS * 8);
  /* per device, so not cached in a static */
  const cudaError_t err = cudaFuncSetAttribute(k_sweep_tiles<T, NT, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (cudaSuccess != err) return (int)err;
  const unsigned int need = (p.ntiles + WARPS - 1) / WARPS;
  const unsigned int blocks = need < (unsigned int)sm_count ? need : (unsigned int)sm_count;
  k_sweep_tiles<T, NT, WARPS><<<blocks, WARPS * 32, smem, s>>>(p);
  return (int)cudaGetLastError();
}

extern "C" int c3d_k_sweep_tiles_ok(const FieldView *f) {
  const size_t eb = (size_t)f->elem_bytes;
  return (0 == ((uintptr_t)f->data % 16)) && (0 == (f->sy * eb) % 16) && (0 == (f->sz * eb) % 16) && (0 == (f->n[0] * eb) % 16);
}

extern "C" int c3d_k_sweep_tiles(const FieldView *f, const SweepRange *r, int nthr, const double *thr, const float *thr_ge, const unsigned int *obj,
                                 unsigned int *tile_cursor, void *records, unsigned long long *count, unsigned long long capacity,
                                 int sm_count, cudaStream_t s) {
  if (nthr < 1 || nthr > C3D_SWEEP_MAX_THR) return (int)cudaErrorInvalidValue;
  TileSweepParams p;
  p.data = f->data;
  p.mx = f->n[0]; p.my = f->n[1]; p.mz = f->n[2];
  p.sy = f->sy; p.sz = f->sz;
  p.nthr = nthr;
  for (int t = 0; t < C3D_SWEEP_MAX_THR; t++) {
    const int u = t < nthr ? t : 0;
    p.thr[t] = thr[u]; p.thr_ge[t] = thr_ge[u]; p.obj[t] = obj[u];
  }
  p.ilo = r->lo[0]; p.ihi = r->hi[0];
  p.jlo = r->lo[1]; p.jhi = r->hi[1];
  p.klo = r->lo[2]; p.khi = r->hi[2];
  if (p.ihi < p.ilo || p.jhi < p.jlo || p.khi < p.klo) return 0;
  p.nxt = (int)((p.ihi + SW_XCUBES) / SW_XCUBES);
  p.nyt = (int)((p.jhi - p.jlo + SW_YCUBES) / SW_YCUBES);
  const long long nk = p.khi - p.klo + 1;
  /* planes per tile: enough tiles for ~6 per resident warp (dynamic hand-out evens the tail), at most 1/32 z re-reads */
  const int warps = 4 == f->elem_bytes ? 16 : 8;
  const long long per_chunk = (long long)p.nxt * p.nyt;
  const long long want_tiles = 6ll * warps * sm_count;
  long long chunks = (want_tiles + per_chunk - 1) / per_chunk;
  long long zc = (nk + chunks - 1) / chunks;
  if (zc < 32) zc = 32;
  if (zc > nk) zc = nk;
  p.zc = (int)zc;
  p.nzc = (int)((nk + zc - 1) / zc);
  p.ntiles = (unsigned int)(per_chunk * p.nzc);
  p.tile_cursor = tile_cursor;
  p.records = static_cast<SweepRecord *>(records);
  p.count = count;
  p.capacity = capacity;
  const cudaError_t err = cudaMemsetAsync(tile_cursor, 0, sizeof(unsigned int), s);
  if (cudaSuccess != err) return (int)err;
  if (4 == f->elem_bytes) {
    switch (nthr) {
      case 1: return launch_tiles<float, 1, 16>(p, sm_count, s);
      case 2: return launch_tiles<float, 2, 16>(p, sm_count, s);
      case 3: return launch_tiles<float, 3, 16>(p, sm_count, s);
      default: return launch_tiles<float, 4, 16>(p, sm_count, s);
    }
  }
  switch (nthr) {
    case 1: return launch_tiles<double, 1, 8>(p, sm_count, s);
    case 2: return launch_tiles<double, 2, 8>(p, sm_count, s);
    case 3: return launch_tiles<double, 3, 8>(p, sm_count, s);
    default: return launch_tiles<double, 4, 8>(p, sm_count, s);
  }
}

extern "C" int c3d_k_expand(const void *records, const unsigned long long *nrecords, unsigned long long rec_capacity, unsigned long long *list,
                            unsigned long long *count, unsigned long long capacity, unsigned long long *per_object, unsigned int nobjects,
                            int sm_count, cudaStream_t s) {
  k_expand<<<sm_count * 8, 256, 0, s>>>(static_cast<const SweepRecord *>(records), nrecords, rec_capacity, list, count, capacity, per_object, nobjects);
  return (int)cudaGetLastError();
}
