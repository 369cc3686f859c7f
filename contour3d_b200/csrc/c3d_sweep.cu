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
 *  - every warp owns tiles of 128 x 31 cubes and marches through `zc` planes;
 *    tiles are handed out by an atomic cursor;
 *  - the field travels HBM -> shared memory through the TMA engine: ONE
 *    cp.async.bulk.tensor.3d (SASS UTMALDG) per 8 rows x (128 + pad) points,
 *    described by a tensor map of the extended array (points outside it are
 *    zero-filled and never reported), completion counted in bytes on an
 *    mbarrier, into a warp-private ring of 3-4 slots, so 12-16 KiB per warp
 *    are always in flight, no register is tied up by a load and no lane
 *    computes an address;
 *  - a lane reads its 4 points of a row with one LDS.128 and folds
 *    `thr < v` of the 4 points into bit 4r+c of a 32-bit word per iso-value
 *    (8 rows x 4 points; FSET + FFMA each), so one word describes 32 points;
 *    the point right of the tile (x + 128, the pad of each row) is classified
 *    by lanes 0, 4, .. 28 and reaches lane 31 through one ballot;
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
#include <stdlib.h>
#include <cuda.h>           /* CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint) */
#include "c3d_internal.h"

#define SW_ROWS 8           /* rows per slot = rows per 32-bit word */
#define SW_GROUPS 4         /* row groups per tile: 32 point rows, 31 cube rows */
#define SW_XCUBES 128       /* cubes per tile row: 4 per lane; the x+1 neighbour of lane 31 is the row's pad point */
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

/* shared-memory geometry of one row: 128 points + one 16-byte unit that carries the point right of the tile */
template <typename T> struct RowGeom {
  static constexpr unsigned int PAD = 16u / sizeof(T);
  static constexpr unsigned int ROWP = SW_XCUBES + PAD;
  static constexpr unsigned int ROWB = ROWP * sizeof(T);
  static constexpr unsigned int SLOTB = SW_ROWS * ROWB;       /* 4224 / 8320 bytes: multiples of 128 */
};

struct TileSweepParams {
  long long mx, my, mz;
  int nthr;
  float thr_ge[C3D_SWEEP_MAX_THR];   /* fp32 fields: thr < (double)v  <=>  v >= thr_ge */
  double thr[C3D_SWEEP_MAX_THR];     /* fp64 fields: thr < v */
  unsigned int obj[C3D_SWEEP_MAX_THR];
  long long ilo, ihi, jlo, jhi, klo, khi;   /* inclusive cube ranges to report */
  int nxt, nyt, nzc, zc;
  int nzc_a, zc_b;                   /* the first nzc_a chunks (cursor order) span zc planes, the rest zc_b: small tiles even out the tail */
  long long nk_a;                    /* planes covered by the first nzc_a chunks */
  int flip[3];                       /* visit tiles (and planes) from the high end of the axis: front-to-back for the camera */
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
/* one box of the tensor map (x, y, z: point coordinates of its first element) global -> shared; the whole box,
   zero-filled where it leaves the array, counts towards the mbarrier's transaction bytes */
__device__ __forceinline__ void tma_box_g2s(const uint32_t dst, const CUtensorMap *map, const int x, const int y, const int z, const uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
               "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

/*
 * Per-iso-value accumulator of one row group: bit 4r+c <=> point c of row r is
 * above the iso-value, i.e. `threshold < value` (triangulate.c:176-179).
 *
 * fp32 fields compare against the smallest float above the threshold, which is
 * the same predicate exactly (NaN is "not above" in both forms).  The ALU pipe
 * (16 lanes/clk/SMSP: FSETP, LOP3, SHF) is the scarce resource of this kernel,
 * so the float version spends one ALU op (FSET.BF -> 1.0f / 0.0f) and one
 * FMA-pipe op (FFMA with the bit weight; sums of distinct powers of two below
 * 2^16 are exact in fp32) per point instead of two ALU ops.
 */
template <typename T> struct GroupBits;
template <> struct GroupBits<float> {
  float lo, hi;               /* rows 0-3 / rows 4-7, weights 2^(4(r&3)+c) */
  __device__ __forceinline__ void clear() { lo = 0.f; hi = 0.f; }
  template <int R, int C> __device__ __forceinline__ void add(const float v, const float thr_ge) {
    float one;
    asm("set.ge.f32.f32 %0, %1, %2;" : "=f"(one) : "f"(v), "f"(thr_ge));
    constexpr float weight = (float)(1u << (4 * (R & 3) + C));
    if (R < 4) lo = fmaf(one, weight, lo);
    else hi = fmaf(one, weight, hi);
  }
  __device__ __forceinline__ unsigned int word() const { return __float2uint_rz(lo) | (__float2uint_rz(hi) << 16); }
  __device__ static __forceinline__ bool above(const float v, const float thr_ge) { return v >= thr_ge; }
};
template <> struct GroupBits<double> {
  unsigned int w;
  __device__ __forceinline__ void clear() { w = 0u; }
  template <int R, int C> __device__ __forceinline__ void add(const double v, const double thr) {
    if (thr < v) w |= 1u << (4 * R + C);
  }
  __device__ __forceinline__ unsigned int word() const { return w; }
  __device__ static __forceinline__ bool above(const double v, const double thr) { return thr < v; }
};

template <typename T> struct Thr;
template <> struct Thr<float> {
  __device__ static __forceinline__ float get(const TileSweepParams &p, const int t) { return p.thr_ge[t]; }
};
template <> struct Thr<double> {
  __device__ static __forceinline__ double get(const TileSweepParams &p, const int t) { return p.thr[t]; }
};

template <typename T, int NT, int R> struct RowFold {
  /* rows R..7 of a slot, unrolled at compile time */
  __device__ static __forceinline__ void run(const T *rows, const unsigned int lane, const TileSweepParams &p, GroupBits<T> (&acc)[NT]);
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

template <typename T, int NT, int R>
__device__ __forceinline__ void RowFold<T, NT, R>::run(const T *rows, const unsigned int lane, const TileSweepParams &p, GroupBits<T> (&acc)[NT]) {
  if constexpr (R < SW_ROWS) {
    const Quad4<T> v = lds_quad(rows + R * RowGeom<T>::ROWP, lane);
#pragma unroll
    for (int t = 0; t < NT; t++) {
      const auto thr = Thr<T>::get(p, t);
      acc[t].template add<R, 0>(v.a, thr);
      acc[t].template add<R, 1>(v.b, thr);
      acc[t].template add<R, 2>(v.c, thr);
      acc[t].template add<R, 3>(v.d, thr);
    }
    RowFold<T, NT, R + 1>::run(rows, lane, p, acc);
  }
}

/*
 * One warp = one tile at a time.  Shared memory per warp:
 *   ring   SW_SLOTS x SW_ROWS rows x (128 + pad) points
 *   stage  SW_STAGE records
 *   bars   SW_SLOTS mbarriers
 */
template <typename T, int NT, int WARPS, int SW_SLOTS>
__global__ void __launch_bounds__(WARPS * 32, 1) k_sweep_tiles(const __grid_constant__ TileSweepParams p, const __grid_constant__ CUtensorMap tmap) {
  constexpr unsigned int ROWB = RowGeom<T>::ROWB;
  constexpr unsigned int SLOTB = RowGeom<T>::SLOTB;
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
    int xt = (int)(tile % (unsigned int)p.nxt);
    int yt = (int)((tile / (unsigned int)p.nxt) % (unsigned int)p.nyt);
    int zt = (int)(tile / ((unsigned int)p.nxt * (unsigned int)p.nyt));
    /* tiles are handed out (and their records appended) in this order, which the raster follows: nearest
       tiles first lets its occlusion test drop most of what lies behind */
    if (p.flip[0]) xt = p.nxt - 1 - xt;
    if (p.flip[1]) yt = p.nyt - 1 - yt;
    const long long x0 = (long long)xt * SW_XCUBES;
    const long long j0 = p.jlo + (long long)yt * SW_YCUBES;
    /* cube planes u0 .. u1 counted from the end of the z axis the sweep starts at */
    const long long nk = p.khi - p.klo + 1;
    long long u0, u1;
    if (zt < p.nzc_a) {
      u0 = (long long)zt * p.zc;
      u1 = (u0 + p.zc < p.nk_a ? u0 + p.zc : p.nk_a) - 1;
    } else {
      u0 = p.nk_a + (long long)(zt - p.nzc_a) * p.zc_b;
      u1 = (u0 + p.zc_b < nk ? u0 + p.zc_b : nk) - 1;
    }
    const long long k0 = p.flip[2] ? p.khi - u1 : p.klo + u0;
    const long long k1 = p.flip[2] ? p.khi - u0 : p.klo + u1;
    const int nplanes = (int)(k1 - k0) + 2;                 /* point planes k0 .. k1 + 1 */
    /* row groups that hold at least one row of the array (a prefix); the rest of a box is zero-filled */
    int ng = 0;
#pragma unroll
    for (int g = 0; g < SW_GROUPS; g++) {
      if (j0 + SW_ROWS * g < p.my) ng = g + 1;
    }
    /* which of my 32 cubes per group may be reported */
    unsigned int valid[SW_GROUPS];
    {
      unsigned int xm = 0u;
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const long long i = x0 + 4 * (long long)lane + c;
        if (i >= p.ilo && i <= p.ihi) xm |= 1u << c;
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
    const int tx = (int)x0, ty = (int)j0;
    const int tz = (int)(p.flip[2] ? k1 + 1 : k0), tdz = p.flip[2] ? -1 : 1;
    auto issue = [&]() {
      if (0u == lane) {
        const uint32_t bar = bars_s + 8u * islot;
        mbar_expect_tx(bar, SLOTB);
        tma_box_g2s(ring_s + islot * SLOTB, &tmap, tx, ty + SW_ROWS * ig, tz + tdz * ipl, bar);
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

    unsigned int UP[NT][SW_GROUPS], UQ[NT][SW_GROUPS];     /* upper plane, being assembled */
#pragma unroll
    for (int t = 0; t < NT; t++) {
#pragma unroll
      for (int g = 0; g < SW_GROUPS; g++) { UP[t][g] = 0u; UQ[t][g] = 0u; }
    }
    for (int pl = 0; pl < nplanes; pl++) {
      /* the row-group loop is kept rolled (the body is large: 8 rows x 4 points x NT iso-values); each new
         group enters at index 3 and the older ones slide down, so after four turns UP[t][g] is group g */
#pragma unroll 1
      for (int g = 0; g < SW_GROUPS; g++) {
        GroupBits<T> acc[NT];
        unsigned int edge[NT];                /* bit 4r: the point right of the tile (x0 + 128) in row r is above */
#pragma unroll
        for (int t = 0; t < NT; t++) { acc[t].clear(); edge[t] = 0u; }
        if (g < ng) {
          mbar_wait(bars_s + 8u * cslot, cphase);
          const T *rows = reinterpret_cast<const T *>(ring + cslot * SLOTB);
          RowFold<T, NT, 0>::run(rows, lane, p, acc);
          const T pad = rows[(lane >> 2) * RowGeom<T>::ROWP + SW_XCUBES];
#pragma unroll
          for (int t = 0; t < NT; t++) edge[t] = __ballot_sync(SW_FULL, 0u == (lane & 3u) && GroupBits<T>::above(pad, Thr<T>::get(p, t)));
          /* the slot is free again: refill it SW_SLOTS steps ahead */
          fence_proxy_async();
          __syncwarp();
          if (issued < nsteps) issue();
          if (++cslot == SW_SLOTS) { cslot = 0u; cphase ^= 1u; }
        }
#pragma unroll
        for (int t = 0; t < NT; t++) {
          const unsigned int w = acc[t].word();
          const unsigned int nb = __shfl_down_sync(SW_FULL, w, 1);
          const unsigned int right = 31u == lane ? edge[t] : nb;
#pragma unroll
          for (int h = 0; h + 1 < SW_GROUPS; h++) { UP[t][h] = UP[t][h + 1]; UQ[t][h] = UQ[t][h + 1]; }
          UP[t][SW_GROUPS - 1] = w;
          UQ[t][SW_GROUPS - 1] = ((w >> 1) & 0x77777777u) | ((right << 3) & 0x88888888u);
        }
      }
      if (pl > 0) {
        /* marching down, the plane just read is the cubes' lower face and the previous one their upper face */
        const bool down = 0 != p.flip[2];
        const long long k = down ? k1 + 1 - pl : k0 + pl - 1;
        /* OR / AND over the 8 corners of each of my 32 cubes per group and iso-value */
        unsigned int active[NT][SW_GROUPS], some = 0u;
#pragma unroll
        for (int t = 0; t < NT; t++) {
          unsigned int c_or[SW_GROUPS], c_and[SW_GROUPS];
#pragma unroll
          for (int g = 0; g < SW_GROUPS; g++) {
            c_or[g] = LP[t][g] | LQ[t][g] | UP[t][g] | UQ[t][g];
            c_and[g] = LP[t][g] & LQ[t][g] & UP[t][g] & UQ[t][g];
          }
#pragma unroll
          for (int g = 0; g < SW_GROUPS; g++) {
            const unsigned int n_or = g + 1 < SW_GROUPS ? c_or[g + 1] : 0u, n_and = g + 1 < SW_GROUPS ? c_and[g + 1] : 0u;
            const unsigned int any = c_or[g] | __funnelshift_r(c_or[g], n_or, 4);      /* rows r and r+1 */
            const unsigned int all = c_and[g] & __funnelshift_r(c_and[g], n_and, 4);
            active[t][g] = any & ~all & valid[g];
            some |= active[t][g];
          }
        }
        if (__any_sync(SW_FULL, 0u != some)) {
#pragma unroll
          for (int t = 0; t < NT; t++) {
#pragma unroll
            for (int g = 0; g < SW_GROUPS; g++) {
              const unsigned int who = __ballot_sync(SW_FULL, 0u != active[t][g]);
              if (0u != who) {
                const unsigned int cnt = __popc(who);
                if (fill + cnt > SW_STAGE) flush();
                if (0u != active[t][g]) {
                  const unsigned int bp = down ? UP[t][g] : LP[t][g], bq = down ? UQ[t][g] : LQ[t][g];   /* z = k */
                  const unsigned int tp = down ? LP[t][g] : UP[t][g], tq = down ? LQ[t][g] : UQ[t][g];   /* z = k + 1 */
                  unsigned int row8 = 0u;
                  if (g + 1 < SW_GROUPS) {
                    const unsigned int bp8 = down ? UP[t][g + 1] : LP[t][g + 1], bq8 = down ? UQ[t][g + 1] : LQ[t][g + 1];
                    const unsigned int tp8 = down ? LP[t][g + 1] : UP[t][g + 1], tq8 = down ? LQ[t][g + 1] : UQ[t][g + 1];
                    row8 = (bp8 & 15u) | ((bq8 & 15u) << 4) | ((tp8 & 15u) << 8) | ((tq8 & 15u) << 12);
                  }
                  const unsigned int at = fill + __popc(who & ((1u << lane) - 1u));
                  uint4 *dst = reinterpret_cast<uint4 *>(stage + at);
                  dst[0] = make_uint4(bp, bq, tp, tq);
                  dst[1] = make_uint4(active[t][g], row8, (unsigned int)(x0 + 4 * (long long)lane) | ((unsigned int)(j0 + SW_ROWS * g) << 16),
                                      (unsigned int)k | (p.obj[t] << 16));
                }
                fill += cnt;
              }
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
 *
 * SORTED: the entries are written nearest-first as a counting sort by view
 * depth (C3D_DEPTH_BUCKETS slabs perpendicular to the screen normal).  The
 * image does not depend on the order; the raster's occlusion test does: drawn
 * front to back, most cubes behind the first surfaces are dropped before any
 * triangle is built.  k_bucket_count histograms the records, k_bucket_scan
 * turns the histogram into cursors, k_expand<true> scatters.
 */
struct DepthOrder {
  float near_depth, scale;          /* bucket = (near_depth - n.(p - cam)) * scale, clamped */
  float last;                       /* highest bucket */
};

__device__ __forceinline__ unsigned int depth_bucket(const FrameParams &p, const DepthOrder &o, const unsigned int obj, const unsigned int x,
                                                     const unsigned int j, const unsigned int k) {
  const FieldView &f = p.view[obj];
  float rx, ry, rz;
  if (nullptr != f.map) {
    mapped_posf(f.map, (long long)x, (long long)j, (long long)k, rx, ry, rz);
  } else {
    rx = (float)__ldg(f.g[0] + x); ry = (float)__ldg(f.g[1] + j); rz = (float)__ldg(f.g[2] + k);
  }
  rx -= p.cull.cam[0]; ry -= p.cull.cam[1]; rz -= p.cull.cam[2];
  const float nr = p.cull.n[0] * rx + p.cull.n[1] * ry + p.cull.n[2] * rz;
  const float b = (o.near_depth - nr) * o.scale;
  return (unsigned int)fminf(fmaxf(b, 0.f), o.last);    /* NaN -> 0 */
}

__global__ void __launch_bounds__(256) k_bucket_count(const __grid_constant__ FrameParams p, const DepthOrder o,
                                                      const SweepRecord *__restrict__ records, const unsigned long long *__restrict__ nrecords,
                                                      const unsigned long long rec_capacity, unsigned int *__restrict__ histogram) {
  __shared__ unsigned int local[C3D_DEPTH_BUCKETS];
  for (unsigned int b = threadIdx.x; b < C3D_DEPTH_BUCKETS; b += blockDim.x) local[b] = 0u;
  __syncthreads();
  unsigned long long n = *nrecords;
  if (n > rec_capacity) n = rec_capacity;
  for (unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (unsigned long long)gridDim.x * blockDim.x) {
    const uint4 b = __ldg(reinterpret_cast<const uint4 *>(records) + 2ull * q + 1ull);
    atomicAdd(local + depth_bucket(p, o, b.w >> 16, b.z & 65535u, b.z >> 16, b.w & 65535u), (unsigned int)__popc(b.x));
  }
  __syncthreads();
  for (unsigned int b = threadIdx.x; b < C3D_DEPTH_BUCKETS; b += blockDim.x) {
    if (0u != local[b]) atomicAdd(histogram + b, local[b]);
  }
}

/* one block: histogram -> start offsets behind what the list already holds; the list cursor moves past all buckets */
__global__ void __launch_bounds__(1024) k_bucket_scan(const unsigned int *__restrict__ histogram, unsigned long long *__restrict__ cursors,
                                                      unsigned long long *__restrict__ count) {
  __shared__ unsigned long long part[1024];
  constexpr unsigned int PER = C3D_DEPTH_BUCKETS / 1024;
  unsigned long long mine = 0ull;
  for (unsigned int q = 0; q < PER; q++) mine += histogram[threadIdx.x * PER + q];
  part[threadIdx.x] = mine;
  __syncthreads();
  for (unsigned int d = 1; d < 1024; d <<= 1) {
    const unsigned long long add = threadIdx.x >= d ? part[threadIdx.x - d] : 0ull;
    __syncthreads();
    part[threadIdx.x] += add;
    __syncthreads();
  }
  const unsigned long long first = *count;
  unsigned long long at = first + part[threadIdx.x] - mine;
  for (unsigned int q = 0; q < PER; q++) {
    cursors[threadIdx.x * PER + q] = at;
    at += histogram[threadIdx.x * PER + q];
  }
  __syncthreads();
  if (1023u == threadIdx.x) *count = first + part[1023];
}

/*
 * Can a triangle of this 4 x 8 patch of cubes have collapsed to a point?  Vertices reach a cube corner only
 * when an interpolation factor (thr - v0) / (v1 - v0) = a0 / (a0 + a1), a = |v - thr|, is within rounding of
 * 0 or 1 (or |v1 - v0| < 1e-8, triangulate.c:62-71).  If min a > 1e-6 max a over the points of the patch that
 * its cut cubes use, every factor of every such cube lies in (5e-7, 1 - 5e-7): no collapse.  The reference's other branch is
 * absolute -- |v1 - v0| < 1e-8 puts the vertex on the tail corner whatever the ratio (triangulate.c:62-71; a field of tiny
 * magnitude) -- so min a > 1e-8 is asked as well: every straddling pair then has |v1 - v0| = a0 + a1 >= 2e-8.  Anything else
 * (including NaN / infinite data) marks the patch, and the raster then skips its visibility shortcuts there.
 */
__device__ __forceinline__ bool patch_irregular(const FieldView &f, const unsigned int x, const unsigned int j, const unsigned int k,
                                                const unsigned int active) {
  /* only the point rows that cut cubes of the patch touch (cube row r uses point rows r and r + 1) */
  unsigned int need = 0u;
#pragma unroll
  for (int r = 0; r < 8; r++) {
    if ((active >> (4 * r)) & 15u) need |= 3u << r;
  }
  const long long nx = f.n[0] - (long long)x < 5 ? f.n[0] - (long long)x : 5;
  if (4 == f.elem_bytes) {
    /* fp32 data: v - thr32 is exact where it matters (Sterbenz: operands within a factor of two), so the only
       error in a = |v - thr| is the rounding of the iso-value itself, thr_err = |thr - (double)thr32|; with
       amin32 > 1.1e-6 amax32 + 1.01 thr_err the exact quantities satisfy amin > 1e-6 amax */
    const float thr32 = (float)f.thr;
    const float thr_err = (float)(fabs(f.thr - (double)thr32) * 1.01) + 0.f;
    const float *data = (const float *)f.data;
    const bool vec = (0 == (f.sy & 3)) && (0 == (f.sz & 3)) && (0 == ((uintptr_t)data & 15)) && nx >= 4;   /* x is a multiple of 4 */
    float amin = FLT_MAX, amax = 0.f, sum = 0.f;
    if (vec) {
      /* all rows of a plane are requested before any is looked at (rows that are not needed, or lie beyond the array, re-read
         a needed one): the test is bound by the latency of these scattered loads, not by their number */
      const int last = (int)((f.n[1] - 1 - (long long)j) < 8 ? (f.n[1] - 1 - (long long)j) : 8);
      const int fallback = __ffs((int)need) - 1;
#pragma unroll
      for (int dz = 0; dz < 2; dz++) {
        const float *plane = data + ((long long)k + dz) * f.sz + (long long)j * f.sy + (long long)x;
        float4 q[9];
        float e[9];
#pragma unroll
        for (int r = 0; r < 9; r++) {
          const int rr = (((need >> r) & 1u) && r <= last) ? r : fallback;
          const float *row = plane + (long long)rr * f.sy;
          q[r] = __ldg(reinterpret_cast<const float4 *>(row));
          e[r] = __ldg(row + (nx > 4 ? 4 : 3));
        }
#pragma unroll
        for (int r = 0; r < 9; r++) {
          const float a0 = fabsf(q[r].x - thr32), a1 = fabsf(q[r].y - thr32), a2 = fabsf(q[r].z - thr32), a3 = fabsf(q[r].w - thr32), a4 = fabsf(e[r] - thr32);
          amin = fminf(amin, fminf(fminf(a0, a1), fminf(fminf(a2, a3), a4)));
          amax = fmaxf(amax, fmaxf(fmaxf(a0, a1), fmaxf(fmaxf(a2, a3), a4)));
          sum += (a0 + a1) + (a2 + a3) + a4;             /* NaN / inf detector (fminf / fmaxf would drop a NaN) */
        }
      }
      return !(sum <= FLT_MAX) || !(amin > 1.1e-6f * amax + thr_err) || !(amin > 1.01e-8f + thr_err);
    }
    for (int dz = 0; dz < 2; dz++) {
      for (int r = 0; r < 9; r++) {
        if ((long long)j + r >= f.n[1]) break;
        if (0u == ((need >> r) & 1u)) continue;
        const float *row = data + ((long long)k + dz) * f.sz + ((long long)j + r) * f.sy + (long long)x;
        float v[5];
#pragma unroll
        for (int c = 0; c < 5; c++) v[c] = __ldg(row + (c < nx ? c : 0));
#pragma unroll
        for (int c = 0; c < 5; c++) {
          const float a = fabsf(v[c] - thr32);
          amin = fminf(amin, a);
          amax = fmaxf(amax, a);
          sum += a;                        /* NaN / inf detector (fminf / fmaxf would drop a NaN) */
        }
      }
    }
    return !(sum <= FLT_MAX) || !(amin > 1.1e-6f * amax + thr_err) || !(amin > 1.01e-8f + thr_err);
  }
  double amin = DBL_MAX, amax = 0.;
  bool bad = false;
  for (int dz = 0; dz < 2; dz++) {
    for (int r = 0; r < 9; r++) {
      if ((long long)j + r >= f.n[1]) break;
      if (0u == ((need >> r) & 1u)) continue;
      const double *row = (const double *)f.data + ((long long)k + dz) * f.sz + ((long long)j + r) * f.sy + (long long)x;
      for (long long c = 0; c < nx; c++) {
        const double a = fabs(__ldg(row + c) - f.thr);
        bad |= !(a <= DBL_MAX);
        amin = fmin(amin, a);
        amax = fmax(amax, a);
      }
    }
  }
  return bad || !(amin > 1.e-6 * amax) || !(amin > 1.e-8);
}

template <bool SORTED>
__global__ void __launch_bounds__(256) k_expand(const __grid_constant__ FrameParams p, const DepthOrder o,
                                                const SweepRecord *__restrict__ records, const unsigned long long *__restrict__ nrecords,
                                                const unsigned long long rec_capacity, unsigned long long *__restrict__ list,
                                                unsigned long long *__restrict__ count, const unsigned long long capacity,
                                                unsigned long long *__restrict__ cursors,
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
    const unsigned int x = b.z & 65535u, j = b.z >> 16, k = b.w & 65535u, obj = b.w >> 16;
    unsigned long long at = 0ull;
    if (SORTED) {
      if (0u != mine) at = atomicAdd(cursors + depth_bucket(p, o, obj, x, j, k), (unsigned long long)mine);
    } else {
      unsigned int incl = mine;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned int up = __shfl_up_sync(SW_FULL, incl, d);
        if (lane >= (unsigned int)d) incl += up;
      }
      const unsigned int total = __shfl_sync(SW_FULL, incl, 31);
      if (0u == total) continue;
      if (0u == lane) at = atomicAdd(count, (unsigned long long)total);
      at = __shfl_sync(SW_FULL, at, 0) + (incl - mine);
    }
    /* wanted whenever the raster may skip cubes, sorted list or not */
    const unsigned long long flag = (p.cull.enabled && 0u != mine && patch_irregular(p.view[obj], x, j, k, active)) ? C3D_ENTRY_IRREGULAR_BIT : 0ull;
    /* rows 0..8 of the four bit planes */
    const unsigned long long pl = a.x | ((unsigned long long)(b.y & 15u) << 32), ql = a.y | ((unsigned long long)((b.y >> 4) & 15u) << 32);
    const unsigned long long pu = a.z | ((unsigned long long)((b.y >> 8) & 15u) << 32), qu = a.w | ((unsigned long long)((b.y >> 12) & 15u) << 32);
    /* statistics: cut cubes per contour object */
    for (unsigned int ob = 0; ob < nobjects; ob++) {
      const unsigned int sum = __reduce_add_sync(SW_FULL, obj == ob ? mine : 0u);
      if (0u == lane && 0u != sum) atomicAdd(per_object + ob, (unsigned long long)sum);
    }
    while (active) {
      const int bit = __ffs((int)active) - 1;
      active &= active - 1u;
      const unsigned int code = (unsigned int)((pl >> bit) & 1ull) | ((unsigned int)((ql >> bit) & 1ull) << 1)
                              | ((unsigned int)((pl >> (bit + 4)) & 1ull) << 2) | ((unsigned int)((ql >> (bit + 4)) & 1ull) << 3)
                              | ((unsigned int)((pu >> bit) & 1ull) << 4) | ((unsigned int)((qu >> bit) & 1ull) << 5)
                              | ((unsigned int)((pu >> (bit + 4)) & 1ull) << 6) | ((unsigned int)((qu >> (bit + 4)) & 1ull) << 7);
      if (at < capacity) list[at] = C3D_ENTRY(obj, x + (unsigned int)(bit & 3), j + (unsigned int)(bit >> 2), k, code) | flag;
      at++;
    }
  }
}

/* ------------------------------------------------- screen-partitioned expand */

/*
 * Which ranks' screen tiles can cube (i, j, k) of view f touch?  Bit r of the result = rank r must draw it.
 * The pixels a cube can reach are the reference's search boxes of its triangles: trunc((0.5 + sx) W -+ 1) clamped to
 * the screen (render.c:81-89), and every vertex lies inside the hull of the eight corners, so the box of the
 * corners, widened by one pixel plus a quarter for the fp32 arithmetic, holds them all.  Cubes the fp32 projection
 * cannot bound (a corner at or behind the screen plane, non-finite data, a huge footprint) go to every rank: each rank
 * only ever draws its own pixels, so a superset costs time, never correctness.
 */
__device__ __forceinline__ unsigned int cube_owners(const FrameParams &p, const FieldView &f, const long long i, const long long j, const long long k) {
  const Partition &part = p.part;
  const unsigned int everyone = (1u << part.nranks) - 1u;
  const CullCam &q = p.cull;
  if (!q.valid || nullptr != f.map) return everyone;
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
  float xlo = 3.e38f, xhi = -3.e38f, ylo = 3.e38f, yhi = -3.e38f, tlo = 3.e38f, thi = -3.e38f;
#pragma unroll
  for (int c = 0; c < 8; c++) {
    const int a = c & 1, b = (c >> 1) & 1, d = c >> 2;
    const float nr = an[a] + bn[b] + cn[d];
    const float t = __fdividef(q.c, nr);
    const float sx = q.ax + t * (al[a] + bl[b] + cl[d]);
    const float sy = q.ay + t * (am[a] + bm[b] + cm[d]);
    xlo = fminf(xlo, sx); xhi = fmaxf(xhi, sx);
    ylo = fminf(ylo, sy); yhi = fmaxf(yhi, sy);
    tlo = fminf(tlo, t); thi = fmaxf(thi, t);
  }
  if (!(tlo > 1.e-4f && thi < 0.9999f)) return everyone;
  if (!(xlo > -1.e6f && xhi < 1.e6f && ylo > -1.e6f && yhi < 1.e6f)) return everyone;
  const float wmax = (float)(q.w - 1), hmax = (float)(q.h - 1);
  const int imin = (int)fminf(fmaxf(floorf((0.5f + xlo) * q.W - 1.25f), 0.f), wmax), imax = (int)fminf(fmaxf(floorf((0.5f + xhi) * q.W + 1.25f), 0.f), wmax);
  const int jmin = (int)fminf(fmaxf(floorf((0.5f + ylo) * q.H - 1.25f), 0.f), hmax), jmax = (int)fminf(fmaxf(floorf((0.5f + yhi) * q.H + 1.25f), 0.f), hmax);
  const int tx0 = imin >> part.tile_shift, tx1 = imax >> part.tile_shift, ty0 = jmin >> part.tile_shift, ty1 = jmax >> part.tile_shift;
  if ((tx1 - tx0 + 1) * (ty1 - ty0 + 1) > 16) return everyone;
  unsigned int mask = 0u;
  for (int ty = ty0; ty <= ty1; ty++) {
    for (int tx = tx0; tx <= tx1; tx++) mask |= 1u << tile_owner(part, (unsigned int)tx, (unsigned int)ty);
  }
  return mask;
}

/*
 * Pass 1: who draws each cut cube?  Warp-cooperative -- a warp takes 32 records and spreads their 1..32 cut cubes over
 * its lanes (the records hold 3-4 on average; one thread per record would idle most lanes through the eight
 * projections of cube_owners()).  The answer is kept, one byte per cube at masks[record * 32 + bit], for pass 2, and
 * counted per (destination rank, depth bucket): histogram[r * C3D_PART_BUCKETS + b].
 */
__global__ void __launch_bounds__(256) k_part_count(const __grid_constant__ FrameParams p, const DepthOrder o,
                                                    const SweepRecord *__restrict__ records, const unsigned long long *__restrict__ nrecords,
                                                    const unsigned long long rec_capacity, unsigned char *__restrict__ masks,
                                                    unsigned int *__restrict__ histogram) {
  extern __shared__ unsigned int local[];
  __shared__ unsigned int s_active[8][32], s_first[8][32], s_xj[8][32], s_kobj[8][32], s_bucket[8][32];
  const unsigned int nbins = (unsigned int)p.part.nranks * C3D_PART_BUCKETS;
  for (unsigned int b = threadIdx.x; b < nbins; b += blockDim.x) local[b] = 0u;
  __syncthreads();
  unsigned long long n = *nrecords;
  if (n > rec_capacity) n = rec_capacity;
  const unsigned int lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long q0 = (unsigned long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); q0 < n; q0 += stride) {
    const unsigned long long q = q0 + lane;
    uint4 b = make_uint4(0u, 0u, 0u, 0u);
    if (q < n) b = __ldg(reinterpret_cast<const uint4 *>(records) + 2ull * q + 1ull);
    const unsigned int active = b.x, mine = __popc(active);
    unsigned int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned int up = __shfl_up_sync(SW_FULL, incl, d);
      if (lane >= (unsigned int)d) incl += up;
    }
    const unsigned int total = __shfl_sync(SW_FULL, incl, 31), excl = incl - mine;
    __syncwarp();
    s_active[wid][lane] = active;
    s_first[wid][lane] = excl;
    s_xj[wid][lane] = b.z;
    s_kobj[wid][lane] = b.w;
    s_bucket[wid][lane] = 0u != mine ? depth_bucket(p, o, b.w >> 16, b.z & 65535u, b.z >> 16, b.w & 65535u) : 0u;
    __syncwarp();
    for (unsigned int t0 = 0; t0 < total; t0 += 32u) {
      /* record of cube t0 + lane: the last record with cubes whose first cube is <= it */
      const bool starts_here = mine > 0u && excl >= t0 && excl < t0 + 32u;
      const unsigned int starts = __reduce_or_sync(SW_FULL, starts_here ? 1u << (excl - t0) : 0u);
      const unsigned int has = __ballot_sync(SW_FULL, mine > 0u);
      const unsigned int before = __popc(__ballot_sync(SW_FULL, mine > 0u && excl < t0));
      if (t0 + lane < total) {
        /* the (before + started so far)-th record that has cubes, counted over the lanes that have any */
        const unsigned int nth = before + __popc(starts & (0xffffffffu >> (31u - lane)));      /* >= 1 */
        const unsigned int src = __fns(has, 0u, (int)nth);
        const unsigned int within = t0 + lane - s_first[wid][src];
        const unsigned int bit = __fns(s_active[wid][src], 0u, (int)within + 1);
        const unsigned int xj = s_xj[wid][src], kobj = s_kobj[wid][src];
        const FieldView &f = p.view[kobj >> 16];
        unsigned int owners = cube_owners(p, f, (long long)(xj & 65535u) + (bit & 3u), (long long)(xj >> 16) + (bit >> 2), (long long)(kobj & 65535u));
        masks[(q0 + src) * 32ull + bit] = (unsigned char)owners;
        const unsigned int bucket = s_bucket[wid][src];
        while (owners) {
          const int r = __ffs((int)owners) - 1;
          owners &= owners - 1u;
          atomicAdd(local + (unsigned int)r * C3D_PART_BUCKETS + bucket, 1u);
        }
      }
    }
  }
  __syncthreads();
  for (unsigned int b = threadIdx.x; b < nbins; b += blockDim.x) {
    if (0u != local[b]) atomicAdd(histogram + b, local[b]);
  }
}

/* one block: histogram -> cursors into the outbox, destination-major; seg_start[r] = where rank r's cubes begin;
   *count = all of them (the host compares it with the outbox capacity) */
__global__ void __launch_bounds__(1024) k_part_scan(const unsigned int *__restrict__ histogram, const int nranks, unsigned long long *__restrict__ cursors,
                                                    unsigned long long *__restrict__ seg_start, unsigned long long *__restrict__ count) {
  __shared__ unsigned long long part[1024];
  const unsigned int per = (unsigned int)(C3D_PART_BUCKETS / 1024) * (unsigned int)nranks;
  const unsigned int first = threadIdx.x * per;
  unsigned long long mine = 0ull;
  for (unsigned int q = 0; q < per; q++) mine += histogram[first + q];
  part[threadIdx.x] = mine;
  __syncthreads();
  for (unsigned int d = 1; d < 1024; d <<= 1) {
    const unsigned long long add = threadIdx.x >= d ? part[threadIdx.x - d] : 0ull;
    __syncthreads();
    part[threadIdx.x] += add;
    __syncthreads();
  }
  unsigned long long at = part[threadIdx.x] - mine;
  for (unsigned int q = 0; q < per; q++) {
    const unsigned int bin = first + q;
    cursors[bin] = at;
    if (0u == bin % C3D_PART_BUCKETS) seg_start[bin / C3D_PART_BUCKETS] = at;
    at += histogram[bin];
  }
  if (1023u == threadIdx.x) *count = part[1023];
}

/* records -> one outbox entry per (cut cube, rank that must draw it), sorted by (rank, depth bucket); entries are tagged
   with this rank (bits 61-63): their cube indices refer to ITS array */
__global__ void __launch_bounds__(256) k_part_expand(const __grid_constant__ FrameParams p, const DepthOrder o,
                                                     const SweepRecord *__restrict__ records, const unsigned long long *__restrict__ nrecords,
                                                     const unsigned long long rec_capacity, const unsigned char *__restrict__ masks,
                                                     unsigned long long *__restrict__ outbox,
                                                     const unsigned long long capacity, unsigned long long *__restrict__ cursors,
                                                     unsigned long long *__restrict__ per_object, const unsigned int nobjects) {
  unsigned long long n = *nrecords;
  if (n > rec_capacity) n = rec_capacity;
  const unsigned int lane = threadIdx.x & 31u;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  const unsigned long long tag = (unsigned long long)p.part.me << C3D_ENTRY_SRC_SHIFT;
  for (unsigned long long q0 = (unsigned long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); q0 < n; q0 += stride) {
    const unsigned long long q = q0 + lane;
    uint4 a = make_uint4(0u, 0u, 0u, 0u), b = make_uint4(0u, 0u, 0u, 0u);
    if (q < n) {
      a = __ldg(reinterpret_cast<const uint4 *>(records) + 2ull * q);
      b = __ldg(reinterpret_cast<const uint4 *>(records) + 2ull * q + 1ull);
    }
    const unsigned int active = b.x;
    const unsigned int mine = __popc(active);
    const unsigned int x = b.z & 65535u, j = b.z >> 16, k = b.w & 65535u, obj = b.w >> 16;
    for (unsigned int ob = 0; ob < nobjects; ob++) {
      const unsigned int sum = __reduce_add_sync(SW_FULL, obj == ob ? mine : 0u);
      if (0u == lane && 0u != sum) atomicAdd(per_object + ob, (unsigned long long)sum);
    }
    if (0u == mine) continue;
    const FieldView &f = p.view[obj];
    const unsigned int bucket = depth_bucket(p, o, obj, x, j, k);
    const unsigned long long flag = (p.cull.enabled && patch_irregular(f, x, j, k, active)) ? C3D_ENTRY_IRREGULAR_BIT : 0ull;
    const unsigned long long pl = a.x | ((unsigned long long)(b.y & 15u) << 32), ql = a.y | ((unsigned long long)((b.y >> 4) & 15u) << 32);
    const unsigned long long pu = a.z | ((unsigned long long)((b.y >> 8) & 15u) << 32), qu = a.w | ((unsigned long long)((b.y >> 12) & 15u) << 32);
    /* who must draw each of my cubes (pass 1 left one byte per cube); then one cursor bump per (record, rank) */
    unsigned char owners[32];
    unsigned int present = 0u;
    {
      const uint4 m0 = __ldg(reinterpret_cast<const uint4 *>(masks) + 2ull * q), m1 = __ldg(reinterpret_cast<const uint4 *>(masks) + 2ull * q + 1ull);
      const unsigned int word[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
      unsigned int todo = active;
      while (todo) {
        const int bit = __ffs((int)todo) - 1;
        todo &= todo - 1u;
        const unsigned int m = (word[bit >> 2] >> (8 * (bit & 3))) & 255u;
        owners[bit] = (unsigned char)m;
        present |= m;
      }
    }
    while (present) {
      const int r = __ffs((int)present) - 1;
      present &= present - 1u;
      unsigned int cnt = 0u;
      {
        unsigned int todo = active;
        while (todo) {
          const int bit = __ffs((int)todo) - 1;
          todo &= todo - 1u;
          cnt += (owners[bit] >> r) & 1u;
        }
      }
      unsigned long long at = atomicAdd(cursors + (unsigned int)r * C3D_PART_BUCKETS + bucket, (unsigned long long)cnt);
      unsigned int todo = active;
      while (todo) {
        const int bit = __ffs((int)todo) - 1;
        todo &= todo - 1u;
        if (0u == ((owners[bit] >> r) & 1u)) continue;
        const unsigned int code = (unsigned int)((pl >> bit) & 1ull) | ((unsigned int)((ql >> bit) & 1ull) << 1)
                                | ((unsigned int)((pl >> (bit + 4)) & 1ull) << 2) | ((unsigned int)((ql >> (bit + 4)) & 1ull) << 3)
                                | ((unsigned int)((pu >> bit) & 1ull) << 4) | ((unsigned int)((qu >> bit) & 1ull) << 5)
                                | ((unsigned int)((pu >> (bit + 4)) & 1ull) << 6) | ((unsigned int)((qu >> (bit + 4)) & 1ull) << 7);
        if (at < capacity) outbox[at] = C3D_ENTRY(obj, x + (unsigned int)(bit & 3), j + (unsigned int)(bit >> 2), k, code) | flag | tag;
        at++;
      }
    }
  }
}

/* ---------------------------------------------------------------- launchers */

template <typename T, int NT, int WARPS, int SW_SLOTS>
static int launch_tiles(const TileSweepParams &p, const CUtensorMap &tmap, const int sm_count, cudaStream_t s) {
  const size_t smem = (size_t)WARPS * (SW_SLOTS * RowGeom<T>::SLOTB + SW_STAGE * sizeof(SweepRecord) + SW_SLOTS * 8);
  /* per device, so not cached in a static */
  const cudaError_t err = cudaFuncSetAttribute(k_sweep_tiles<T, NT, WARPS, SW_SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (cudaSuccess != err) return (int)err;
  const unsigned int need = (p.ntiles + WARPS - 1) / WARPS;
  const unsigned int blocks = need < (unsigned int)sm_count ? need : (unsigned int)sm_count;
  k_sweep_tiles<T, NT, WARPS, SW_SLOTS><<<blocks, WARPS * 32, smem, s>>>(p, tmap);
  return (int)cudaGetLastError();
}

/* cuTensorMapEncodeTiled through the runtime (libcuda is not a link-time dependency) */
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled(void) {
  static EncodeTiledFn fn = NULL;
  if (NULL == fn) {
    void *sym = NULL;
    cudaDriverEntryPointQueryResult found = cudaDriverEntryPointSymbolNotFound;
    if (cudaSuccess == cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &found) && cudaDriverEntryPointSuccess == found) {
      fn = (EncodeTiledFn)sym;
    }
  }
  return fn;
}

/* tensor map of the extended array: boxes of 8 rows x (128 + pad) points of one plane, zero fill outside */
static int make_tensor_map(const FieldView *f, CUtensorMap *map) {
  const EncodeTiledFn encode = encode_tiled();
  if (NULL == encode) return (int)cudaErrorNotSupported;
  const size_t eb = (size_t)f->elem_bytes;
  const cuuint64_t dims[3] = {(cuuint64_t)f->n[0], (cuuint64_t)f->n[1], (cuuint64_t)f->n[2]};
  const cuuint64_t strides[2] = {(cuuint64_t)f->sy * eb, (cuuint64_t)f->sz * eb};
  const cuuint32_t box[3] = {4 == eb ? RowGeom<float>::ROWP : RowGeom<double>::ROWP, SW_ROWS, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  static int promo = -1;
  if (promo < 0) {
    const char *env = getenv("CONTOUR3D_SWEEP_L2PROMO");
    promo = (NULL != env && '\0' != *env) ? atoi(env) : 2;
  }
  const CUtensorMapL2promotion l2 = 0 == promo ? CU_TENSOR_MAP_L2_PROMOTION_NONE : 1 == promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                  : 2 == promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  const CUresult rc = encode(map, 4 == eb ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3u, const_cast<void *>(f->data), dims,
                             strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return CUDA_SUCCESS == rc ? 0 : (int)cudaErrorInvalidValue;
}

extern "C" int c3d_k_sweep_tiles_ok(const FieldView *f) {
  /* what a tensor map asks of the array: 16-byte aligned base and strides, strides below 2^40 bytes */
  const size_t eb = (size_t)f->elem_bytes;
  return (0 == ((uintptr_t)f->data % 16)) && (0 == (f->sy * eb) % 16) && (0 == (f->sz * eb) % 16) && (f->sz * eb < (1ull << 40))
      && (f->sy >= f->n[0]) && (NULL != encode_tiled());
}

extern "C" int c3d_k_sweep_tiles(const FieldView *f, const SweepRange *r, int nthr, const double *thr, const float *thr_ge, const unsigned int *obj,
                                 const int flip[3], unsigned int *tile_cursor, void *records, unsigned long long *count, unsigned long long capacity,
                                 int sm_count, cudaStream_t s) {
  if (nthr < 1 || nthr > C3D_SWEEP_MAX_THR) return (int)cudaErrorInvalidValue;
  TileSweepParams p;
  p.mx = f->n[0]; p.my = f->n[1]; p.mz = f->n[2];
  p.nthr = nthr;
  for (int t = 0; t < C3D_SWEEP_MAX_THR; t++) {
    const int u = t < nthr ? t : 0;
    p.thr[t] = thr[u]; p.thr_ge[t] = thr_ge[u]; p.obj[t] = obj[u];
  }
  for (int d = 0; d < 3; d++) p.flip[d] = (NULL != flip && 0 != flip[d]) ? 1 : 0;
  p.ilo = r->lo[0]; p.ihi = r->hi[0];
  p.jlo = r->lo[1]; p.jhi = r->hi[1];
  p.klo = r->lo[2]; p.khi = r->hi[2];
  if (p.ihi < p.ilo || p.jhi < p.jlo || p.khi < p.klo) return 0;
  p.nxt = (int)((p.ihi + SW_XCUBES) / SW_XCUBES);
  p.nyt = (int)((p.jhi - p.jlo + SW_YCUBES) / SW_YCUBES);
  const long long nk = p.khi - p.klo + 1;
  /* planes per tile: enough tiles for ~6 per resident warp (dynamic hand-out evens the tail), at most 1/32 z re-reads */
  static int wide = -1;         /* experiment switch: 16 warps x 3 slots instead of 12 x 4 for 3-4 fused iso-values */
  if (wide < 0) {
    const char *env = getenv("CONTOUR3D_SWEEP_WARPS");
    wide = (NULL != env && 16 == atoi(env)) ? 1 : 0;
  }
  const int warps = 4 == f->elem_bytes ? ((nthr >= 3 && !wide) ? 12 : 16) : 8;
  const long long per_chunk = (long long)p.nxt * p.nyt;
  const long long want_tiles = 6ll * warps * sm_count;
  long long chunks = (want_tiles + per_chunk - 1) / per_chunk;
  long long zc = (nk + chunks - 1) / chunks;
  if (zc < 32) {
    /* thin arrays (a slab of an 8-rank run is 130 planes): 32-plane chunks would leave warps without a tile; accept up to
       1/8 z re-reads for at least ~3 tiles per resident warp */
    const long long chunks3 = (want_tiles / 2 + per_chunk - 1) / per_chunk;
    zc = (nk + chunks3 - 1) / chunks3;
    if (zc > 32) zc = 32;
    if (zc < 8) zc = 8;
  }
  if (zc > nk) zc = nk;
  /* Tiles are handed out dynamically, but one tile of zc = 32 planes is a fifth of a warp's whole share: the last ones
     finish unevenly.  The final ~10 % of the planes (in hand-out order) are therefore cut into chunks an eighth as deep
     (measured at 1024^3: 0.759 -> 0.722 ms, 87.6 -> 92.0 % of the HBM copy peak; 20 % / a quarter: 0.731 ms). */
  static int tail_pct = -1;
  if (tail_pct < 0) {
    const char *env = getenv("CONTOUR3D_SWEEP_TAIL");
    tail_pct = (NULL != env && '\0' != *env) ? atoi(env) : 10;
    if (tail_pct < 0 || tail_pct > 100) tail_pct = 10;
  }
  static int tail_div = -1;
  if (tail_div < 0) {
    const char *env = getenv("CONTOUR3D_SWEEP_TAILDIV");
    tail_div = (NULL != env && atoi(env) >= 1) ? atoi(env) : 8;
  }
  long long zcb = zc / tail_div < 4 ? 4 : zc / tail_div;
  if (zcb > zc) zcb = zc;
  long long nkb = (nk * tail_pct / 100 + zcb - 1) / zcb * zcb;
  if (nkb > nk) nkb = nk;
  if (zcb == zc) nkb = 0;
  p.zc = (int)zc;
  p.zc_b = (int)zcb;
  p.nk_a = nk - nkb;
  p.nzc_a = (int)((p.nk_a + zc - 1) / zc);
  p.nzc = p.nzc_a + (int)((nkb + zcb - 1) / zcb);
  p.ntiles = (unsigned int)(per_chunk * p.nzc);
  p.tile_cursor = tile_cursor;
  p.records = static_cast<SweepRecord *>(records);
  p.count = count;
  p.capacity = capacity;
  CUtensorMap tmap;
  const int rc = make_tensor_map(f, &tmap);
  if (0 != rc) return rc;
  const cudaError_t err = cudaMemsetAsync(tile_cursor, 0, sizeof(unsigned int), s);
  if (cudaSuccess != err) return (int)err;
  if (4 == f->elem_bytes) {
    switch (nthr) {
      case 1: return launch_tiles<float, 1, 16, 3>(p, tmap, sm_count, s);
      case 2: return launch_tiles<float, 2, 16, 3>(p, tmap, sm_count, s);
      case 3: return wide ? launch_tiles<float, 3, 16, 3>(p, tmap, sm_count, s) : launch_tiles<float, 3, 12, 4>(p, tmap, sm_count, s);
      default: return launch_tiles<float, 4, 12, 4>(p, tmap, sm_count, s);
    }
  }
  switch (nthr) {
    case 1: return launch_tiles<double, 1, 8, 3>(p, tmap, sm_count, s);
    case 2: return launch_tiles<double, 2, 8, 3>(p, tmap, sm_count, s);
    case 3: return launch_tiles<double, 3, 8, 3>(p, tmap, sm_count, s);
    default: return launch_tiles<double, 4, 8, 3>(p, tmap, sm_count, s);
  }
}

extern "C" int c3d_k_expand(const FrameParams *fp, int sorted, float near_depth, float far_depth,
                            const void *records, const unsigned long long *nrecords, unsigned long long rec_capacity, unsigned long long *list,
                            unsigned long long *count, unsigned long long capacity, unsigned int *histogram, unsigned long long *cursors,
                            unsigned long long *per_object, unsigned int nobjects, int sm_count, cudaStream_t s) {
  const SweepRecord *rec = static_cast<const SweepRecord *>(records);
  DepthOrder o;
  o.near_depth = near_depth;
  o.scale = (near_depth > far_depth) ? (float)C3D_DEPTH_BUCKETS / (near_depth - far_depth) : 0.f;
  o.last = (float)(C3D_DEPTH_BUCKETS - 1);
  if (sorted) {
    const cudaError_t err = cudaMemsetAsync(histogram, 0, C3D_DEPTH_BUCKETS * sizeof(unsigned int), s);
    if (cudaSuccess != err) return (int)err;
    k_bucket_count<<<sm_count * 4, 256, 0, s>>>(*fp, o, rec, nrecords, rec_capacity, histogram);
    k_bucket_scan<<<1, 1024, 0, s>>>(histogram, cursors, count);
    k_expand<true><<<sm_count * 8, 256, 0, s>>>(*fp, o, rec, nrecords, rec_capacity, list, count, capacity, cursors, per_object, nobjects);
  } else {
    k_expand<false><<<sm_count * 8, 256, 0, s>>>(*fp, o, rec, nrecords, rec_capacity, list, count, capacity, cursors, per_object, nobjects);
  }
  return (int)cudaGetLastError();
}

extern "C" int c3d_k_expand_part(const FrameParams *fp, float near_depth, float far_depth, const void *records, const unsigned long long *nrecords,
                                 unsigned long long rec_capacity, unsigned long long *outbox, unsigned long long *count, unsigned long long capacity,
                                 unsigned int *histogram, unsigned long long *cursors, unsigned long long *seg_start, unsigned char *masks,
                                 unsigned long long *per_object, unsigned int nobjects, int sm_count, cudaStream_t s) {
  const SweepRecord *rec = static_cast<const SweepRecord *>(records);
  const int nranks = fp->part.nranks;
  if (nranks < 2 || nranks > C3D_MAX_RANKS) return (int)cudaErrorInvalidValue;
  DepthOrder o;
  o.near_depth = near_depth;
  o.scale = (near_depth > far_depth) ? (float)C3D_PART_BUCKETS / (near_depth - far_depth) : 0.f;
  o.last = (float)(C3D_PART_BUCKETS - 1);
  const size_t bins = (size_t)nranks * C3D_PART_BUCKETS;
  cudaError_t err = cudaMemsetAsync(histogram, 0, bins * sizeof(unsigned int), s);
  if (cudaSuccess != err) return (int)err;
  err = cudaFuncSetAttribute(k_part_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bins * sizeof(unsigned int)));
  if (cudaSuccess != err) return (int)err;
  k_part_count<<<sm_count * 2, 256, bins * sizeof(unsigned int), s>>>(*fp, o, rec, nrecords, rec_capacity, masks, histogram);
  k_part_scan<<<1, 1024, 0, s>>>(histogram, nranks, cursors, seg_start, count);
  k_part_expand<<<sm_count * 8, 256, 0, s>>>(*fp, o, rec, nrecords, rec_capacity, masks, outbox, capacity, cursors, per_object, nobjects);
  return (int)cudaGetLastError();
}
