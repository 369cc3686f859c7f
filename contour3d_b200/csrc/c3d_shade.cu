/*
 * c3d_shade.cu -- deferred shading of the winning triangle of every pixel.
 *
 * Replaces contour3d_contour_compute_vertex_normals + average (contour/normal.c:6-165)
 * and the colour half of contour3d_contour_render_triangle (contour/render.c:137-159).
 * The reference smooths EVERY triangle of every slice; here a vertex normal is
 * computed only for vertices of triangles that own a pixel, and only ONCE per
 * vertex: a vertex lives on a cube edge (triangulate.c:15-32), its normal is the
 * normalised sum of `area * face_normal` over the triangles that touch that edge
 * in the 1, 2 or 4 cubes sharing it, visited in a fixed global order
 * (normal.c:78-158) -- a function of the GLOBAL edge alone.  At 1024^3 on a 4K
 * screen ~20 pixels reference the same vertex, so the pixels first agree on the
 * set of distinct vertices through a hash table keyed by the global edge.
 * Sums are evaluated in the reference's order: bit-identical, and independent of
 * any atomic ordering (the table only decides WHERE a normal is stored).
 *
 *   k_shade_select   one thread per pixel: background, lines and pixels lost to
 *                    another rank are written at once; pixels won by a contour
 *                    triangle are compacted into a work list
 *   k_shade_setup    one thread per listed pixel: rebuilds the winning triangle,
 *                    the barycentric weights of the pixel centre (render.c:102-110)
 *                    and enters its three vertices into the table
 *   k_shade_normals  eight lanes per distinct vertex: the face terms of the
 *                    touching triangles in parallel, added up in order, normalised
 *   k_shade_blend    one thread per listed pixel: blend, light, truncate
 *                    (render.c:137-159)
 *
 * File:line references are to NaokiHori/Contour3D @ c66854e.
 */
#include <float.h>
#include "c3d_device.cuh"


/* ------------------------------------------------------------------ select */

__global__ void __launch_bounds__(256) k_shade_select(const __grid_constant__ FrameParams p, const LineDev *__restrict__ lines,
                                                      const Pixel *__restrict__ canvas, const unsigned long long *__restrict__ depth_global,
                                                      unsigned int *__restrict__ rgba, const size_t npix, WorkItem *__restrict__ work,
                                                      unsigned int *__restrict__ nwork) {
  /* a warp takes a tile of 8 x 4 pixels, so that the work list is ordered by screen tiles: the 32 pixels a warp of
     k_shade_setup handles then share cubes, vertices and cache lines (rows of 3840 pixels do not) */
  const unsigned int lane = threadIdx.x & 31u;
  const unsigned int w = (unsigned int)p.cam.w, h = (unsigned int)p.cam.h;
  const unsigned int tiles_x = (w + 7u) >> 3, tiles_y = (h + 3u) >> 2;
  const unsigned int ntiles = tiles_x * tiles_y;
  const unsigned int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (unsigned int tile = warp; tile < ntiles; tile += nwarps) {
    const unsigned int pi = (tile % tiles_x) * 8u + (lane & 7u), pj = (tile / tiles_x) * 4u + (lane >> 3);
    const size_t pix = (size_t)pj * w + pi;
    bool mine = false;
    unsigned long long mykey = 0ull;
    if (pi < w && pj < h && pix < npix) {
      const ulonglong2 px = *reinterpret_cast<const ulonglong2 *>(canvas + pix);
      unsigned int out = 0u;     /* 0 = nothing from this rank: background, or another rank is nearer */
      mykey = px.y;
      if (0ull != px.y && (nullptr == depth_global || px.x == depth_global[pix])) {
        const unsigned int obj = (unsigned int)(px.y >> C3D_KEY_OBJ_SHIFT) - 1u;
        if (obj >= p.ncontours) {
          const LineDev &ln = lines[obj - p.ncontours];
          out = 0xFF000000u | ((unsigned int)ln.color[2] << 16) | ((unsigned int)ln.color[1] << 8) | ln.color[0];
        } else {
          mine = true;
        }
      }
      if (!mine) rgba[pix] = out;
    }
    const unsigned int who = __ballot_sync(C3D_FULL_MASK, mine);
    if (0u != who) {
      unsigned int at = 0u;
      if (0u == lane) at = atomicAdd(nwork, (unsigned int)__popc(who));
      at = __shfl_sync(C3D_FULL_MASK, at, 0);
      if (mine) work[at + __popc(who & ((1u << lane) - 1u))] = WorkItem{(unsigned int)pix, 0u, mykey};
    }
  }
}

/*
 * Screen-partitioned frames: this rank resolves only its own tiles.  Background and line pixels are final and go
 * straight into the writer's image (rank 0's, through peer memory); a pixel won by a contour triangle is queued for
 * the rank that holds the winning cube's array -- it alone can form the vertex normals (normal.c:78-158 needs the
 * neighbouring cubes) -- in the segment that rank will pull: seg[r * seg_cap ...], nseg[r].
 */
#define SEL_UNIT 4096u          /* pixels a block resolves at a time: a whole tile of 64 x 64, or a band of a larger one */
__global__ void __launch_bounds__(256) k_shade_select_part(const __grid_constant__ FrameParams p, const LineDev *__restrict__ lines,
                                                           const Pixel *__restrict__ canvas, unsigned int *__restrict__ rgba_out,
                                                           WorkItem *__restrict__ seg, const unsigned int seg_cap, unsigned int *__restrict__ nseg,
                                                           unsigned long long *__restrict__ overflowed) {
  /* One block per unit; the pixels are counted per destination rank in shared memory first, so that this rank's eight
     segment counters see one atomic per (unit, rank) instead of one per warp. */
  __shared__ unsigned char s_home[SEL_UNIT];
  __shared__ unsigned int s_count[C3D_MAX_RANKS], s_base[C3D_MAX_RANKS];
  const Partition &part = p.part;
  const unsigned int w = (unsigned int)p.cam.w, h = (unsigned int)p.cam.h;
  const unsigned int tile = 1u << part.tile_shift;
  const unsigned int unit_px = tile * tile < SEL_UNIT ? tile * tile : SEL_UNIT;
  const unsigned int rows = unit_px >> part.tile_shift;                 /* rows of the tile per unit */
  const unsigned int per_tile = (tile * tile) / unit_px;
  const unsigned int ntiles = (unsigned int)(part.tiles_x * part.tiles_y);
  for (unsigned long long unit = blockIdx.x; unit < (unsigned long long)ntiles * per_tile; unit += gridDim.x) {
    const unsigned int t = (unsigned int)(unit / per_tile), sub = (unsigned int)(unit % per_tile);
    const unsigned int tx = t % (unsigned int)part.tiles_x, ty = t / (unsigned int)part.tiles_x;
    if (tile_owner(part, tx, ty) != (unsigned int)part.me) continue;      /* block-uniform */
    if (threadIdx.x < C3D_MAX_RANKS) s_count[threadIdx.x] = 0u;
    __syncthreads();
    for (unsigned int q = threadIdx.x; q < unit_px; q += blockDim.x) {
      const unsigned int pi = tx * tile + (q & (tile - 1u)), pj = ty * tile + sub * rows + (q >> part.tile_shift);
      unsigned int home = 0xffu;
      if (pi < w && pj < h) {
        const size_t pix = (size_t)pj * w + pi;
        const ulonglong2 px = *reinterpret_cast<const ulonglong2 *>(canvas + pix);
        unsigned int out = 0u;
        if (0ull != px.y) {
          const unsigned int obj = (unsigned int)(px.y >> C3D_KEY_OBJ_SHIFT) - 1u;
          if (obj >= p.ncontours) {
            const LineDev &ln = lines[obj - p.ncontours];
            out = 0xFF000000u | ((unsigned int)ln.color[2] << 16) | ((unsigned int)ln.color[1] << 8) | ln.color[0];
          } else {
            const unsigned long long prim = px.y & C3D_KEY_PRIM_MASK;
            home = cube_home(part, (unsigned int)((prim >> 20) & 65535ull), (unsigned int)((prim >> 36) & 65535ull));
            atomicAdd(&s_count[home], 1u);
          }
        }
        if (0xffu == home) rgba_out[pix] = out;
      }
      s_home[q] = (unsigned char)home;
    }
    __syncthreads();
    if (threadIdx.x < (unsigned int)part.nranks) {
      const unsigned int cnt = s_count[threadIdx.x];
      s_base[threadIdx.x] = 0u != cnt ? atomicAdd(nseg + threadIdx.x, cnt) : 0u;
      s_count[threadIdx.x] = 0u;
    }
    __syncthreads();
    for (unsigned int q = threadIdx.x; q < unit_px; q += blockDim.x) {
      const unsigned int home = s_home[q];
      if (0xffu == home) continue;
      const unsigned int pi = tx * tile + (q & (tile - 1u)), pj = ty * tile + sub * rows + (q >> part.tile_shift);
      const size_t pix = (size_t)pj * w + pi;
      const unsigned int slot = s_base[home] + atomicAdd(&s_count[home], 1u);
      if (slot < seg_cap) seg[(size_t)home * seg_cap + slot] = WorkItem{(unsigned int)pix, home, canvas[pix].key};
      else atomicAdd(overflowed, 1ull);
    }
    __syncthreads();
  }
}

/* ----------------------------------------------------------------- geometry */

/* the c-th cube (reference visiting order) sharing edge e of cube (i, j, k), and the edge's id in that cube;
   the 19-way branch of normal.c:78-158 */
__device__ __forceinline__ bool edge_neighbour(const unsigned int e, const unsigned int c, int &di, int &dj, int &dk, unsigned int &local) {
  di = dj = dk = 0;
  if (e < 12u) {
    const unsigned int axis = e >> 2, a = e & 1u, b = (e >> 1) & 1u, pp = c & 1u, qq = c >> 1;
    const int da = (int)a - 1 + (int)pp, db = (int)b - 1 + (int)qq;
    local = 4u * axis + (1u - pp) + 2u * (1u - qq);
    if (0u == axis) { dj = da; dk = db; }
    else if (1u == axis) { di = da; dk = db; }
    else { di = da; dj = db; }
    return true;
  }
  if (e < 18u) {
    if (c >= 2u) return false;
    const unsigned int kind = (e - 12u) >> 1, a = e & 1u;
    const int da = (int)a - 1 + (int)c;
    local = 12u + 2u * kind + (1u - c);
    if (0u == kind) di = da;
    else if (1u == kind) dj = da;
    else dk = da;
    return true;
  }
  local = 18u;
  return 0u == c;
}

__device__ __forceinline__ unsigned int cube_code(const FieldView &f, const long long i, const long long j, const long long k) {
  const long long base = k * f.sz + j * f.sy + i;
  unsigned int code = 0u;
#pragma unroll
  for (int n = 0; n < 8; n++) {
    code |= (unsigned int)(f.thr < ld_value(f, base + (n & 1) + ((n >> 1) & 1) * f.sy + (n >> 2) * f.sz)) << n;
  }
  return code;
}

/* area * face_normal of triangle s of a cube (triangulate.c:75-93, normal.c:39-40).  The same arithmetic as
   cut_edge_global() + face_normal() (c3d_device.cuh), with the loads hoisted: the cube's six coordinates once instead of
   per vertex, element type decided once. */
template <typename T>
__device__ __forceinline__ V3 face_term_of(const FieldView &f, const unsigned int pairs, const long long i, const long long j, const long long k) {
  const T *__restrict__ at = static_cast<const T *>(f.data) + (k * f.sz + j * f.sy + i);
  const double x0 = __ldg(f.g[0] + i), x1 = __ldg(f.g[0] + i + 1);
  const double y0 = __ldg(f.g[1] + j), y1 = __ldg(f.g[1] + j + 1);
  const double z0 = __ldg(f.g[2] + k), z1 = __ldg(f.g[2] + k + 1);
  V3 v[3];
#pragma unroll
  for (int m = 0; m < 3; m++) {
    const unsigned int tail = (pairs >> (6 * m)) & 7u, head = (pairs >> (6 * m + 3)) & 7u;
    const double v0 = (double)__ldg(at + ((tail & 1u) + ((tail & 2u) ? f.sy : 0ll) + ((tail & 4u) ? f.sz : 0ll)));
    const double v1 = (double)__ldg(at + ((head & 1u) + ((head & 2u) ? f.sy : 0ll) + ((head & 4u) ? f.sz : 0ll)));
    const V3 p0 = V3{(tail & 1u) ? x1 : x0, (tail & 2u) ? y1 : y0, (tail & 4u) ? z1 : z0};
    /* triangulate.c:40-73 */
    const double dv = r_sub(v1, v0);
    if (fabs(dv) < 1.e-8) {
      v[m] = p0;
    } else {
      const double factor = r_div(r_sub(f.thr, v0), dv);
      const V3 p1 = V3{(head & 1u) ? x1 : x0, (head & 2u) ? y1 : y0, (head & 4u) ? z1 : z0};
      v[m] = v_add(p0, v_mul(factor, v_sub(p1, p0)));
    }
  }
  V3 fn;
  double area;
  face_normal(v, fn, area);
  return v_mul(area, fn);
}

/* the same with the cube's eight corner values already at hand (k_shade_normals keeps what the corner-bit test of a slot
   has loaded: 6 shared-memory reads instead of 6 global ones with their 64-bit address arithmetic); identity converter only */
__device__ __forceinline__ V3 face_term_held(const FieldView &f, const float *__restrict__ val, const unsigned int pairs, const long long i,
                                             const long long j, const long long k) {
  const double x0 = __ldg(f.g[0] + i), x1 = __ldg(f.g[0] + i + 1);
  const double y0 = __ldg(f.g[1] + j), y1 = __ldg(f.g[1] + j + 1);
  const double z0 = __ldg(f.g[2] + k), z1 = __ldg(f.g[2] + k + 1);
  V3 v[3];
#pragma unroll
  for (int m = 0; m < 3; m++) {
    const unsigned int tail = (pairs >> (6 * m)) & 7u, head = (pairs >> (6 * m + 3)) & 7u;
    const double v0 = (double)val[tail], v1 = (double)val[head];
    const V3 p0 = V3{(tail & 1u) ? x1 : x0, (tail & 2u) ? y1 : y0, (tail & 4u) ? z1 : z0};
    const double dv = r_sub(v1, v0);
    if (fabs(dv) < 1.e-8) {
      v[m] = p0;
    } else {
      const double factor = r_div(r_sub(f.thr, v0), dv);
      const V3 p1 = V3{(head & 1u) ? x1 : x0, (head & 2u) ? y1 : y0, (head & 4u) ? z1 : z0};
      v[m] = v_add(p0, v_mul(factor, v_sub(p1, p0)));
    }
  }
  V3 fn;
  double area;
  face_normal(v, fn, area);
  return v_mul(area, fn);
}

/* object `obj` as rank `owner` holds it: this rank's own array, or (a partitioned frame's shading helper) the owner's
   through peer memory */
template <bool HELP>
__device__ __forceinline__ const FieldView &owner_view(const FrameParams &p, const unsigned int owner, const unsigned int obj) {
  if (HELP) return p.pviews[C3D_PEER_VIEW(owner, obj)];
  return p.view[obj];
}

/*
 * The shading kernels run over a RANGE of the work list / vertex list, read from the counters on the device:
 *   counters[0] work items, [1] vertices listed so far, [2] work items of the first pass, [3] vertices listed by the first pass.
 * HELP = false: items [0, counters[2]) -- pixels won by cubes of this rank's own array (all of them on one rank);
 * HELP = true:  items [counters[2], counters[0]) -- pixels another rank has handed over; their cubes are read from the
 *               owner's array through peer memory (a separate instantiation: the common pass keeps its views in the constant bank).
 */
template <bool HELP> __device__ __forceinline__ void item_range(const unsigned int *counters, unsigned int &begin, unsigned int &end) {
  begin = HELP ? counters[2] : 0u;
  end = HELP ? counters[0] : counters[2];
}
template <bool HELP> __device__ __forceinline__ void vertex_range(const unsigned int *counters, unsigned int &begin, unsigned int &end) {
  begin = HELP ? counters[3] : 0u;
  end = counters[1];
}

__device__ __forceinline__ V3 face_term(const FrameParams &p, const FieldView &f, const long long i, const long long j, const long long k,
                                        const unsigned int code, const unsigned int s) {
  const unsigned int pairs = __ldg(p.tab.pairs + code * C3D_MAX_TRIS + s);
  if (nullptr != f.map) {      /* tabulated converter: corner positions come from the tables, one vertex at a time */
    V3 v[3], fn;
    double area;
#pragma unroll
    for (int m = 0; m < 3; m++) v[m] = cut_edge_global(f, i, j, k, f.thr, (pairs >> (6 * m)) & 7u, (pairs >> (6 * m + 3)) & 7u);
    face_normal(v, fn, area);
    return v_mul(area, fn);
  }
  return 4 == f.elem_bytes ? face_term_of<float>(f, pairs, i, j, k) : face_term_of<double>(f, pairs, i, j, k);
}

/* ------------------------------------------------------------------- setup */

struct ShadePix {              /* 56 bytes per listed pixel */
  double w[3];                 /* barycentric weights of the pixel centre (render.c:102-110) */
  unsigned int vslot[3];       /* table slots of the three vertices; C3D_NO_SLOT: table full, normal computed on the spot */
  unsigned int pix;
  unsigned short i, j, k;      /* cube of the winning triangle (indices on the owner's array) */
  unsigned char obj, edges[3]; /* cube-edge ids of its three vertices */
  unsigned char owner;         /* rank whose array the cube indices refer to (partitioned frames; else 0) */
};
static_assert(sizeof(ShadePix) == C3D_SHADE_PIX_BYTES, "ShadePix layout");

#define C3D_NO_SLOT 0xffffffffu
#define SH_PROBES 32
#define SH_CHUNK 256u                     /* vertex-list entries a warp reserves at a time (a multiple of 32, >= 64) */

/* the global edge a cube-local edge id names: the cube whose edge 0 / 4 / 8 / 12 / 14 / 16 / 18 it is (the numbering of
   triangulate.c:15-32: 0-3 x edges at (dj, dk), 4-7 y edges at (di, dk), 8-11 z edges at (di, dj), 12/13 yz-face
   diagonals at di, 14/15 xz at dj, 16/17 xy at dk, 18 the body diagonal), packed as
   1:1 | owner:3 | object:4 | type:3 | k:16 | j:16 | i:16 (never 0 = empty slot); owner = the rank on whose array (i, j, k) count */
__device__ __forceinline__ unsigned long long vertex_key(const unsigned int owner, const unsigned int obj, const unsigned int i, const unsigned int j,
                                                         const unsigned int k, const unsigned int e) {
  unsigned int ci = i, cj = j, ck = k, type;
  if (e < 12u) {
    const unsigned int axis = e >> 2, a = e & 1u, b = (e >> 1) & 1u;
    type = axis;
    if (0u == axis) { cj += a; ck += b; }
    else if (1u == axis) { ci += a; ck += b; }
    else { ci += a; cj += b; }
  } else if (e < 18u) {
    const unsigned int kind = (e - 12u) >> 1, a = e & 1u;
    type = 3u + kind;
    if (0u == kind) ci += a;
    else if (1u == kind) cj += a;
    else ck += a;
  } else {
    type = 6u;
  }
  return (1ull << 63) | ((unsigned long long)owner << 55) | ((unsigned long long)obj << 51) | ((unsigned long long)type << 48) | ((unsigned long long)ck << 32)
       | ((unsigned long long)cj << 16) | (unsigned long long)ci;
}

template <bool HELP>
__global__ void __launch_bounds__(128) k_shade_setup(const __grid_constant__ FrameParams p,
                                                     const WorkItem *__restrict__ work, unsigned int *__restrict__ counters,
                                                     ShadePix *__restrict__ out, unsigned long long *__restrict__ table, const unsigned int mask,
                                                     unsigned int *__restrict__ vlist, unsigned long long *__restrict__ overflowed) {
  const unsigned int lane = threadIdx.x & 31u;
  unsigned int first, n;
  item_range<HELP>(counters, first, n);
  /* The vertex list is handed out in chunks of SH_CHUNK entries per warp: one counter for everybody costs a serialised
     round trip per append (measured: 58 % of this kernel's stall samples).  Unused entries of a chunk hold C3D_NO_SLOT. */
  unsigned int chunk_next = 0u, chunk_end = 0u;
  /* whole warps iterate together (ballots inside) */
  for (unsigned int base = first + ((blockIdx.x * blockDim.x + threadIdx.x) & ~31u); base < n; base += gridDim.x * blockDim.x) {
    const unsigned int idx = base + lane;
    const bool valid = idx < n;
    ShadePix me;
    me.w[0] = me.w[1] = me.w[2] = 0.;
    me.vslot[0] = me.vslot[1] = me.vslot[2] = C3D_NO_SLOT;
    me.pix = 0u; me.i = me.j = me.k = 0; me.obj = 0; me.edges[0] = me.edges[1] = me.edges[2] = 0; me.owner = 0;
    unsigned long long vkey[3] = {0ull, 0ull, 0ull};
    if (valid) {
      const uint4 item = __ldg(reinterpret_cast<const uint4 *>(work) + idx);
      const unsigned int pix = item.x;
      const unsigned long long key = (unsigned long long)item.z | ((unsigned long long)item.w << 32);
      const unsigned int obj = (unsigned int)(key >> C3D_KEY_OBJ_SHIFT) - 1u;
      const unsigned long long prim = key & C3D_KEY_PRIM_MASK;
      const unsigned int s = (unsigned int)(prim & 15ull);
      const unsigned int owner = item.y;
      const FieldView &f = owner_view<HELP>(p, owner, obj);
      /* the key names the cube by its GLOBAL indices; the owner's array starts at f.off */
      const long long i = (long long)((prim >> 4) & 65535ull) - f.off[0], j = (long long)((prim >> 20) & 65535ull) - f.off[1],
                      k = (long long)((prim >> 36) & 65535ull) - f.off[2];
      const unsigned int code = cube_code(f, i, j, k);
      const unsigned int pairs = __ldg(p.tab.pairs + code * C3D_MAX_TRIS + s);
      const unsigned int edges = __ldg(p.tab.edges + code * C3D_MAX_TRIS + s);
      Tri2D t;
      bool ok = true;
#pragma unroll
      for (int m = 0; m < 3; m++) {
        const ProjectedVertex v = exact_vertex_t<true>(p.cam, f, i, j, k, (pairs >> (6 * m)) & 7u, (pairs >> (6 * m + 3)) & 7u);
        ok = ok && (0. != v.w);
        t.q[m] = V3{v.x, v.y, v.z};
      }
      if (ok) {
        setup_projected(p.cam, t);
        pixel_weights(t, __ldg(p.px + (pix % (unsigned int)p.cam.w)), __ldg(p.py + (pix / (unsigned int)p.cam.w)), me.w);
      }
      me.pix = pix;
      me.i = (unsigned short)i; me.j = (unsigned short)j; me.k = (unsigned short)k;
      me.obj = (unsigned char)obj;
      me.owner = (unsigned char)owner;
#pragma unroll
      for (int m = 0; m < 3; m++) {
        me.edges[m] = (unsigned char)((edges >> (5 * m)) & 31u);
        vkey[m] = vertex_key(owner, obj, (unsigned int)i, (unsigned int)j, (unsigned int)k, me.edges[m]);
      }
    }
    /* the three vertices probe the table side by side: an atomic round trip costs about a microsecond, and a warp waits
       for its slowest lane, so the probe sequences must overlap rather than add up */
    unsigned int h[3];
    bool done[3], inserted[3];
#pragma unroll
    for (int m = 0; m < 3; m++) {
      h[m] = (unsigned int)((vkey[m] * 0x9E3779B97F4A7C15ull) >> 32) & mask;
      done[m] = !valid;
      inserted[m] = false;
    }
    for (int probe = 0; probe < SH_PROBES; probe++) {
      unsigned long long old[3] = {0ull, 0ull, 0ull};
#pragma unroll
      for (int m = 0; m < 3; m++) {
        if (!done[m]) old[m] = atomicCAS(table + h[m], 0ull, vkey[m]);
      }
#pragma unroll
      for (int m = 0; m < 3; m++) {
        if (!done[m]) {
          if (0ull == old[m]) { inserted[m] = true; me.vslot[m] = h[m]; done[m] = true; }
          else if (old[m] == vkey[m]) { me.vslot[m] = h[m]; done[m] = true; }
          else h[m] = (h[m] + 1u) & mask;
        }
      }
      if (done[0] && done[1] && done[2]) break;
    }
#pragma unroll
    for (int m = 0; m < 3; m++) {
      if (valid && C3D_NO_SLOT == me.vslot[m]) atomicAdd(overflowed, 1ull);
      /* the first pixel to name a vertex lists it for k_shade_normals */
      const unsigned int who = __ballot_sync(C3D_FULL_MASK, inserted[m]);
      if (0u != who) {
        const unsigned int cnt = (unsigned int)__popc(who);
        if (chunk_next + cnt > chunk_end) {                     /* warp-uniform */
          if (chunk_next + lane < chunk_end) vlist[chunk_next + lane] = C3D_NO_SLOT;      /* at most 31 entries are left */
          unsigned int at = 0u;
          if (0u == lane) at = atomicAdd(counters + 1, (unsigned int)SH_CHUNK);
          chunk_next = __shfl_sync(C3D_FULL_MASK, at, 0);
          chunk_end = chunk_next + SH_CHUNK;
        }
        if (inserted[m]) vlist[chunk_next + __popc(who & ((1u << lane) - 1u))] = me.vslot[m];
        chunk_next += cnt;
      }
    }
    if (valid) out[idx] = me;
  }
  for (unsigned int at = chunk_next + lane; at < chunk_end; at += 32u) vlist[at] = C3D_NO_SLOT;
}

/* ----------------------------------------------------------------- normals */

#define SH_THREADS 128
#define SH_VERTS 96                        /* vertices per batch */
#define SH_SLOTS (SH_VERTS * 4)            /* up to 4 cubes share an edge */
#define SH_ITEMS 1024                      /* face terms kept in shared memory per batch */

struct ShadeSlot {             /* one cube sharing the edge of one vertex */
  unsigned short i, j, k;
  unsigned short touch;        /* triangles of that cube with a vertex on the shared edge (bit s) */
  unsigned char obj, code, owner;
};

/*
 * Cooperative: a block takes SH_VERTS listed vertices; it enumerates the (vertex, sharing cube) slots, expands them into
 * one item per touching triangle, computes the items' face terms with ALL threads (the expensive part: three edge cuts,
 * a cross product, a square root and four divisions in fp64 each), and adds them up per vertex in the reference's order.
 */
/* the normal of listed vertex v goes to vn[v]; its table slot, whose key nobody needs any more, is overwritten with v so
   that the pixels (which hold slots) find it.  Vertices beyond the array (v >= vn_cap) are marked C3D_NO_SLOT and
   computed per pixel. */
__device__ __forceinline__ void store_normal(const V3 unit, const unsigned int v, const unsigned int slot, double *__restrict__ vn,
                                             const unsigned int vn_cap, unsigned long long *__restrict__ table) {
  if (v < vn_cap) {
    vn[3ull * v + 0ull] = unit.x; vn[3ull * v + 1ull] = unit.y; vn[3ull * v + 2ull] = unit.z;
    table[slot] = (unsigned long long)v;
  } else {
    table[slot] = (unsigned long long)C3D_NO_SLOT;
  }
}

template <bool HELP>
__global__ void __launch_bounds__(SH_THREADS) k_shade_normals(const __grid_constant__ FrameParams p, unsigned long long *__restrict__ table,
                                                              const unsigned int *__restrict__ vlist, const unsigned int *__restrict__ counters,
                                                              double *__restrict__ vn, const unsigned int vn_cap, unsigned long long *__restrict__ overflowed) {
  __shared__ ShadeSlot s_slot[SH_SLOTS];
  __shared__ unsigned int s_first[SH_SLOTS + 1];          /* first item of each slot */
  __shared__ float s_val[SH_SLOTS][8];                    /* fp32 arrays: the corner values the slot's corner-bit test has loaded */
  __shared__ unsigned short s_item[SH_ITEMS];             /* slot << 4 | triangle */
  __shared__ double s_term[SH_ITEMS][3];
  __shared__ unsigned int s_warp_total[SH_THREADS / 32];
  const unsigned int tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
  unsigned int v0, nv;
  vertex_range<HELP>(counters, v0, nv);
  for (unsigned int base = v0 + blockIdx.x * SH_VERTS; base < nv; base += gridDim.x * SH_VERTS) {
    /* ---- slots = (vertex, sharing cube) in the reference's visiting order; count the triangles touching the edge ---- */
    unsigned int mycount[SH_SLOTS / SH_THREADS];
#pragma unroll
    for (unsigned int r = 0; r < SH_SLOTS / SH_THREADS; r++) {
      const unsigned int q = r * SH_THREADS + tid;
      const unsigned int v = base + (q >> 2), c = q & 3u;
      ShadeSlot slot;
      slot.i = slot.j = slot.k = 0; slot.touch = 0; slot.obj = 0; slot.code = 0; slot.owner = 0;
      const unsigned int at = v < nv ? __ldg(vlist + v) : C3D_NO_SLOT;
      if (C3D_NO_SLOT != at) {
        const unsigned long long key = table[at];
        const unsigned int type = (unsigned int)(key >> 48) & 7u, obj = (unsigned int)(key >> 51) & 15u, owner = (unsigned int)(key >> 55) & 7u;
        const unsigned int e = type < 3u ? 4u * type : (type < 6u ? 6u + 2u * type : 18u);   /* edge 0 / 4 / 8 / 12 / 14 / 16 / 18 of that cube */
        int di, dj, dk;
        unsigned int local;
        slot.obj = (unsigned char)obj;
        slot.owner = (unsigned char)owner;
        if (edge_neighbour(e, c, di, dj, dk, local)) {
          const long long ii = (long long)(key & 65535ull) + di, jj = (long long)((key >> 16) & 65535ull) + dj, kk = (long long)((key >> 32) & 65535ull) + dk;
          const FieldView &f = owner_view<HELP>(p, owner, obj);
          unsigned int code;
          if (4 == f.elem_bytes) {
            /* the eight values once: corner bits now, cut edges of the touching triangles later */
            const float *at8 = static_cast<const float *>(f.data) + (kk * f.sz + jj * f.sy + ii);
            float c8[8];
#pragma unroll
            for (int n = 0; n < 8; n++) c8[n] = __ldg(at8 + ((n & 1) + ((n >> 1) & 1) * f.sy + (n >> 2) * f.sz));
            code = 0u;
#pragma unroll
            for (int n = 0; n < 8; n++) {
              code |= (unsigned int)(f.thr < (double)c8[n]) << n;
              s_val[q][n] = c8[n];
            }
          } else {
            code = cube_code(f, ii, jj, kk);
          }
          slot.i = (unsigned short)ii; slot.j = (unsigned short)jj; slot.k = (unsigned short)kk;
          slot.code = (unsigned char)code;
          slot.touch = __ldg(p.tab.touch + code * 19u + local);
        }
      }
      s_slot[q] = slot;
      mycount[r] = __popc((unsigned int)slot.touch);
    }
    /* ---- exclusive scan of the slot counts in slot order (q = r * 128 + tid) ---- */
    unsigned int total_items = 0u;
#pragma unroll
    for (unsigned int r = 0; r < SH_SLOTS / SH_THREADS; r++) {
      unsigned int incl = mycount[r];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned int up = __shfl_up_sync(C3D_FULL_MASK, incl, d);
        if (lane >= (unsigned int)d) incl += up;
      }
      if (31u == lane) s_warp_total[wid] = incl;
      __syncthreads();
      unsigned int before = total_items;
      for (unsigned int w = 0; w < wid; w++) before += s_warp_total[w];
      unsigned int round_total = 0u;
      for (unsigned int w = 0; w < SH_THREADS / 32; w++) round_total += s_warp_total[w];
      s_first[r * SH_THREADS + tid] = before + incl - mycount[r];
      total_items += round_total;
      __syncthreads();
    }
    if (0u == tid) s_first[SH_SLOTS] = total_items;
    const bool fits = total_items <= SH_ITEMS;                   /* uniform */
    if (fits) {
      /* ---- one item per touching triangle, in (slot, triangle) order ---- */
#pragma unroll
      for (unsigned int r = 0; r < SH_SLOTS / SH_THREADS; r++) {
        const unsigned int q = r * SH_THREADS + tid;
        unsigned int touch = s_slot[q].touch, at = s_first[q];
        while (touch) {
          const unsigned int s = (unsigned int)__ffs((int)touch) - 1u;
          touch &= touch - 1u;
          s_item[at++] = (unsigned short)((q << 4) | s);
        }
      }
      __syncthreads();
      /* ---- the face terms, all threads ---- */
      for (unsigned int it = tid; it < total_items; it += SH_THREADS) {
        const unsigned int q = s_item[it] >> 4, s = s_item[it] & 15u;
        const ShadeSlot slot = s_slot[q];
        const FieldView &f = owner_view<HELP>(p, slot.owner, slot.obj);
        const V3 term = (4 == f.elem_bytes && nullptr == f.map)
                            ? face_term_held(f, s_val[q], __ldg(p.tab.pairs + slot.code * C3D_MAX_TRIS + s), slot.i, slot.j, slot.k)
                            : face_term(p, f, slot.i, slot.j, slot.k, slot.code, s);
        s_term[it][0] = term.x; s_term[it][1] = term.y; s_term[it][2] = term.z;
      }
      __syncthreads();
      /* ---- per vertex, add the terms in order and normalise (average(), normal.c:6-43, 159) ---- */
      if (tid < SH_VERTS && base + tid < nv && C3D_NO_SLOT != __ldg(vlist + base + tid)) {
        const unsigned int q0 = tid * 4u;
        V3 sum = V3{0., 0., 0.};
        for (unsigned int it = s_first[q0]; it < s_first[q0 + 4u]; it++) {
          sum = v_add(sum, V3{s_term[it][0], s_term[it][1], s_term[it][2]});
        }
        store_normal(v_unit(sum), base + tid, __ldg(vlist + base + tid), vn, vn_cap, table);
      }
    } else {
      /* more touching triangles than the shared buffer holds: the same sums, one thread per vertex */
      if (0u == tid) atomicAdd(overflowed, 1ull);
      if (tid < SH_VERTS && base + tid < nv && C3D_NO_SLOT != __ldg(vlist + base + tid)) {
        V3 sum = V3{0., 0., 0.};
        for (unsigned int q = tid * 4u; q < tid * 4u + 4u; q++) {
          const ShadeSlot slot = s_slot[q];
          unsigned int touch = slot.touch;
          while (touch) {
            const unsigned int s = (unsigned int)__ffs((int)touch) - 1u;
            touch &= touch - 1u;
            sum = v_add(sum, face_term(p, owner_view<HELP>(p, slot.owner, slot.obj), slot.i, slot.j, slot.k, slot.code, s));
          }
        }
        store_normal(v_unit(sum), base + tid, __ldg(vlist + base + tid), vn, vn_cap, table);
      }
    }
    __syncthreads();
  }
}

/* ------------------------------------------------------------------- blend */

template <bool HELP>
__global__ void __launch_bounds__(128) k_shade_blend(const __grid_constant__ FrameParams p, const ShadePix *__restrict__ in,
                                                     const unsigned int *__restrict__ counters, const unsigned long long *__restrict__ table,
                                                     const double *__restrict__ vn, unsigned int *__restrict__ rgba) {
  unsigned int first, n;
  item_range<HELP>(counters, first, n);
  for (unsigned int idx = first + blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
    const ShadePix sp = in[idx];
    const FieldView &f = owner_view<HELP>(p, sp.owner, sp.obj);
    V3 nv[3];
#pragma unroll
    for (int m = 0; m < 3; m++) {
      const unsigned int slot = sp.vslot[m];
      const unsigned int v = C3D_NO_SLOT != slot ? (unsigned int)table[slot] : C3D_NO_SLOT;
      if (C3D_NO_SLOT != v) nv[m] = V3{vn[3ull * v + 0ull], vn[3ull * v + 1ull], vn[3ull * v + 2ull]};
      else nv[m] = vertex_normal(f, p.tab, (long long)sp.i, (long long)sp.j, (long long)sp.k, (unsigned int)sp.edges[m]);
    }
    /* w0*n0 + (w1*n1 + w2*n2), render.c:137-146 */
    const V3 nrm = v_unit(v_add(v_mul(sp.w[0], nv[0]), v_add(v_mul(sp.w[1], nv[1]), v_mul(sp.w[2], nv[2]))));
    const double factor = fmax(0.15, fabs(v_dot(nrm, v_of(p.cam.light))));
    const unsigned int r = (unsigned int)__double2int_rz(r_mul(factor, (double)f.color[0])) & 255u;
    const unsigned int g = (unsigned int)__double2int_rz(r_mul(factor, (double)f.color[1])) & 255u;
    const unsigned int b = (unsigned int)__double2int_rz(r_mul(factor, (double)f.color[2])) & 255u;
    rgba[sp.pix] = 0xFF000000u | (b << 16) | (g << 8) | r;
  }
}

/* resident blocks of a kernel on this device (grid-stride kernels are launched with exactly that many) */
template <typename K> static int resident_blocks(K kernel, const int threads, const int sm_count) {
  int per_sm = 0;
  if (cudaSuccess != cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) || per_sm < 1) per_sm = 4;
  return per_sm * sm_count;
}

/* clears the counters and the vertex table of a frame's shading */
extern "C" int c3d_k_shade_begin(const ShadeScratch *sc, cudaStream_t s) {
  cudaError_t err = cudaMemsetAsync(sc->counters, 0, 4 * sizeof(unsigned int), s);
  if (cudaSuccess != err) return (int)err;
  err = cudaMemsetAsync(sc->table, 0, ((size_t)sc->table_mask + 1u) * sizeof(unsigned long long), s);
  return (int)err;
}

__global__ void k_shade_mark(unsigned int *counters, const int after_select) {
  if (after_select) counters[2] = counters[0];        /* one pass over everything */
  else counters[3] = counters[1];                     /* the vertices of the first pass end here */
}

/* shades the work items of sc->work into rgba (which may be another rank's image, through peer memory).  helpers != 0:
   counters[2] (set by the work plan) separates this rank's own pixels from those handed over by other ranks, which take
   a second pass that reads the owners' arrays through p->pviews. */
extern "C" int c3d_k_shade_finish(const FrameParams *p, unsigned int *rgba, const ShadeScratch *sc, unsigned long long *overflowed,
                                  int helpers, int sm_count, cudaStream_t s) {
  if (!helpers) k_shade_mark<<<1, 1, 0, s>>>(sc->counters, 1);
  k_shade_setup<false><<<resident_blocks(k_shade_setup<false>, 128, sm_count), 128, 0, s>>>(*p, sc->work, sc->counters, static_cast<ShadePix *>(sc->pix),
                                                                                            sc->table, sc->table_mask, sc->vlist, overflowed);
  k_shade_normals<false><<<resident_blocks(k_shade_normals<false>, SH_THREADS, sm_count), SH_THREADS, 0, s>>>(*p, sc->table, sc->vlist, sc->counters, sc->vn, sc->vn_cap, overflowed);
  k_shade_blend<false><<<resident_blocks(k_shade_blend<false>, 128, sm_count), 128, 0, s>>>(*p, static_cast<const ShadePix *>(sc->pix), sc->counters, sc->table, sc->vn, rgba);
  if (helpers) {
    k_shade_mark<<<1, 1, 0, s>>>(sc->counters, 0);
    k_shade_setup<true><<<resident_blocks(k_shade_setup<true>, 128, sm_count), 128, 0, s>>>(*p, sc->work, sc->counters, static_cast<ShadePix *>(sc->pix),
                                                                                            sc->table, sc->table_mask, sc->vlist, overflowed);
    k_shade_normals<true><<<resident_blocks(k_shade_normals<true>, SH_THREADS, sm_count), SH_THREADS, 0, s>>>(*p, sc->table, sc->vlist, sc->counters, sc->vn, sc->vn_cap, overflowed);
    k_shade_blend<true><<<resident_blocks(k_shade_blend<true>, 128, sm_count), 128, 0, s>>>(*p, static_cast<const ShadePix *>(sc->pix), sc->counters, sc->table, sc->vn, rgba);
  }
  return (int)cudaGetLastError();
}

extern "C" int c3d_k_shade(const FrameParams *p, const LineDev *lines_dev, const Pixel *canvas, const unsigned long long *depth_global,
                           unsigned int *rgba, size_t npix, const ShadeScratch *sc, unsigned long long *overflowed,
                           int sm_count, cudaStream_t s) {
  const int rc = c3d_k_shade_begin(sc, s);
  if (0 != rc) return rc;
  k_shade_select<<<sm_count * 8, 256, 0, s>>>(*p, lines_dev, canvas, depth_global, rgba, npix, sc->work, sc->counters);
  return c3d_k_shade_finish(p, rgba, sc, overflowed, 0, sm_count, s);
}

extern "C" int c3d_k_shade_select_part(const FrameParams *p, const LineDev *lines_dev, const Pixel *canvas, unsigned int *rgba_out,
                                       WorkItem *seg, unsigned int seg_cap, unsigned int *nseg, unsigned long long *overflowed,
                                       int sm_count, cudaStream_t s) {
  k_shade_select_part<<<sm_count * 8, 256, 0, s>>>(*p, lines_dev, canvas, rgba_out, seg, seg_cap, nseg, overflowed);
  return (int)cudaGetLastError();
}
