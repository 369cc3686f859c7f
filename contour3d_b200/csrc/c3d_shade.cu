/*
 * c3d_shade.cu -- deferred shading of the winning triangle of every pixel.
 *
 * Replaces contour3d_contour_compute_vertex_normals + average (contour/normal.c:6-165)
 * and the colour half of contour3d_contour_render_triangle (contour/render.c:137-159).
 * The reference smooths EVERY triangle of every slice; here vertex normals are
 * computed only for the three vertices of the one triangle that owns a pixel,
 * by gathering `area * face_normal` over the triangles that touch the same cube
 * edge in the 1, 2 or 4 cubes that share it, in the reference's fixed visiting
 * order (normal.c:78-158) -- so the sums are bit-identical and do not depend
 * on any atomic ordering.
 *
 *   k_shade_select   one thread per pixel: background, lines and pixels lost to
 *                    another rank are written at once; pixels won by a contour
 *                    triangle are compacted into a work list
 *   k_shade_smooth   cooperative: a block takes 128 listed pixels; every thread
 *                    rebuilds its pixel's triangle (barycentric weights), then in
 *                    sub-batches of 32 pixels the block enumerates the
 *                    (pixel, vertex, neighbour cube) slots, expands them into one
 *                    item per touching triangle, computes the items' face terms
 *                    with ALL threads (the expensive, previously divergent part),
 *                    adds them up per vertex in order, and shades
 *
 * File:line references are to NaokiHori/Contour3D @ c66854e.
 */
#include <float.h>
#include "c3d_device.cuh"

#define SH_THREADS 128
#define SH_BATCH 32                        /* pixels per sub-batch */
#define SH_SLOTS (SH_BATCH * 12)           /* 3 vertices x up to 4 cubes */
#define SH_ITEMS 1024                      /* face terms kept in shared memory per sub-batch */

/* ------------------------------------------------------------------ select */

__global__ void __launch_bounds__(256) k_shade_select(const __grid_constant__ FrameParams p, const LineDev *__restrict__ lines,
                                                      const Pixel *__restrict__ canvas, const unsigned long long *__restrict__ depth_global,
                                                      unsigned int *__restrict__ rgba, const size_t npix, unsigned int *__restrict__ work,
                                                      unsigned int *__restrict__ nwork) {
  const unsigned int lane = threadIdx.x & 31u;
  for (size_t base = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) & ~(size_t)31; base < npix; base += (size_t)gridDim.x * blockDim.x) {
    const size_t pix = base + lane;
    bool mine = false;
    if (pix < npix) {
      const ulonglong2 px = *reinterpret_cast<const ulonglong2 *>(canvas + pix);
      unsigned int out = 0u;     /* 0 = nothing from this rank: background, or another rank is nearer */
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
      if (mine) work[at + __popc(who & ((1u << lane) - 1u))] = (unsigned int)pix;
    }
  }
}

/* ------------------------------------------------------------------ smooth */

struct ShadePixel {
  double w[3];                 /* barycentric weights of the pixel centre (render.c:102-110) */
  unsigned int pix;
  unsigned short i, j, k;      /* cube of the winning triangle */
  unsigned char obj, edges[3]; /* cube-edge ids of its three vertices */
  unsigned char valid;
};

struct ShadeSlot {             /* one neighbour cube of one vertex */
  unsigned short i, j, k;
  unsigned short touch;        /* triangles of that cube with a vertex on the shared edge (bit s) */
  unsigned char obj, code;
};

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

/* area * face_normal of triangle s of a cube (triangulate.c:75-93, normal.c:39-40) */
__device__ __forceinline__ V3 face_term(const FrameParams &p, const unsigned int obj, const long long i, const long long j, const long long k,
                                        const unsigned int code, const unsigned int s) {
  const FieldView &f = p.view[obj];
  const unsigned int pairs = __ldg(p.tab.pairs + code * C3D_MAX_TRIS + s);
  V3 v[3], fn;
  double area;
#pragma unroll
  for (int m = 0; m < 3; m++) v[m] = cut_edge_global(f, i, j, k, f.thr, (pairs >> (6 * m)) & 7u, (pairs >> (6 * m + 3)) & 7u);
  face_normal(v, fn, area);
  return v_mul(area, fn);
}

__global__ void __launch_bounds__(SH_THREADS) k_shade_smooth(const __grid_constant__ FrameParams p, const Pixel *__restrict__ canvas,
                                                             const unsigned int *__restrict__ work, const unsigned int *__restrict__ nwork,
                                                             unsigned int *__restrict__ rgba, unsigned long long *__restrict__ overflowed) {
  __shared__ ShadePixel s_pix[SH_THREADS];
  __shared__ ShadeSlot s_slot[SH_SLOTS];
  __shared__ unsigned int s_first[SH_SLOTS + 1];          /* first item of each slot */
  __shared__ unsigned short s_item[SH_ITEMS];             /* slot << 4 | triangle */
  __shared__ double s_term[SH_ITEMS][3];
  __shared__ double s_vn[SH_BATCH * 3][3];
  __shared__ unsigned int s_warp_total[SH_THREADS / 32];
  const unsigned int tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
  const unsigned int n = *nwork;
  for (unsigned int base = blockIdx.x * SH_THREADS; base < n; base += gridDim.x * SH_THREADS) {
    /* ---- S1: every thread rebuilds the winning triangle of its pixel and the weights of the pixel centre ---- */
    {
      ShadePixel me;
      me.valid = 0;
      me.pix = 0u; me.i = me.j = me.k = 0; me.obj = 0; me.edges[0] = me.edges[1] = me.edges[2] = 0;
      me.w[0] = me.w[1] = me.w[2] = 0.;
      if (base + tid < n) {
        const unsigned int pix = __ldg(work + base + tid);
        const unsigned long long key = canvas[pix].key;
        const unsigned int obj = (unsigned int)(key >> C3D_KEY_OBJ_SHIFT) - 1u;
        const unsigned long long prim = key & C3D_KEY_PRIM_MASK;
        const unsigned int s = (unsigned int)(prim & 15ull);
        const long long i = (long long)((prim >> 4) & 65535ull), j = (long long)((prim >> 20) & 65535ull), k = (long long)((prim >> 36) & 65535ull);
        const FieldView &f = p.view[obj];
        const unsigned int code = cube_code(f, i, j, k);
        const unsigned int pairs = __ldg(p.tab.pairs + code * C3D_MAX_TRIS + s);
        const unsigned int edges = __ldg(p.tab.edges + code * C3D_MAX_TRIS + s);
        Tri2D t;
        bool ok = true;
#pragma unroll
        for (int m = 0; m < 3; m++) {
          const ProjectedVertex v = exact_vertex(p, obj, i, j, k, (pairs >> (6 * m)) & 7u, (pairs >> (6 * m + 3)) & 7u);
          ok = ok && (0. != v.w);
          t.q[m] = V3{v.x, v.y, v.z};
        }
        if (ok) {
          setup_projected(p.cam, t);
          pixel_weights(t, __ldg(p.px + (pix % (unsigned int)p.cam.w)), __ldg(p.py + (pix / (unsigned int)p.cam.w)), me.w);
        }
        me.valid = 1;
        me.pix = pix;
        me.i = (unsigned short)i; me.j = (unsigned short)j; me.k = (unsigned short)k;
        me.obj = (unsigned char)obj;
        for (int m = 0; m < 3; m++) me.edges[m] = (unsigned char)((edges >> (5 * m)) & 31u);
      }
      s_pix[tid] = me;
    }
    __syncthreads();
    for (unsigned int sub = 0; sub < SH_THREADS / SH_BATCH; sub++) {
      if (base + sub * SH_BATCH >= n) break;                       /* uniform */
      const ShadePixel *batch = s_pix + sub * SH_BATCH;
      /* ---- S2: slots = (pixel, vertex, neighbour cube); count the triangles touching the shared edge ---- */
      unsigned int mycount[SH_SLOTS / SH_THREADS];
#pragma unroll
      for (unsigned int r = 0; r < SH_SLOTS / SH_THREADS; r++) {
        const unsigned int q = r * SH_THREADS + tid;
        const unsigned int px = q / 12u, m = (q % 12u) >> 2, c = q & 3u;
        const ShadePixel &sp = batch[px];
        ShadeSlot slot;
        slot.i = slot.j = slot.k = 0; slot.touch = 0; slot.obj = sp.obj; slot.code = 0;
        int di, dj, dk;
        unsigned int local;
        if (sp.valid && edge_neighbour(sp.edges[m], c, di, dj, dk, local)) {
          const long long ii = (long long)sp.i + di, jj = (long long)sp.j + dj, kk = (long long)sp.k + dk;
          const unsigned int code = cube_code(p.view[sp.obj], ii, jj, kk);
          slot.i = (unsigned short)ii; slot.j = (unsigned short)jj; slot.k = (unsigned short)kk;
          slot.code = (unsigned char)code;
          slot.touch = __ldg(p.tab.touch + code * 19u + local);
        }
        s_slot[q] = slot;
        mycount[r] = __popc((unsigned int)slot.touch);
      }
      /* ---- S3: exclusive scan of the slot counts in slot order (q = r * 128 + tid) ---- */
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
        /* ---- S3b: one item per touching triangle, in (slot, triangle) order ---- */
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
        /* ---- S4: the face terms, all threads ---- */
        for (unsigned int it = tid; it < total_items; it += SH_THREADS) {
          const unsigned int q = s_item[it] >> 4, s = s_item[it] & 15u;
          const ShadeSlot slot = s_slot[q];
          const V3 term = face_term(p, slot.obj, slot.i, slot.j, slot.k, slot.code, s);
          s_term[it][0] = term.x; s_term[it][1] = term.y; s_term[it][2] = term.z;
        }
        __syncthreads();
        /* ---- S5: per vertex, add the terms in order and normalise (normal.c:159) ---- */
        if (tid < SH_BATCH * 3) {
          const unsigned int q0 = tid * 4u;
          V3 sum = V3{0., 0., 0.};
          for (unsigned int it = s_first[q0]; it < s_first[q0 + 4u]; it++) {
            sum = v_add(sum, V3{s_term[it][0], s_term[it][1], s_term[it][2]});
          }
          const V3 vn = v_unit(sum);
          s_vn[tid][0] = vn.x; s_vn[tid][1] = vn.y; s_vn[tid][2] = vn.z;
        }
      } else {
        /* more touching triangles than the shared buffer holds: the same sums, one thread per vertex */
        if (0u == tid) atomicAdd(overflowed, 1ull);
        if (tid < SH_BATCH * 3) {
          V3 sum = V3{0., 0., 0.};
          for (unsigned int q = tid * 4u; q < tid * 4u + 4u; q++) {
            const ShadeSlot slot = s_slot[q];
            unsigned int touch = slot.touch;
            while (touch) {
              const unsigned int s = (unsigned int)__ffs((int)touch) - 1u;
              touch &= touch - 1u;
              sum = v_add(sum, face_term(p, slot.obj, slot.i, slot.j, slot.k, slot.code, s));
            }
          }
          const V3 vn = v_unit(sum);
          s_vn[tid][0] = vn.x; s_vn[tid][1] = vn.y; s_vn[tid][2] = vn.z;
        }
      }
      __syncthreads();
      /* ---- S6: blend, light, truncate (render.c:137-159) ---- */
      if (tid < SH_BATCH && batch[tid].valid) {
        const ShadePixel &sp = batch[tid];
        const V3 n0 = V3{s_vn[3 * tid][0], s_vn[3 * tid][1], s_vn[3 * tid][2]};
        const V3 n1 = V3{s_vn[3 * tid + 1][0], s_vn[3 * tid + 1][1], s_vn[3 * tid + 1][2]};
        const V3 n2 = V3{s_vn[3 * tid + 2][0], s_vn[3 * tid + 2][1], s_vn[3 * tid + 2][2]};
        /* w0*n0 + (w1*n1 + w2*n2), render.c:137-146 */
        const V3 nrm = v_unit(v_add(v_mul(sp.w[0], n0), v_add(v_mul(sp.w[1], n1), v_mul(sp.w[2], n2))));
        const double factor = fmax(0.15, fabs(v_dot(nrm, v_of(p.cam.light))));
        const FieldView &f = p.view[sp.obj];
        const unsigned int r = (unsigned int)__double2int_rz(r_mul(factor, (double)f.color[0])) & 255u;
        const unsigned int g = (unsigned int)__double2int_rz(r_mul(factor, (double)f.color[1])) & 255u;
        const unsigned int b = (unsigned int)__double2int_rz(r_mul(factor, (double)f.color[2])) & 255u;
        rgba[sp.pix] = 0xFF000000u | (b << 16) | (g << 8) | r;
      }
      __syncthreads();
    }
  }
}

extern "C" int c3d_k_shade(const FrameParams *p, const LineDev *lines_dev, const Pixel *canvas, const unsigned long long *depth_global,
                           unsigned int *rgba, size_t npix, unsigned int *work, unsigned int *nwork, unsigned long long *overflowed,
                           int sm_count, cudaStream_t s) {
  const cudaError_t err = cudaMemsetAsync(nwork, 0, sizeof(unsigned int), s);
  if (cudaSuccess != err) return (int)err;
  k_shade_select<<<sm_count * 8, 256, 0, s>>>(*p, lines_dev, canvas, depth_global, rgba, npix, work, nwork);
  k_shade_smooth<<<sm_count * 6, SH_THREADS, 0, s>>>(*p, canvas, work, nwork, rgba, overflowed);
  return (int)cudaGetLastError();
}
