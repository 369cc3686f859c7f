/*
 * contour3d_b200.h -- C-ABI of the CUDA (sm_100a) implementation.
 *
 * This is the layer the drop-in entry point contour3d_execute()
 * (include/contour3d.h) is written on, and what foreign-language hosts bind
 * (ctypes: contour3d_b200/api.py).  Plain pointers and sizes only; every
 * function returns 0 on success and non-zero on failure after printing one
 * "[CONTOUR3D ERROR] ..." line to stderr (reference: src/contour3d/logger.c:5-18).
 * There is no CPU fallback anywhere behind this interface.
 *
 * Each entry point names the reference code it replaces (paths relative to
 * the reference repository NaokiHori/Contour3D @ c66854e).
 */
#if !defined(CONTOUR3D_B200_H)
#define CONTOUR3D_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__cplusplus)
extern "C" {
#endif

#define C3D_NCCL_ID_BYTES 128

typedef struct c3d_context c3d_context_t;   /* device, streams, NCCL communicator, canvases, work lists */
typedef struct c3d_field c3d_field_t;       /* one device-resident local array with halo room */

/* Position of this rank in the process grid: what contour/extend.c:216-252
   obtains from sdecomp.get_nprocs / get_myrank / get_neighbours.  Direction 0
   (x) is never split by SimpleDecomp (sdecomp/main.c:65-73,133-139). */
typedef struct {
  int nprocs[3];
  int myrank[3];
  int neighbours[3][2];       /* communicator ranks of the -/+ neighbours, -1 = none */
} c3d_decomp_t;

/* reference: arguments of contour3d_execute (include/contour3d.h:58-86) and
   init_camera / init_light / init_screen (src/contour3d/main.c:13-63) */
typedef struct {
  double camera[3];
  double light[3];            /* any length; normalised internally */
  double center[3];
  double local_x[3];
  double local_y[3];
  size_t width, height;
  uint8_t bg[3];
} c3d_scene_t;

/* reference: the `converter` callback of a contour object (include/contour3d.h:31-33), which the reference calls
   on every cube corner (triangulate.c:156-173).  Device code cannot call a host function, so the host tabulates it:
   component c of the Cartesian position of point (i, j, k) of the field's EXTENDED local array (c3d_field_geometry)
   is table[c][i * stride[c][0] + j * stride[c][1] + k * stride[c][2]]; stride 0 = independent of that direction.
   contour3d_execute() builds these by probing the callback (c3d_positions_probe); the values are the callback's own,
   so the geometry is bit-identical to the reference's whatever the function computes. */
typedef struct {
  const double *table[3];     /* HOST pointers */
  size_t stride[3][3];
} c3d_positions_t;

/* reference: contour3d_contour_obj_t (include/contour3d.h:22-40); the array is a c3d_field_t */
typedef struct {
  const c3d_field_t *field;
  const double *grids[3];     /* HOST pointers, global coordinates (glsizes[d] each) */
  double threshold;
  uint8_t color[3];
  const c3d_positions_t *positions;   /* NULL: identity converter, positions are the grid coordinates themselves */
} c3d_contour_t;

/* reference: contour3d_line_obj_t (include/contour3d.h:42-56) */
typedef struct {
  double edges[2][3];
  size_t nitems;
  double width;
  uint8_t color[3];
  const double *points;       /* NULL: identity converter; else HOST array of nitems x 3 doubles: converter(edges[0] + n * delta),
                                 n = 0 .. nitems-1, delta as in line.c:138-140,157-158 (contour3d_execute fills it) */
} c3d_line_t;

/* same 208-byte record as the oracle's dump (oracle/c3d_oracle.h) */
typedef struct {
  uint64_t cube;              /* (k*(my-1)+j)*(mx-1)+i on the extended local array */
  uint32_t tri;
  uint32_t has_vn;
  uint32_t edge[3];
  uint32_t pad;
  double area;
  double v[9];
  double fn[3];
  double vn[9];
} c3d_triangle_t;

/* per-stage device times of the last frame, milliseconds (CUDA events) */
typedef struct {
  float upload, halo, clear, sweep, expand, raster, lines, resolve, reduce, shade, pack, download, total;
  float sweep_each[16];          /* per sweep launch */
  uint64_t sweep_bytes[16];      /* algorithmic bytes of each sweep launch (elem_bytes x points of the extended array) */
  uint32_t sweep_thresholds[16]; /* iso-values classified by each sweep launch */
  uint64_t active_cubes[16];     /* per contour object, rendered range */
  uint64_t culled_cubes;         /* cut cubes the raster skipped because the z-buffer already hid them */
  uint64_t culled_triangles;     /* triangles of the remaining cubes dropped by the per-triangle pre-test */
  uint64_t shade_spills;         /* shading sub-batches that took the one-thread-per-vertex path */
  uint32_t nsweeps;
  uint32_t ncontours;
  uint32_t launches;             /* kernels launched by this library in the frame */
  uint32_t exchange_mode;        /* 0 one rank; 1 every rank draws its own cubes, whole images reduced through NCCL;
                                    2 screen-partitioned: cube lists, shading work and colours move through peer memory */
  uint64_t list_entries;         /* cubes this rank's raster walked (mode 2: from all ranks, for its screen tiles) */
  float detail[16];              /* mode 2, ms: [0] barrier at the start, [1] halo pull, [2] staging + clear + sweeps, [3] cubes -> outboxes,
                                    [4] barrier, [5] pull + merge of the lists, [6] raster + lines + select, [7] barrier, [8] pull of the
                                    shading work, [9] shading, [10] barrier at the end */
  float tile_share[8];           /* mode 2: fraction of the screen tiles each rank draws (follows the ranks' measured loads) */
} c3d_timings_t;

/* ---- life cycle ---- */

/* fills a 128-byte ncclUniqueId on one rank; the host distributes it
   (MPI_Bcast in contour3d_execute.c, torch.distributed in bench.py) */
int c3d_nccl_unique_id(void *id128);

/* device < 0 selects comm_rank % device_count.  nccl_id may be NULL when comm_size == 1 */
int c3d_context_create(c3d_context_t **ctx, int device, int comm_rank, int comm_size, const void *nccl_id128);
int c3d_context_destroy(c3d_context_t *ctx);
int c3d_context_device(const c3d_context_t *ctx);
/* cudaStream_t the pipeline runs on, as an integer (for event timing from outside) */
uintptr_t c3d_context_stream(const c3d_context_t *ctx);
int c3d_context_synchronize(c3d_context_t *ctx);
/* how the ranks of this context exchange data: 0 one rank; 1 NCCL (halo send/recv, whole-image reductions); 2 peer memory
   over NVLink (CUDA IPC): halo planes, cube lists, shading work and colours are read / written in place, screen-partitioned */
int c3d_context_exchange_mode(const c3d_context_t *ctx);

/* block decomposition of SimpleDecomp (sdecomp/kernel.c:89,111-116) for a
   process grid dims = {1, Py, Pz}, ranks row-major with z fastest (MPI_Cart_create) */
int c3d_decomp_make(const int dims[3], int comm_rank, c3d_decomp_t *out);
int c3d_decomp_local(const c3d_decomp_t *d, const size_t glsizes[3], size_t mysizes[3], size_t offsets[3]);

/* ---- fields: replaces the host-side padding + copy of contour/extend.c:182-278 ---- */

/* elem_bytes 4 (float) or 8 (double).  The device buffer already has the
   extended shape (two extra points on every side that has a neighbour) */
int c3d_field_create(c3d_context_t *ctx, const c3d_decomp_t *decomp, const size_t glsizes[3], int elem_bytes, c3d_field_t **field);
int c3d_field_destroy(c3d_context_t *ctx, c3d_field_t *field);
/* copy the local pencil (mysizes, x fastest; host or device pointer, same element type) into the interior */
int c3d_field_upload(c3d_context_t *ctx, c3d_field_t *field, const void *local);
/* extended shape / global offsets of the first extended point / device pointer + strides of the INTERIOR */
int c3d_field_geometry(const c3d_field_t *field, size_t ext_sizes[3], size_t ext_offsets[3], size_t mysizes[3], size_t offsets[3]);
int c3d_field_interior(const c3d_field_t *field, void **device_ptr, size_t strides_elems[2]);
/* One rank (no halo room needed): sweep a DEVICE array of the caller where it lies instead of copying it.  Returns 0 and a
   field that borrows device_array (read only; c3d_field_destroy leaves it alone), 2 when that does not apply (host pointer,
   other device, several ranks) -- then create + upload as usual.  reference: the copy of contour/extend.c:199-208 */
int c3d_field_wrap(c3d_context_t *ctx, const c3d_decomp_t *decomp, const size_t glsizes[3], int elem_bytes,
                   const void *device_array, c3d_field_t **field);
/* tell the library the interior was written through c3d_field_interior (halo must be re-exchanged) */
int c3d_field_mark_dirty(c3d_field_t *field);
/* halo exchange: replaces communicate_in_{x,y,z}, contour/extend.c:10-147 (ncclSend/ncclRecv) */
int c3d_field_exchange(c3d_context_t *ctx, c3d_field_t *field);
/* copy the extended array to the host (tests) */
int c3d_field_download_ext(c3d_context_t *ctx, const c3d_field_t *field, void *host);

/* synthetic benchmark fields, generated on the device (kind: 0 gyroid, 1 multi-octave value noise, 2 curl magnitude of a
   three-component noise field); kinds 1 and 2 have a bit-identical host twin (c3d_synth_host.c) for the CPU reference arm */
int c3d_field_synthesize(c3d_context_t *ctx, c3d_field_t *field, int kind, uint32_t seed, int octaves, double gain);

/* ---- whole frame: replaces contour3d_execute's body, src/contour3d/main.c:68-146 ---- */

/* Renders all contour objects and lines, reduces across ranks, and on
   communicator rank 0 optionally copies the y-flipped RGB bytes (3*W*H, PPM
   payload order, output.c:113-123) to host_rgb and/or writes the PPM file.
   host_rgb and fname may be NULL (result stays on the device). */
int c3d_frame_render(c3d_context_t *ctx, const c3d_scene_t *scene,
                     size_t num_contours, const c3d_contour_t *contours,
                     size_t num_lines, const c3d_line_t *lines,
                     uint8_t *host_rgb, const char *fname);
int c3d_frame_timings(const c3d_context_t *ctx, c3d_timings_t *out);
/* canvas of the last frame on this rank, canvas order (row j, column i; not flipped):
   depth as double (-DBL_MAX where empty), winner key (0 = background) */
int c3d_frame_canvas(c3d_context_t *ctx, double *host_depth, uint64_t *host_key);

/* The frame's two image collectives on their own (measurement: NVLink / NVSwitch throughput of the reduction that replaces
   MPI_Reduce of output.c:82): `iters` x ncclAllReduce(max) over a width x height plane of 64-bit depths, then `iters` x
   ncclReduce(max) of the 32-bit colour plane to rank 0, each timed with CUDA events after a warm-up round; milliseconds
   per call.  Collective: every rank of the context's communicator must call it.  comm_size 1: both times are 0. */
int c3d_frame_collectives_probe(c3d_context_t *ctx, size_t width, size_t height, int iters, float *allreduce_ms, float *reduce_ms);

/* ---- stage-level entry points (parity tests, mesh export) ---- */

/* replaces the classification of triangulate_tetrahedron (triangulate.c:145-152,176-192):
   per cube of the extended array the six pre-fold 4-bit masks (24 bits) and the triangle count */
int c3d_case_codes(c3d_context_t *ctx, const c3d_field_t *field, double threshold, uint32_t *host_codes, uint8_t *host_counts);
/* the streaming sweep: compact list of active cubes, (cube << 8) | corner bits.
   rendered_only = 1 restricts to the cubes the reference renders (contour/main.c:92-93).
   Writes min(*count, capacity) entries in unspecified order. */
int c3d_sweep(c3d_context_t *ctx, const c3d_field_t *field, double threshold, int rendered_only,
              size_t capacity, uint64_t *host_entries, uint64_t *count);
/* replaces contour3d_contour_triangulate_slice + compute_vertex_normals for the whole array
   (triangulate.c:287-342, normal.c:46-165): all triangles, vertex normals for interior cubes,
   sorted by (cube, tri).  Writes min(*count, capacity) records. */
int c3d_tessellate(c3d_context_t *ctx, const c3d_field_t *field, const double *const grids[3], double threshold,
                   size_t capacity, c3d_triangle_t *host_out, uint64_t *count);

/* ---- coordinate converters (host side, needs no device) ---- */

/* the callback type of include/contour3d.h:31-33 / 49-51 on plain doubles: xyz in, xyz out */
typedef void (*c3d_converter_fn)(const double in[3], double out[3], void *user);

/* Tabulates `convert` over the grid points [offset[d], offset[d] + count[d]) of `grids` (global coordinates):
   finds which output component depends on which direction (bitwise comparison over sample lines through the block),
   fills one table per component of the smallest dimensionality that reproduces the function on a scattered
   verification sample, and returns them in *out (malloc'ed: release with c3d_positions_free).  *identity = 1 and
   out->table[] = NULL when the function returns its argument unchanged everywhere it was evaluated.
   max_bytes bounds the total size of the tables (a full 3-D dependence costs 8 bytes x points per component). */
int c3d_positions_probe(c3d_converter_fn convert, void *user, const double *const grids[3], const size_t offset[3],
                        const size_t count[3], size_t max_bytes, c3d_positions_t *out, int *identity);
void c3d_positions_free(c3d_positions_t *p);

/* ---- camera helper (host side) ---- */

/* Fills camera, center, local_x, local_y of *scene from viewing angles in degrees, the way the reference's
   CameraAdjuster tool does (CameraAdjuster/src/config.ts:140-192): camera at distance camera_distance from the origin,
   tilted by elevation about x, turned by azimuth about z; screen vectors of lengths screen_width / screen_height with
   the same rotations plus roll about the screen normal; screen centre on the origin-camera line at screen_distance.
   light, width, height and bg are left as they are.  Returns non-zero on non-positive lengths. */
int c3d_camera_from_angles(double elevation_deg, double azimuth_deg, double roll_deg, double camera_distance, double screen_distance,
                           double screen_width, double screen_height, c3d_scene_t *scene);

/* ---- diagnostics (host side, needs no device) ---- */

/* Self test of the bootstrap the ranks of one node use to exchange the handles of each other's device memory (a POSIX
   shared-memory board named after the 128-byte id of the run; it carries what the reference's MPI communicator carries
   between ranks, contour/extend.c:10-147, output.c:82, before any device is involved).  Collective over `size` processes
   that share id128: every rank contributes `value` and receives in *sum the sum over the ranks of (rank + 1) * value.
   Returns 0, or non-zero when the board could not be set up (the library then exchanges through NCCL only). */
int c3d_peer_selftest(int rank, int size, const void *id128, unsigned long long value, unsigned long long *sum);

/* The two plans the ranks of a screen-partitioned frame each compute for themselves from shared numbers, run on the host
   (what the reference leaves to MPI_Reduce, output.c:82, is here a division of labour that every rank must derive identically):
   - which pixels rank `me` shades, from counts[d * n + o] = pixels of rank d's tiles won by cubes of rank o's array;
     chunks = nchunks x 5 words (owner, d, first, count, dest), room for 80; nlocal = pixels of its own array (they come
     first), total = all it takes.  remote_cost < 1: no helping.
   - which rank draws which screen tile (ty * tiles_x + tx), from the ranks' shares of the screen. */
int c3d_selftest_work_plan(const unsigned int *counts, int n, int me, double remote_cost, unsigned int *chunks, unsigned int *nchunks,
                           unsigned int *nlocal, unsigned int *total);
int c3d_selftest_tile_map(const double *share, int n, int tiles_x, int tiles_y, unsigned char *out);

/* library / build information, e.g. "contour3d_b200 0.1 sm_100a" */
const char *c3d_version(void);

#if defined(__cplusplus)
}
#endif

#endif /* CONTOUR3D_B200_H */
