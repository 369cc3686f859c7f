/*
 * c3d_runtime.cu -- host side of the C-ABI (include/contour3d_b200.h):
 * persistent context (stream, NCCL communicator, canvases, work lists),
 * device-resident fields with halo room, halo exchange, frame orchestration.
 *
 * No CPU fallback: every entry point needs a CUDA device and fails loudly
 * ("[CONTOUR3D ERROR] ...", return 1) otherwise.
 *
 * File:line references are to NaokiHori/Contour3D @ c66854e.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#include <string.h>
#include <errno.h>
#include <math.h>
#include <float.h>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>
#include <nccl.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/uio.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include "contour3d_b200.h"
#include "c3d_internal.h"
#include "c3d_peer.h"

#define C3D_STR__(x) #x
#define C3D_STR_(x) C3D_STR__(x)

/* same format as the reference's logger (src/contour3d/logger.c:5-18) */
static int c3d_error(const char *fmt, ...) {
  va_list args;
  va_start(args, fmt);
  fprintf(stderr, "[CONTOUR3D ERROR] ");
  vfprintf(stderr, fmt, args);
  fprintf(stderr, "\n");
  fflush(stderr);
  va_end(args);
  return 1;
}

#define CUDA_OK(call) \
  do { \
    const cudaError_t err_ = (call); \
    if (cudaSuccess != err_) return c3d_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(err_), __FILE__, __LINE__); \
  } while (0)

#define LAUNCH_OK(call) \
  do { \
    const int err_ = (call); \
    if (0 != err_) return c3d_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString((cudaError_t)err_), __FILE__, __LINE__); \
  } while (0)

#define NCCL_OK(call) \
  do { \
    const ncclResult_t err_ = (call); \
    if (ncclSuccess != err_) return c3d_error("%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(err_), __FILE__, __LINE__); \
  } while (0)

/*
 * NCCL is bound lazily with dlopen the first time a multi-rank context is
 * created: a single-GPU process never loads it, and a process that already
 * holds an NCCL (e.g. the one bundled with PyTorch, same soname libnccl.so.2)
 * re-uses that copy instead of pulling in a second, possibly older, one.
 */
struct NcclApi {
  void *handle;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*GroupStart)(void);
  ncclResult_t (*GroupEnd)(void);
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t);
  const char *(*GetErrorString)(ncclResult_t);
};
static NcclApi g_nccl;

static int nccl_load(void) {
  if (NULL != g_nccl.handle) return 0;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (NULL == h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (NULL == h) return c3d_error("multi-GPU runs need NCCL, but libnccl.so.2 could not be loaded: %s", dlerror());
#define C3D_BIND(name) \
  *(void **)(&g_nccl.name) = dlsym(h, "nccl" #name); \
  if (NULL == g_nccl.name) return c3d_error("libnccl.so.2 lacks nccl" #name)
  C3D_BIND(GetUniqueId); C3D_BIND(CommInitRank); C3D_BIND(CommDestroy); C3D_BIND(GroupStart); C3D_BIND(GroupEnd);
  C3D_BIND(Send); C3D_BIND(Recv); C3D_BIND(AllReduce); C3D_BIND(Reduce); C3D_BIND(GetErrorString);
#undef C3D_BIND
  g_nccl.handle = h;
  return 0;
}

/* what a rank tells the others about its block of a field (exchanged once, when the field is created) */
struct FieldGeom {
  int coords[3];                      /* position in the process grid */
  unsigned int mysizes[3], offsets[3], ext[3], ext_off[3], lo[3];
};

struct c3d_field {
  c3d_decomp_t decomp;
  size_t gl[3], mysizes[3], offsets[3];
  size_t ext[3], ext_off[3], lo[3];   /* extended sizes, global index of the first extended point, low halo widths */
  int elem_bytes;
  void *data;
  bool halo_valid;
  bool borrowed;                      /* data is the caller's device array (c3d_field_wrap): never written, never freed */
  bool shared;                        /* every rank has every other rank's array mapped (peer[]) and knows its shape (geom[]) */
  void *peer[C3D_MAX_RANKS];
  FieldGeom geom[C3D_MAX_RANKS];
};

enum { CNT_LIST = 0, CNT_REC = 1, CNT_CULLED = 2, CNT_CULLED_TRIS = 3, CNT_SHADE_SPILL = 4, CNT_OVERFLOW = 5, CNT_OBJ = 6, CNT_SNAP = CNT_OBJ + C3D_MAX_OBJECTS,
       CNT_OUT = CNT_SNAP + C3D_MAX_OBJECTS + 1,      /* screen-partitioned frames: entries this rank addressed to the ranks (outbox cursor) */
       CNT_SEG_OVER,                                  /* ... pixels / work items that found no room (sized for the worst case: stays 0) */
       CNT_PEER,                                      /* ... host copy only: a flag barrier timed out */
       CNT_SHADE_ITEMS,                               /* pixels this rank shaded (sizes next frame's vertex table) */
       CNT_LOADS,                                     /* ... C3D_MAX_RANKS words: what every rank's previous frame spent drawing / shading (two floats each) */
       CNT_LOADS_END = CNT_LOADS + C3D_MAX_RANKS - 1,
       CNT_TOTAL };
#define C3D_HIZ_CUTS_MULTI ""
#define C3D_HIZ_CUTS_SINGLE ""
#define C3D_DETAIL_EVENTS 20
enum { EV_START, EV_HALO, EV_CLEAR, EV_SWEEP, EV_EXPAND, EV_RASTER, EV_LINES, EV_RESOLVE, EV_REDUCE0, EV_SHADE, EV_REDUCE1, EV_PACK, EV_DOWNLOAD, EV_COUNT };

/*
 * Device memory of a screen-partitioned frame that OTHER ranks read or write: one allocation per rank, mapped by all.
 * The layout up to the outbox depends only on (ranks, screen, seg_cap), so every rank can address every other rank's;
 * the outbox comes last because its capacity is the owner's business (readers go by seg_start and the histogram).
 */
struct Exchange {
  char *base;
  void *peer[C3D_MAX_RANKS];
  size_t bytes, npix, width, height, seg_cap, out_cap;
  int nranks;
  bool stale;                          /* must be re-created (and re-mapped by everybody) before the next frame */
  size_t off_flags, off_seg_start, off_nseg, off_hist, off_rgba, off_seg, off_out;
};

struct c3d_context {
  int device, rank, size, sm_count;
  cudaStream_t stream;
  ncclComm_t comm;
  bool has_comm;
  c3d_peer peer;                                   /* device-to-device exchange between the ranks of this node */
  Exchange xa;
  unsigned long long *merge_base, *part_cursors;   /* C3D_PART_BUCKETS x ranks each: merged-list offsets / outbox cursors */
  unsigned long long shade_items_prev;             /* pixels shaded in the previous frame; 0: unknown */
  unsigned char *part_masks;                       /* 32 bytes per sweep record: which ranks draw each of its cubes */
  size_t part_masks_cap;                           /* in records */
  int tile_shift;                                  /* log2 of the screen-tile edge of partitioned frames */
  PullPlan *pull_plan;                             /* device: which shading work this rank pulls (k_work_plan) */
  /* Screen tiles are dealt out in proportion to `tile_share`: a rank whose cubes win much of the image (the slab facing the
     camera) has more to shade and gets fewer tiles to draw.  The shares follow the ranks' measured times of earlier frames
     (published in the exchange arena, gathered at the end of every frame: the same numbers on every rank, so the same table). */
  double tile_share[C3D_MAX_RANKS];
  unsigned char *tile_map_dev;
  std::vector<unsigned char> *tile_map_host;
  size_t tile_map_w, tile_map_h;
  int tile_balance;                                /* 0: fixed pattern */
  float my_loads[2];                               /* my last frame: ms drawing (raster + lines + select), ms shading */
  /* The PPM of the previous call, kept mapped and registered with the driver: a solver writes the same name every output
     step, and the packed image can then travel from the device straight into the file's pages (RAM-backed file systems
     and the page cache alike) -- no staging copy, no write().  Dropped as soon as name, size or inode differ. */
  char *out_path;
  unsigned char *out_map;
  size_t out_len;
  unsigned long long out_ino, out_dev;
  bool out_registered;
  float shade_share;                               /* relative cost of a pixel shaded by a helper rank (0: no helping) */
  void *table_mem;
  DevTables tab;
  /* canvases */
  size_t npix_cap;
  Pixel *canvas;                                   /* (depth, winner key) per pixel */
  unsigned long long *depth_local, *depth_global;  /* multi-rank only: contiguous depth planes for ncclMax */
  unsigned long long *hiz;                         /* coarse z-buffer: farthest depth per C3D_HIZ_TILE^2 pixels (c3d_k_hiz_build) */
  int raster_cuts[C3D_RASTER_PARTS];               /* where the raster's list is cut for a Hi-Z refresh, in 1/64ths, ascending */
  int nraster_cuts;
  double *pixel_xy;                                /* pixel-centre tables: width + height doubles */
  size_t table_w, table_h;
  unsigned int *rgba;
  ShadeScratch shade;                              /* work list, vertex table and normals of the deferred shading */
  unsigned char *ppm_dev, *ppm_host;
  size_t width, height;
  /* active-cube list of the frame (all contour objects, tagged) */
  unsigned long long *list;
  size_t list_cap;
  void *records;                 /* 32-byte patch records written by the tiled sweep */
  size_t rec_cap;
  unsigned int *tile_cursor;
  unsigned int *bucket_hist;             /* C3D_DEPTH_BUCKETS counters + cursors of the depth ordering */
  unsigned long long *bucket_cursors;
  /* counters: [CNT_LIST] list cursor, [CNT_REC] record cursor, [CNT_OBJ + n] cut cubes of object n (expand),
     [CNT_SNAP + m] list cursor after direct-sweep launch m */
  unsigned long long *counts_dev, *counts_host;
  /* per-frame descriptors */
  FieldView *views_dev;
  LineDev *lines_dev;
  size_t lines_cap;
  double *grids_dev;             /* stage-level entry points: one view's grid coordinates */
  size_t grids_cap;
  char *arena_dev, *arena_host;  /* per-frame descriptors: converter maps, lines, grid coordinates, converter tables (device; pinned staging of the small part) */
  size_t arena_cap, arena_host_cap;
  void *halo_buf;
  size_t halo_cap;
  cudaEvent_t ev[EV_COUNT];
  cudaEvent_t ev_upload[2 * C3D_MAX_OBJECTS];   /* start/stop pairs of the uploads since the last frame */
  int nuploads;
  cudaEvent_t ev_sweep[C3D_MAX_OBJECTS + 1];
  cudaEvent_t ev_detail[C3D_DETAIL_EVENTS];      /* screen-partitioned frames: the steps between the stage events */
  int ndetail;
  c3d_timings_t last;
};

extern "C" const char *c3d_version(void) { return "contour3d_b200 0.1 (sm_100a, CUDA " C3D_STR_(CUDART_VERSION) ")"; }

extern "C" int c3d_nccl_unique_id(void *id128) {
  if (NULL == id128) return c3d_error("c3d_nccl_unique_id: NULL output");
  static_assert(sizeof(ncclUniqueId) == C3D_NCCL_ID_BYTES, "ncclUniqueId size");
  if (0 != nccl_load()) {
    /* no NCCL on this host: the id still names the ranks' shared-memory board (c3d_peer.cu); any 128 random bytes will do */
    FILE *fp = fopen("/dev/urandom", "rb");
    const size_t got = NULL != fp ? fread(id128, 1, C3D_NCCL_ID_BYTES, fp) : 0;
    if (NULL != fp) fclose(fp);
    return C3D_NCCL_ID_BYTES == got ? 0 : c3d_error("could not create a unique id");
  }
  ncclUniqueId id;
  NCCL_OK(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

static int context_init(c3d_context *ctx, const int comm_rank, const int comm_size, const void *nccl_id128) {
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, ctx->device));
  if (prop.major < 10) return c3d_error("device %d is sm_%d%d; this library is built for sm_100a only", ctx->device, prop.major, prop.minor);
  ctx->sm_count = prop.multiProcessorCount;
  CUDA_OK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  for (int e = 0; e < EV_COUNT; e++) CUDA_OK(cudaEventCreate(&ctx->ev[e]));
  for (int e = 0; e <= C3D_MAX_OBJECTS; e++) CUDA_OK(cudaEventCreate(&ctx->ev_sweep[e]));
  for (int e = 0; e < 2 * C3D_MAX_OBJECTS; e++) CUDA_OK(cudaEventCreate(&ctx->ev_upload[e]));
  for (int e = 0; e < C3D_DETAIL_EVENTS; e++) CUDA_OK(cudaEventCreate(&ctx->ev_detail[e]));
  /* case tables -> device */
  c3d_tables_t *host_tab = (c3d_tables_t *)malloc(sizeof(c3d_tables_t));
  if (NULL == host_tab) return c3d_error("out of host memory");
  c3d_tables_build(host_tab);
  cudaError_t err = cudaMalloc(&ctx->table_mem, sizeof(c3d_tables_t));
  if (cudaSuccess == err) err = cudaMemcpy(ctx->table_mem, host_tab, sizeof(c3d_tables_t), cudaMemcpyHostToDevice);
  free(host_tab);
  if (cudaSuccess != err) return c3d_error("case tables: %s", cudaGetErrorString(err));
  const char *tb = (const char *)ctx->table_mem;
  ctx->tab.count = (const uint8_t *)(tb + offsetof(c3d_tables_t, count));
  ctx->tab.code24 = (const uint32_t *)(tb + offsetof(c3d_tables_t, code24));
  ctx->tab.pairs = (const uint32_t *)(tb + offsetof(c3d_tables_t, pairs));
  ctx->tab.emask = (const uint32_t *)(tb + offsetof(c3d_tables_t, emask));
  ctx->tab.edges = (const uint32_t *)(tb + offsetof(c3d_tables_t, edges));
  ctx->tab.touch = (const uint16_t *)(tb + offsetof(c3d_tables_t, touch));
  CUDA_OK(cudaMalloc(&ctx->counts_dev, CNT_TOTAL * sizeof(unsigned long long)));
  CUDA_OK(cudaMallocHost(&ctx->counts_host, CNT_TOTAL * sizeof(unsigned long long)));
  CUDA_OK(cudaMalloc(&ctx->tile_cursor, sizeof(unsigned int)));
  CUDA_OK(cudaMalloc(&ctx->shade.counters, 4 * sizeof(unsigned int)));
  CUDA_OK(cudaMalloc(&ctx->bucket_hist, C3D_DEPTH_BUCKETS * sizeof(unsigned int)));
  CUDA_OK(cudaMalloc(&ctx->bucket_cursors, C3D_DEPTH_BUCKETS * sizeof(unsigned long long)));
  CUDA_OK(cudaMalloc(&ctx->views_dev, C3D_MAX_OBJECTS * sizeof(FieldView)));
  /* Hi-Z refresh points of the raster (see frame_once): "a,b,..." in 1/64ths of the nearest-first list, "" or "0" = none.
     Off by default (measured: the frame is bounded by the front rank, which has nothing to cull and only waits at the extra
     collectives); every rank must use the same cuts (each refresh is a collective). */
  {
    const char *env = getenv("CONTOUR3D_HIZ_CUTS");
    const char *spec = NULL != env ? env : (comm_size > 1 ? C3D_HIZ_CUTS_MULTI : C3D_HIZ_CUTS_SINGLE);
    int last = 0;
    while ('\0' != *spec && ctx->nraster_cuts < (int)C3D_RASTER_PARTS - 1) {
      char *end = NULL;
      const long v = strtol(spec, &end, 10);
      if (end == spec) break;
      if (v > last && v < (long)C3D_RASTER_PARTS) ctx->raster_cuts[ctx->nraster_cuts++] = last = (int)v;
      spec = (',' == *end) ? end + 1 : end;
    }
  }
  {
    const char *env = getenv("CONTOUR3D_TILE_SHIFT");
    ctx->tile_shift = (NULL != env && atoi(env) >= 3 && atoi(env) <= 10) ? atoi(env) : 6;
    /* Tile shares that follow the ranks' measured loads: built and measured (the slab facing the camera ends up with half the
       tiles of the others), but the frame did not get shorter -- the shading of a rank's cubes can only start when EVERY rank
       has resolved its tiles, so the stages are balanced one by one, not as a sum.  Off unless asked for. */
    env = getenv("CONTOUR3D_TILE_BALANCE");
    ctx->tile_balance = (NULL != env && '1' == env[0]) ? 1 : 0;
    for (int r = 0; r < C3D_MAX_RANKS; r++) ctx->tile_share[r] = 1. / (double)comm_size;
    ctx->tile_map_host = new std::vector<unsigned char>();
  }
  if (comm_size > 1) {
    if (NULL == nccl_id128) return c3d_error("c3d_context_create: %d ranks need the shared 128-byte id (c3d_nccl_unique_id)", comm_size);
    /* 1. device-to-device exchange between the ranks of this node (CUDA IPC): the data path of multi-rank frames */
    if (0 != c3d_peer_init(&ctx->peer, comm_rank, comm_size, nccl_id128, ctx->device)) return 1;
    if (ctx->peer.active) {
      CUDA_OK(cudaMalloc(&ctx->merge_base, (size_t)C3D_PART_BUCKETS * C3D_MAX_RANKS * sizeof(unsigned long long)));
      CUDA_OK(cudaMalloc(&ctx->part_cursors, (size_t)C3D_PART_BUCKETS * C3D_MAX_RANKS * sizeof(unsigned long long)));
      CUDA_OK(cudaMalloc(&ctx->pull_plan, sizeof(PullPlan)));
      /* what shading a pixel costs a helper (reading the owner's array through peer memory) relative to the owner itself (measured at 8 ranks: 8 gives the shortest frame);
         0 = no helping: every rank shades exactly the pixels its own cubes won */
      const char *env = getenv("CONTOUR3D_SHADE_REMOTE_COST");
      ctx->shade_share = (NULL != env && '\0' != *env) ? (float)atof(env) : 8.f;
      if (!(ctx->shade_share >= 1.f)) ctx->shade_share = 0.f;
    }
    /* 2. NCCL communicator: frames with coordinate converters and nodes without peer access reduce whole images through it.
          Ranks that share a GPU (test boxes) cannot form one: they have the peer path only. */
    if (!ctx->peer.shared_device) {
      if (0 != nccl_load()) return 1;
      ncclUniqueId id;
      memcpy(&id, nccl_id128, sizeof(id));
      NCCL_OK(g_nccl.CommInitRank(&ctx->comm, comm_size, id, comm_rank));
      ctx->has_comm = true;
    } else if (!ctx->peer.active) {
      return c3d_error("ranks share a GPU and cannot map each other's memory: no way to exchange data");
    }
  }
  return 0;
}

extern "C" int c3d_context_create(c3d_context_t **out, int device, int comm_rank, int comm_size, const void *nccl_id128) {
  if (NULL == out) return c3d_error("c3d_context_create: NULL output");
  *out = NULL;
  if (comm_size < 1 || comm_rank < 0 || comm_rank >= comm_size) return c3d_error("c3d_context_create: bad rank %d of %d", comm_rank, comm_size);
  int ndev = 0;
  const cudaError_t probe = cudaGetDeviceCount(&ndev);
  if (cudaSuccess != probe || ndev < 1) {
    return c3d_error("no usable CUDA device (%s); this library has no CPU path", cudaSuccess != probe ? cudaGetErrorString(probe) : "device count is 0");
  }
  if (device < 0) device = comm_rank % ndev;
  if (device >= ndev) return c3d_error("device %d requested but only %d present", device, ndev);
  CUDA_OK(cudaSetDevice(device));
  c3d_context *ctx = (c3d_context *)calloc(1, sizeof(c3d_context));
  if (NULL == ctx) return c3d_error("out of host memory");
  ctx->device = device;
  ctx->rank = comm_rank;
  ctx->size = comm_size;
  if (0 != context_init(ctx, comm_rank, comm_size, nccl_id128)) {
    c3d_context_destroy(ctx);        /* releases whatever had been created */
    return 1;
  }
  *out = ctx;
  return 0;
}

static void release_output_map(c3d_context *ctx);

static void release_exchange(c3d_context *ctx) {
  if (NULL != ctx->xa.base) {
    c3d_peer_unshare(&ctx->peer, ctx->xa.peer);
    cudaFree(ctx->xa.base);
  }
  memset(&ctx->xa, 0, sizeof(ctx->xa));
}

extern "C" int c3d_context_destroy(c3d_context_t *ctx) {
  if (NULL == ctx) return 0;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->has_comm) g_nccl.CommDestroy(ctx->comm);
  release_exchange(ctx);
  release_output_map(ctx);
  cudaFree(ctx->merge_base); cudaFree(ctx->part_cursors); cudaFree(ctx->part_masks); cudaFree(ctx->pull_plan); cudaFree(ctx->tile_map_dev);
  delete ctx->tile_map_host;
  c3d_peer_destroy(&ctx->peer);
  cudaFree(ctx->table_mem);
  cudaFree(ctx->shade.work); cudaFree(ctx->shade.counters); cudaFree(ctx->shade.pix); cudaFree(ctx->shade.table); cudaFree(ctx->shade.vlist); cudaFree(ctx->shade.vn);
  cudaFree(ctx->canvas); cudaFree(ctx->depth_local); cudaFree(ctx->depth_global); cudaFree(ctx->hiz); cudaFree(ctx->pixel_xy); cudaFree(ctx->rgba); cudaFree(ctx->ppm_dev);
  if (ctx->ppm_host) cudaFreeHost(ctx->ppm_host);
  cudaFree(ctx->list); cudaFree(ctx->records); cudaFree(ctx->tile_cursor); cudaFree(ctx->bucket_hist); cudaFree(ctx->bucket_cursors); cudaFree(ctx->counts_dev);
  if (ctx->counts_host) cudaFreeHost(ctx->counts_host);
  cudaFree(ctx->views_dev); cudaFree(ctx->lines_dev); cudaFree(ctx->grids_dev); cudaFree(ctx->halo_buf); cudaFree(ctx->arena_dev);
  if (ctx->arena_host) cudaFreeHost(ctx->arena_host);
  for (int e = 0; e < EV_COUNT; e++) if (ctx->ev[e]) cudaEventDestroy(ctx->ev[e]);
  for (int e = 0; e <= C3D_MAX_OBJECTS; e++) if (ctx->ev_sweep[e]) cudaEventDestroy(ctx->ev_sweep[e]);
  for (int e = 0; e < 2 * C3D_MAX_OBJECTS; e++) if (ctx->ev_upload[e]) cudaEventDestroy(ctx->ev_upload[e]);
  for (int e = 0; e < C3D_DETAIL_EVENTS; e++) if (ctx->ev_detail[e]) cudaEventDestroy(ctx->ev_detail[e]);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  cudaGetLastError();
  free(ctx);
  return 0;
}

/* 0: one rank; 1: whole-image reduction through NCCL; 2: screen-partitioned over peer memory */
extern "C" int c3d_context_exchange_mode(const c3d_context_t *ctx) {
  if (NULL == ctx || ctx->size < 2) return 0;
  return ctx->peer.active ? 2 : 1;
}

extern "C" int c3d_context_device(const c3d_context_t *ctx) { return ctx ? ctx->device : -1; }
extern "C" uintptr_t c3d_context_stream(const c3d_context_t *ctx) { return ctx ? (uintptr_t)ctx->stream : 0; }
extern "C" int c3d_context_synchronize(c3d_context_t *ctx) {
  if (NULL == ctx) return c3d_error("NULL context");
  CUDA_OK(cudaSetDevice(ctx->device));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

/* ---------------------------------------------------------- decomposition */

extern "C" int c3d_decomp_make(const int dims[3], int comm_rank, c3d_decomp_t *out) {
  if (NULL == dims || NULL == out) return c3d_error("c3d_decomp_make: NULL argument");
  if (1 != dims[0]) return c3d_error("the x direction is never decomposed (dims[0] = %d)", dims[0]);
  if (dims[1] < 1 || dims[2] < 1) return c3d_error("bad process grid {%d, %d, %d}", dims[0], dims[1], dims[2]);
  const int size = dims[1] * dims[2];
  if (comm_rank < 0 || comm_rank >= size) return c3d_error("rank %d outside the %d-process grid", comm_rank, size);
  /* row-major coordinates, last dimension fastest (MPI_Cart_create) */
  const int cy = comm_rank / dims[2], cz = comm_rank % dims[2];
  for (int d = 0; d < 3; d++) out->nprocs[d] = dims[d];
  out->myrank[0] = 0; out->myrank[1] = cy; out->myrank[2] = cz;
  out->neighbours[0][0] = out->neighbours[0][1] = -1;
  out->neighbours[1][0] = cy > 0 ? (cy - 1) * dims[2] + cz : -1;
  out->neighbours[1][1] = cy < dims[1] - 1 ? (cy + 1) * dims[2] + cz : -1;
  out->neighbours[2][0] = cz > 0 ? cy * dims[2] + cz - 1 : -1;
  out->neighbours[2][1] = cz < dims[2] - 1 ? cy * dims[2] + cz + 1 : -1;
  return 0;
}

extern "C" int c3d_decomp_local(const c3d_decomp_t *d, const size_t glsizes[3], size_t mysizes[3], size_t offsets[3]) {
  if (NULL == d || NULL == glsizes || NULL == mysizes || NULL == offsets) return c3d_error("c3d_decomp_local: NULL argument");
  for (int dir = 0; dir < 3; dir++) {
    const size_t p = (size_t)d->nprocs[dir], r = (size_t)d->myrank[dir];
    if (p < 1 || r >= p) return c3d_error("bad decomposition in direction %d (%zu of %zu)", dir, r, p);
    if (glsizes[dir] < p) return c3d_error("glsize (%zu) should be equal to or more than nprocs (%zu)", glsizes[dir], p);
    /* sdecomp/kernel.c:89, 111-116 */
    mysizes[dir] = (glsizes[dir] + r) / p;
    size_t off = 0;
    for (size_t q = 0; q < r; q++) off += (glsizes[dir] + q) / p;
    offsets[dir] = off;
  }
  return 0;
}

/* ------------------------------------------------------------------ fields */

extern "C" int c3d_field_create(c3d_context_t *ctx, const c3d_decomp_t *decomp, const size_t glsizes[3], int elem_bytes, c3d_field_t **out) {
  if (NULL == ctx || NULL == decomp || NULL == glsizes || NULL == out) return c3d_error("c3d_field_create: NULL argument");
  *out = NULL;
  if (4 != elem_bytes && 8 != elem_bytes) return c3d_error("element size must be 4 or 8 bytes, received %d", elem_bytes);
  if (1 != decomp->nprocs[0]) return c3d_error("the x direction is never decomposed (nprocs[0] = %d)", decomp->nprocs[0]);
  if ((long long)decomp->nprocs[1] * decomp->nprocs[2] != ctx->size)
    return c3d_error("process grid %dx%dx%d does not match the %d-rank context", decomp->nprocs[0], decomp->nprocs[1], decomp->nprocs[2], ctx->size);
  c3d_field *f = (c3d_field *)calloc(1, sizeof(c3d_field));
  if (NULL == f) return c3d_error("out of host memory");
  f->decomp = *decomp;
  f->elem_bytes = elem_bytes;
  if (0 != c3d_decomp_local(decomp, glsizes, f->mysizes, f->offsets)) { free(f); return 1; }
  size_t total = 1;
  for (int d = 0; d < 3; d++) {
    f->gl[d] = glsizes[d];
    /* two extra points towards every neighbour, none at the domain boundary (contour/extend.c:216-252) */
    const size_t lo = decomp->myrank[d] > 0 ? 2 : 0;
    const size_t hi = decomp->myrank[d] < decomp->nprocs[d] - 1 ? 2 : 0;
    if ((lo || hi) && f->mysizes[d] < 2) {
      const size_t have = f->mysizes[d];
      free(f);
      return c3d_error("local size %zu in direction %d is too small for a 2-point halo", have, d);
    }
    f->lo[d] = lo;
    f->ext[d] = f->mysizes[d] + lo + hi;
    f->ext_off[d] = f->offsets[d] - lo;
    total *= f->ext[d];
  }
  CUDA_OK(cudaSetDevice(ctx->device));
  const cudaError_t err = cudaMalloc(&f->data, total * (size_t)elem_bytes);
  if (cudaSuccess != err) {
    cudaGetLastError();
    f->data = NULL;
    c3d_error("failed to allocate %zu bytes of device memory for a field: %s", total * (size_t)elem_bytes, cudaGetErrorString(err));
  }
  if (ctx->peer.active) {
    /* collective: every rank maps every other rank's array and learns its shape (a rank that could not allocate says so,
       and then all of them fail together instead of one waiting for the others for ever) */
    const int failed = c3d_peer_share(&ctx->peer, f->data, f->peer);
    FieldGeom mine;
    for (int d = 0; d < 3; d++) {
      mine.coords[d] = decomp->myrank[d];
      mine.mysizes[d] = (unsigned int)f->mysizes[d]; mine.offsets[d] = (unsigned int)f->offsets[d];
      mine.ext[d] = (unsigned int)f->ext[d]; mine.ext_off[d] = (unsigned int)f->ext_off[d]; mine.lo[d] = (unsigned int)f->lo[d];
    }
    if (0 != failed || 0 != c3d_peer_allgather(&ctx->peer, &mine, sizeof(mine), f->geom)) {
      if (NULL != f->data) { c3d_peer_unshare(&ctx->peer, f->peer); cudaFree(f->data); }
      free(f);
      return c3d_error("the ranks could not map each other's copy of a field");
    }
    f->shared = true;
  } else if (NULL == f->data) {
    free(f);
    return 1;
  }
  f->halo_valid = (1 == ctx->size);
  *out = f;
  return 0;
}

/* One rank, or a rank without neighbours: the array needs no halo room, so a DEVICE array of the caller can be swept where it
   lies.  Returns 0 and a field that borrows `device_array` (never written, never freed), or 2 when that does not apply
   (host pointer, another device, halo room needed) -- the caller then creates a field and uploads as usual. */
extern "C" int c3d_field_wrap(c3d_context_t *ctx, const c3d_decomp_t *decomp, const size_t glsizes[3], int elem_bytes,
                              const void *device_array, c3d_field_t **out) {
  if (NULL == ctx || NULL == decomp || NULL == glsizes || NULL == device_array || NULL == out) return c3d_error("c3d_field_wrap: NULL argument");
  *out = NULL;
  if (4 != elem_bytes && 8 != elem_bytes) return c3d_error("element size must be 4 or 8 bytes, received %d", elem_bytes);
  if (1 != ctx->size) return 2;
  for (int d = 0; d < 3; d++) {
    if (1 != decomp->nprocs[d]) return 2;
  }
  cudaPointerAttributes attr;
  if (cudaSuccess != cudaPointerGetAttributes(&attr, device_array)) { cudaGetLastError(); return 2; }
  if (cudaMemoryTypeDevice != attr.type || attr.device != ctx->device) return 2;
  c3d_field *f = (c3d_field *)calloc(1, sizeof(c3d_field));
  if (NULL == f) return c3d_error("out of host memory");
  f->decomp = *decomp;
  f->elem_bytes = elem_bytes;
  for (int d = 0; d < 3; d++) {
    f->gl[d] = f->mysizes[d] = f->ext[d] = glsizes[d];
    f->offsets[d] = f->ext_off[d] = f->lo[d] = 0;
  }
  f->data = const_cast<void *>(device_array);
  f->borrowed = true;
  f->halo_valid = true;
  *out = f;
  return 0;
}

extern "C" int c3d_field_destroy(c3d_context_t *ctx, c3d_field_t *field) {
  if (NULL == field) return 0;
  if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
  if (ctx && field->shared) c3d_peer_unshare(&ctx->peer, field->peer);
  if (!field->borrowed) cudaFree(field->data);
  free(field);
  return 0;
}

static inline char *interior_ptr(const c3d_field *f) {
  return (char *)f->data + ((f->lo[2] * f->ext[1] + f->lo[1]) * f->ext[0] + f->lo[0]) * (size_t)f->elem_bytes;
}

extern "C" int c3d_field_upload(c3d_context_t *ctx, c3d_field_t *f, const void *local) {
  if (NULL == ctx || NULL == f || NULL == local) return c3d_error("c3d_field_upload: NULL argument");
  if (f->borrowed) return c3d_error("c3d_field_upload: the field borrows the caller's array (c3d_field_wrap)");
  CUDA_OK(cudaSetDevice(ctx->device));
  const size_t eb = (size_t)f->elem_bytes;
  const bool timed = ctx->nuploads < C3D_MAX_OBJECTS;
  if (timed) CUDA_OK(cudaEventRecord(ctx->ev_upload[2 * ctx->nuploads], ctx->stream));
  if (f->mysizes[0] == f->ext[0] && f->mysizes[1] == f->ext[1]) {
    /* x and y undivided (one rank, or z slabs): the interior is one contiguous block -> a single flat copy,
       which runs at the link rate; the pitched 3-D form below pays per-row descriptors */
    CUDA_OK(cudaMemcpyAsync(interior_ptr(f), local, f->mysizes[0] * f->mysizes[1] * f->mysizes[2] * eb, cudaMemcpyDefault, ctx->stream));
  } else {
    cudaMemcpy3DParms p;
    memset(&p, 0, sizeof(p));
    p.srcPtr = make_cudaPitchedPtr((void *)local, f->mysizes[0] * eb, f->mysizes[0] * eb, f->mysizes[1]);
    p.dstPtr = make_cudaPitchedPtr(f->data, f->ext[0] * eb, f->ext[0] * eb, f->ext[1]);
    p.dstPos = make_cudaPos(f->lo[0] * eb, f->lo[1], f->lo[2]);
    p.extent = make_cudaExtent(f->mysizes[0] * eb, f->mysizes[1], f->mysizes[2]);
    p.kind = cudaMemcpyDefault;
    CUDA_OK(cudaMemcpy3DAsync(&p, ctx->stream));
  }
  if (timed) {
    CUDA_OK(cudaEventRecord(ctx->ev_upload[2 * ctx->nuploads + 1], ctx->stream));
    ctx->nuploads++;
  }
  f->halo_valid = (1 == ctx->size);
  return 0;
}

extern "C" int c3d_field_geometry(const c3d_field_t *f, size_t ext_sizes[3], size_t ext_offsets[3], size_t mysizes[3], size_t offsets[3]) {
  if (NULL == f) return c3d_error("c3d_field_geometry: NULL field");
  for (int d = 0; d < 3; d++) {
    if (ext_sizes) ext_sizes[d] = f->ext[d];
    if (ext_offsets) ext_offsets[d] = f->ext_off[d];
    if (mysizes) mysizes[d] = f->mysizes[d];
    if (offsets) offsets[d] = f->offsets[d];
  }
  return 0;
}

extern "C" int c3d_field_interior(const c3d_field_t *f, void **device_ptr, size_t strides_elems[2]) {
  if (NULL == f || NULL == device_ptr || NULL == strides_elems) return c3d_error("c3d_field_interior: NULL argument");
  *device_ptr = interior_ptr(f);
  strides_elems[0] = f->ext[0];
  strides_elems[1] = f->ext[0] * f->ext[1];
  return 0;
}

extern "C" int c3d_field_mark_dirty(c3d_field_t *f) {
  if (NULL == f) return c3d_error("c3d_field_mark_dirty: NULL field");
  f->halo_valid = false;
  return 0;
}

static int ensure_bytes(void **ptr, size_t *cap, const size_t need) {
  if (need <= *cap && NULL != *ptr) return 0;
  if (*ptr) CUDA_OK(cudaFree(*ptr));
  *ptr = NULL;
  *cap = 0;
  CUDA_OK(cudaMalloc(ptr, need));
  *cap = need;
  return 0;
}

/*
 * Halo exchange, y first and z second so that z planes carry the freshly
 * received y halos into the corners (same ordering argument as
 * contour/extend.c:211-213).  Domain-boundary ranks have no halo at all
 * (the reference exchanges and then clips, extend.c:216-278).
 *   y: 2 rows x nx per interior plane, strided -> pack kernel + ncclSend/Recv
 *   z: 2 full extended planes, contiguous      -> ncclSend/Recv in place
 */
/*
 * The same exchange over peer memory: every rank PULLS its halo out of the neighbours' arrays (rows of the neighbour's
 * interior for y, then -- after all y halos are in place -- two whole planes for z, which carry the fresh y halos into
 * the corners).  `synced` = a flag barrier already stands between the neighbours' last writes and this call.
 */
static int field_exchange_peer(c3d_context *ctx, c3d_field *f, const bool synced, uint32_t *launches) {
  cudaStream_t s = ctx->stream;
  if (!synced) { LAUNCH_OK(c3d_peer_barrier(&ctx->peer, s)); *launches += 1; }
  const long long nx = (long long)f->ext[0], sy = (long long)f->ext[0], sz = (long long)(f->ext[0] * f->ext[1]);
  const bool ysplit = f->decomp.neighbours[1][0] >= 0 || f->decomp.neighbours[1][1] >= 0 || f->decomp.nprocs[1] > 1;
  const bool zsplit = f->decomp.nprocs[2] > 1;
  for (int side = 0; side < 2; side++) {
    const int nb = f->decomp.neighbours[1][side];
    if (nb < 0) continue;
    const FieldGeom &g = f->geom[nb];
    /* the two rows next to the shared face, planes of the common interior (same z block) */
    const long long their_j0 = 0 == side ? (long long)g.lo[1] + g.mysizes[1] - 2 : (long long)g.lo[1];
    const long long my_j0 = 0 == side ? 0 : (long long)(f->lo[1] + f->mysizes[1]);
    LAUNCH_OK(c3d_k_halo_pull_y(f->data, f->peer[nb], f->elem_bytes, nx, sy, sz, (long long)g.ext[0], (long long)g.ext[0] * g.ext[1],
                                (long long)f->lo[2], (long long)g.lo[2], (long long)f->mysizes[2], my_j0, their_j0, s));
    *launches += 1;
  }
  if (zsplit) {
    if (ysplit) { LAUNCH_OK(c3d_peer_barrier(&ctx->peer, s)); *launches += 1; }       /* everybody's y halos are in place */
    for (int side = 0; side < 2; side++) {
      const int nb = f->decomp.neighbours[2][side];
      if (nb < 0) continue;
      const FieldGeom &g = f->geom[nb];
      const size_t their_plane = (size_t)g.ext[0] * g.ext[1], my_plane = (size_t)sz;
      if (their_plane != my_plane) return c3d_error("z neighbour %d holds planes of another shape", nb);
      const size_t their_k0 = 0 == side ? (size_t)g.lo[2] + g.mysizes[2] - 2 : (size_t)g.lo[2];
      const size_t my_k0 = 0 == side ? 0 : f->lo[2] + f->mysizes[2];
      LAUNCH_OK(c3d_k_halo_pull_z((char *)f->data + my_k0 * my_plane * (size_t)f->elem_bytes,
                                  (const char *)f->peer[nb] + their_k0 * their_plane * (size_t)f->elem_bytes, f->elem_bytes, 2 * my_plane, s));
      *launches += 1;
    }
  }
  f->halo_valid = true;
  return 0;
}

extern "C" int c3d_field_exchange(c3d_context_t *ctx, c3d_field_t *f) {
  if (NULL == ctx || NULL == f) return c3d_error("c3d_field_exchange: NULL argument");
  if (1 == ctx->size) { f->halo_valid = true; return 0; }
  CUDA_OK(cudaSetDevice(ctx->device));
  if (f->shared) {
    /* a barrier on both sides: the neighbours' arrays are complete before, and stay untouched until everybody has read */
    uint32_t launches = 0;
    if (0 != field_exchange_peer(ctx, f, false, &launches)) return 1;
    LAUNCH_OK(c3d_peer_barrier(&ctx->peer, ctx->stream));
    return 0;
  }
  if (!ctx->has_comm) return c3d_error("no communicator: the halo cannot be exchanged");
  const size_t eb = (size_t)f->elem_bytes;
  const long long nx = (long long)f->ext[0], sy = (long long)f->ext[0], sz = (long long)(f->ext[0] * f->ext[1]);
  const ncclDataType_t dt = 4 == f->elem_bytes ? ncclFloat : ncclDouble;
  {
    const int lo = f->decomp.neighbours[1][0], hi = f->decomp.neighbours[1][1];
    if (lo >= 0 || hi >= 0) {
      const long long nk = (long long)f->mysizes[2], k0 = (long long)f->lo[2];
      const size_t chunk = (size_t)nk * 2 * (size_t)nx;           /* elements per message */
      if (0 != ensure_bytes(&ctx->halo_buf, &ctx->halo_cap, 4 * chunk * eb)) return 1;
      char *buf = (char *)ctx->halo_buf;
      void *send_lo = buf, *send_hi = buf + chunk * eb, *recv_lo = buf + 2 * chunk * eb, *recv_hi = buf + 3 * chunk * eb;
      const long long jfirst = (long long)f->lo[1], jlast = (long long)(f->lo[1] + f->mysizes[1]) - 2;
      if (lo >= 0) LAUNCH_OK(c3d_k_halo_pack_y(f->data, f->elem_bytes, nx, sy, sz, k0, nk, jfirst, send_lo, ctx->stream));
      if (hi >= 0) LAUNCH_OK(c3d_k_halo_pack_y(f->data, f->elem_bytes, nx, sy, sz, k0, nk, jlast, send_hi, ctx->stream));
      NCCL_OK(g_nccl.GroupStart());
      if (lo >= 0) { NCCL_OK(g_nccl.Send(send_lo, chunk, dt, lo, ctx->comm, ctx->stream)); NCCL_OK(g_nccl.Recv(recv_lo, chunk, dt, lo, ctx->comm, ctx->stream)); }
      if (hi >= 0) { NCCL_OK(g_nccl.Send(send_hi, chunk, dt, hi, ctx->comm, ctx->stream)); NCCL_OK(g_nccl.Recv(recv_hi, chunk, dt, hi, ctx->comm, ctx->stream)); }
      NCCL_OK(g_nccl.GroupEnd());
      if (lo >= 0) LAUNCH_OK(c3d_k_halo_unpack_y(f->data, f->elem_bytes, nx, sy, sz, k0, nk, 0, recv_lo, ctx->stream));
      if (hi >= 0) LAUNCH_OK(c3d_k_halo_unpack_y(f->data, f->elem_bytes, nx, sy, sz, k0, nk, (long long)(f->lo[1] + f->mysizes[1]), recv_hi, ctx->stream));
    }
  }
  {
    const int lo = f->decomp.neighbours[2][0], hi = f->decomp.neighbours[2][1];
    if (lo >= 0 || hi >= 0) {
      const size_t plane = (size_t)sz, chunk = 2 * plane;
      char *base = (char *)f->data;
      const size_t kfirst = f->lo[2], klast = f->lo[2] + f->mysizes[2] - 2;
      NCCL_OK(g_nccl.GroupStart());
      if (lo >= 0) {
        NCCL_OK(g_nccl.Send(base + kfirst * plane * eb, chunk, dt, lo, ctx->comm, ctx->stream));
        NCCL_OK(g_nccl.Recv(base, chunk, dt, lo, ctx->comm, ctx->stream));
      }
      if (hi >= 0) {
        NCCL_OK(g_nccl.Send(base + klast * plane * eb, chunk, dt, hi, ctx->comm, ctx->stream));
        NCCL_OK(g_nccl.Recv(base + (f->lo[2] + f->mysizes[2]) * plane * eb, chunk, dt, hi, ctx->comm, ctx->stream));
      }
      NCCL_OK(g_nccl.GroupEnd());
    }
  }
  f->halo_valid = true;
  return 0;
}

extern "C" int c3d_field_download_ext(c3d_context_t *ctx, const c3d_field_t *f, void *host) {
  if (NULL == ctx || NULL == f || NULL == host) return c3d_error("c3d_field_download_ext: NULL argument");
  CUDA_OK(cudaSetDevice(ctx->device));
  CUDA_OK(cudaMemcpyAsync(host, f->data, f->ext[0] * f->ext[1] * f->ext[2] * (size_t)f->elem_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int c3d_field_synthesize(c3d_context_t *ctx, c3d_field_t *f, int kind, uint32_t seed, int octaves, double gain) {
  if (NULL == ctx || NULL == f) return c3d_error("c3d_field_synthesize: NULL argument");
  if (kind < 0 || kind > 2) return c3d_error("unknown synthetic field kind %d", kind);
  CUDA_OK(cudaSetDevice(ctx->device));
  /* the generator is a function of the GLOBAL index, so it can fill the halo too: no exchange needed */
  const long long n[3] = {(long long)f->ext[0], (long long)f->ext[1], (long long)f->ext[2]};
  const long long off[3] = {(long long)f->ext_off[0], (long long)f->ext_off[1], (long long)f->ext_off[2]};
  const long long gl[3] = {(long long)f->gl[0], (long long)f->gl[1], (long long)f->gl[2]};
  LAUNCH_OK(c3d_k_synth(f->data, f->elem_bytes, n, n[0], n[0] * n[1], off, gl, kind, seed, octaves, gain, ctx->stream));
  f->halo_valid = true;
  return 0;
}

/* ------------------------------------------------------------------- frame */

static float smallest_float_above(const double thr) {
  /* smallest float f with (double)f > thr: then, for any float v, thr < (double)v <=> v >= f */
  if (thr >= (double)FLT_MAX) return INFINITY;
  if (thr < -(double)FLT_MAX) return -FLT_MAX;
  float f = (float)thr;
  if ((double)f > thr) {
    const float below = nextafterf(f, -INFINITY);
    return (double)below > thr ? below : f;
  }
  return nextafterf(f, INFINITY);
}

static size_t grid_slot_doubles(const c3d_field *f) {
  const size_t total = f->ext[0] + f->ext[1] + f->ext[2];
  return total + (total + 1) / 2 + 1;
}

/* grid_slot: room for sum(ext) doubles followed by sum(ext) floats */
static int make_view(c3d_context *ctx, const c3d_field *f, const double *const grids[3], const double threshold,
                     const uint8_t color[3], double *grid_slot, FieldView *v) {
  v->data = f->data;
  v->elem_bytes = f->elem_bytes;
  for (int d = 0; d < 3; d++) v->n[d] = (long long)f->ext[d];
  v->sy = (long long)f->ext[0];
  v->sz = (long long)(f->ext[0] * f->ext[1]);
  const size_t total = f->ext[0] + f->ext[1] + f->ext[2];
  double *at = grid_slot;
  float *atf = (float *)(grid_slot + total);
  for (int d = 0; d < 3; d++) {
    if (NULL == grids || NULL == grids[d]) return c3d_error("grid coordinates of direction %d are missing", d);
    CUDA_OK(cudaMemcpyAsync(at, grids[d] + f->ext_off[d], f->ext[d] * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    v->g[d] = at;
    v->gf[d] = atf;
    at += f->ext[d];
    atf += f->ext[d];
  }
  LAUNCH_OK(c3d_k_to_float(grid_slot, (float *)(grid_slot + total), total, ctx->stream));
  v->map = NULL;
  for (int d = 0; d < 3; d++) v->off[d] = (long long)f->ext_off[d];
  v->thr = threshold;
  v->thr_ge = smallest_float_above(threshold);
  v->color[0] = color[0]; v->color[1] = color[1]; v->color[2] = color[2];
  return 0;
}

static size_t round_up(const size_t v, const size_t a) { return (v + a - 1) / a * a; }

/* entries of the converter table of component c (c3d_positions_t, include/contour3d_b200.h) */
static size_t table_length(const c3d_field *f, const c3d_positions_t *pos, const int c) {
  size_t len = 1;
  for (int a = 0; a < 3; a++) len += (f->ext[a] - 1) * pos->stride[c][a];
  return len;
}

/*
 * Everything the kernels read besides the arrays, in ONE device arena filled with as few copies as possible
 * (a dozen small pageable copies cost more host time than the clear of a 4K canvas):
 *   [PosMap x nc][LineDev x nl][FieldView x ranks x 16 (partitioned frames)]
 *   [doubles: grid coordinates of every object | converter tables | converted line points]
 *   [floats: the same grid coordinates and tables rounded, for the visibility pre-tests]
 * The maps, lines, peer views and grid coordinates are staged in pinned memory and travel in one copy; tables and line
 * points (arbitrarily large, caller-owned) are copied from where they lie.
 * part: a screen-partitioned frame -- the GLOBAL coordinates are uploaded (every rank indexes them with the offsets of
 * whichever rank's array a cube comes from) and *pviews receives the table of those arrays as seen through peer memory.
 */
static int stage_frame(c3d_context *ctx, const size_t nc, const c3d_contour_t *contours, const size_t nl, const c3d_line_t *lines,
                       std::vector<LineDev> &hl, FieldView *views, const LineDev **lines_dev, const bool part, const FieldView **pviews) {
  cudaStream_t s = ctx->stream;
  const size_t maps_bytes = round_up(nc * sizeof(PosMap), 256), lines_bytes = round_up(nl * sizeof(LineDev), 256);
  const size_t pviews_bytes = part ? round_up((size_t)C3D_MAX_RANKS * 16 * sizeof(FieldView), 256) : 0;
  size_t grid_doubles = 0, table_doubles = 0, point_doubles = 0;
  for (size_t n = 0; n < nc; n++) {
    const c3d_field *f = contours[n].field;
    for (int d = 0; d < 3; d++) {
      if ((long long)f->ext[d] - 1 > C3D_MAX_AXIS) return c3d_error("extended array too large in direction %d (%zu points)", d, f->ext[d]);
      if ((long long)f->gl[d] - 1 > C3D_MAX_AXIS) return c3d_error("array too large in direction %d (%zu points; winner keys hold 16-bit cube indices)", d, f->gl[d]);
      if (NULL == contours[n].grids[d]) return c3d_error("grid coordinates of direction %d are missing", d);
    }
    grid_doubles += part ? f->gl[0] + f->gl[1] + f->gl[2] : f->ext[0] + f->ext[1] + f->ext[2];
    if (NULL != contours[n].positions) {
      for (int c = 0; c < 3; c++) {
        if (NULL == contours[n].positions->table[c]) return c3d_error("contour object %zu: converter table %d is missing", n, c);
        table_doubles += table_length(f, contours[n].positions, c);
      }
    }
  }
  for (size_t n = 0; n < nl; n++) {
    if (NULL != lines[n].points) point_doubles += 3 * lines[n].nitems;
  }
  const size_t doubles_at = maps_bytes + lines_bytes + pviews_bytes;
  const size_t floats_at = round_up(doubles_at + (grid_doubles + table_doubles + point_doubles) * sizeof(double), 256);
  const size_t total = floats_at + (grid_doubles + table_doubles) * sizeof(float) + 256;
  const size_t staged = doubles_at + grid_doubles * sizeof(double);
  if (0 != ensure_bytes((void **)&ctx->arena_dev, &ctx->arena_cap, total)) return 1;
  if (staged > ctx->arena_host_cap) {
    if (ctx->arena_host) cudaFreeHost(ctx->arena_host);
    ctx->arena_host = NULL;
    ctx->arena_host_cap = 0;
    CUDA_OK(cudaMallocHost(&ctx->arena_host, staged + 4096));
    ctx->arena_host_cap = staged + 4096;
  }
  PosMap *maps_host = (PosMap *)ctx->arena_host;
  PosMap *maps_dev = (PosMap *)ctx->arena_dev;
  FieldView *pviews_host = (FieldView *)(ctx->arena_host + maps_bytes + lines_bytes);
  if (part) memset(pviews_host, 0, pviews_bytes);
  double *dbl_host = (double *)(ctx->arena_host + doubles_at);
  double *dbl_dev = (double *)(ctx->arena_dev + doubles_at);
  float *flt_dev = (float *)(ctx->arena_dev + floats_at);
  size_t at = 0, tab_at = grid_doubles;
  for (size_t n = 0; n < nc; n++) {
    const c3d_field *f = contours[n].field;
    FieldView *v = &views[n];
    v->data = f->data;
    v->elem_bytes = f->elem_bytes;
    for (int d = 0; d < 3; d++) v->n[d] = (long long)f->ext[d];
    v->sy = (long long)f->ext[0];
    v->sz = (long long)(f->ext[0] * f->ext[1]);
    for (int d = 0; d < 3; d++) {
      if (part) {
        memcpy(dbl_host + at, contours[n].grids[d], f->gl[d] * sizeof(double));
        v->g[d] = dbl_dev + at + f->ext_off[d];
        v->gf[d] = flt_dev + at + f->ext_off[d];
        for (int r = 0; r < ctx->size; r++) {
          FieldView *pv = &pviews_host[C3D_PEER_VIEW(r, n)];
          pv->g[d] = dbl_dev + at + f->geom[r].ext_off[d];
          pv->gf[d] = flt_dev + at + f->geom[r].ext_off[d];
        }
        at += f->gl[d];
      } else {
        memcpy(dbl_host + at, contours[n].grids[d] + f->ext_off[d], f->ext[d] * sizeof(double));
        v->g[d] = dbl_dev + at;
        v->gf[d] = flt_dev + at;
        at += f->ext[d];
      }
      v->off[d] = (long long)f->ext_off[d];
    }
    v->map = NULL;
    if (NULL != contours[n].positions) {
      const c3d_positions_t *pos = contours[n].positions;
      PosMap *m = &maps_host[n];
      for (int c = 0; c < 3; c++) {
        const size_t len = table_length(f, pos, c);
        CUDA_OK(cudaMemcpyAsync(dbl_dev + tab_at, pos->table[c], len * sizeof(double), cudaMemcpyHostToDevice, s));
        m->pos[c] = dbl_dev + tab_at;
        m->posf[c] = flt_dev + tab_at;
        for (int a = 0; a < 3; a++) m->ps[c][a] = (long long)pos->stride[c][a];
        tab_at += len;
      }
      v->map = &maps_dev[n];
    }
    v->thr = contours[n].threshold;
    v->thr_ge = smallest_float_above(contours[n].threshold);
    v->color[0] = contours[n].color[0]; v->color[1] = contours[n].color[1]; v->color[2] = contours[n].color[2];
    if (part) {
      /* the same object as each rank holds it: that rank's block of the array, through peer memory */
      for (int r = 0; r < ctx->size; r++) {
        FieldView *pv = &pviews_host[C3D_PEER_VIEW(r, n)];
        const FieldGeom &g = f->geom[r];
        pv->data = f->peer[r];
        pv->elem_bytes = f->elem_bytes;
        for (int d = 0; d < 3; d++) { pv->n[d] = (long long)g.ext[d]; pv->off[d] = (long long)g.ext_off[d]; }
        pv->sy = (long long)g.ext[0];
        pv->sz = (long long)g.ext[0] * (long long)g.ext[1];
        pv->map = NULL;
        pv->thr = v->thr; pv->thr_ge = v->thr_ge;
        pv->color[0] = v->color[0]; pv->color[1] = v->color[1]; pv->color[2] = v->color[2];
      }
    }
  }
  /* lines: deltas = (1 / (npoints - 1)) * (tail - head), line.c:138-140 */
  LineDev *lines_host = (LineDev *)(ctx->arena_host + maps_bytes);
  unsigned long long first = 0;
  size_t pts_at = grid_doubles + table_doubles;
  for (size_t n = 0; n < nl; n++) {
    const c3d_line_t *l = &lines[n];
    if (l->nitems < 2) return c3d_error("give more than two points to draw a line, received %zu", l->nitems);
    const double inv = 1. / (double)(l->nitems - 1);
    for (int d = 0; d < 3; d++) {
      hl[n].e0[d] = l->edges[0][d];
      hl[n].delta[d] = inv * (l->edges[1][d] - l->edges[0][d]);
    }
    hl[n].half_width_sq = pow(0.5 * l->width, 2.);
    hl[n].points = NULL;
    if (NULL != l->points) {
      CUDA_OK(cudaMemcpyAsync(dbl_dev + pts_at, l->points, 3 * l->nitems * sizeof(double), cudaMemcpyHostToDevice, s));
      hl[n].points = dbl_dev + pts_at;
      pts_at += 3 * l->nitems;
    }
    hl[n].first_segment = first;
    hl[n].nsegments = l->nitems - 1;
    hl[n].color[0] = l->color[0]; hl[n].color[1] = l->color[1]; hl[n].color[2] = l->color[2];
    first += l->nitems - 1;
    lines_host[n] = hl[n];
  }
  *lines_dev = (const LineDev *)(ctx->arena_dev + maps_bytes);
  *pviews = part ? (const FieldView *)(ctx->arena_dev + maps_bytes + lines_bytes) : NULL;
  if (staged > 0) CUDA_OK(cudaMemcpyAsync(ctx->arena_dev, ctx->arena_host, staged, cudaMemcpyHostToDevice, s));
  LAUNCH_OK(c3d_k_to_float(dbl_dev, flt_dev, grid_doubles + table_doubles, s));
  return 0;
}

static void rendered_range(const FieldView *v, const int rendered_only, SweepRange *r) {
  /* cubes 1..m-3 per direction are rendered (contour/main.c:71,92-93); all cubes are 0..m-2 */
  for (int d = 0; d < 3; d++) {
    r->lo[d] = rendered_only ? 1 : 0;
    r->hi[d] = rendered_only ? v->n[d] - 3 : v->n[d] - 2;
  }
}

static int ensure_canvas(c3d_context *ctx, const size_t w, const size_t h) {
  const size_t npix = w * h;
  if (npix > ctx->npix_cap) {
    cudaFree(ctx->canvas); cudaFree(ctx->depth_local); cudaFree(ctx->depth_global); cudaFree(ctx->rgba); cudaFree(ctx->ppm_dev);
    cudaFree(ctx->shade.work); cudaFree(ctx->shade.pix); cudaFree(ctx->shade.table); cudaFree(ctx->shade.vlist); cudaFree(ctx->shade.vn);
    cudaFree(ctx->hiz);
    ctx->hiz = NULL;
    ctx->shade.work = NULL; ctx->shade.pix = NULL; ctx->shade.table = NULL; ctx->shade.vlist = NULL; ctx->shade.vn = NULL;
    if (ctx->ppm_host) cudaFreeHost(ctx->ppm_host);
    ctx->canvas = NULL; ctx->depth_local = ctx->depth_global = NULL; ctx->rgba = NULL; ctx->ppm_dev = NULL; ctx->ppm_host = NULL;
    ctx->npix_cap = 0;
    CUDA_OK(cudaMalloc(&ctx->canvas, npix * sizeof(Pixel)));
    if (ctx->size > 1) {
      CUDA_OK(cudaMalloc(&ctx->depth_local, npix * 8));
      CUDA_OK(cudaMalloc(&ctx->depth_global, npix * 8));
    }
    CUDA_OK(cudaMalloc(&ctx->rgba, npix * 4));
    /* one word per pixel is an upper bound for any screen shape (a tile row / column may be partial) */
    CUDA_OK(cudaMalloc(&ctx->hiz, (npix / 8 + w + h + 64) * sizeof(unsigned long long)));
    CUDA_OK(cudaMalloc(&ctx->shade.work, npix * sizeof(WorkItem)));
    CUDA_OK(cudaMalloc(&ctx->shade.pix, npix * C3D_SHADE_PIX_BYTES));
    {
      /* four to eight slots per pixel: at the benchmark's resolution triangles are smaller than a pixel and 3.7 M winning
         pixels name 6 M distinct vertices (load 0.18: probe sequences stay short, each probe is an atomic round trip);
         3 per pixel is the worst case.  A vertex that finds no slot within SH_PROBES, or no room in the normal array
         (1.5 per pixel), is computed per pixel instead (k_shade_blend) */
      size_t slots = 1u << 16;
      while (slots < 4 * npix && slots < (1ull << 28)) slots <<= 1;
      /* the list is handed out in 256-entry chunks per warp (k_shade_setup): room for every resident warp's unused tail */
      /* a chunk holds at least 225 useful entries (a warp abandons at most 31 when its next 32 do not fit), the last chunk of
         every warp may stay almost empty: entries <= vertices * 256 / 225 + warps * 256 */
      const size_t most = std::min<size_t>(slots, 3 * npix + 32);
      const size_t listed = most + most / 7 + (size_t)ctx->sm_count * 64 * 256;
      const size_t normals = std::min<size_t>(listed, npix + npix / 2 + (size_t)ctx->sm_count * 64 * 256);
      CUDA_OK(cudaMalloc(&ctx->shade.table, slots * sizeof(unsigned long long)));
      CUDA_OK(cudaMalloc(&ctx->shade.vlist, listed * sizeof(unsigned int)));
      CUDA_OK(cudaMalloc(&ctx->shade.vn, normals * 3 * sizeof(double)));
      ctx->shade.vn_cap = (unsigned int)normals;
      {
        /* test hook: a tiny normal array pushes vertices onto the per-pixel path of k_shade_blend */
        const char *env = getenv("CONTOUR3D_SHADE_NORMALS_CAP");
        if (NULL != env && '\0' != *env && (size_t)strtoull(env, NULL, 10) < normals) ctx->shade.vn_cap = (unsigned int)strtoull(env, NULL, 10);
      }
      ctx->shade.table_mask = (unsigned int)(slots - 1);
      ctx->shade.table_slots = (unsigned int)slots;
    }
    CUDA_OK(cudaMalloc(&ctx->ppm_dev, npix * 3));
    CUDA_OK(cudaMallocHost(&ctx->ppm_host, npix * 3));
    ctx->npix_cap = npix;
  }
  if (w != ctx->table_w || h != ctx->table_h) {
    cudaFree(ctx->pixel_xy);
    ctx->pixel_xy = NULL;
    ctx->table_w = ctx->table_h = 0;
    /* doubles for w columns + h rows, then the same as floats */
    CUDA_OK(cudaMalloc(&ctx->pixel_xy, (w + h) * (sizeof(double) + sizeof(float))));
    float *as_float = (float *)(ctx->pixel_xy + w + h);
    LAUNCH_OK(c3d_k_pixel_tables(ctx->pixel_xy, ctx->pixel_xy + w, as_float, as_float + w, (long long)w, (long long)h, ctx->stream));
    ctx->table_w = w;
    ctx->table_h = h;
  }
  ctx->width = w;
  ctx->height = h;
  return 0;
}

static void make_cam(const c3d_scene_t *s, Cam *c) {
  /* init_light / init_screen, src/contour3d/main.c:24-63: host fp64 (x86-64 without -march: no FMA) */
  auto unit = [](const double v[3], double out[3]) {
    const double inv = 1. / sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    out[0] = v[0] * inv; out[1] = v[1] * inv; out[2] = v[2] * inv;
  };
  auto dot = [](const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
  for (int d = 0; d < 3; d++) {
    c->cam[d] = s->camera[d]; c->center[d] = s->center[d]; c->lx[d] = s->local_x[d]; c->ly[d] = s->local_y[d];
  }
  unit(s->light, c->light);
  const double cross[3] = {
    s->local_x[1] * s->local_y[2] - s->local_x[2] * s->local_y[1],
    s->local_x[2] * s->local_y[0] - s->local_x[0] * s->local_y[2],
    s->local_x[0] * s->local_y[1] - s->local_x[1] * s->local_y[0],
  };
  unit(cross, c->normal);
  c->lxlx = dot(c->lx, c->lx);
  c->lyly = dot(c->ly, c->ly);
  c->nc = dot(c->normal, c->cam);
  c->d = dot(c->normal, c->center);
  c->W = (double)s->width;
  c->H = (double)s->height;
  c->w = (long long)s->width;
  c->h = (long long)s->height;
}

/* fp32 projection constants of cube_hidden() (c3d_kernels.cu) */
static void make_cull(const Cam *c, CullCam *q) {
  const double cx = (c->cam[0] - c->center[0]) * c->lx[0] + (c->cam[1] - c->center[1]) * c->lx[1] + (c->cam[2] - c->center[2]) * c->lx[2];
  const double cy = (c->cam[0] - c->center[0]) * c->ly[0] + (c->cam[1] - c->center[1]) * c->ly[1] + (c->cam[2] - c->center[2]) * c->ly[2];
  for (int d = 0; d < 3; d++) {
    q->cam[d] = (float)c->cam[d];
    q->n[d] = (float)c->normal[d];
    q->lxs[d] = (float)(c->lx[d] / c->lxlx);
    q->lys[d] = (float)(c->ly[d] / c->lyly);
  }
  q->ax = (float)(cx / c->lxlx);
  q->ay = (float)(cy / c->lyly);
  q->c = (float)(c->d - c->nc);
  q->inv_abs_c = (float)(1. / fabs(c->d - c->nc));
  q->lx_norm = (float)(sqrt(c->lxlx) / c->lxlx * 1.000001);
  q->ly_norm = (float)(sqrt(c->lyly) / c->lyly * 1.000001);
  q->W = (float)c->W;
  q->H = (float)c->H;
  q->w = (int)c->w;
  q->h = (int)c->h;
  const char *env = getenv("CONTOUR3D_NO_CULL");
  const bool off = NULL != env && '\0' != *env && '0' != *env;
  /* needs finite, non-degenerate constants and a screen that fits the int / float arithmetic of the test */
  q->valid = fabs(c->d - c->nc) > 1.e-9 && isfinite(q->ax) && isfinite(q->ay) && c->w < 65535 && c->h < 65535;   /* pixel boxes travel as 16-bit fields (C3D_NO_BOX, c3d_kernels.cu) */
  q->enabled = !off && q->valid;
}

/* ---- screen-partitioned frames: the memory the ranks share ---- */

/* tiles of a w x h screen owned by rank r (the pattern of tile_owner(), c3d_internal.h) */
static size_t own_tiles(const int nranks, const int r, const int tiles_x, const int tiles_y) {
  size_t n = 0;
  for (int ty = 0; ty < tiles_y; ty++) {
    for (int tx = 0; tx < tiles_x; tx++) n += ((tx + 3 * ty) % nranks == r) ? 1 : 0;
  }
  return n;
}

#define C3D_SHARE_CAP 1.75

/* the tile -> rank table for the current shares: tiles are visited in a scattered order (a stride coprime to their number) and
   each goes to the rank that is furthest behind its share, so every rank's tiles are spread over the whole screen */
static void tile_map_build(const double *share, const int n, const int tiles_x, const int tiles_y, std::vector<unsigned char> &map) {
  const long long nt = (long long)tiles_x * tiles_y;
  map.assign((size_t)nt, 0);
  long long stride = (long long)((double)nt * 0.6180339887498949) | 1ll;
  auto gcd = [](long long a, long long b) { while (b) { const long long t = a % b; a = b; b = t; } return a; };
  while (stride > 1 && 1 != gcd(stride, nt)) stride += 2;
  if (stride >= nt) stride = 1;
  double credit[C3D_MAX_RANKS] = {0.};
  for (long long u = 0; u < nt; u++) {
    const long long t = (u * stride) % nt;
    int best = 0;
    for (int r = 0; r < n; r++) {
      credit[r] += share[r];
      if (credit[r] > credit[best]) best = r;
    }
    credit[best] -= 1.;
    map[(size_t)t] = (unsigned char)best;
  }
}

/* host-side self test hook (needs no device): the table tile_map_build() derives from `share` */
extern "C" int c3d_selftest_tile_map(const double *share, int n, int tiles_x, int tiles_y, unsigned char *out) {
  if (NULL == share || NULL == out || n < 1 || n > C3D_MAX_RANKS || tiles_x < 1 || tiles_y < 1) return 1;
  std::vector<unsigned char> map;
  tile_map_build(share, n, tiles_x, tiles_y, map);
  memcpy(out, map.data(), map.size());
  return 0;
}

/* New shares from what every rank's frame spent drawing (d_r, proportional to its share of the tiles) and shading (h_r, given by
   whose cubes win the image -- not by the tiles): the level L with  D s_r + h_r = L,  sum s_r = 1,  D = sum d_r.  Damped, clamped. */
static bool tile_shares_update(c3d_context *ctx, const unsigned long long *gathered) {
  const int n = ctx->size;
  double d[C3D_MAX_RANKS], h[C3D_MAX_RANKS], D = 0., H = 0.;
  for (int r = 0; r < n; r++) {
    float two[2];
    memcpy(two, &gathered[r], sizeof(two));
    if (!(two[0] > 0.f) || !(two[1] >= 0.f) || !(two[0] < 1.e6f) || !(two[1] < 1.e6f)) return false;     /* somebody has no measurement yet */
    d[r] = two[0]; h[r] = two[1];
    D += d[r]; H += h[r];
  }
  const double level = (D + H) / (double)n;
  double want[C3D_MAX_RANKS], sum = 0., worst = 0.;
  for (int r = 0; r < n; r++) {
    double s = (level - h[r]) / D;
    s = 0.5 * ctx->tile_share[r] + 0.5 * s;
    s = std::min(std::max(s, 0.25 / (double)n), C3D_SHARE_CAP / (double)n);
    want[r] = s;
    sum += s;
  }
  for (int r = 0; r < n; r++) {
    want[r] /= sum;
    worst = std::max(worst, fabs(want[r] - ctx->tile_share[r]) * (double)n);
  }
  if (worst < 0.03) return false;                    /* within 3 % of an even share: leave the table alone */
  for (int r = 0; r < n; r++) ctx->tile_share[r] = want[r];
  return true;
}

/* collective: (re)creates this rank's exchange arena and maps everybody's */
static int ensure_exchange(c3d_context *ctx, const size_t w, const size_t h, const size_t out_cap) {
  Exchange &x = ctx->xa;
  if (NULL != x.base && !x.stale && x.width == w && x.height == h && x.nranks == ctx->size) return 0;
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  const size_t keep_out = (NULL != x.base && x.out_cap > out_cap) ? x.out_cap : out_cap;
  release_exchange(ctx);
  const int n = ctx->size;
  const int tiles_x = (int)((w + ((size_t)1 << ctx->tile_shift) - 1) >> ctx->tile_shift), tiles_y = (int)((h + ((size_t)1 << ctx->tile_shift) - 1) >> ctx->tile_shift);
  size_t most = 0;
  for (int r = 0; r < n; r++) most = std::max(most, own_tiles(n, r, tiles_x, tiles_y));
  /* the shares move with the load (tile_shares_update): at most C3D_SHARE_CAP times the even share */
  most = std::max(most, (size_t)(C3D_SHARE_CAP * (double)tiles_x * (double)tiles_y / (double)n) + 2);
  most = std::min(most, (size_t)tiles_x * (size_t)tiles_y);
  x.nranks = n;
  x.width = w; x.height = h; x.npix = w * h;
  x.seg_cap = most << (2 * ctx->tile_shift);          /* every pixel of my tiles may be won by cubes of ONE rank */
  x.out_cap = keep_out;
  size_t at = 0;
  x.off_flags = at; at += 256;
  x.off_seg_start = at; at += 256;
  x.off_nseg = at; at += 256;
  x.off_hist = at; at += round_up((size_t)n * C3D_PART_BUCKETS * sizeof(unsigned int), 256);
  x.off_rgba = at; at += round_up(x.npix * 4, 256);
  x.off_seg = at; at += round_up((size_t)n * x.seg_cap * sizeof(WorkItem), 256);
  x.off_out = at; at += x.out_cap * 8 + 256;     /* readers fetch aligned pairs: room for one entry beyond the last */
  x.bytes = at;
  void *base = NULL;
  const cudaError_t err = cudaMalloc(&base, x.bytes);
  if (cudaSuccess != err) { cudaGetLastError(); base = NULL; c3d_error("failed to allocate %zu bytes for the ranks' exchange buffers: %s", x.bytes, cudaGetErrorString(err)); }
  if (NULL != base) {
    cudaMemset(base, 0, x.off_hist);
    const unsigned long long cap = (unsigned long long)x.out_cap;       /* readers of my outbox stop here (flags word 9) */
    cudaMemcpy((char *)base + x.off_flags + 72, &cap, sizeof(cap), cudaMemcpyHostToDevice);
  }
  if (0 != c3d_peer_share(&ctx->peer, base, x.peer)) {
    if (NULL != base) cudaFree(base);
    memset(&x, 0, sizeof(x));
    return c3d_error("the ranks could not map each other's exchange buffers");
  }
  x.base = (char *)base;
  x.stale = false;
  return 0;
}

static bool frame_can_partition(const c3d_context *ctx, const size_t nc, const c3d_contour_t *contours) {
  if (ctx->size < 2 || !ctx->peer.active || ctx->peer.broken) return false;
  for (size_t n = 0; n < nc; n++) {
    const c3d_field *f = contours[n].field;
    if (!f->shared || NULL != contours[n].positions) return false;
    if (0 != memcmp(f->gl, contours[0].field->gl, sizeof(f->gl)) || 0 != memcmp(&f->decomp, &contours[0].field->decomp, sizeof(f->decomp))) return false;
    /* the tiled sweep (records) is the only producer of the partitioned lists: needs 16-byte rows, the same on every rank */
    if (0 != (f->ext[0] * (size_t)f->elem_bytes) % 16) return false;
  }
  return true;
}

static void release_output_map(c3d_context *ctx) {
  if (NULL != ctx->out_map) {
    if (ctx->out_registered) cudaHostUnregister(ctx->out_map);
    munmap(ctx->out_map, ctx->out_len);
  }
  free(ctx->out_path);
  ctx->out_path = NULL;
  ctx->out_map = NULL;
  ctx->out_len = 0;
  ctx->out_registered = false;
  cudaGetLastError();
}

/* maps `fname` at its final length (header + payload) for direct device-to-file copies; false: use write_ppm() */
static bool acquire_output_map(c3d_context *ctx, const char *fname, const char *header, const size_t hlen, const size_t bytes) {
  static int enabled = -1;
  if (enabled < 0) {
    const char *env = getenv("CONTOUR3D_MAP_OUTPUT");
    enabled = (NULL != env && '0' == env[0]) ? 0 : 1;
  }
  if (!enabled) return false;
  const size_t total = hlen + bytes;
  struct stat st;
  if (NULL != ctx->out_map && NULL != ctx->out_path && 0 == strcmp(ctx->out_path, fname) && ctx->out_len == total && 0 == stat(fname, &st)
      && (unsigned long long)st.st_ino == ctx->out_ino && (unsigned long long)st.st_dev == ctx->out_dev && (size_t)st.st_size == total) {
    return true;                                    /* the same file as last time: header and length are already in place */
  }
  release_output_map(ctx);
  const int fd = open(fname, O_RDWR | O_CREAT, 0644);
  if (fd < 0) return false;                         /* write_ppm() reports the error */
  bool ok = 0 == ftruncate(fd, (off_t)total) && 0 == fstat(fd, &st) && S_ISREG(st.st_mode);
  void *map = MAP_FAILED;
  if (ok) map = mmap(NULL, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (!ok || MAP_FAILED == map) return false;
  if (cudaSuccess != cudaHostRegister(map, total, cudaHostRegisterDefault)) {
    cudaGetLastError();
    munmap(map, total);
    enabled = 0;                                    /* this file system's pages cannot be pinned: do not try again */
    return false;
  }
  memcpy(map, header, hlen);
  ctx->out_path = strdup(fname);
  ctx->out_map = (unsigned char *)map;
  ctx->out_len = total;
  ctx->out_ino = (unsigned long long)st.st_ino;
  ctx->out_dev = (unsigned long long)st.st_dev;
  ctx->out_registered = true;
  return NULL != ctx->out_path;
}

static int write_ppm(const char *fname, const size_t width, const size_t height, const unsigned char *rgb, const size_t bytes) {
  errno = 0;
  const int fd = open(fname, O_WRONLY | O_CREAT, 0644);
  if (fd < 0) return c3d_error("%s: %s", fname, strerror(errno));
  char header[64];
  const int hlen = snprintf(header, sizeof(header), "P6\n%zu %zu\n255\n", width, height);
  struct iovec io[2];
  io[0].iov_base = header; io[0].iov_len = (size_t)hlen;
  io[1].iov_base = const_cast<unsigned char *>(rgb); io[1].iov_len = bytes;
  size_t done = 0;
  const size_t total = (size_t)hlen + bytes;
  while (done < total) {
    /* writev may stop short: continue where it stopped */
    struct iovec rest[2];
    int nrest = 0;
    size_t skip = done;
    for (int q = 0; q < 2; q++) {
      if (skip >= io[q].iov_len) { skip -= io[q].iov_len; continue; }
      rest[nrest].iov_base = (char *)io[q].iov_base + skip;
      rest[nrest].iov_len = io[q].iov_len - skip;
      skip = 0;
      nrest++;
    }
    const ssize_t w = writev(fd, rest, nrest);
    if (w <= 0) {
      if (w < 0 && EINTR == errno) continue;
      const int err = errno;
      close(fd);
      return c3d_error("%s: write failed after %zu of %zu bytes (%s)", fname, done, total, strerror(err));
    }
    done += (size_t)w;
  }
  if (0 != ftruncate(fd, (off_t)total)) { const int err = errno; close(fd); return c3d_error("%s: %s", fname, strerror(err)); }
  if (0 != close(fd)) return c3d_error("%s: %s", fname, strerror(errno));
  return 0;
}

static int frame_once(c3d_context *ctx, const c3d_scene_t *scene, const size_t nc, const c3d_contour_t *contours,
                      const size_t nl, const c3d_line_t *lines, uint8_t *host_rgb, const char *fname, bool *overflow) {
  *overflow = false;
  cudaStream_t s = ctx->stream;
  const size_t npix = scene->width * scene->height;
  Cam cam;
  make_cam(scene, &cam);
  uint32_t launches = 0;
  /* Several ranks: either the screen is partitioned (every rank draws its tiles from everybody's cubes, read through peer
     memory; decided from properties all ranks share, so they agree without asking) or, with coordinate converters or without
     peer access, every rank draws its own cubes and whole images are reduced through NCCL. */
  const bool part = frame_can_partition(ctx, nc, contours);
  if (ctx->size > 1 && !part && !ctx->has_comm) {
    return c3d_error("this frame needs the NCCL path (coordinate converter, unaligned rows or objects of different shapes), but the ranks share a GPU");
  }
  c3d_peer *peer = &ctx->peer;
  ctx->ndetail = 0;
  /* detail mark n closes step n of c3d_timings_t.detail (partitioned frames only) */
  auto mark = [&]() { if (part && ctx->ndetail < C3D_DETAIL_EVENTS) cudaEventRecord(ctx->ev_detail[ctx->ndetail++], s); };
  CUDA_OK(cudaEventRecord(ctx->ev[EV_START], s));
  mark();
  /* halos */
  if (part && NULL != ctx->xa.base) {
    /* what my previous frame spent drawing / shading, where the others can read it (gathered at the end of this frame) */
    memcpy(ctx->counts_host + CNT_LOADS, ctx->my_loads, sizeof(ctx->my_loads));
    CUDA_OK(cudaMemcpyAsync(ctx->xa.base + ctx->xa.off_flags + 64, ctx->counts_host + CNT_LOADS, 8, cudaMemcpyHostToDevice, s));
  }
  if (part) { LAUNCH_OK(c3d_peer_barrier(peer, s)); launches += 1; }    /* every rank's arrays are complete; last frame's buffers are free */
  mark();                                                               /* detail[0]: barrier at the start */
  for (size_t n = 0; n < nc; n++) {
    c3d_field *f = (c3d_field *)contours[n].field;
    if (!f->halo_valid) {
      if (part) {
        if (0 != field_exchange_peer(ctx, f, true, &launches)) return 1;
      } else {
        if (0 != c3d_field_exchange(ctx, f)) return 1;
        launches += 4;
      }
    }
  }
  CUDA_OK(cudaEventRecord(ctx->ev[EV_HALO], s));
  mark();                                                               /* detail[1]: halo pull */
  /* descriptors */
  FrameParams fp;
  memset(&fp, 0, sizeof(fp));
  fp.cam = cam;
  make_cull(&cam, &fp.cull);
  fp.tab = ctx->tab;
  fp.px = ctx->pixel_xy;
  fp.py = ctx->pixel_xy + scene->width;
  fp.pxf = (const float *)(ctx->pixel_xy + scene->width + scene->height);
  fp.pyf = fp.pxf + scene->width;
  fp.ncontours = (unsigned int)nc;
  FieldView *views = fp.view;
  std::vector<LineDev> hl(nl);
  const LineDev *lines_dev = NULL;
  if (0 != stage_frame(ctx, nc, contours, nl, lines, hl, views, &lines_dev, part, &fp.pviews)) return 1;
  if (part) {
    Partition &q = fp.part;
    q.nranks = ctx->size;
    q.me = ctx->rank;
    q.tile_shift = ctx->tile_shift;
    q.tiles_x = (int)((scene->width + ((size_t)1 << q.tile_shift) - 1) >> q.tile_shift);
    q.tiles_y = (int)((scene->height + ((size_t)1 << q.tile_shift) - 1) >> q.tile_shift);
    q.recip = 0xffffffffu / (unsigned int)q.nranks + 1u;
    q.py = q.pz = 1;
    q.tile_map = NULL;
    if (ctx->tile_balance) {
      const size_t nt = (size_t)q.tiles_x * (size_t)q.tiles_y;
      if (ctx->tile_map_w != scene->width || ctx->tile_map_h != scene->height || ctx->tile_map_host->size() != nt) {
        /* new screen: even shares again (every rank gets here in the same frame) */
        for (int r = 0; r < ctx->size; r++) ctx->tile_share[r] = 1. / (double)ctx->size;
        tile_map_build(ctx->tile_share, ctx->size, q.tiles_x, q.tiles_y, *ctx->tile_map_host);
        cudaFree(ctx->tile_map_dev);
        ctx->tile_map_dev = NULL;
        CUDA_OK(cudaMalloc(&ctx->tile_map_dev, nt));
        CUDA_OK(cudaMemcpyAsync(ctx->tile_map_dev, ctx->tile_map_host->data(), nt, cudaMemcpyHostToDevice, s));
        ctx->tile_map_w = scene->width; ctx->tile_map_h = scene->height;
      }
      q.tile_map = ctx->tile_map_dev;
    }
    if (nc > 0) {
      /* which rank holds which block: told by the ranks themselves when the field was created (any Cartesian numbering) */
      const c3d_field *f = contours[0].field;
      q.py = f->decomp.nprocs[1];
      q.pz = f->decomp.nprocs[2];
      if (q.py * q.pz != ctx->size) return c3d_error("process grid %d x %d does not match %d ranks", q.py, q.pz, ctx->size);
      for (int r = 0; r < ctx->size; r++) {
        const FieldGeom &g = f->geom[r];
        if (g.coords[1] < 0 || g.coords[1] >= q.py || g.coords[2] < 0 || g.coords[2] >= q.pz) return c3d_error("rank %d reports a position outside the process grid", r);
        q.yoff[g.coords[1]] = (unsigned short)g.offsets[1];
        q.zoff[g.coords[2]] = (unsigned short)g.offsets[2];
        q.home[g.coords[1] * q.pz + g.coords[2]] = (unsigned char)r;
      }
      q.yoff[q.py] = (unsigned short)f->gl[1];
      q.zoff[q.pz] = (unsigned short)f->gl[2];
    }
  }
  /* work buffers of the frame: patch records from the tiled sweeps, then one list entry per cut cube */
  if (nc > 0 && 0 == ctx->list_cap) {
    size_t want = 0;
    for (size_t n = 0; n < nc; n++) {
      const c3d_field *f = contours[n].field;
      const size_t cubes = f->ext[0] * f->ext[1] * f->ext[2];
      want += std::max<size_t>(1u << 20, cubes / 16);
    }
    {
      /* test hook: a deliberately small first allocation exercises the grow-and-redo path below */
      const char *env = getenv("CONTOUR3D_LIST_CAP");
      if (NULL != env && '\0' != *env && strtoull(env, NULL, 10) >= 64) want = (size_t)strtoull(env, NULL, 10);
    }
    CUDA_OK(cudaMalloc(&ctx->list, want * 8));
    ctx->list_cap = want;
  }
  if (nc > 0 && 0 == ctx->rec_cap) {
    const size_t want = NULL != getenv("CONTOUR3D_LIST_CAP") ? std::max<size_t>(64, ctx->list_cap / 4) : std::max<size_t>(1u << 18, ctx->list_cap / 4);
    CUDA_OK(cudaMalloc(&ctx->records, want * 32));
    ctx->rec_cap = want;
  }
  if (part && 0 != ensure_exchange(ctx, scene->width, scene->height, std::max<size_t>(ctx->list_cap, 64))) return 1;
  if (part && ctx->part_masks_cap < ctx->rec_cap) {
    cudaFree(ctx->part_masks);
    ctx->part_masks = NULL;
    ctx->part_masks_cap = 0;
    CUDA_OK(cudaMalloc(&ctx->part_masks, std::max<size_t>(ctx->rec_cap, 1) * 32));
    ctx->part_masks_cap = ctx->rec_cap;
  }
  Exchange &xa = ctx->xa;
  unsigned int *rgba_out = ctx->rgba;                /* where finished colours go: the writer's image */
  if (part) rgba_out = (unsigned int *)((char *)xa.peer[0] + xa.off_rgba);
  CUDA_OK(cudaMemsetAsync(ctx->counts_dev, 0, CNT_TOTAL * sizeof(unsigned long long), s));
  {
    /* The vertex table is cleared every frame: size it by what the previous frame shaded (a rank of a partitioned frame shades a
       fraction of the screen; three vertices per pixel at most, load below 1/2) instead of by the whole screen.  Too small costs
       time, never correctness: a vertex that finds no slot is smoothed per pixel (k_shade_blend). */
    size_t slots = ctx->shade.table_slots;
    if (ctx->shade_items_prev > 0) {
      const size_t want = 8 * (size_t)ctx->shade_items_prev + (1u << 16);
      size_t fit = 1u << 16;
      while (fit < want && fit < ctx->shade.table_slots) fit <<= 1;
      slots = std::min<size_t>(fit, ctx->shade.table_slots);
    }
    ctx->shade.table_mask = (unsigned int)(slots - 1);
  }
  if (part) {
    CUDA_OK(cudaMemsetAsync(xa.base + xa.off_nseg, 0, 256, s));
    LAUNCH_OK(c3d_k_clear_part(ctx->canvas, &fp.part, cam.w, cam.h, ctx->sm_count, s));
  } else {
    LAUNCH_OK(c3d_k_clear(ctx->canvas, npix, s));
  }
  launches += 1;
  CUDA_OK(cudaEventRecord(ctx->ev[EV_CLEAR], s));
  /* 1. classification sweeps: contour objects that share an array share ONE pass over it (up to
        C3D_SWEEP_MAX_THR iso-values per launch); arrays whose rows are not 16-byte aligned take the
        direct kernel, one launch per object */
  uint32_t nsweeps = 0;
  uint32_t sweep_nthr[C3D_MAX_OBJECTS];
  uint64_t sweep_bytes[C3D_MAX_OBJECTS];
  int snap_of[C3D_MAX_OBJECTS];
  bool used_records = false;
  uint32_t ndirect = 0;
  {
    bool done[C3D_MAX_OBJECTS] = {false};
    for (size_t n = 0; n < nc; n++) {
      if (done[n]) continue;
      const c3d_field *f = contours[n].field;
      SweepRange r;
      rendered_range(&views[n], 1, &r);
      const uint64_t bytes = (uint64_t)f->ext[0] * f->ext[1] * f->ext[2] * (uint64_t)f->elem_bytes;
      if (c3d_k_sweep_tiles_ok(&views[n])) {
        double thr[C3D_SWEEP_MAX_THR];
        float thr_ge[C3D_SWEEP_MAX_THR];
        unsigned int obj[C3D_SWEEP_MAX_THR];
        int cnt = 0;
        for (size_t m = n; m < nc && cnt < C3D_SWEEP_MAX_THR; m++) {
          if (done[m] || contours[m].field != contours[n].field) continue;
          thr[cnt] = views[m].thr; thr_ge[cnt] = views[m].thr_ge; obj[cnt] = (unsigned int)m;
          snap_of[m] = -1;
          done[m] = true;
          cnt++;
        }
        CUDA_OK(cudaEventRecord(ctx->ev_sweep[nsweeps], s));
        /* front-to-back for this camera: start at the end of each axis the camera is nearer to */
        int flip[3];
        for (int d = 0; d < 3; d++) {
          const double lo = contours[n].grids[d][f->ext_off[d]], hi = contours[n].grids[d][f->ext_off[d] + f->ext[d] - 1];
          flip[d] = fabs(cam.cam[d] - hi) < fabs(cam.cam[d] - lo) ? 1 : 0;
        }
        LAUNCH_OK(c3d_k_sweep_tiles(&views[n], &r, cnt, thr, thr_ge, obj, flip, ctx->tile_cursor, ctx->records, ctx->counts_dev + CNT_REC,
                                    ctx->rec_cap, ctx->sm_count, s));
        sweep_nthr[nsweeps] = (uint32_t)cnt;
        used_records = true;
      } else {
        if (part) return c3d_error("internal: a partitioned frame met an array the tiled sweep cannot take");
        done[n] = true;
        CUDA_OK(cudaEventRecord(ctx->ev_sweep[nsweeps], s));
        LAUNCH_OK(c3d_k_sweep(&views[n], &r, (unsigned int)n, ctx->list, ctx->counts_dev + CNT_LIST, ctx->list_cap, ctx->sm_count, s));
        CUDA_OK(cudaMemcpyAsync(ctx->counts_dev + CNT_SNAP + 1 + ndirect, ctx->counts_dev + CNT_LIST, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
        snap_of[n] = (int)ndirect++;
        sweep_nthr[nsweeps] = 1;
      }
      sweep_bytes[nsweeps] = bytes;
      nsweeps++;
      launches += 1;
    }
  }
  CUDA_OK(cudaEventRecord(ctx->ev_sweep[nsweeps], s));
  CUDA_OK(cudaEventRecord(ctx->ev[EV_SWEEP], s));
  mark();                                                               /* detail[2]: staging, clear, sweeps */
  if (used_records || part) {
    /* range of the view depth n.(p - cam) over the corners of every array: the slabs of the front-to-back order.
       Partitioned frames take the corners of the GLOBAL domain, so that a bucket means the same depth on every rank. */
    double near_depth = -DBL_MAX, far_depth = DBL_MAX;
    for (size_t n = 0; n < nc; n++) {
      const c3d_field *f = contours[n].field;
      if (NULL != contours[n].positions) {
        /* converted coordinates: bound each component by the extremes of its table (a box around the mapped block) */
        double lo = 0., hi = 0.;
        for (int c = 0; c < 3; c++) {
          const size_t len = table_length(f, contours[n].positions, c);
          const double *tab = contours[n].positions->table[c];
          double tmin = DBL_MAX, tmax = -DBL_MAX;
          for (size_t q = 0; q < len; q++) {
            if (tab[q] < tmin) tmin = tab[q];
            if (tab[q] > tmax) tmax = tab[q];
          }
          const double a = cam.normal[c] * (tmin - cam.cam[c]), b = cam.normal[c] * (tmax - cam.cam[c]);
          lo += std::min(a, b);
          hi += std::max(a, b);
        }
        if (hi >= lo) {         /* NaN-free */
          near_depth = std::max(near_depth, hi);
          far_depth = std::min(far_depth, lo);
        }
        continue;
      }
      for (int c = 0; c < 8; c++) {
        double depth = 0.;
        for (int d = 0; d < 3; d++) {
          const size_t first = part ? 0 : f->ext_off[d], count = part ? f->gl[d] : f->ext[d];
          const size_t at = first + (((c >> d) & 1) ? count - 1 : 0);
          depth += cam.normal[d] * (contours[n].grids[d][at] - cam.cam[d]);
        }
        near_depth = std::max(near_depth, depth);
        far_depth = std::min(far_depth, depth);
      }
    }
    if (part) {
      /* 2a. every cut cube goes to the outbox of each rank whose screen tiles it can touch, nearest first */
      if (!(near_depth > far_depth)) { near_depth = 1.; far_depth = 0.; }      /* degenerate / NaN coordinates: one bucket */
      LAUNCH_OK(c3d_k_expand_part(&fp, (float)near_depth, (float)far_depth, ctx->records, ctx->counts_dev + CNT_REC, ctx->rec_cap,
                                  (unsigned long long *)(xa.base + xa.off_out), ctx->counts_dev + CNT_OUT, xa.out_cap,
                                  (unsigned int *)(xa.base + xa.off_hist), ctx->part_cursors, (unsigned long long *)(xa.base + xa.off_seg_start), ctx->part_masks,
                                  ctx->counts_dev + CNT_OBJ, (unsigned int)nc, ctx->sm_count, s));
      launches += 3;
      mark();                                                           /* detail[3]: cubes -> outboxes (count, scan, scatter) */
    } else {
      const int sorted = fp.cull.enabled && near_depth > far_depth;
      LAUNCH_OK(c3d_k_expand(&fp, sorted, (float)near_depth, (float)far_depth, ctx->records, ctx->counts_dev + CNT_REC, ctx->rec_cap, ctx->list,
                             ctx->counts_dev + CNT_LIST, ctx->list_cap, ctx->bucket_hist, ctx->bucket_cursors,
                             ctx->counts_dev + CNT_OBJ, (unsigned int)nc, ctx->sm_count, s));
      launches += sorted ? 3 : 1;
    }
  }
  if (part) {
    /* 2b. all outboxes are complete -> every rank pulls what is addressed to it into ONE nearest-first list */
    LAUNCH_OK(c3d_peer_barrier(peer, s));
    mark();                                                             /* detail[4]: barrier (all outboxes complete) */
    MergeSources ms;
    memset(&ms, 0, sizeof(ms));
    for (int r = 0; r < ctx->size; r++) {
      const char *pb = (const char *)xa.peer[r];
      ms.hist[r] = (const unsigned int *)(pb + xa.off_hist);
      ms.seg_start[r] = (const unsigned long long *)(pb + xa.off_seg_start);
      ms.outbox[r] = (const unsigned long long *)(pb + xa.off_out);
      ms.out_cap[r] = (const unsigned long long *)(pb + xa.off_flags + 72);
    }
    LAUNCH_OK(c3d_k_merge_lists(&ms, ctx->size, ctx->rank, ctx->merge_base, ctx->list, ctx->counts_dev + CNT_LIST, ctx->list_cap, s));
    launches += 3;
    mark();                                                             /* detail[5]: pull + merge of the lists */
  }
  CUDA_OK(cudaEventRecord(ctx->ev[EV_EXPAND], s));
  /* 2. tessellate + rasterise every active cube, contours in one launch (src/contour3d/main.c:105-117) */
  /*    The nearest-first list is drawn in parts; between parts the coarse z-buffer is rebuilt from the canvas and,
        with several ranks, maximised across them (1 MB at 4K): from then on a rank also drops the cubes that ANOTHER
        rank has already hidden -- without this a slab behind its neighbours rasterises everything it owns.  The table is
        only ever a lower bound of the final depths, so the image cannot change (cube_hidden, c3d_kernels.cu). */
  if (nc > 0) {
    const int ncuts = (fp.cull.enabled && !part) ? ctx->nraster_cuts : 0;
    const size_t ntiles = ((scene->width + C3D_HIZ_TILE - 1) / C3D_HIZ_TILE) * ((scene->height + C3D_HIZ_TILE - 1) / C3D_HIZ_TILE);
    for (int cut = 0; cut <= ncuts; cut++) {
      const unsigned int begin = 0 == cut ? 0u : (unsigned int)ctx->raster_cuts[cut - 1];
      const unsigned int end = cut == ncuts ? C3D_RASTER_PARTS : (unsigned int)ctx->raster_cuts[cut];
      LAUNCH_OK(c3d_k_raster(&fp, ctx->list, ctx->counts_dev + CNT_LIST, ctx->list_cap, ctx->canvas, ctx->counts_dev + CNT_CULLED,
                             cut > 0 ? ctx->hiz : NULL, begin, end, ctx->sm_count, s));
      launches += 1;
      if (cut < ncuts) {
        LAUNCH_OK(c3d_k_hiz_build(ctx->canvas, cam.w, cam.h, ctx->hiz, ctx->sm_count, s));
        launches += 1;
        if (ctx->size > 1) NCCL_OK(g_nccl.AllReduce(ctx->hiz, ctx->hiz, ntiles, ncclUint64, ncclMax, ctx->comm, s));
      }
    }
  }
  CUDA_OK(cudaEventRecord(ctx->ev[EV_RASTER], s));
  /* 3. lines: drawn after the contours in the reference (main.c:119-129); their keys are larger, so they win ties */
  LAUNCH_OK(c3d_k_lines(&cam, lines_dev, hl.data(), nl, nc, ctx->canvas, s));
  if (nl > 0) launches += 1;
  CUDA_OK(cudaEventRecord(ctx->ev[EV_LINES], s));
  if (part) {
    /* 4'. my tiles are final: background and lines go to the writer's image, pixels won by a contour triangle to the rank
           that holds the winning cube's array (replaces MPI_Reduce, output.c:10-30,82: nothing image-sized is reduced) */
    LAUNCH_OK(c3d_k_shade_begin(&ctx->shade, s));
    LAUNCH_OK(c3d_k_shade_select_part(&fp, lines_dev, ctx->canvas, rgba_out, (WorkItem *)(xa.base + xa.off_seg), (unsigned int)xa.seg_cap,
                                      (unsigned int *)(xa.base + xa.off_nseg), ctx->counts_dev + CNT_SEG_OVER, ctx->sm_count, s));
    launches += 1;
    CUDA_OK(cudaEventRecord(ctx->ev[EV_RESOLVE], s));
    mark();                                                             /* detail[6]: raster, lines, select */
    LAUNCH_OK(c3d_peer_barrier(peer, s));
    mark();                                                             /* detail[7]: barrier (all work segments complete) */
    WorkExchange ws;
    memset(&ws, 0, sizeof(ws));
    for (int r = 0; r < ctx->size; r++) {
      const char *pb = (const char *)xa.peer[r];
      ws.seg[r] = (const WorkItem *)(pb + xa.off_seg);
      ws.nseg[r] = (const unsigned int *)(pb + xa.off_nseg);
    }
    ws.seg_cap = (unsigned int)xa.seg_cap;
    LAUNCH_OK(c3d_k_work_exchange(&ws, ctx->size, ctx->rank, ctx->shade_share, ctx->pull_plan, ctx->shade.work, ctx->shade.counters,
                                  (unsigned int)std::min<size_t>(npix, 0xffffffffu), ctx->counts_dev + CNT_SEG_OVER, s));
    launches += 2;
    CUDA_OK(cudaEventRecord(ctx->ev[EV_REDUCE0], s));
    mark();                                                             /* detail[8]: pull of the shading work */
    /* 5'. shade what this rank's cubes have won, anywhere on the screen, straight into the writer's image */
    LAUNCH_OK(c3d_k_shade_finish(&fp, rgba_out, &ctx->shade, ctx->counts_dev + CNT_SHADE_SPILL, 1, ctx->sm_count, s));
    launches += 7;
    CUDA_OK(cudaEventRecord(ctx->ev[EV_SHADE], s));
    mark();                                                             /* detail[9]: shading */
    /* 6'. the image is complete when everybody is done; a buffer that overflowed anywhere makes everybody redo the frame */
    LAUNCH_OK(c3d_k_overflow_flag(ctx->counts_dev + CNT_REC, ctx->rec_cap, ctx->counts_dev + CNT_LIST, ctx->list_cap, ctx->counts_dev + CNT_OUT, xa.out_cap,
                                  ctx->counts_dev + CNT_SEG_OVER, (unsigned long long *)(xa.base + xa.off_flags), s));
    LAUNCH_OK(c3d_peer_barrier(peer, s));
    FlagSources fs;
    memset(&fs, 0, sizeof(fs));
    for (int r = 0; r < ctx->size; r++) fs.flag[r] = (const unsigned long long *)((const char *)xa.peer[r] + xa.off_flags);
    LAUNCH_OK(c3d_k_flags_max(&fs, ctx->size, ctx->counts_dev + CNT_OVERFLOW, ctx->counts_dev + CNT_LOADS, s));
    launches += 3;
    CUDA_OK(cudaEventRecord(ctx->ev[EV_REDUCE1], s));
    mark();                                                             /* detail[10]: barrier at the end (+ overflow agreement) */
    if (0 == ctx->rank) {
      LAUNCH_OK(c3d_k_pack((const unsigned int *)(xa.base + xa.off_rgba), scene->bg, cam.w, cam.h, ctx->ppm_dev, s));
      launches += 1;
    }
  } else {
    CUDA_OK(cudaEventRecord(ctx->ev[EV_RESOLVE], s));
    /* 4. nearest depth across ranks (replaces the depth half of MPI_Reduce, output.c:10-30,82) */
    const unsigned long long *depth_global = NULL;
    if (ctx->size > 1) {
      LAUNCH_OK(c3d_k_depth_extract(ctx->canvas, ctx->depth_local, npix, s));
      launches += 1;
      NCCL_OK(g_nccl.AllReduce(ctx->depth_local, ctx->depth_global, npix, ncclUint64, ncclMax, ctx->comm, s));
      depth_global = ctx->depth_global;
    }
    CUDA_OK(cudaEventRecord(ctx->ev[EV_REDUCE0], s));
    /* 5. shade only what survived */
    LAUNCH_OK(c3d_k_shade(&fp, lines_dev, ctx->canvas, depth_global, ctx->rgba, npix, &ctx->shade,
                          ctx->counts_dev + CNT_SHADE_SPILL, ctx->sm_count, s));
    launches += 5;
    CUDA_OK(cudaEventRecord(ctx->ev[EV_SHADE], s));
    /* 6. colours to the writer (the colour half of the reference's reduce) */
    if (ctx->size > 1) {
      NCCL_OK(g_nccl.Reduce(ctx->rgba, ctx->rgba, npix, ncclUint32, ncclMax, 0, ctx->comm, s));
    }
    CUDA_OK(cudaEventRecord(ctx->ev[EV_REDUCE1], s));
    if (0 == ctx->rank) {
      LAUNCH_OK(c3d_k_pack(ctx->rgba, scene->bg, cam.w, cam.h, ctx->ppm_dev, s));
      launches += 1;
    }
  }
  CUDA_OK(cudaEventRecord(ctx->ev[EV_PACK], s));
  const bool want_host = (0 == ctx->rank) && (NULL != host_rgb || NULL != fname);
  bool mapped_out = false;
  if (want_host && NULL != fname && NULL == host_rgb) {
    char header[64];
    const int hlen = snprintf(header, sizeof(header), "P6\n%zu %zu\n255\n", scene->width, scene->height);
    mapped_out = acquire_output_map(ctx, fname, header, (size_t)hlen, 3 * npix);
    if (mapped_out) CUDA_OK(cudaMemcpyAsync(ctx->out_map + hlen, ctx->ppm_dev, 3 * npix, cudaMemcpyDeviceToHost, s));
  }
  if (want_host && !mapped_out) CUDA_OK(cudaMemcpyAsync(ctx->ppm_host, ctx->ppm_dev, 3 * npix, cudaMemcpyDeviceToHost, s));
  if (ctx->size > 1 && !part) {
    /* the frame is redone when a work buffer overflowed: all ranks must agree, the frame is full of collectives */
    LAUNCH_OK(c3d_k_overflow_flag(ctx->counts_dev + CNT_REC, ctx->rec_cap, ctx->counts_dev + CNT_LIST, ctx->list_cap, NULL, 0, NULL, ctx->counts_dev + CNT_OVERFLOW, s));
    NCCL_OK(g_nccl.AllReduce(ctx->counts_dev + CNT_OVERFLOW, ctx->counts_dev + CNT_OVERFLOW, 1, ncclUint64, ncclMax, ctx->comm, s));
    launches += 1;
  }
  CUDA_OK(cudaMemcpyAsync(ctx->counts_dev + CNT_SHADE_ITEMS, ctx->shade.counters, sizeof(unsigned int), cudaMemcpyDeviceToDevice, s));
  CUDA_OK(cudaMemcpyAsync(ctx->counts_host, ctx->counts_dev, CNT_TOTAL * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  if (part) CUDA_OK(cudaMemcpyAsync(ctx->counts_host + CNT_PEER, peer->status, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaEventRecord(ctx->ev[EV_DOWNLOAD], s));
  CUDA_OK(cudaStreamSynchronize(s));
  if (part && 0ull != ctx->counts_host[CNT_PEER]) {
    peer->broken = true;
    return c3d_error("rank %d waited in vain for another rank at a device barrier: the ranks are out of step", ctx->rank);
  }
  /* capacity policy: a buffer that overflowed is grown and the frame is redone */
  {
    const size_t need_rec = (size_t)ctx->counts_host[CNT_REC], need_list = (size_t)ctx->counts_host[CNT_LIST];
    if (need_rec > ctx->rec_cap) {
      *overflow = true;
      cudaFree(ctx->records);
      ctx->records = NULL;
      ctx->rec_cap = 0;
      const size_t want = need_rec + need_rec / 4 + 1024;
      const cudaError_t err = cudaMalloc(&ctx->records, want * 32);
      if (cudaSuccess != err) return c3d_error("failed to grow the sweep record buffer to %zu entries: %s", want, cudaGetErrorString(err));
      ctx->rec_cap = want;
    }
    if (need_list > ctx->list_cap) {
      *overflow = true;
      cudaFree(ctx->list);
      ctx->list = NULL;
      ctx->list_cap = 0;
      const size_t want = need_list + need_list / 4 + 1024;
      const cudaError_t err = cudaMalloc(&ctx->list, want * 8);
      if (cudaSuccess != err) return c3d_error("failed to grow the active-cube list to %zu entries: %s", want, cudaGetErrorString(err));
      ctx->list_cap = want;
    }
    if (part && (size_t)ctx->counts_host[CNT_OUT] > xa.out_cap) {
      *overflow = true;
      xa.out_cap = (size_t)ctx->counts_host[CNT_OUT] + (size_t)ctx->counts_host[CNT_OUT] / 4 + 1024;    /* taken up when the arena is re-created */
    }
    if (part && 0ull != ctx->counts_host[CNT_SEG_OVER]) return c3d_error("internal: a shading work segment overflowed");
    if (0ull != ctx->counts_host[CNT_OVERFLOW]) *overflow = true;      /* another rank's buffer: redo with it */
    if (*overflow) {
      /* everybody redoes the frame; the shared arena is re-created (and re-mapped by all) because one rank's outbox may have grown */
      if (part) xa.stale = true;
      return 0;
    }
  }
  /* timings */
  c3d_timings_t *t = &ctx->last;
  memset(t, 0, sizeof(*t));
  auto ms = [&](int a, int b) { float v = 0.f; cudaEventElapsedTime(&v, ctx->ev[a], ctx->ev[b]); return v; };
  for (int u = 0; u < ctx->nuploads; u++) {
    float v = 0.f;
    if (cudaSuccess == cudaEventElapsedTime(&v, ctx->ev_upload[2 * u], ctx->ev_upload[2 * u + 1])) t->upload += v;
  }
  ctx->nuploads = 0;
  t->halo = ms(EV_START, EV_HALO);
  t->clear = ms(EV_HALO, EV_CLEAR);
  t->sweep = ms(EV_CLEAR, EV_SWEEP);
  t->expand = ms(EV_SWEEP, EV_EXPAND);
  t->raster = ms(EV_EXPAND, EV_RASTER);
  t->lines = ms(EV_RASTER, EV_LINES);
  t->resolve = ms(EV_LINES, EV_RESOLVE);
  t->reduce = ms(EV_RESOLVE, EV_REDUCE0) + ms(EV_SHADE, EV_REDUCE1);
  t->shade = ms(EV_REDUCE0, EV_SHADE);
  t->pack = ms(EV_REDUCE1, EV_PACK);
  t->download = ms(EV_PACK, EV_DOWNLOAD);
  t->total = ms(EV_START, EV_DOWNLOAD);
  t->ncontours = (uint32_t)nc;
  t->launches = launches;
  t->nsweeps = nsweeps;
  t->exchange_mode = (uint32_t)(ctx->size < 2 ? 0 : (part ? 2 : 1));
  ctx->shade_items_prev = ctx->counts_host[CNT_SHADE_ITEMS] & 0xffffffffull;
  if (part) {
    ctx->my_loads[0] = t->raster + t->lines + t->resolve;
    ctx->my_loads[1] = t->shade;
    /* every rank holds the same gathered numbers, so every rank changes the table in the same frame, or none does */
    if (ctx->tile_balance && nc > 0 && tile_shares_update(ctx, ctx->counts_host + CNT_LOADS)) {
      tile_map_build(ctx->tile_share, ctx->size, fp.part.tiles_x, fp.part.tiles_y, *ctx->tile_map_host);
      CUDA_OK(cudaMemcpyAsync(ctx->tile_map_dev, ctx->tile_map_host->data(), ctx->tile_map_host->size(), cudaMemcpyHostToDevice, s));
    }
    for (int r = 0; r < ctx->size && r < 8; r++) t->tile_share[r] = (float)ctx->tile_share[r];
  }
  for (int d = 0; d + 1 < ctx->ndetail && d < 16; d++) cudaEventElapsedTime(&t->detail[d], ctx->ev_detail[d], ctx->ev_detail[d + 1]);
  t->list_entries = ctx->counts_host[CNT_LIST];
  t->culled_cubes = ctx->counts_host[CNT_CULLED];
  t->culled_triangles = ctx->counts_host[CNT_CULLED_TRIS];
  t->shade_spills = ctx->counts_host[CNT_SHADE_SPILL];
  for (uint32_t m = 0; m < nsweeps; m++) {
    cudaEventElapsedTime(&t->sweep_each[m], ctx->ev_sweep[m], ctx->ev_sweep[m + 1]);
    t->sweep_bytes[m] = sweep_bytes[m];
    t->sweep_thresholds[m] = sweep_nthr[m];
  }
  for (size_t n = 0; n < nc; n++) {
    t->active_cubes[n] = ctx->counts_host[CNT_OBJ + n];
    if (snap_of[n] >= 0) t->active_cubes[n] += ctx->counts_host[CNT_SNAP + 1 + snap_of[n]] - ctx->counts_host[CNT_SNAP + snap_of[n]];
  }
  if (want_host && !mapped_out) {
    if (NULL != host_rgb) memcpy(host_rgb, ctx->ppm_host, 3 * npix);
    if (NULL != fname) {
      /* binary P6, output.c:125-146.  The file is complete when this returns, like the reference's; it is written in place
         (no O_TRUNC: a solver writes the same name every output step, and on a RAM-backed file system giving all the pages
         back and faulting them in again costs more than the copy) and cut to its length afterwards. */
      if (0 != write_ppm(fname, scene->width, scene->height, ctx->ppm_host, 3 * npix)) return 1;
    }
  }
  return 0;
}

extern "C" int c3d_frame_render(c3d_context_t *ctx, const c3d_scene_t *scene, size_t nc, const c3d_contour_t *contours,
                                size_t nl, const c3d_line_t *lines, uint8_t *host_rgb, const char *fname) {
  if (NULL == ctx || NULL == scene) return c3d_error("c3d_frame_render: NULL argument");
  if (nc > C3D_MAX_OBJECTS) return c3d_error("at most %d contour objects per frame, received %zu", C3D_MAX_OBJECTS, nc);
  if (nc > 0 && NULL == contours) return c3d_error("contour objects are missing");
  if (nl > 0 && NULL == lines) return c3d_error("line objects are missing");
  if (0 == scene->width || 0 == scene->height) return c3d_error("empty screen (%zu x %zu)", scene->width, scene->height);
  if (nc + nl >= 4095) return c3d_error("too many objects (%zu)", nc + nl);
  for (size_t n = 0; n < nc; n++) {
    if (NULL == contours[n].field) return c3d_error("contour object %zu has no field", n);
  }
  CUDA_OK(cudaSetDevice(ctx->device));
  if (0 != ensure_canvas(ctx, scene->width, scene->height)) return 1;
  for (int attempt = 0; attempt < 4; attempt++) {
    bool overflow = false;
    if (0 != frame_once(ctx, scene, nc, contours, nl, lines, host_rgb, fname, &overflow)) return 1;
    if (!overflow) return 0;
  }
  return c3d_error("active-cube list kept overflowing");
}

extern "C" int c3d_frame_timings(const c3d_context_t *ctx, c3d_timings_t *out) {
  if (NULL == ctx || NULL == out) return c3d_error("c3d_frame_timings: NULL argument");
  *out = ctx->last;
  return 0;
}

extern "C" int c3d_frame_canvas(c3d_context_t *ctx, double *host_depth, uint64_t *host_key) {
  if (NULL == ctx) return c3d_error("c3d_frame_canvas: NULL context");
  const size_t npix = ctx->width * ctx->height;
  if (0 == npix) return c3d_error("no frame has been rendered yet");
  CUDA_OK(cudaSetDevice(ctx->device));
  double *tmp_depth = NULL;
  unsigned long long *tmp_key = NULL;
  if (host_depth) CUDA_OK(cudaMalloc(&tmp_depth, npix * sizeof(double)));
  if (host_key) CUDA_OK(cudaMalloc(&tmp_key, npix * 8));
  LAUNCH_OK(c3d_k_canvas_split(ctx->canvas, tmp_depth, tmp_key, npix, ctx->stream));
  if (host_depth) CUDA_OK(cudaMemcpyAsync(host_depth, tmp_depth, npix * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (host_key) CUDA_OK(cudaMemcpyAsync(host_key, tmp_key, npix * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  cudaFree(tmp_depth);
  cudaFree(tmp_key);
  return 0;
}

extern "C" int c3d_frame_collectives_probe(c3d_context_t *ctx, size_t width, size_t height, int iters, float *allreduce_ms, float *reduce_ms) {
  if (NULL == ctx || NULL == allreduce_ms || NULL == reduce_ms || iters < 1 || 0 == width * height) return c3d_error("c3d_frame_collectives_probe: bad argument");
  *allreduce_ms = *reduce_ms = 0.f;
  if (ctx->size < 2) return 0;
  CUDA_OK(cudaSetDevice(ctx->device));
  if (0 != ensure_canvas(ctx, width, height)) return 1;
  const size_t npix = width * height;
  cudaStream_t s = ctx->stream;
  CUDA_OK(cudaMemsetAsync(ctx->depth_local, 0, npix * 8, s));
  CUDA_OK(cudaMemsetAsync(ctx->rgba, 0, npix * 4, s));
  /* warm-up round (also lines the ranks up), then the timed rounds */
  NCCL_OK(g_nccl.AllReduce(ctx->depth_local, ctx->depth_global, npix, ncclUint64, ncclMax, ctx->comm, s));
  NCCL_OK(g_nccl.Reduce(ctx->rgba, ctx->rgba, npix, ncclUint32, ncclMax, 0, ctx->comm, s));
  CUDA_OK(cudaEventRecord(ctx->ev[EV_START], s));
  for (int it = 0; it < iters; it++) NCCL_OK(g_nccl.AllReduce(ctx->depth_local, ctx->depth_global, npix, ncclUint64, ncclMax, ctx->comm, s));
  CUDA_OK(cudaEventRecord(ctx->ev[EV_REDUCE0], s));
  for (int it = 0; it < iters; it++) NCCL_OK(g_nccl.Reduce(ctx->rgba, ctx->rgba, npix, ncclUint32, ncclMax, 0, ctx->comm, s));
  CUDA_OK(cudaEventRecord(ctx->ev[EV_REDUCE1], s));
  CUDA_OK(cudaStreamSynchronize(s));
  float a = 0.f, b = 0.f;
  CUDA_OK(cudaEventElapsedTime(&a, ctx->ev[EV_START], ctx->ev[EV_REDUCE0]));
  CUDA_OK(cudaEventElapsedTime(&b, ctx->ev[EV_REDUCE0], ctx->ev[EV_REDUCE1]));
  *allreduce_ms = a / (float)iters;
  *reduce_ms = b / (float)iters;
  return 0;
}

/* ------------------------------------------------------------ stage level */

static int stage_view(c3d_context *ctx, const c3d_field *f, const double *const grids[3], const double thr, FieldView *v) {
  const uint8_t white[3] = {255, 255, 255};
  size_t cap = ctx->grids_cap * sizeof(double);
  const size_t elems = grid_slot_doubles(f);
  if (0 != ensure_bytes((void **)&ctx->grids_dev, &cap, (elems + 8) * sizeof(double))) return 1;
  ctx->grids_cap = cap / sizeof(double);
  if (NULL == grids) {
    /* classification does not need coordinates */
    v->data = f->data; v->elem_bytes = f->elem_bytes;
    for (int d = 0; d < 3; d++) { v->n[d] = (long long)f->ext[d]; v->g[d] = ctx->grids_dev; v->gf[d] = (const float *)ctx->grids_dev; }
    v->sy = (long long)f->ext[0]; v->sz = (long long)(f->ext[0] * f->ext[1]);
    v->map = NULL;
    for (int d = 0; d < 3; d++) v->off[d] = (long long)f->ext_off[d];
    v->thr = thr; v->thr_ge = smallest_float_above(thr);
    return 0;
  }
  return make_view(ctx, f, grids, thr, white, ctx->grids_dev, v);
}

extern "C" int c3d_case_codes(c3d_context_t *ctx, const c3d_field_t *f, double threshold, uint32_t *host_codes, uint8_t *host_counts) {
  if (NULL == ctx || NULL == f || NULL == host_codes || NULL == host_counts) return c3d_error("c3d_case_codes: NULL argument");
  CUDA_OK(cudaSetDevice(ctx->device));
  if (f->ext[0] < 2 || f->ext[1] < 2 || f->ext[2] < 2) return c3d_error("array too small");
  const size_t ncubes = (f->ext[0] - 1) * (f->ext[1] - 1) * (f->ext[2] - 1);
  FieldView v;
  if (0 != stage_view(ctx, f, NULL, threshold, &v)) return 1;
  unsigned int *codes = NULL;
  unsigned char *counts = NULL;
  CUDA_OK(cudaMalloc(&codes, ncubes * 4));
  CUDA_OK(cudaMalloc(&counts, ncubes));
  LAUNCH_OK(c3d_k_case_codes(&v, &ctx->tab, codes, counts, ctx->stream));
  CUDA_OK(cudaMemcpyAsync(host_codes, codes, ncubes * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaMemcpyAsync(host_counts, counts, ncubes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  cudaFree(codes);
  cudaFree(counts);
  return 0;
}

static int sweep_to_device(c3d_context *ctx, const FieldView *v, const int rendered_only, unsigned long long **list, unsigned long long *count) {
  SweepRange r;
  rendered_range(v, rendered_only, &r);
  size_t cap = 1u << 20, rec_cap = 1u << 18;
  const bool tiles = 0 != c3d_k_sweep_tiles_ok(v);
  for (int attempt = 0; attempt < 4; attempt++) {
    void *records = NULL;
    CUDA_OK(cudaMalloc(list, cap * 8));
    CUDA_OK(cudaMemsetAsync(ctx->counts_dev, 0, CNT_TOTAL * sizeof(unsigned long long), ctx->stream));
    if (tiles) {
      const unsigned int obj = 0u;
      CUDA_OK(cudaMalloc(&records, rec_cap * 32));
      /* the result is a set, so any visiting order will do: alternate them so that the stage-level parity
         tests cover both marching directions */
      static int turn = 0;
      const int flip[3] = {turn & 1, (turn >> 1) & 1, (turn >> 2) & 1};
      turn = (turn + 3) & 7;
      LAUNCH_OK(c3d_k_sweep_tiles(v, &r, 1, &v->thr, &v->thr_ge, &obj, flip, ctx->tile_cursor, records, ctx->counts_dev + CNT_REC, rec_cap,
                                  ctx->sm_count, ctx->stream));
      FrameParams none;
      memset(&none, 0, sizeof(none));
      LAUNCH_OK(c3d_k_expand(&none, 0, 0.f, 0.f, records, ctx->counts_dev + CNT_REC, rec_cap, *list, ctx->counts_dev + CNT_LIST, cap,
                             ctx->bucket_hist, ctx->bucket_cursors, ctx->counts_dev + CNT_OBJ, 1u, ctx->sm_count, ctx->stream));
    } else {
      LAUNCH_OK(c3d_k_sweep(v, &r, 0u, *list, ctx->counts_dev + CNT_LIST, cap, ctx->sm_count, ctx->stream));
    }
    CUDA_OK(cudaMemcpyAsync(ctx->counts_host, ctx->counts_dev, CNT_TOTAL * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    if (records) CUDA_OK(cudaFree(records));
    *count = ctx->counts_host[CNT_LIST];
    const size_t nrec = (size_t)ctx->counts_host[CNT_REC];
    if (*count <= cap && nrec <= rec_cap) return 0;
    CUDA_OK(cudaFree(*list));
    *list = NULL;
    if (*count > cap) cap = (size_t)*count + 1024;
    if (nrec > rec_cap) rec_cap = nrec + 1024;
  }
  return c3d_error("sweep buffers did not converge");
}

extern "C" int c3d_sweep(c3d_context_t *ctx, const c3d_field_t *f, double threshold, int rendered_only, size_t capacity,
                         uint64_t *host_entries, uint64_t *count) {
  if (NULL == ctx || NULL == f || NULL == count) return c3d_error("c3d_sweep: NULL argument");
  CUDA_OK(cudaSetDevice(ctx->device));
  FieldView v;
  if (0 != stage_view(ctx, f, NULL, threshold, &v)) return 1;
  unsigned long long *list = NULL, n = 0;
  if (0 != sweep_to_device(ctx, &v, rendered_only, &list, &n)) return 1;
  *count = n;
  const size_t take = std::min<size_t>((size_t)n, capacity);
  if (take > 0 && NULL != host_entries) {
    CUDA_OK(cudaMemcpyAsync(host_entries, list, take * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    /* device entries carry (i, j, k); the C-ABI documents the flat cube index */
    const unsigned long long cx = (unsigned long long)(f->ext[0] - 1), cy = (unsigned long long)(f->ext[1] - 1);
    for (size_t q = 0; q < take; q++) {
      const unsigned long long e = host_entries[q];
      const unsigned long long cube = ((unsigned long long)C3D_ENTRY_K(e) * cy + (unsigned long long)C3D_ENTRY_J(e)) * cx + (unsigned long long)C3D_ENTRY_I(e);
      host_entries[q] = (cube << 8) | C3D_ENTRY_CODE(e);
    }
  }
  cudaFree(list);
  return 0;
}

extern "C" int c3d_tessellate(c3d_context_t *ctx, const c3d_field_t *f, const double *const grids[3], double threshold,
                              size_t capacity, c3d_triangle_t *host_out, uint64_t *count) {
  if (NULL == ctx || NULL == f || NULL == grids || NULL == count) return c3d_error("c3d_tessellate: NULL argument");
  static_assert(sizeof(c3d_triangle_t) == 208, "record layout");
  CUDA_OK(cudaSetDevice(ctx->device));
  FieldView v;
  if (0 != stage_view(ctx, f, grids, threshold, &v)) return 1;
  unsigned long long *list = NULL, ncubes = 0;
  if (0 != sweep_to_device(ctx, &v, 0, &list, &ncubes)) return 1;
  /* at most 12 triangles per cube: size the record buffer from the exact count */
  std::vector<unsigned long long> entries((size_t)ncubes);
  if (ncubes) CUDA_OK(cudaMemcpy(entries.data(), list, (size_t)ncubes * 8, cudaMemcpyDeviceToHost));
  c3d_tables_t *tab = (c3d_tables_t *)malloc(sizeof(c3d_tables_t));
  c3d_tables_build(tab);
  unsigned long long total = 0;
  for (unsigned long long e : entries) total += tab->count[e & 255ull];
  free(tab);
  *count = total;
  if (0 == total || NULL == host_out || 0 == capacity) { cudaFree(list); return 0; }
  void *records = NULL;
  CUDA_OK(cudaMalloc(&records, (size_t)total * sizeof(c3d_triangle_t)));
  CUDA_OK(cudaMemsetAsync(ctx->counts_dev, 0, 8, ctx->stream));
  LAUNCH_OK(c3d_k_emit(&v, &ctx->tab, list, ncubes, ctx->counts_dev, total, records, ctx->stream));
  std::vector<c3d_triangle_t> all((size_t)total);
  CUDA_OK(cudaMemcpyAsync(all.data(), records, (size_t)total * sizeof(c3d_triangle_t), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  cudaFree(records);
  cudaFree(list);
  std::sort(all.begin(), all.end(), [](const c3d_triangle_t &a, const c3d_triangle_t &b) {
    return a.cube != b.cube ? a.cube < b.cube : a.tri < b.tri;
  });
  memcpy(host_out, all.data(), std::min<size_t>((size_t)total, capacity) * sizeof(c3d_triangle_t));
  return 0;
}
