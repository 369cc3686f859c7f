/*
 * c3d_peer.cu -- device-to-device exchange between the ranks of one node (see c3d_peer.h):
 * bootstrap board in POSIX shared memory, CUDA IPC mappings, the flag barrier, and the kernels that pull
 * halo planes, cube lists and shading work out of the other ranks' memory.
 *
 * Replaces the transport of the reference's MPI_Sendrecv (contour/extend.c:10-147) and MPI_Reduce (output.c:82).
 * File:line references are to NaokiHori/Contour3D @ c66854e.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#include <string.h>
#include <errno.h>
#include <fcntl.h>
#include <time.h>
#include <unistd.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include "c3d_peer.h"

#define BOARD_MAGIC 0xC3DB0A2Du
#define BOARD_SLOT 256

struct PeerBoard {
  unsigned int magic, size, arrived, generation, attached, pad[3];
  unsigned char slot[C3D_MAX_RANKS][BOARD_SLOT];
};

static int peer_error(const char *fmt, ...) {
  va_list args;
  va_start(args, fmt);
  fprintf(stderr, "[CONTOUR3D ERROR] ");
  vfprintf(stderr, fmt, args);
  fprintf(stderr, "\n");
  fflush(stderr);
  va_end(args);
  return 1;
}

static double wall_seconds(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1.e-9 * (double)ts.tv_nsec;
}

static double peer_timeout_seconds(void) {
  const char *env = getenv("CONTOUR3D_PEER_TIMEOUT");
  const double v = (NULL != env && '\0' != *env) ? atof(env) : 0.;
  return v > 0. ? v : 60.;
}

/* all ranks of the board; generation counting, so that it can be used any number of times */
static int board_barrier(c3d_peer *peer) {
  PeerBoard *b = peer->board;
  const unsigned int gen = __atomic_load_n(&b->generation, __ATOMIC_ACQUIRE);
  if (__atomic_add_fetch(&b->arrived, 1u, __ATOMIC_ACQ_REL) == (unsigned int)peer->size) {
    __atomic_store_n(&b->arrived, 0u, __ATOMIC_RELAXED);
    __atomic_add_fetch(&b->generation, 1u, __ATOMIC_RELEASE);
    return 0;
  }
  const double t0 = wall_seconds(), limit = peer_timeout_seconds();
  unsigned int spins = 0;
  while (__atomic_load_n(&b->generation, __ATOMIC_ACQUIRE) == gen) {
    if (++spins > 2000u) {
      usleep(50);
      if (wall_seconds() - t0 > limit) {
        peer->broken = true;
        return peer_error("rank %d waited %.0f s for the other ranks (host barrier): are all %d ranks calling the library collectively?", peer->rank, limit, peer->size);
      }
    } else {
      sched_yield();
    }
  }
  return 0;
}

int c3d_peer_allgather(c3d_peer *peer, const void *mine, size_t bytes, void *all) {
  if (!peer->attached || peer->broken) return peer_error("rank %d: the ranks' shared board is not available", peer->rank);
  if (bytes > BOARD_SLOT) return peer_error("board message of %zu bytes", bytes);
  memcpy(peer->board->slot[peer->rank], mine, bytes);
  if (0 != board_barrier(peer)) return 1;
  for (int r = 0; r < peer->size; r++) memcpy((char *)all + (size_t)r * bytes, peer->board->slot[r], bytes);
  return board_barrier(peer);
}

int c3d_peer_share(c3d_peer *peer, void *base, void *mapped[C3D_MAX_RANKS]) {
  struct Msg { int ok; int pad; cudaIpcMemHandle_t h; };
  for (int r = 0; r < C3D_MAX_RANKS; r++) mapped[r] = NULL;
  Msg mine, all[C3D_MAX_RANKS];
  memset(&mine, 0, sizeof(mine));
  if (NULL != base) {
    const cudaError_t err = cudaIpcGetMemHandle(&mine.h, base);
    mine.ok = cudaSuccess == err ? 1 : 0;
    if (cudaSuccess != err) { cudaGetLastError(); peer_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(err)); }
  }
  if (0 != c3d_peer_allgather(peer, &mine, sizeof(mine), all)) return 1;
  int ok = 1;
  for (int r = 0; r < peer->size; r++) ok &= all[r].ok;
  int opened = ok;
  if (ok) {
    for (int r = 0; r < peer->size && opened; r++) {
      if (r == peer->rank) continue;
      const cudaError_t err = cudaIpcOpenMemHandle(&mapped[r], all[r].h, cudaIpcMemLazyEnablePeerAccess);
      if (cudaSuccess != err) {
        cudaGetLastError();
        peer_error("rank %d cannot map device memory of rank %d: %s", peer->rank, r, cudaGetErrorString(err));
        mapped[r] = NULL;
        opened = 0;
      }
    }
  }
  int flags[C3D_MAX_RANKS];
  if (0 != c3d_peer_allgather(peer, &opened, sizeof(opened), flags)) return 1;
  for (int r = 0; r < peer->size; r++) ok &= flags[r];
  if (!ok) {
    c3d_peer_unshare(peer, mapped);
    return 1;
  }
  mapped[peer->rank] = base;
  return 0;
}

void c3d_peer_unshare(c3d_peer *peer, void *mapped[C3D_MAX_RANKS]) {
  for (int r = 0; r < C3D_MAX_RANKS; r++) {
    if (r != peer->rank && NULL != mapped[r]) cudaIpcCloseMemHandle(mapped[r]);
    mapped[r] = NULL;
  }
}

/* Joins the ranks' board: rank 0 creates the shared-memory segment named after the run's id, the others open it; everybody
   waits until all have arrived, then the name is removed.  0: attached; 1: no board (another node, no /dev/shm, a rank
   missing): the caller goes on without device-side exchange; 2: the ranks are out of step. */
static int board_attach(c3d_peer *peer, const void *id128) {
  const int rank = peer->rank, size = peer->size;
  unsigned long long w[2];
  memcpy(w, id128, sizeof(w));
  char name[64];
  snprintf(name, sizeof(name), "/c3d_%016llx%016llx", w[0], w[1]);
  const double t0 = wall_seconds(), limit = peer_timeout_seconds();
  int fd = -1;
  if (0 == rank) {
    shm_unlink(name);
    fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0 || 0 != ftruncate(fd, (off_t)sizeof(PeerBoard))) {
      if (fd >= 0) close(fd);
      fprintf(stderr, "[CONTOUR3D] no shared-memory board (%s): ranks exchange through NCCL only\n", strerror(errno));
      return 1;
    }
  } else {
    for (;;) {
      fd = shm_open(name, O_RDWR, 0600);
      struct stat st;
      if (fd >= 0 && 0 == fstat(fd, &st) && (size_t)st.st_size >= sizeof(PeerBoard)) break;
      if (fd >= 0) { close(fd); fd = -1; }
      if (wall_seconds() - t0 > limit) {
        fprintf(stderr, "[CONTOUR3D] rank %d found no shared-memory board within %.0f s (another node?): NCCL only\n", rank, limit);
        return 1;
      }
      usleep(200);
    }
  }
  PeerBoard *b = (PeerBoard *)mmap(NULL, sizeof(PeerBoard), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (MAP_FAILED == (void *)b) return 1;
  if (0 == rank) {
    memset(b, 0, sizeof(*b));
    b->size = (unsigned int)size;
    __atomic_store_n(&b->magic, BOARD_MAGIC, __ATOMIC_RELEASE);
  } else {
    while (BOARD_MAGIC != __atomic_load_n(&b->magic, __ATOMIC_ACQUIRE)) {
      if (wall_seconds() - t0 > limit) { munmap(b, sizeof(PeerBoard)); return 1; }
      usleep(100);
    }
  }
  peer->board = b;
  peer->board_bytes = sizeof(PeerBoard);
  peer->attached = true;
  /* rank 0 waits for everybody before the name disappears (a rank on another node never arrives: then nobody uses the board) */
  __atomic_add_fetch(&b->attached, 1u, __ATOMIC_ACQ_REL);
  while (__atomic_load_n(&b->attached, __ATOMIC_ACQUIRE) < (unsigned int)size) {
    if (wall_seconds() - t0 > limit) {
      fprintf(stderr, "[CONTOUR3D] rank %d: only %u of %d ranks reached the shared-memory board: NCCL only\n", rank, b->attached, size);
      if (0 == rank) shm_unlink(name);
      munmap(b, sizeof(PeerBoard));
      peer->board = NULL;
      peer->attached = false;
      return 1;
    }
    usleep(100);
  }
  if (0 != board_barrier(peer)) return 2;
  if (0 == rank) shm_unlink(name);
  return 0;
}

/* Host-side self test of the bootstrap (no device needed): every rank joins the board named after id128, contributes
   `value`, and receives the sum over the ranks of (rank + 1) * value -- three allgathers in a row, so that the generation
   counting of the board's barrier is exercised.  Collective.  Returns 0, or 1 when the board is not available. */
extern "C" int c3d_peer_selftest(int rank, int size, const void *id128, unsigned long long value, unsigned long long *sum) {
  if (NULL == id128 || NULL == sum || size < 2 || size > C3D_MAX_RANKS || rank < 0 || rank >= size) return 1;
  c3d_peer peer;
  memset(&peer, 0, sizeof(peer));
  peer.rank = rank;
  peer.size = size;
  if (0 != board_attach(&peer, id128)) return 1;
  unsigned long long total = 0ull;
  int rc = 0;
  for (int round = 0; round < 3 && 0 == rc; round++) {
    unsigned long long mine = (unsigned long long)(rank + 1) * value + (unsigned long long)round, all[C3D_MAX_RANKS];
    rc = c3d_peer_allgather(&peer, &mine, sizeof(mine), all);
    total = 0ull;
    for (int r = 0; r < size && 0 == rc; r++) total += all[r] - (unsigned long long)round;
  }
  *sum = total;
  munmap(peer.board, peer.board_bytes);
  return rc;
}

int c3d_peer_init(c3d_peer *peer, int rank, int size, const void *id128, int device) {
  memset(peer, 0, sizeof(*peer));
  peer->rank = rank;
  peer->size = size;
  if (size < 2) return 0;
  {
    const char *env = getenv("CONTOUR3D_P2P");
    if (NULL != env && '0' == env[0]) return 0;                 /* exchange through NCCL only */
  }
  if (size > C3D_MAX_RANKS) return 0;                           /* list entries name the source rank in 3 bits */
  if (NULL == id128) return 0;
  {
    const int rc = board_attach(peer, id128);
    if (2 == rc) return 1;
    if (0 != rc) return 0;
  }
  /* who shares a GPU with whom */
  struct Ident { char bus[48]; };
  Ident mine, all[C3D_MAX_RANKS];
  memset(&mine, 0, sizeof(mine));
  if (cudaSuccess != cudaDeviceGetPCIBusId(mine.bus, (int)sizeof(mine.bus), device)) { cudaGetLastError(); snprintf(mine.bus, sizeof(mine.bus), "rank%d", rank); }
  if (0 != c3d_peer_allgather(peer, &mine, sizeof(mine), all)) return 1;
  for (int a = 0; a < size; a++) {
    for (int c = a + 1; c < size; c++) {
      if (0 == strncmp(all[a].bus, all[c].bus, sizeof(mine.bus))) peer->shared_device = true;
    }
  }
  /* signal blocks */
  unsigned long long *sig = NULL;
  if (cudaSuccess != cudaMalloc(&sig, 4096) || cudaSuccess != cudaMemset(sig, 0, 4096)) {
    cudaGetLastError();
    if (sig) cudaFree(sig);
    sig = NULL;
  }
  void *mapped[C3D_MAX_RANKS];
  if (0 != c3d_peer_share(peer, sig, mapped)) {
    if (peer->broken) return 1;
    if (sig) cudaFree(sig);
    fprintf(stderr, "[CONTOUR3D] rank %d: no peer access between the ranks' GPUs: NCCL only\n", rank);
    return 0;
  }
  peer->sig = sig;
  peer->status = sig + 64;
  for (int r = 0; r < size; r++) peer->sig_peer[r] = (unsigned long long *)mapped[r];
  peer->active = true;
  return 0;
}

void c3d_peer_destroy(c3d_peer *peer) {
  if (peer->active) {
    void *mapped[C3D_MAX_RANKS];
    for (int r = 0; r < C3D_MAX_RANKS; r++) mapped[r] = r < peer->size ? (void *)peer->sig_peer[r] : NULL;
    c3d_peer_unshare(peer, mapped);
    cudaFree(peer->sig);
  }
  if (NULL != peer->board) munmap(peer->board, peer->board_bytes);
  memset(peer, 0, sizeof(*peer));
}

/* ------------------------------------------------------------ flag barrier */

struct PeerSignals { unsigned long long *peer[C3D_MAX_RANKS]; unsigned long long *mine; };

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

/*
 * Thread r announces this rank's epoch in rank r's block and waits until rank r has announced at least the same
 * epoch in ours.  Everything this rank wrote before (own memory or a peer's) is ordered before the announcement
 * (fence + release at system scope); what the peers wrote before theirs is visible to the kernels that follow.
 */
__global__ void k_peer_barrier(const PeerSignals sig, const int me, const int n, const unsigned long long epoch, unsigned long long *status,
                               const unsigned long long timeout_ns) {
  const int r = (int)threadIdx.x;
  if (r < n && r != me) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(sig.peer[r] + me), "l"(epoch) : "memory");
    const unsigned long long t0 = global_timer_ns();
    unsigned long long seen = 0ull;
    unsigned int spins = 0u;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(sig.mine + r) : "memory");
      if (seen >= epoch) break;
      if (0u == (++spins & 1023u) && global_timer_ns() - t0 > timeout_ns) {
        atomicExch(status, 1ull);
        break;
      }
      __nanosleep(64);
    }
    __threadfence_system();
  }
}

int c3d_peer_barrier(c3d_peer *peer, cudaStream_t s) {
  if (!peer->active) return (int)cudaErrorNotReady;
  PeerSignals sig;
  for (int r = 0; r < C3D_MAX_RANKS; r++) sig.peer[r] = r < peer->size ? peer->sig_peer[r] : NULL;
  sig.mine = peer->sig;
  peer->epoch++;
  k_peer_barrier<<<1, 32, 0, s>>>(sig, peer->rank, peer->size, peer->epoch, peer->status, (unsigned long long)(peer_timeout_seconds() * 1.e9));
  return (int)cudaGetLastError();
}

/* --------------------------------------------------------------- halo pull */

template <typename T>
__global__ void __launch_bounds__(256) k_halo_pull_y(T *__restrict__ mine, const T *__restrict__ theirs, const long long nx, const long long my_sy,
                                                     const long long my_sz, const long long their_sy, const long long their_sz, const long long my_k0,
                                                     const long long their_k0, const long long nk, const long long my_j0, const long long their_j0) {
  const long long total = nk * 2 * nx;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const long long i = q % nx, r = (q / nx) % 2, k = q / (2 * nx);
    mine[(my_k0 + k) * my_sz + (my_j0 + r) * my_sy + i] = theirs[(their_k0 + k) * their_sz + (their_j0 + r) * their_sy + i];
  }
}

int c3d_k_halo_pull_y(void *mine, const void *theirs, int elem_bytes, long long nx, long long my_sy, long long my_sz, long long their_sy,
                      long long their_sz, long long my_k0, long long their_k0, long long nk, long long my_j0, long long their_j0, cudaStream_t s) {
  if (4 == elem_bytes)
    k_halo_pull_y<float><<<148 * 4, 256, 0, s>>>((float *)mine, (const float *)theirs, nx, my_sy, my_sz, their_sy, their_sz, my_k0, their_k0, nk, my_j0, their_j0);
  else
    k_halo_pull_y<double><<<148 * 4, 256, 0, s>>>((double *)mine, (const double *)theirs, nx, my_sy, my_sz, their_sy, their_sz, my_k0, their_k0, nk, my_j0, their_j0);
  return (int)cudaGetLastError();
}

template <typename V>
__global__ void __launch_bounds__(256) k_copy_from_peer(V *__restrict__ mine, const V *__restrict__ theirs, const size_t n) {
  /* several independent 16-byte requests per thread in flight: a peer read is a round trip over NVLink */
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; q + 3 * stride < n; q += 4 * stride) {
    const V a = theirs[q], b = theirs[q + stride], c = theirs[q + 2 * stride], d = theirs[q + 3 * stride];
    mine[q] = a; mine[q + stride] = b; mine[q + 2 * stride] = c; mine[q + 3 * stride] = d;
  }
  for (; q < n; q += stride) mine[q] = theirs[q];
}

int c3d_k_halo_pull_z(void *mine, const void *theirs, int elem_bytes, size_t count, cudaStream_t s) {
  const size_t bytes = count * (size_t)elem_bytes;
  if (0 == bytes) return 0;
  if (0 == (((uintptr_t)mine | (uintptr_t)theirs | bytes) & 15u)) {
    const size_t n = bytes / 16;
    const unsigned int blocks = (unsigned int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    k_copy_from_peer<uint4><<<blocks, 256, 0, s>>>((uint4 *)mine, (const uint4 *)theirs, n);
  } else if (4 == elem_bytes) {
    k_copy_from_peer<float><<<148 * 8, 256, 0, s>>>((float *)mine, (const float *)theirs, count);
  } else {
    k_copy_from_peer<double><<<148 * 8, 256, 0, s>>>((double *)mine, (const double *)theirs, count);
  }
  return (int)cudaGetLastError();
}

/* -------------------------------------------------------------- list merge */

/*
 * Every source rank s has sorted the cubes it addresses to me by view depth (counting sort, C3D_PART_BUCKETS slabs
 * of the SAME global depth range on all ranks).  The merged list is bucket-major, source-minor: still nearest first,
 * which is all the raster's occlusion test wants.  merge_base[b * n + s] = first merged index of (bucket b, source s).
 */
__global__ void __launch_bounds__(1024) k_merge_scan(const MergeSources src, const int n, const int me, unsigned long long *__restrict__ merge_base,
                                                     unsigned long long *__restrict__ count) {
  __shared__ unsigned long long part[1024];
  const unsigned int per = (unsigned int)(C3D_PART_BUCKETS / 1024) * (unsigned int)n;     /* consecutive bins of a thread */
  const unsigned int first = threadIdx.x * per;
  unsigned long long mine = 0ull;
  for (unsigned int q = 0; q < per; q++) {
    const unsigned int bin = first + q, b = bin / (unsigned int)n, s = bin % (unsigned int)n;
    mine += src.hist[s][(unsigned int)me * C3D_PART_BUCKETS + b];
  }
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
    const unsigned int bin = first + q, b = bin / (unsigned int)n, s = bin % (unsigned int)n;
    merge_base[bin] = at;
    at += src.hist[s][(unsigned int)me * C3D_PART_BUCKETS + b];
  }
  if (1023u == threadIdx.x) *count = part[1023];
}

#define MERGE_THREADS 256
#define MERGE_QUADS 4          /* 16-byte requests a thread keeps in flight: a peer read is a round trip over NVLink */
__global__ void __launch_bounds__(MERGE_THREADS) k_merge_copy(const MergeSources src, const int n, const int me,
                                                              const unsigned long long *__restrict__ merge_base,
                                                              unsigned long long *__restrict__ list, const unsigned long long capacity) {
  __shared__ unsigned int prefix[C3D_PART_BUCKETS + 1];       /* of source s: entries for me before bucket b */
  __shared__ unsigned int warp_sum[MERGE_THREADS / 32];
  const int s = (int)blockIdx.y;
  const unsigned int *hist = src.hist[s] + (unsigned int)me * C3D_PART_BUCKETS;
  constexpr unsigned int PER = C3D_PART_BUCKETS / MERGE_THREADS;
  unsigned int v[PER], sum = 0u;
#pragma unroll
  for (unsigned int q = 0; q < PER; q++) { v[q] = hist[threadIdx.x * PER + q]; sum += v[q]; }
  unsigned int incl = sum;
  const unsigned int lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned int up = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= (unsigned int)d) incl += up;
  }
  if (31u == lane) warp_sum[wid] = incl;
  __syncthreads();
  unsigned int before = 0u;
  for (unsigned int w = 0; w < wid; w++) before += warp_sum[w];
  unsigned int at = before + incl - sum;
#pragma unroll
  for (unsigned int q = 0; q < PER; q++) { prefix[threadIdx.x * PER + q] = at; at += v[q]; }
  if (MERGE_THREADS - 1 == threadIdx.x) prefix[C3D_PART_BUCKETS] = at;
  __syncthreads();
  const unsigned int total = prefix[C3D_PART_BUCKETS];
  const unsigned long long start = src.seg_start[s][me];
  const unsigned long long readable = *src.out_cap[s];         /* entries rank s's outbox really holds (the counts may exceed it: overflow, frame redone) */
  const unsigned long long *outbox = src.outbox[s] + start;
  /* entries travel in aligned pairs (16 bytes); the segment may start on an odd entry */
  const unsigned int odd = (unsigned int)(start & 1ull);
  const unsigned int npairs = (total + odd + 1u) / 2u;
  const ulonglong2 *pairs = reinterpret_cast<const ulonglong2 *>(outbox - odd);
  auto place = [&](const unsigned int e, const unsigned long long entry) {
    /* largest b with prefix[b] <= e */
    unsigned int lo = 0u, hi = C3D_PART_BUCKETS;
    while (hi - lo > 1u) {
      const unsigned int mid = (lo + hi) >> 1;
      if (prefix[mid] <= e) lo = mid; else hi = mid;
    }
    const unsigned long long to = merge_base[lo * (unsigned int)n + (unsigned int)s] + (e - prefix[lo]);
    if (to < capacity) list[to] = entry;
  };
  const unsigned int stride = gridDim.x * MERGE_THREADS;
  for (unsigned int p0 = blockIdx.x * MERGE_THREADS + threadIdx.x; p0 < npairs; p0 += MERGE_QUADS * stride) {
    ulonglong2 got[MERGE_QUADS];
#pragma unroll
    for (int q = 0; q < MERGE_QUADS; q++) {
      const unsigned int p = p0 + (unsigned int)q * stride;
      got[q] = (p < npairs && start - odd + 2ull * p + 1ull < readable) ? pairs[p] : make_ulonglong2(0ull, 0ull);
    }
#pragma unroll
    for (int q = 0; q < MERGE_QUADS; q++) {
      const unsigned int p = p0 + (unsigned int)q * stride;
      if (p >= npairs) break;
      const unsigned int first = 2u * p;                       /* index of got.x counted from the aligned start */
      if (first >= odd && first - odd < total) place(first - odd, got[q].x);
      if (first + 1u >= odd && first + 1u - odd < total) place(first + 1u - odd, got[q].y);
    }
  }
}

int c3d_k_merge_lists(const MergeSources *src, int nranks, int me, unsigned long long *merge_base, unsigned long long *list,
                      unsigned long long *count, unsigned long long capacity, cudaStream_t s) {
  k_merge_scan<<<1, 1024, 0, s>>>(*src, nranks, me, merge_base, count);
  k_merge_copy<<<dim3(148, (unsigned int)nranks), MERGE_THREADS, 0, s>>>(*src, nranks, me, merge_base, list, capacity);
  return (int)cudaGetLastError();
}

/* --------------------------------------------------------------- work pull */

/*
 * Who shades what.  Rank d has queued, per owner o, the pixels of its tiles won by cubes of o's array: count C[d][o], items
 * in seg[d][o].  Owner o's pixels in the order d = 0, 1, .. form its list of T[o] = sum_d C[d][o] items.  The front slab of
 * a z-decomposed volume wins a third of the image on its own; ranks with little to shade take some of it over, reading the
 * owner's array through peer memory -- at `remote_cost` times the price of a pixel of their own (measured: about five).
 * The plan levels the ESTIMATED TIME: the level tau solves  sum_{T > tau} (T - tau) = sum_{T < tau} (tau - T) / remote_cost;
 * owners above keep their first tau items, ranks below fill up to tau with foreign ones, dealt out in rank order.
 * Every rank runs this same plan on the same counts and writes down only what IT has to pull.
 * counters[0] <- all items of this rank, counters[2] <- those of its own array (they come first).
 */
/* the plan itself: plain arithmetic on the n x n counts, the same on the host (c3d_selftest_work_plan) and on the device.
   *nlocal = items of this rank's own array (they come first), *total = all items it pulls */
__host__ __device__ static void work_plan_compute(const unsigned int (*C)[C3D_MAX_RANKS], const int n, const int me, const double remote_cost,
                                                  PullPlan *plan, unsigned int *nlocal, unsigned int *total) {
  unsigned int T[C3D_MAX_RANKS], keep[C3D_MAX_RANKS], room[C3D_MAX_RANKS];
  unsigned int most = 0u;
  for (int o = 0; o < n; o++) {
    T[o] = 0u;
    for (int d = 0; d < n; d++) T[o] += C[d][o];
    most = T[o] > most ? T[o] : most;
  }
  const bool help = remote_cost >= 1.;
  double tau = (double)most;                     /* counts are far below 2^53: exact in double */
  if (help) {
    /* one thread of one block runs this every frame: 36 halvings bring the level to a fraction of a pixel, the division is hoisted */
    const double inv_cost = 1. / remote_cost;
    double lo = 0., hi = (double)most;
    for (int it = 0; it < 36; it++) {
      const double mid = 0.5 * (lo + hi);
      double give = 0., takes = 0.;
      for (int o = 0; o < n; o++) {
        if ((double)T[o] > mid) give += (double)T[o] - mid; else takes += (mid - (double)T[o]) * inv_cost;
      }
      if (give > takes) lo = mid; else hi = mid;
    }
    tau = hi;
  }
  auto room_of = [&](const int o) -> unsigned int {
    return (help && (double)T[o] < tau) ? (unsigned int)((tau - (double)T[o]) / remote_cost) : 0u;
  };
  for (int o = 0; o < n; o++) {
    const unsigned int level = (unsigned int)tau;
    keep[o] = T[o] > level ? level : T[o];
    room[o] = room_of(o);
  }
  unsigned int nchunks = 0u, dest = 0u;
  /* a range [a, a + cnt) of owner o's list, as pieces of the ranks' segments */
  auto take = [&](const int o, unsigned int a, unsigned int cnt) {
    unsigned int before = 0u;
    for (int d = 0; d < n && cnt > 0u; d++) {
      const unsigned int end = before + C[d][o];
      if (a < end) {
        const unsigned int piece = (end - a) < cnt ? (end - a) : cnt;
        PullChunk c;
        c.owner = (unsigned int)o; c.d = (unsigned int)d; c.first = a - before; c.count = piece; c.dest = dest;
        if (nchunks < C3D_PULL_CHUNKS) plan->chunk[nchunks++] = c;
        dest += piece; a += piece; cnt -= piece;
      }
      before = end;
    }
  };
  /* this rank's own: what it keeps, plus what nobody had room for (rounding) */
  unsigned int left_over[C3D_MAX_RANKS], handed[C3D_MAX_RANKS];
  {
    int r = 0;
    for (int o = 0; o < n; o++) {
      unsigned int left = T[o] - keep[o];
      handed[o] = 0u;
      while (left > 0u) {
        while (r < n && 0u == room[r]) r++;
        if (r >= n) break;
        const unsigned int give = left < room[r] ? left : room[r];
        room[r] -= give; left -= give; handed[o] += give;
      }
      left_over[o] = left;
    }
  }
  take(me, 0u, keep[me]);
  if (left_over[me] > 0u) take(me, keep[me] + handed[me], left_over[me]);
  *nlocal = dest;
  /* the same dealing once more, now writing down the pieces that come to this rank */
  for (int o = 0; o < n; o++) room[o] = room_of(o);
  {
    int r = 0;
    for (int o = 0; o < n; o++) {
      unsigned int a = keep[o], left = T[o] - keep[o];
      while (left > 0u) {
        while (r < n && 0u == room[r]) r++;
        if (r >= n) break;
        const unsigned int give = left < room[r] ? left : room[r];
        if (r == me) take(o, a, give);
        room[r] -= give; a += give; left -= give;
      }
    }
  }
  plan->nchunks = nchunks;
  *total = dest;
}

__global__ void k_work_plan(const WorkExchange ex, const int n, const int me, const float remote_cost, PullPlan *__restrict__ plan,
                            unsigned int *__restrict__ counters, const unsigned int capacity, unsigned long long *__restrict__ overflowed) {
  /* the counts live on the other ranks: one thread each fetches them, then one thread plans */
  __shared__ unsigned int s_C[C3D_MAX_RANKS][C3D_MAX_RANKS];
  if ((int)threadIdx.x < n * n) s_C[threadIdx.x / n][threadIdx.x % n] = ex.nseg[threadIdx.x / n][threadIdx.x % n];
  __syncthreads();
  if (0 != threadIdx.x || 0 != blockIdx.x) return;
  unsigned int nlocal = 0u, total = 0u;
  work_plan_compute(s_C, n, me, (double)remote_cost, plan, &nlocal, &total);
  plan->nwork = total < capacity ? total : capacity;
  counters[0] = plan->nwork;
  counters[2] = nlocal < plan->nwork ? nlocal : plan->nwork;
  if (total > capacity) atomicAdd(overflowed, 1ull);
}

/* host-side self test hook: the plan of rank `me` for the counts matrix counts[d * n + o]; chunks as 5 words each
   (owner, d, first, count, dest).  Needs no device. */
extern "C" int c3d_selftest_work_plan(const unsigned int *counts, int n, int me, double remote_cost, unsigned int *chunks, unsigned int *nchunks,
                                      unsigned int *nlocal, unsigned int *total) {
  if (NULL == counts || NULL == chunks || NULL == nchunks || NULL == nlocal || NULL == total || n < 1 || n > C3D_MAX_RANKS || me < 0 || me >= n) return 1;
  unsigned int C[C3D_MAX_RANKS][C3D_MAX_RANKS];
  memset(C, 0, sizeof(C));
  for (int d = 0; d < n; d++) {
    for (int o = 0; o < n; o++) C[d][o] = counts[d * n + o];
  }
  PullPlan plan;
  memset(&plan, 0, sizeof(plan));
  work_plan_compute(C, n, me, remote_cost, &plan, nlocal, total);
  *nchunks = plan.nchunks;
  for (unsigned int c = 0; c < plan.nchunks; c++) {
    chunks[5 * c + 0] = plan.chunk[c].owner; chunks[5 * c + 1] = plan.chunk[c].d; chunks[5 * c + 2] = plan.chunk[c].first;
    chunks[5 * c + 3] = plan.chunk[c].count; chunks[5 * c + 4] = plan.chunk[c].dest;
  }
  return 0;
}

__global__ void __launch_bounds__(256) k_work_copy(const WorkExchange ex, const PullPlan *__restrict__ plan, WorkItem *__restrict__ work,
                                                   const unsigned int capacity) {
  const unsigned int nchunks = plan->nchunks;
  uint4 *to = reinterpret_cast<uint4 *>(work);
  const unsigned int stride = gridDim.x * blockDim.x;
  for (unsigned int c = 0; c < nchunks; c++) {
    const PullChunk ch = plan->chunk[c];
    const uint4 *from = reinterpret_cast<const uint4 *>(ex.seg[ch.d] + (size_t)ch.owner * ex.seg_cap + ch.first);
    unsigned int q = blockIdx.x * blockDim.x + threadIdx.x;
    /* four independent 16-byte requests per thread in flight (peer reads) */
    for (; q + 3u * stride < ch.count; q += 4u * stride) {
      const uint4 a = from[q], b = from[q + stride], cc = from[q + 2u * stride], d = from[q + 3u * stride];
      if (ch.dest + q + 3u * stride < capacity) { to[ch.dest + q] = a; to[ch.dest + q + stride] = b; to[ch.dest + q + 2u * stride] = cc; to[ch.dest + q + 3u * stride] = d; }
    }
    for (; q < ch.count; q += stride) {
      if (ch.dest + q < capacity) to[ch.dest + q] = from[q];
    }
  }
}

int c3d_k_work_exchange(const WorkExchange *ex, int nranks, int me, float remote_cost, PullPlan *plan, WorkItem *work, unsigned int *counters,
                        unsigned int capacity, unsigned long long *overflowed, cudaStream_t s) {
  k_work_plan<<<1, 64, 0, s>>>(*ex, nranks, me, remote_cost, plan, counters, capacity, overflowed);
  k_work_copy<<<148 * 2, 256, 0, s>>>(*ex, plan, work, capacity);
  return (int)cudaGetLastError();
}

__global__ void k_flags_max(const FlagSources src, const int n, unsigned long long *mine, unsigned long long *loads) {
  const int r = (int)threadIdx.x;
  if (r < n) loads[r] = src.flag[r][8];
  if (0 == r) {
    unsigned long long m = 0ull;
    for (int q = 0; q < n; q++) {
      const unsigned long long v = *src.flag[q];
      m = v > m ? v : m;
    }
    *mine = m;
  }
}

int c3d_k_flags_max(const FlagSources *src, int nranks, unsigned long long *mine, unsigned long long *loads, cudaStream_t s) {
  k_flags_max<<<1, 32, 0, s>>>(*src, nranks, mine, loads);
  return (int)cudaGetLastError();
}
