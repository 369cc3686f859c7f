/*
 * c3d_peer.h -- ranks of one node working on each other's device memory.
 *
 * The reference's two exchange steps are MPI calls between CPU ranks: six MPI_Sendrecv per object for the halo
 * (contour/extend.c:10-147, 211-213) and one MPI_Reduce of the whole canvas (output.c:82).  Here every rank maps the
 * device buffers of the others (CUDA IPC over NVLink / NVSwitch) and the kernels read or write them directly:
 * halo planes are pulled from the neighbour's array, cube lists are pulled by the rank that owns the screen tile,
 * shading work is pulled by the rank that owns the cube, colours are stored straight into the writer's image.
 * Ranks are lined up by a flag barrier in device memory (one st.release.sys per peer, one spinning ld.acquire.sys),
 * not by host calls.  Bootstrap (exchange of the IPC handles) goes through a small POSIX shared-memory board keyed
 * by the 128-byte id the ranks already share.
 *
 * Not part of the C-ABI.
 */
#if !defined(C3D_PEER_H)
#define C3D_PEER_H

#include <stddef.h>
#include <cuda_runtime.h>
#include "c3d_internal.h"

struct PeerBoard;

struct c3d_peer {
  int rank, size;
  bool attached;               /* the board is mapped: host-side allgather / barrier work */
  bool active;                 /* every rank has every other rank's signal block mapped: device-side exchange works */
  bool shared_device;          /* at least two ranks sit on the same GPU (test boxes): no NCCL communicator */
  bool broken;                 /* a collective step failed or timed out: the ranks are no longer in step */
  PeerBoard *board;
  size_t board_bytes;
  unsigned long long *sig;                    /* my signal block: [r] = last epoch rank r has announced to me */
  unsigned long long *sig_peer[C3D_MAX_RANKS]; /* the others' blocks, mapped */
  unsigned long long *status;                 /* device word: 1 = a flag barrier timed out */
  unsigned long long epoch;
};

/* joins the board named after id128 (rank 0 creates it) and exchanges signal blocks; on return peer->active tells
   whether device-side exchange is available.  Returns non-zero only for errors that leave the ranks out of step. */
int c3d_peer_init(c3d_peer *peer, int rank, int size, const void *id128, int device);
void c3d_peer_destroy(c3d_peer *peer);
/* host side, collective: every rank contributes `bytes` (<= 240), all[r * bytes ...] receives rank r's */
int c3d_peer_allgather(c3d_peer *peer, const void *mine, size_t bytes, void *all);
/* collective: maps the allocation `base` (a pointer returned by cudaMalloc, or NULL when this rank could not allocate)
   of every rank; mapped[rank] = base.  Non-zero when any rank failed: nothing stays mapped then. */
int c3d_peer_share(c3d_peer *peer, void *base, void *mapped[C3D_MAX_RANKS]);
/* local: closes this rank's mappings of the other ranks' allocations */
void c3d_peer_unshare(c3d_peer *peer, void *mapped[C3D_MAX_RANKS]);
/* device side, asynchronous on `s`: returns (on the stream) when every rank has reached the same barrier */
int c3d_peer_barrier(c3d_peer *peer, cudaStream_t s);

/* ---- kernels that work across ranks (c3d_peer.cu) ---- */

/* rows [their_j0, +2) of planes [k0, +nk) of the neighbour's array -> rows [my_j0, +2) of mine (communicate_in_y, extend.c:52-99) */
int c3d_k_halo_pull_y(void *mine, const void *theirs, int elem_bytes, long long nx, long long my_sy, long long my_sz, long long their_sy,
                      long long their_sz, long long my_k0, long long their_k0, long long nk, long long my_j0, long long their_j0, cudaStream_t s);
/* `count` contiguous elements (two whole planes: communicate_in_z, extend.c:101-147) */
int c3d_k_halo_pull_z(void *mine, const void *theirs, int elem_bytes, size_t count, cudaStream_t s);

struct MergeSources {
  const unsigned int *hist[C3D_MAX_RANKS];             /* rank s: cubes per (destination, depth bucket) */
  const unsigned long long *seg_start[C3D_MAX_RANKS];  /* rank s: where each destination's segment starts in its outbox */
  const unsigned long long *outbox[C3D_MAX_RANKS];
  const unsigned long long *out_cap[C3D_MAX_RANKS];    /* rank s: entries its outbox has room for */
};
/* the cube lists every rank has addressed to this one -> ONE nearest-first list (bucket-major, source-minor) */
int c3d_k_merge_lists(const MergeSources *src, int nranks, int me, unsigned long long *merge_base, unsigned long long *list,
                      unsigned long long *count, unsigned long long capacity, cudaStream_t s);

struct WorkExchange {
  const unsigned int *nseg[C3D_MAX_RANKS];             /* rank d: pixels it has resolved, per owner rank */
  const WorkItem *seg[C3D_MAX_RANKS];                  /* rank d: the items, owner o's at + o * seg_cap */
  unsigned int seg_cap;
};
#define C3D_PULL_CHUNKS (C3D_MAX_RANKS * (C3D_MAX_RANKS + 2))
struct PullChunk { unsigned int owner, d, first, count, dest; };
struct PullPlan { unsigned int nchunks, nwork; PullChunk chunk[C3D_PULL_CHUNKS]; };
/* decides which pixels this rank shades (those its own cubes won, up to a level; ranks below the level take over what the
   busiest hand off, at remote_cost times the price: remote_cost < 1 = everybody shades exactly its own) and pulls them:
   counters[0] <- items pulled, counters[2] <- those of this rank's own array (first in `work`) */
int c3d_k_work_exchange(const WorkExchange *ex, int nranks, int me, float remote_cost, PullPlan *plan, WorkItem *work, unsigned int *counters,
                        unsigned int capacity, unsigned long long *overflowed, cudaStream_t s);

/* flag[0] of every rank -> max into mine; flag[8] of every rank (two floats: what its last frame spent drawing / shading)
   -> loads[r] */
struct FlagSources { const unsigned long long *flag[C3D_MAX_RANKS]; };
int c3d_k_flags_max(const FlagSources *src, int nranks, unsigned long long *mine, unsigned long long *loads, cudaStream_t s);

#endif /* C3D_PEER_H */
