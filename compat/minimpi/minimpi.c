/*
 * minimpi.c -- single-node MPI stand-in; see mpi.h for scope and launch modes.
 *
 * Mechanics: one shared arena = header (process-shared barrier) + one message
 * slot per rank.  Every communication call is executed collectively:
 *   writer side : pack my contribution into MY slot
 *   barrier
 *   reader side : read the slot(s) of the rank(s) I depend on
 *   barrier
 * Slots are virtual-memory reservations (MAP_NORESERVE / sparse tmpfs), so a
 * generous size costs nothing until it is touched.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <unistd.h>
#include <errno.h>
#include <fcntl.h>
#include <time.h>
#include <pthread.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include "mpi.h"

#define MM_MAX_TYPES 256
#define MM_MAX_OPS   32
#define MM_MAX_COMMS 32
#define MM_MAX_DIMS  8
#define MM_MAGIC     0x6d696e696d706921ull

typedef struct {
  uint64_t magic;
  int nranks;
  size_t slot_bytes;
  pthread_barrier_t barrier;
} mm_header_t;

/* datatype = list of (displacement, blocklength, oldtype) + size/extent;
   "dense" types (size == extent, no holes) move with one memcpy */
typedef struct {
  int used;
  int dense;
  size_t size;    /* bytes of payload */
  size_t extent;  /* bytes spanned by one element */
  /* hvector form */
  int count, blocklen;
  long stride;
  MPI_Datatype oldtype;
} mm_type_t;

typedef struct {
  int used;
  int ndims;
  int dims[MM_MAX_DIMS];
  int periods[MM_MAX_DIMS];
} mm_comm_t;

static struct {
  int initialised;
  int finalised;
  int rank, size;
  int forked;            /* fork mode: children _exit in Finalize */
  pid_t *children;
  mm_header_t *hdr;
  unsigned char *slots;
  size_t slot_bytes;
  size_t arena_bytes;
  char shm_name[128];
  mm_type_t types[MM_MAX_TYPES];
  MPI_User_function *ops[MM_MAX_OPS];
  mm_comm_t comms[MM_MAX_COMMS];
} mm = { .rank = 0, .size = 1 };

static void mm_die(const char *msg) {
  fprintf(stderr, "[minimpi rank %d] fatal: %s\n", mm.rank, msg);
  fflush(stderr);
  _exit(86);
}

static void mm_builtin_types(void) {
  static const size_t sizes[] = {0, 1, 1, 1, sizeof(int), sizeof(float), sizeof(double), sizeof(unsigned long), 8};
  for (int t = 1; t <= 8; t++) {
    mm.types[t].used = 1;
    mm.types[t].dense = 1;
    mm.types[t].size = sizes[t];
    mm.types[t].extent = sizes[t];
    mm.types[t].count = 1;
    mm.types[t].blocklen = 1;
    mm.types[t].stride = (long)sizes[t];
    mm.types[t].oldtype = MPI_DATATYPE_NULL;
  }
}

static inline unsigned char *mm_slot(int rank) {
  return mm.slots + (size_t)rank * mm.slot_bytes;
}

static void mm_barrier(void) {
  if (mm.size > 1) {
    pthread_barrier_wait(&mm.hdr->barrier);
  }
}

static int mm_env_int(const char *name, int fallback) {
  const char *s = getenv(name);
  if (NULL == s || '\0' == *s) return fallback;
  return atoi(s);
}

static void mm_setup_types_once(void) {
  if (!mm.types[MPI_BYTE].used) {
    mm_builtin_types();
    mm.comms[MPI_COMM_WORLD].used = 1;
    mm.comms[MPI_COMM_SELF].used = 1;
  }
}

int MPI_Init(int *argc, char ***argv) {
  (void)argc; (void)argv;
  if (mm.initialised) return MPI_SUCCESS;
  mm_setup_types_once();
  int np_fork = mm_env_int("MINIMPI_NP", 0);
  int at_size = mm_env_int("MINIMPI_SIZE", 0);
  int at_rank = mm_env_int("MINIMPI_RANK", -1);
  const char *session = getenv("MINIMPI_SESSION");
  if (0 == np_fork && 0 == at_size && NULL != getenv("WORLD_SIZE") && NULL != getenv("RANK")) {
    /* launched by torchrun: one process per rank already exists */
    at_size = mm_env_int("WORLD_SIZE", 1);
    at_rank = mm_env_int("RANK", 0);
    if (NULL == session) session = getenv("MASTER_PORT");
  }
  const size_t slot_mb = (size_t)mm_env_int("MINIMPI_SLOT_MB", 1024);
  mm.slot_bytes = slot_mb << 20;
  if (np_fork > 1) {
    /* ---- fork mode ---- */
    mm.size = np_fork;
    mm.arena_bytes = sizeof(mm_header_t) + 4096 + mm.slot_bytes * (size_t)mm.size;
    void *arena = mmap(NULL, mm.arena_bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (MAP_FAILED == arena) mm_die("mmap of the shared arena failed");
    mm.hdr = arena;
    mm.slots = (unsigned char *)arena + 4096;
    mm.hdr->nranks = mm.size;
    mm.hdr->slot_bytes = mm.slot_bytes;
    pthread_barrierattr_t attr;
    pthread_barrierattr_init(&attr);
    pthread_barrierattr_setpshared(&attr, PTHREAD_PROCESS_SHARED);
    pthread_barrier_init(&mm.hdr->barrier, &attr, (unsigned)mm.size);
    pthread_barrierattr_destroy(&attr);
    mm.hdr->magic = MM_MAGIC;
    mm.children = calloc((size_t)mm.size, sizeof(pid_t));
    fflush(stdout);
    fflush(stderr);
    mm.rank = 0;
    for (int r = 1; r < mm.size; r++) {
      pid_t pid = fork();
      if (pid < 0) mm_die("fork failed");
      if (0 == pid) {
        mm.rank = r;
        mm.forked = 1;
        free(mm.children);
        mm.children = NULL;
        break;
      }
      mm.children[r] = pid;
    }
  } else if (at_size > 1) {
    /* ---- attach mode ---- */
    if (at_rank < 0 || at_rank >= at_size) mm_die("attach mode: bad rank");
    mm.size = at_size;
    mm.rank = at_rank;
    snprintf(mm.shm_name, sizeof(mm.shm_name), "/minimpi_%s_%ld", session ? session : "default", (long)getppid());
    mm.arena_bytes = sizeof(mm_header_t) + 4096 + mm.slot_bytes * (size_t)mm.size;
    int fd = -1;
    if (0 == mm.rank) {
      shm_unlink(mm.shm_name);
      fd = shm_open(mm.shm_name, O_CREAT | O_EXCL | O_RDWR, 0600);
      if (fd < 0) mm_die("shm_open (create) failed");
      if (0 != ftruncate(fd, (off_t)mm.arena_bytes)) mm_die("ftruncate failed");
    } else {
      for (int tries = 0; tries < 60000; tries++) {
        fd = shm_open(mm.shm_name, O_RDWR, 0600);
        if (fd >= 0) {
          struct stat st;
          if (0 == fstat(fd, &st) && (size_t)st.st_size >= mm.arena_bytes) break;
          close(fd);
          fd = -1;
        }
        usleep(1000);
      }
      if (fd < 0) mm_die("shm_open (attach) timed out");
    }
    void *arena = mmap(NULL, mm.arena_bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_NORESERVE, fd, 0);
    close(fd);
    if (MAP_FAILED == arena) mm_die("mmap of the named arena failed");
    mm.hdr = arena;
    mm.slots = (unsigned char *)arena + 4096;
    if (0 == mm.rank) {
      mm.hdr->nranks = mm.size;
      mm.hdr->slot_bytes = mm.slot_bytes;
      pthread_barrierattr_t attr;
      pthread_barrierattr_init(&attr);
      pthread_barrierattr_setpshared(&attr, PTHREAD_PROCESS_SHARED);
      pthread_barrier_init(&mm.hdr->barrier, &attr, (unsigned)mm.size);
      pthread_barrierattr_destroy(&attr);
      __sync_synchronize();
      mm.hdr->magic = MM_MAGIC;
    } else {
      volatile uint64_t *magic = &mm.hdr->magic;
      for (int tries = 0; *magic != MM_MAGIC; tries++) {
        if (tries > 60000) mm_die("arena was never initialised by rank 0");
        usleep(1000);
      }
      __sync_synchronize();
    }
    /* everyone has mapped the arena once this barrier passes -> safe to unlink the name */
    mm_barrier();
    if (0 == mm.rank) shm_unlink(mm.shm_name);
  } else {
    /* ---- single rank ---- */
    mm.size = 1;
    mm.rank = 0;
    mm.slot_bytes = 0;
  }
  mm.initialised = 1;
  return MPI_SUCCESS;
}

int MPI_Initialized(int *flag) {
  *flag = mm.initialised;
  return MPI_SUCCESS;
}

int MPI_Finalize(void) {
  if (!mm.initialised || mm.finalised) return MPI_SUCCESS;
  mm_barrier();
  mm.finalised = 1;
  if (mm.forked) {
    fflush(stdout);
    fflush(stderr);
    _exit(0);
  }
  if (NULL != mm.children) {
    for (int r = 1; r < mm.size; r++) {
      int status = 0;
      waitpid(mm.children[r], &status, 0);
    }
    free(mm.children);
    mm.children = NULL;
  }
  if (NULL != mm.hdr) {
    munmap(mm.hdr, mm.arena_bytes);
    mm.hdr = NULL;
  }
  return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm comm, int errorcode) {
  (void)comm;
  fprintf(stderr, "[minimpi rank %d] MPI_Abort(%d)\n", mm.rank, errorcode);
  fflush(stderr);
  if (NULL != mm.children) {
    for (int r = 1; r < mm.size; r++) kill(mm.children[r], 15);
  }
  _exit(errorcode ? errorcode : 1);
}

double MPI_Wtime(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1.e-9 * (double)ts.tv_nsec;
}

/* ---------------- communicators / topology ---------------- */

static int mm_comm_ok(MPI_Comm comm) {
  mm_setup_types_once();
  return comm > 0 && comm < MM_MAX_COMMS && mm.comms[comm].used;
}

int MPI_Comm_size(MPI_Comm comm, int *size) {
  if (!mm_comm_ok(comm)) return MPI_ERR_OTHER;
  *size = MPI_COMM_SELF == comm ? 1 : mm.size;
  return MPI_SUCCESS;
}

int MPI_Comm_rank(MPI_Comm comm, int *rank) {
  if (!mm_comm_ok(comm)) return MPI_ERR_OTHER;
  *rank = MPI_COMM_SELF == comm ? 0 : mm.rank;
  return MPI_SUCCESS;
}

static MPI_Comm mm_comm_new(void) {
  for (int c = 3; c < MM_MAX_COMMS; c++) {
    if (!mm.comms[c].used) {
      memset(&mm.comms[c], 0, sizeof(mm_comm_t));
      mm.comms[c].used = 1;
      return c;
    }
  }
  mm_die("out of communicator handles");
  return MPI_COMM_NULL;
}

int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *newcomm) {
  if (!mm_comm_ok(comm)) return MPI_ERR_OTHER;
  const MPI_Comm c = mm_comm_new();
  mm.comms[c] = mm.comms[comm];
  *newcomm = c;
  return MPI_SUCCESS;
}

int MPI_Comm_free(MPI_Comm *comm) {
  if (*comm >= 3 && *comm < MM_MAX_COMMS) mm.comms[*comm].used = 0;
  *comm = MPI_COMM_NULL;
  return MPI_SUCCESS;
}

/* balanced factorisation: primes of the remaining process count, largest
   first, each onto the currently smallest free dimension; free dimensions are
   then sorted non-increasing (2 -> {2,1}, 4 -> {2,2}, 8 -> {4,2}, 12 -> {4,3}) */
int MPI_Dims_create(int nnodes, int ndims, int dims[]) {
  int fixed = 1, nfree = 0;
  for (int d = 0; d < ndims; d++) {
    if (dims[d] > 0) fixed *= dims[d]; else nfree++;
  }
  if (0 == nfree) return MPI_SUCCESS;
  if (0 != nnodes % fixed) return MPI_ERR_OTHER;
  int rest = nnodes / fixed;
  int primes[64], nprimes = 0;
  for (int p = 2; p * p <= rest; p++) {
    while (0 == rest % p) { primes[nprimes++] = p; rest /= p; }
  }
  if (rest > 1) primes[nprimes++] = rest;
  int bins[MM_MAX_DIMS];
  for (int b = 0; b < nfree; b++) bins[b] = 1;
  for (int n = nprimes - 1; n >= 0; n--) {
    int smallest = 0;
    for (int b = 1; b < nfree; b++) if (bins[b] < bins[smallest]) smallest = b;
    bins[smallest] *= primes[n];
  }
  for (int a = 0; a < nfree; a++) {
    for (int b = a + 1; b < nfree; b++) {
      if (bins[b] > bins[a]) { const int t = bins[a]; bins[a] = bins[b]; bins[b] = t; }
    }
  }
  for (int d = 0, b = 0; d < ndims; d++) {
    if (dims[d] <= 0) dims[d] = bins[b++];
  }
  return MPI_SUCCESS;
}

int MPI_Cart_create(MPI_Comm comm_old, int ndims, const int dims[], const int periods[], int reorder, MPI_Comm *comm_cart) {
  (void)reorder;
  if (!mm_comm_ok(comm_old) || ndims > MM_MAX_DIMS) return MPI_ERR_OTHER;
  int total = 1;
  for (int d = 0; d < ndims; d++) total *= dims[d];
  int size = 0;
  MPI_Comm_size(comm_old, &size);
  if (total != size) {
    fprintf(stderr, "[minimpi] MPI_Cart_create: product of dims (%d) != communicator size (%d)\n", total, size);
    *comm_cart = MPI_COMM_NULL;
    return MPI_ERR_OTHER;
  }
  const MPI_Comm c = mm_comm_new();
  mm.comms[c].ndims = ndims;
  for (int d = 0; d < ndims; d++) {
    mm.comms[c].dims[d] = dims[d];
    mm.comms[c].periods[d] = periods[d] ? 1 : 0;
  }
  *comm_cart = c;
  return MPI_SUCCESS;
}

static void mm_coords(const mm_comm_t *c, int rank, int coords[]) {
  for (int d = c->ndims - 1; d >= 0; d--) {
    coords[d] = rank % c->dims[d];
    rank /= c->dims[d];
  }
}

static int mm_rank_of(const mm_comm_t *c, const int coords[]) {
  int rank = 0;
  for (int d = 0; d < c->ndims; d++) rank = rank * c->dims[d] + coords[d];
  return rank;
}

int MPI_Cart_get(MPI_Comm comm, int maxdims, int dims[], int periods[], int coords[]) {
  if (!mm_comm_ok(comm)) return MPI_ERR_OTHER;
  const mm_comm_t *c = &mm.comms[comm];
  int mine[MM_MAX_DIMS] = {0};
  mm_coords(c, mm.rank, mine);
  for (int d = 0; d < maxdims && d < c->ndims; d++) {
    dims[d] = c->dims[d];
    periods[d] = c->periods[d];
    coords[d] = mine[d];
  }
  return MPI_SUCCESS;
}

int MPI_Cart_shift(MPI_Comm comm, int direction, int disp, int *rank_source, int *rank_dest) {
  if (!mm_comm_ok(comm)) return MPI_ERR_OTHER;
  const mm_comm_t *c = &mm.comms[comm];
  if (direction < 0 || direction >= c->ndims) return MPI_ERR_OTHER;
  int mine[MM_MAX_DIMS] = {0};
  mm_coords(c, mm.rank, mine);
  const int n = c->dims[direction];
  int out[2];
  for (int side = 0; side < 2; side++) {
    int coords[MM_MAX_DIMS];
    memcpy(coords, mine, sizeof(coords));
    int x = mine[direction] + (0 == side ? -disp : disp);
    if (x < 0 || x >= n) {
      if (!c->periods[direction]) { out[side] = MPI_PROC_NULL; continue; }
      x = ((x % n) + n) % n;
    }
    coords[direction] = x;
    out[side] = mm_rank_of(c, coords);
  }
  *rank_source = out[0];
  *rank_dest = out[1];
  return MPI_SUCCESS;
}

int MPI_Cart_sub(MPI_Comm comm, const int remain_dims[], MPI_Comm *newcomm) {
  /* sub-communicators are only used by the pencil transposes, which are stubs here */
  (void)remain_dims;
  return MPI_Comm_dup(comm, newcomm);
}

/* ---------------- datatypes ---------------- */

static MPI_Datatype mm_type_new(void) {
  mm_setup_types_once();
  for (int t = 16; t < MM_MAX_TYPES; t++) {
    if (!mm.types[t].used) {
      memset(&mm.types[t], 0, sizeof(mm_type_t));
      mm.types[t].used = 1;
      return t;
    }
  }
  mm_die("out of datatype handles");
  return MPI_DATATYPE_NULL;
}

int MPI_Type_create_hvector(int count, int blocklength, MPI_Aint stride, MPI_Datatype oldtype, MPI_Datatype *newtype) {
  mm_setup_types_once();
  const mm_type_t *old = &mm.types[oldtype];
  const MPI_Datatype t = mm_type_new();
  mm_type_t *nt = &mm.types[t];
  nt->count = count;
  nt->blocklen = blocklength;
  nt->stride = stride;
  nt->oldtype = oldtype;
  nt->size = (size_t)count * (size_t)blocklength * old->size;
  nt->extent = count > 0 ? (size_t)(count - 1) * (size_t)stride + (size_t)blocklength * old->extent : 0;
  nt->dense = old->dense && (1 == count || (size_t)stride == (size_t)blocklength * old->extent);
  *newtype = t;
  return MPI_SUCCESS;
}

int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype *newtype) {
  mm_setup_types_once();
  return MPI_Type_create_hvector(1, count, 0, oldtype, newtype);
}

/* struct types are handled as opaque blobs of their C extent (holes are
   copied too, which is harmless for the reduce-with-user-op use case) */
int MPI_Type_create_struct(int count, const int blocklengths[], const MPI_Aint displacements[], const MPI_Datatype types[], MPI_Datatype *newtype) {
  mm_setup_types_once();
  size_t ub = 0, align = 1;
  for (int n = 0; n < count; n++) {
    const mm_type_t *old = &mm.types[types[n]];
    const size_t end = (size_t)displacements[n] + (size_t)blocklengths[n] * old->extent;
    if (end > ub) ub = end;
    size_t a = 1;
    if (MPI_DOUBLE == types[n] || MPI_UNSIGNED_LONG == types[n] || MPI_UINT64_T == types[n]) a = 8;
    else if (MPI_INT == types[n] || MPI_FLOAT == types[n]) a = 4;
    else if (types[n] >= 16) a = old->extent % 8 == 0 ? 8 : (old->extent % 4 == 0 ? 4 : 1);
    if (a > align) align = a;
  }
  const size_t extent = (ub + align - 1) / align * align;
  const MPI_Datatype t = mm_type_new();
  mm_type_t *nt = &mm.types[t];
  nt->count = 1;
  nt->blocklen = 1;
  nt->stride = (long)extent;
  nt->oldtype = MPI_DATATYPE_NULL;
  nt->size = extent;
  nt->extent = extent;
  nt->dense = 1;
  *newtype = t;
  return MPI_SUCCESS;
}

int MPI_Type_commit(MPI_Datatype *datatype) {
  (void)datatype;
  return MPI_SUCCESS;
}

int MPI_Type_free(MPI_Datatype *datatype) {
  if (*datatype >= 16 && *datatype < MM_MAX_TYPES) mm.types[*datatype].used = 0;
  *datatype = MPI_DATATYPE_NULL;
  return MPI_SUCCESS;
}

/* copy `count` elements of `type` between a typed (strided) buffer and a
   packed byte stream; returns the number of packed bytes */
static size_t mm_pack(int to_stream, MPI_Datatype type, size_t count, unsigned char *typed, unsigned char *stream) {
  const mm_type_t *t = &mm.types[type];
  if (t->dense) {
    const size_t bytes = count * t->size;
    if (to_stream) memcpy(stream, typed, bytes); else memcpy(typed, stream, bytes);
    return bytes;
  }
  const mm_type_t *old = &mm.types[t->oldtype];
  size_t done = 0;
  for (size_t n = 0; n < count; n++) {
    unsigned char *elem = typed + n * t->extent;
    for (int c = 0; c < t->count; c++) {
      done += mm_pack(to_stream, t->oldtype, (size_t)t->blocklen, elem + (size_t)c * (size_t)t->stride, stream + done);
    }
  }
  (void)old;
  return done;
}

static void mm_check_fit(size_t bytes) {
  if (bytes > mm.slot_bytes) {
    fprintf(stderr, "[minimpi] message of %zu bytes exceeds the slot size %zu; raise MINIMPI_SLOT_MB\n", bytes, mm.slot_bytes);
    mm_die("slot overflow");
  }
}

/* ---------------- communication (collective protocol) ---------------- */

int MPI_Barrier(MPI_Comm comm) {
  (void)comm;
  mm_barrier();
  return MPI_SUCCESS;
}

int MPI_Bcast(void *buffer, int count, MPI_Datatype datatype, int root, MPI_Comm comm) {
  (void)comm;
  if (mm.size < 2) return MPI_SUCCESS;
  const size_t bytes = (size_t)count * mm.types[datatype].size;
  mm_check_fit(bytes);
  if (mm.rank == root) mm_pack(1, datatype, (size_t)count, buffer, mm_slot(root));
  mm_barrier();
  if (mm.rank != root) mm_pack(0, datatype, (size_t)count, buffer, mm_slot(root));
  mm_barrier();
  return MPI_SUCCESS;
}

int MPI_Sendrecv(const void *sendbuf, int sendcount, MPI_Datatype sendtype, int dest, int sendtag,
                 void *recvbuf, int recvcount, MPI_Datatype recvtype, int source, int recvtag,
                 MPI_Comm comm, MPI_Status *status) {
  (void)sendtag; (void)recvtag; (void)comm; (void)status;
  mm_setup_types_once();
  if (mm.size < 2) {
    /* a single rank can only talk to itself (periodic direction with one process) */
    if (0 == dest && 0 == source) {
      const size_t bytes = (size_t)sendcount * mm.types[sendtype].size;
      unsigned char *tmp = malloc(bytes ? bytes : 1);
      mm_pack(1, sendtype, (size_t)sendcount, (unsigned char *)(uintptr_t)sendbuf, tmp);
      mm_pack(0, recvtype, (size_t)recvcount, recvbuf, tmp);
      free(tmp);
    }
    return MPI_SUCCESS;
  }
  if (MPI_PROC_NULL != dest) {
    mm_check_fit((size_t)sendcount * mm.types[sendtype].size);
    mm_pack(1, sendtype, (size_t)sendcount, (unsigned char *)(uintptr_t)sendbuf, mm_slot(mm.rank));
  }
  mm_barrier();
  if (MPI_PROC_NULL != source) {
    mm_pack(0, recvtype, (size_t)recvcount, recvbuf, mm_slot(source));
  }
  mm_barrier();
  return MPI_SUCCESS;
}

int MPI_Op_create(MPI_User_function *fn, int commute, MPI_Op *op) {
  (void)commute;
  for (int o = 8; o < MM_MAX_OPS; o++) {
    if (NULL == mm.ops[o]) {
      mm.ops[o] = fn;
      *op = o;
      return MPI_SUCCESS;
    }
  }
  return MPI_ERR_OTHER;
}

int MPI_Op_free(MPI_Op *op) {
  if (*op >= 8 && *op < MM_MAX_OPS) mm.ops[*op] = NULL;
  *op = MPI_OP_NULL;
  return MPI_SUCCESS;
}

static void mm_apply(MPI_Op op, MPI_Datatype type, int count, void *in, void *inout) {
  if (op >= 8) {
    mm.ops[op](in, inout, &count, &type);
    return;
  }
#define MM_LOOP(T) \
  for (int n = 0; n < count; n++) { \
    const T a = ((T *)in)[n]; T *b = (T *)inout + n; \
    if (MPI_MIN == op) { if (a < *b) *b = a; } \
    else if (MPI_MAX == op) { if (a > *b) *b = a; } \
    else { *b += a; } \
  }
  if (MPI_DOUBLE == type) { MM_LOOP(double) }
  else if (MPI_FLOAT == type) { MM_LOOP(float) }
  else if (MPI_INT == type) { MM_LOOP(int) }
  else if (MPI_UNSIGNED_LONG == type || MPI_UINT64_T == type) { MM_LOOP(uint64_t) }
  else if (MPI_UNSIGNED_CHAR == type || MPI_BYTE == type) { MM_LOOP(unsigned char) }
  else mm_die("builtin reduction on an unsupported datatype");
#undef MM_LOOP
}

int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype datatype, MPI_Op op, int root, MPI_Comm comm) {
  (void)comm;
  mm_setup_types_once();
  const size_t bytes = (size_t)count * mm.types[datatype].extent;
  if (mm.size < 2) {
    if (MPI_IN_PLACE != sendbuf && sendbuf != recvbuf) memcpy(recvbuf, sendbuf, bytes);
    return MPI_SUCCESS;
  }
  mm_check_fit(bytes);
  if (mm.rank != root) {
    memcpy(mm_slot(mm.rank), sendbuf, bytes);
  } else if (MPI_IN_PLACE != sendbuf) {
    memcpy(recvbuf, sendbuf, bytes);
  }
  mm_barrier();
  if (mm.rank == root) {
    for (int q = 0; q < mm.size; q++) {
      if (q != root) mm_apply(op, datatype, count, mm_slot(q), recvbuf);
    }
  }
  mm_barrier();
  return MPI_SUCCESS;
}

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype datatype, MPI_Op op, MPI_Comm comm) {
  (void)comm;
  mm_setup_types_once();
  const size_t bytes = (size_t)count * mm.types[datatype].extent;
  if (mm.size < 2) {
    if (MPI_IN_PLACE != sendbuf && sendbuf != recvbuf) memcpy(recvbuf, sendbuf, bytes);
    return MPI_SUCCESS;
  }
  mm_check_fit(bytes);
  memcpy(mm_slot(mm.rank), MPI_IN_PLACE == sendbuf ? recvbuf : sendbuf, bytes);
  mm_barrier();
  memcpy(recvbuf, mm_slot(0), bytes);
  for (int q = 1; q < mm.size; q++) {
    mm_apply(op, datatype, count, mm_slot(q), recvbuf);
  }
  mm_barrier();
  return MPI_SUCCESS;
}

int MPI_Alltoallw(const void *sendbuf, const int sendcounts[], const int sdispls[], const MPI_Datatype sendtypes[],
                  void *recvbuf, const int recvcounts[], const int rdispls[], const MPI_Datatype recvtypes[], MPI_Comm comm) {
  (void)sendbuf; (void)sendcounts; (void)sdispls; (void)sendtypes;
  (void)recvbuf; (void)recvcounts; (void)rdispls; (void)recvtypes; (void)comm;
  fprintf(stderr, "[minimpi] MPI_Alltoallw is not provided (pencil transposes are outside the hot path)\n");
  return MPI_ERR_OTHER;
}
