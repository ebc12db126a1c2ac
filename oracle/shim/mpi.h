/*
 * mpi.h -- a tiny stand-in for the MPI C API, just large enough to compile and run the
 * UNMODIFIED reference solver (/root/reference/d2q9-bgk.c) where no MPI runtime exists.
 *
 * TEST INFRASTRUCTURE ONLY.  This file lives under oracle/ and is used solely by
 * oracle/Makefile to build the reference into oracle/_ref/ (the parity checker and the
 * CPU baseline).  Nothing in the shipped product includes it.
 *
 * It implements the 17 entry points the reference calls (d2q9-bgk.c:185-187, 295-313,
 * 327, 364, 396, 427-437, 755, 828-832, 966-970, 1029, 1051, 276/405) for N ranks that
 * are plain fork()ed processes sharing one anonymous mmap region:
 *
 *   - MPI_Init forks N-1 children (N = $MPI_SHIM_RANKS, default 1).
 *   - Persistent point-to-point requests move whole messages through per-(sender,
 *     direction) FIFO mailboxes in the shared region.  Sends are copied eagerly at
 *     MPI_Startall, receives are drained at MPI_Waitall, so MPI's non-overtaking rule
 *     holds: the first send posted to a peer matches the first receive posted for it.
 *     With one rank the ring neighbours are the rank itself and this yields exactly the
 *     periodic wrap of the reference (halo above the last row <- first row, halo below
 *     the first row <- last row).
 *   - Reduce/Bcast/Scatterv stream through a shared staging buffer between barriers.
 *
 * Only what the reference needs is supported: MPI_COMM_WORLD, MPI_INT/MPI_FLOAT and
 * contiguous types built from them, MPI_SUM on floats, neighbour-only point-to-point.
 */
#ifndef LBM_ORACLE_MPI_SHIM_H
#define LBM_ORACLE_MPI_SHIM_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stddef.h>
#include <sched.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/wait.h>

typedef size_t MPI_Datatype;      /* a datatype is its extent in bytes */
typedef int MPI_Comm;
typedef int MPI_Op;
typedef struct { int unused; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_INT   ((MPI_Datatype)sizeof(int))
#define MPI_FLOAT ((MPI_Datatype)sizeof(float))
#define MPI_SUM 1
#define MPI_SUCCESS 0
#define MPI_STATUS_IGNORE   ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)

#define SHIM_MAX_RANKS   64
#define SHIM_SLOTS       8                 /* messages in flight per mailbox */
#define SHIM_SLOT_BYTES  (1u << 20)        /* one row of cells: nx*36 B, so nx <= 29127 */
#define SHIM_STAGE_BYTES (64u << 20)       /* staging buffer for the collectives */

typedef struct {
  void* buf;
  size_t bytes;
  int peer;
  int is_send;
} MPI_Request;

typedef struct {
  volatile unsigned long head;             /* messages written  (sender only)   */
  volatile unsigned long tail;             /* messages consumed (receiver only) */
  char pad[48];
  char slot[SHIM_SLOTS][SHIM_SLOT_BYTES];
} shim_mailbox;

typedef struct {
  volatile int bar_count;
  volatile int bar_sense;
  char pad[56];
  shim_mailbox box[SHIM_MAX_RANKS][2];     /* [sender][0 = to rank-1, 1 = to rank+1] */
  char stage[SHIM_STAGE_BYTES];
} shim_world;

static shim_world* shim_w = NULL;
static int shim_size = 1, shim_rank = 0, shim_local_sense = 0;
static pid_t shim_kids[SHIM_MAX_RANKS];

static inline void shim_fail(const char* what)
{
  fprintf(stderr, "mpi shim: %s\n", what);
  exit(EXIT_FAILURE);
}

static inline void shim_relax(unsigned* spins)
{
  if (++*spins > 2000) { sched_yield(); *spins = 0; }
  else __builtin_ia32_pause();
}

/* which of the sender's two mailboxes carries sender -> receiver traffic */
static inline int shim_direction(int sender, int receiver)
{
  if (receiver == (sender - 1 + shim_size) % shim_size) return 0;
  if (receiver == (sender + 1) % shim_size) return 1;
  shim_fail("point-to-point between non-neighbouring ranks is not supported");
  return 0;
}

static inline int MPI_Barrier(MPI_Comm comm)
{
  (void)comm;
  if (shim_size == 1) return MPI_SUCCESS;
  unsigned spins = 0;
  shim_local_sense ^= 1;
  if (__sync_add_and_fetch(&shim_w->bar_count, 1) == shim_size) {
    shim_w->bar_count = 0;
    __sync_synchronize();
    shim_w->bar_sense = shim_local_sense;
  } else {
    while (shim_w->bar_sense != shim_local_sense) shim_relax(&spins);
  }
  __sync_synchronize();
  return MPI_SUCCESS;
}

static inline int MPI_Init(int* argc, char*** argv)
{
  (void)argc; (void)argv;
  const char* env = getenv("MPI_SHIM_RANKS");
  shim_size = env ? atoi(env) : 1;
  if (shim_size < 1 || shim_size > SHIM_MAX_RANKS) shim_fail("MPI_SHIM_RANKS must be 1..64");
  shim_w = (shim_world*)mmap(NULL, sizeof(shim_world), PROT_READ | PROT_WRITE,
                             MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (shim_w == MAP_FAILED) shim_fail("mmap of the shared region failed");
  fflush(stdout); fflush(stderr);
  for (int r = 1; r < shim_size; r++) {
    pid_t pid = fork();
    if (pid < 0) shim_fail("fork failed");
    if (pid == 0) { shim_rank = r; break; }
    shim_kids[r] = pid;
  }
  return MPI_SUCCESS;
}

static inline int MPI_Finalize(void)
{
  fflush(stdout); fflush(stderr);
  if (shim_rank != 0) _exit(EXIT_SUCCESS);
  int bad = 0;
  for (int r = 1; r < shim_size; r++) {
    int status = 0;
    waitpid(shim_kids[r], &status, 0);
    if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) bad = 1;
  }
  if (bad) shim_fail("a rank exited abnormally");
  return MPI_SUCCESS;
}

static inline int MPI_Comm_size(MPI_Comm c, int* size) { (void)c; *size = shim_size; return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int* rank) { (void)c; *rank = shim_rank; return MPI_SUCCESS; }

static inline int MPI_Type_contiguous(int count, MPI_Datatype old, MPI_Datatype* out)
{
  *out = (size_t)count * old;
  return MPI_SUCCESS;
}
static inline int MPI_Type_commit(MPI_Datatype* t) { (void)t; return MPI_SUCCESS; }
static inline int MPI_Type_free(MPI_Datatype* t) { *t = 0; return MPI_SUCCESS; }

static inline int MPI_Recv_init(void* buf, int count, MPI_Datatype t, int source, int tag,
                                MPI_Comm c, MPI_Request* req)
{
  (void)tag; (void)c;
  req->buf = buf; req->bytes = (size_t)count * t; req->peer = source; req->is_send = 0;
  if (req->bytes > SHIM_SLOT_BYTES) shim_fail("message larger than a mailbox slot");
  return MPI_SUCCESS;
}

static inline int MPI_Send_init(const void* buf, int count, MPI_Datatype t, int dest, int tag,
                                MPI_Comm c, MPI_Request* req)
{
  (void)tag; (void)c;
  req->buf = (void*)buf; req->bytes = (size_t)count * t; req->peer = dest; req->is_send = 1;
  if (req->bytes > SHIM_SLOT_BYTES) shim_fail("message larger than a mailbox slot");
  return MPI_SUCCESS;
}

/* sends are buffered eagerly, in the order they appear in the request array */
static inline int MPI_Startall(int n, MPI_Request* reqs)
{
  for (int i = 0; i < n; i++) {
    if (!reqs[i].is_send) continue;
    shim_mailbox* mb = &shim_w->box[shim_rank][shim_direction(shim_rank, reqs[i].peer)];
    unsigned spins = 0;
    while (mb->head - mb->tail >= SHIM_SLOTS) shim_relax(&spins);
    memcpy(mb->slot[mb->head % SHIM_SLOTS], reqs[i].buf, reqs[i].bytes);
    __sync_synchronize();
    mb->head = mb->head + 1;
  }
  return MPI_SUCCESS;
}

/* receives are matched in the order they appear in the request array */
static inline int MPI_Waitall(int n, MPI_Request* reqs, MPI_Status* st)
{
  (void)st;
  for (int i = 0; i < n; i++) {
    if (reqs[i].is_send) continue;
    shim_mailbox* mb = &shim_w->box[reqs[i].peer][shim_direction(reqs[i].peer, shim_rank)];
    unsigned spins = 0;
    while (mb->head == mb->tail) shim_relax(&spins);
    __sync_synchronize();
    memcpy(reqs[i].buf, mb->slot[mb->tail % SHIM_SLOTS], reqs[i].bytes);
    __sync_synchronize();
    mb->tail = mb->tail + 1;
  }
  return MPI_SUCCESS;
}

static inline int MPI_Request_free(MPI_Request* r) { r->buf = NULL; return MPI_SUCCESS; }

/* float SUM only; ranks are added in rank order, chunk by chunk through the staging buffer */
static inline int MPI_Reduce(const void* send, void* recv, int count, MPI_Datatype t, MPI_Op op,
                             int root, MPI_Comm c)
{
  (void)c;
  if (t != MPI_FLOAT || op != MPI_SUM) shim_fail("MPI_Reduce supports float SUM only");
  if (shim_size == 1) { memcpy(recv, send, (size_t)count * sizeof(float)); return MPI_SUCCESS; }
  const size_t chunk = SHIM_STAGE_BYTES / sizeof(float) / (size_t)shim_size;
  float* stage = (float*)shim_w->stage;
  for (size_t done = 0; done < (size_t)count; done += chunk) {
    size_t len = (size_t)count - done < chunk ? (size_t)count - done : chunk;
    memcpy(stage + (size_t)shim_rank * chunk, (const float*)send + done, len * sizeof(float));
    MPI_Barrier(MPI_COMM_WORLD);
    if (shim_rank == root) {
      float* out = (float*)recv + done;
      for (size_t i = 0; i < len; i++) {
        float acc = stage[i];
        for (int r = 1; r < shim_size; r++) acc += stage[(size_t)r * chunk + i];
        out[i] = acc;
      }
    }
    MPI_Barrier(MPI_COMM_WORLD);
  }
  return MPI_SUCCESS;
}

static inline int MPI_Bcast(void* buf, int count, MPI_Datatype t, int root, MPI_Comm c)
{
  (void)c;
  size_t bytes = (size_t)count * t;
  if (shim_size == 1) return MPI_SUCCESS;
  if (bytes > SHIM_STAGE_BYTES) shim_fail("MPI_Bcast payload larger than the staging buffer");
  if (shim_rank == root) memcpy(shim_w->stage, buf, bytes);
  MPI_Barrier(MPI_COMM_WORLD);
  if (shim_rank != root) memcpy(buf, shim_w->stage, bytes);
  MPI_Barrier(MPI_COMM_WORLD);
  return MPI_SUCCESS;
}

static inline int MPI_Scatterv(const void* send, const int* counts, const int* displs,
                               MPI_Datatype st, void* recv, int rcount, MPI_Datatype rt,
                               int root, MPI_Comm c)
{
  (void)c; (void)rcount; (void)rt;
  for (int r = 0; r < shim_size; r++) {
    size_t bytes = (size_t)counts[r] * st;
    const char* src = (shim_rank == root) ? (const char*)send + (size_t)displs[r] * st : NULL;
    if (r == root) {
      if (shim_rank == root) memcpy(recv, src, bytes);
      continue;
    }
    for (size_t done = 0; done < bytes; done += SHIM_STAGE_BYTES) {
      size_t len = bytes - done < SHIM_STAGE_BYTES ? bytes - done : SHIM_STAGE_BYTES;
      if (shim_rank == root) memcpy(shim_w->stage, src + done, len);
      MPI_Barrier(MPI_COMM_WORLD);
      if (shim_rank == r) memcpy((char*)recv + done, shim_w->stage, len);
      MPI_Barrier(MPI_COMM_WORLD);
    }
  }
  return MPI_SUCCESS;
}

static inline int MPI_Pcontrol(const int level, ...) { (void)level; return MPI_SUCCESS; }

#endif /* LBM_ORACLE_MPI_SHIM_H */
