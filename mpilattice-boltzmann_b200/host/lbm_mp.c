/*
 * d2q9-bgk-mp -- the reference's `mpirun -np N ./d2q9-bgk <paramfile> <obstaclefile>` on N B200s.
 *
 *   d2q9-bgk-mp [-np N] <paramfile> <obstaclefile>        (N defaults to $LBM_RANKS, else to every visible GPU)
 *
 * One PROCESS per GPU, like the reference's one MPI rank per core (d2q9-bgk.c:185-187; mpi_submit:63): the launcher
 * reads the deck once (the reference's rank 0, 208-209), forks N ranks before anything touches CUDA, and every rank
 *   1. takes its row slab            lbm_b200_decompose            (834-862)
 *   2. creates it on its own GPU     lbm_b200_create_slab          (MPI_Scatterv of the obstacle rows, 968-970)
 *   3. publishes / maps the halo     lbm_b200_ipc_export/_connect  (MPI_Send_init / MPI_Recv_init, 295-313)
 *   4. runs the timestep loop        lbm_b200_run                  (315-394; halo rows travel as NVLink stores)
 *   5. hands its partial av_vels and its rows of the final fields to rank 0 through shared memory
 *                                                                   (MPI_Reduce 396; the rank-ordered write 1034-1143)
 * Rank 0 prints the reference's five lines and writes av_vels.dat / final_state.dat.  The only transport between
 * the ranks on the host is one anonymous shared mapping (the IPC blobs, two barriers, the results): no MPI runtime
 * is needed, and none exists in this image.
 *
 * Environment: LBM_DEVICES (comma list, one CUDA device per rank; default rank r -> device r), LBM_FINAL_STATE=0,
 * LBM_INPLACE=1, LBM_VERBOSE=1 as for d2q9-bgk.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <sched.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/resource.h>
#include <sys/time.h>
#include <sys/wait.h>
#include <unistd.h>

#include "deck_io.h"

#define MAX_RANKS 64
#define BLOB_MAX 512

typedef struct {
  volatile int failed;               /* a rank died: everybody leaves */
  volatile int arrived;              /* barrier: ranks that reached the current generation */
  volatile int generation;
  char blobs[MAX_RANKS][BLOB_MAX];   /* lbm_b200_ipc_export of every rank */
  double elapsed[MAX_RANKS];         /* tic..toc of every rank */
  float device_ms[MAX_RANKS];
} shared_head;

static shared_head* g_shm;
static int g_rank = -1, g_ranks = 0;

static void rank_die(const char* message, const int line, const char* file)
{
  if (g_shm) g_shm->failed = 1;
  fprintf(stderr, "rank %d: ", g_rank);
  die(message, line, file);
}
#define RANK_TRY(call) do { if ((call) != LBM_B200_OK) rank_die(lbm_b200_last_error(), __LINE__, __FILE__); } while (0)

/* sense-reversing barrier over the shared mapping; leaves (exit) as soon as any rank has failed */
static void barrier(void)
{
  const int gen = g_shm->generation;
  if (__atomic_add_fetch(&g_shm->arrived, 1, __ATOMIC_ACQ_REL) == g_ranks) {
    g_shm->arrived = 0;
    __atomic_store_n(&g_shm->generation, gen + 1, __ATOMIC_RELEASE);
    return;
  }
  while (__atomic_load_n(&g_shm->generation, __ATOMIC_ACQUIRE) == gen) {
    if (g_shm->failed) _exit(EXIT_FAILURE);
    sched_yield();
  }
}

static double now(void)
{
  struct timeval t;
  gettimeofday(&t, NULL);
  return t.tv_sec + (t.tv_usec / 1000000.0);
}

int main(int argc, char* argv[])
{
  int ranks = 0, arg = 1;
  if (argc >= 3 && strcmp(argv[1], "-np") == 0) { ranks = atoi(argv[2]); arg = 3; }
  if (argc - arg != 2) {
    fprintf(stderr, "Usage: %s [-np N] <paramfile> <obstaclefile>\n", argv[0]);
    exit(EXIT_FAILURE);
  }
  if (ranks == 0) {
    const char* env = getenv("LBM_RANKS");
    if (env && *env) ranks = atoi(env);
  }
  int devices[MAX_RANKS], n_listed = 0;
  {
    const char* env = getenv("LBM_DEVICES");
    if (env && *env) {
      char* copy = strdup(env);
      for (char* tok = strtok(copy, ","); tok && n_listed < MAX_RANKS; tok = strtok(NULL, ",")) devices[n_listed++] = atoi(tok);
      free(copy);
    }
  }
  if (ranks == 0) {
    /* every visible GPU -- asked in a child, so that the launcher itself never initialises CUDA before it forks */
    int fd[2];
    if (pipe(fd) != 0) die("pipe failed", __LINE__, __FILE__);
    const pid_t pid = fork();
    if (pid == 0) { int n = lbm_b200_device_count(); if (write(fd[1], &n, sizeof n) != sizeof n) _exit(1); _exit(0); }
    if (read(fd[0], &ranks, sizeof ranks) != sizeof ranks) ranks = 0;
    waitpid(pid, NULL, 0);
    close(fd[0]); close(fd[1]);
    if (ranks < 1) die("no CUDA device available (there is no CPU fallback)", __LINE__, __FILE__);
  }
  if (ranks < 1 || ranks > MAX_RANKS) die("the number of ranks must be 1..64", __LINE__, __FILE__);
  if (n_listed && n_listed != ranks) die("LBM_DEVICES must list one device per rank", __LINE__, __FILE__);
  if (lbm_b200_ipc_blob_bytes() > BLOB_MAX) die("IPC blob larger than the launcher's slot", __LINE__, __FILE__);

  deck_params p;
  read_params(argv[arg], &p);
  int* obstacles = read_obstacles(argv[arg + 1], &p);
  const float free_cells_inv = lbm_b200_free_cells_inv(obstacles, (long)p.nx * p.ny);
  int rows[MAX_RANKS], first[MAX_RANKS];
  if (lbm_b200_decompose(p.ny, ranks, rows, first) != LBM_B200_OK) die(lbm_b200_last_error(), __LINE__, __FILE__);

  /* shared mapping: head, per-rank av_vels partials, the four final fields */
  const size_t n = (size_t)p.nx * p.ny;
  const size_t iters = (size_t)(p.max_iters > 0 ? p.max_iters : 1);
  const size_t bytes = sizeof(shared_head) + sizeof(float) * (iters * (size_t)ranks + 4 * n);
  g_shm = (shared_head*)mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (g_shm == MAP_FAILED) die("cannot map the shared result area", __LINE__, __FILE__);
  memset(g_shm, 0, sizeof(shared_head));
  float* av_parts = (float*)(g_shm + 1);
  float* fields = av_parts + iters * (size_t)ranks;
  g_ranks = ranks;

  pid_t pids[MAX_RANKS];
  for (int r = 0; r < ranks; r++) {
    fflush(NULL);
    pids[r] = fork();
    if (pids[r] < 0) { g_shm->failed = 1; die("fork failed", __LINE__, __FILE__); }
    if (pids[r] == 0) { g_rank = r; break; }
  }

  if (g_rank < 0) {
    /* the launcher: wait for the ranks; the first failure releases everybody who sits in a barrier */
    int status = EXIT_SUCCESS;
    for (int left = ranks; left > 0; left--) {
      int st = 0;
      if (wait(&st) < 0) break;
      if (!WIFEXITED(st) || WEXITSTATUS(st) != EXIT_SUCCESS) { g_shm->failed = 1; status = EXIT_FAILURE; }
    }
    return status;
  }

  /* ---------------------------------------------------------------------------------------------- a rank */
  const int r = g_rank;
  const int device = n_listed ? devices[r] : r;
  const char* env = getenv("LBM_INPLACE");
  const int inplace = env && atoi(env) == 1;
  lbm_b200* sim = NULL;
  const int* my_rows = obstacles + (size_t)first[r] * p.nx;            /* this rank's share of MPI_Scatterv (968-970) */
  RANK_TRY(lbm_b200_create_slab_ex(&sim, p.nx, p.ny, first[r], rows[r], r, ranks, p.density, p.accel, p.omega,
                                   free_cells_inv, my_rows, LBM_B200_OBST_INT32, device, inplace));
  if (ranks > 1) {
    RANK_TRY(lbm_b200_ipc_export(sim, g_shm->blobs[r]));
    barrier();
    RANK_TRY(lbm_b200_ipc_connect(sim, g_shm->blobs[(r + ranks - 1) % ranks], g_shm->blobs[(r + 1) % ranks]));
  }
  barrier();

  struct rusage ru;
  const double tic = now();
  RANK_TRY(lbm_b200_run(sim, p.max_iters, av_parts + iters * (size_t)r));
  barrier();                                                           /* every rank's partials are in: MPI_Reduce (396) */
  const double toc = now();
  getrusage(RUSAGE_SELF, &ru);
  g_shm->elapsed[r] = toc - tic;
  lbm_b200_elapsed_ms(sim, (float*)&g_shm->device_ms[r]);

  const size_t off = (size_t)first[r] * p.nx;
  RANK_TRY(lbm_b200_get_final_state(sim, fields + off, fields + n + off, fields + 2 * n + off, fields + 3 * n + off));
  barrier();

  if (r == 0) {
    float* av_vels = av_parts;                                         /* rank-order float sum, as MPI_Reduce */
    for (int q = 1; q < ranks; q++)
      for (int t = 0; t < p.max_iters; t++) av_vels[t] += av_parts[iters * (size_t)q + t];
    const float *u_x = fields, *u_y = fields + n, *u = fields + 2 * n, *pressure = fields + 3 * n;
    printf("==done==\n");
    printf("Reynolds number:\t\t%.12E\n", calc_reynolds(&p, obstacles, u_x, u_y, free_cells_inv));
    printf("Elapsed time:\t\t\t%.6lf (s)\n", toc - tic);
    printf("Elapsed user CPU time:\t\t%.6lf (s)\n", ru.ru_utime.tv_sec + (ru.ru_utime.tv_usec / 1000000.0));
    printf("Elapsed system CPU time:\t%.6lf (s)\n", ru.ru_stime.tv_sec + (ru.ru_stime.tv_usec / 1000000.0));
    fflush(stdout);
    env = getenv("LBM_VERBOSE");
    if (env && atoi(env)) {
      float ms = 0.f;
      for (int q = 0; q < ranks; q++) ms = g_shm->device_ms[q] > ms ? g_shm->device_ms[q] : ms;
      const double lups = (double)n * p.max_iters;
      fprintf(stderr, "ranks: %d  wall: %.1f MLUPS  device: %.1f MLUPS, %.1f GB/s at 72 B/cell/step\n", ranks,
              lups / (toc - tic) / 1e6, lups / (ms * 1e-3) / 1e6, lups * 72.0 / (ms * 1e-3) / 1e9);
    }
    env = getenv("LBM_FINAL_STATE");
    write_values(&p, obstacles, u_x, u_y, u, pressure, av_vels, !(env && strcmp(env, "0") == 0));
  }
  barrier();                                                           /* nobody unmaps a neighbour that is still read */
  lbm_b200_destroy(sim);
  free(obstacles);
  return EXIT_SUCCESS;
}
