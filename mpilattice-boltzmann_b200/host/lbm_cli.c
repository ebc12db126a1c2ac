/*
 * d2q9-bgk -- drop-in command-line front end of the B200 D2Q9-BGK solver.
 *
 *   d2q9-bgk[.exe] <paramfile> <obstaclefile>
 *
 * Same contract as the reference program (ag14774/MPILattice-Boltzmann, d2q9-bgk.c): the
 * 7-line .params file and the "x y 1" obstacle file go in, ./av_vels.dat and
 * ./final_state.dat come out, and five lines are printed on stdout ("==done==", Reynolds
 * number, elapsed wall / user / system time).  The reference's check/check.py validates the
 * two output files unchanged.
 *
 * Plain C on the host, like the reference; all device work goes through the C-ABI of
 * liblbm_b200.so (include/lbm_b200.h).  There is no CPU fallback.
 *
 * Environment (all optional; the two-argument command line is untouched):
 *   LBM_GPUS          number of row slabs = GPUs to use (default 1; "all" = every visible GPU)
 *   LBM_DEVICES       comma list of CUDA device ids, one per slab (default 0,1,2,...)
 *   LBM_FINAL_STATE   0 = do not write final_state.dat (for decks whose text would be GBs;
 *                     the reference does the same under -DPROFILE, d2q9-bgk.c:419-421)
 *   LBM_INPLACE       1 = one population buffer per slab, streamed in place (lbm_b200_create_inplace):
 *                     half the device memory, same results; needs nx % 4 == 0
 *   LBM_VERBOSE       1 = also print MLUPS and effective GB/s on stderr
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/resource.h>
#include <sys/time.h>

#include "../../include/lbm_b200.h"
#include "fast_format.h"

#define FINALSTATEFILE "final_state.dat"
#define AVVELSFILE     "av_vels.dat"

typedef struct {
  int nx, ny, max_iters, reynolds_dim;
  float density, accel, omega;
} deck_params;

/* same wording and exit status as the reference's die()/usage() (d2q9-bgk.c:1145-1157) */
static void die(const char* message, const int line, const char* file)
{
  fprintf(stderr, "Error at line %d of file %s:\n", line, file);
  fprintf(stderr, "%s\n", message);
  fflush(stderr);
  exit(EXIT_FAILURE);
}

static void usage(const char* exe)
{
  fprintf(stderr, "Usage: %s <paramfile> <obstaclefile>\n", exe);
  exit(EXIT_FAILURE);
}

#define LBM_TRY(call) do { if ((call) != LBM_B200_OK) die(lbm_b200_last_error(), __LINE__, __FILE__); } while (0)

/* the .params format of d2q9-bgk.c:781-800: four ints then three floats, one per line */
static void read_params(const char* path, deck_params* p)
{
  char message[1024];
  FILE* fp = fopen(path, "r");
  if (fp == NULL) {
    snprintf(message, sizeof message, "could not open input parameter file: %s", path);
    die(message, __LINE__, __FILE__);
  }
  int* ints[4] = {&p->nx, &p->ny, &p->max_iters, &p->reynolds_dim};
  const char* int_names[4] = {"nx", "ny", "maxIters", "reynolds_dim"};
  float* reals[3] = {&p->density, &p->accel, &p->omega};
  const char* real_names[3] = {"density", "accel", "omega"};
  for (int i = 0; i < 4; i++) {
    if (fscanf(fp, "%d\n", ints[i]) != 1) {
      snprintf(message, sizeof message, "could not read param file: %s", int_names[i]);
      die(message, __LINE__, __FILE__);
    }
  }
  for (int i = 0; i < 3; i++) {
    if (fscanf(fp, "%f\n", reals[i]) != 1) {
      snprintf(message, sizeof message, "could not read param file: %s", real_names[i]);
      die(message, __LINE__, __FILE__);
    }
  }
  fclose(fp);
}

/* the obstacle format and checks of d2q9-bgk.c:924-953 */
static int* read_obstacles(const char* path, const deck_params* p)
{
  char message[1024];
  int* obstacles = (int*)calloc((size_t)p->nx * p->ny, sizeof(int));
  if (obstacles == NULL) die("cannot allocate column memory for obstacles", __LINE__, __FILE__);
  FILE* fp = fopen(path, "r");
  if (fp == NULL) {
    snprintf(message, sizeof message, "could not open input obstacles file: %s", path);
    die(message, __LINE__, __FILE__);
  }
  int xx, yy, blocked, got;
  while ((got = fscanf(fp, "%d %d %d\n", &xx, &yy, &blocked)) != EOF) {
    if (got != 3) die("expected 3 values per line in obstacle file", __LINE__, __FILE__);
    if (xx < 0 || xx > p->nx - 1) die("obstacle x-coord out of range", __LINE__, __FILE__);
    if (yy < 0 || yy > p->ny - 1) die("obstacle y-coord out of range", __LINE__, __FILE__);
    if (blocked != 1) die("obstacle blocked value should be 1", __LINE__, __FILE__);
    obstacles[(size_t)yy * p->nx + xx] = blocked;
  }
  fclose(fp);
  return obstacles;
}

/* final_state.dat and av_vels.dat in the reference's formats (d2q9-bgk.c:1115, 1136) */
static void write_values(const deck_params* p, const int* obstacles, const float* u_x, const float* u_y,
                         const float* u, const float* pressure, const float* av_vels, int write_final_state)
{
  FILE* fp;
  if (write_final_state) {
    fp = fopen(FINALSTATEFILE, "w");
    if (fp == NULL) die("could not open file output file", __LINE__, __FILE__);
    const size_t cap = 1u << 20;
    char* buf = (char*)malloc(cap + 256);
    if (buf == NULL) die("cannot allocate the output buffer", __LINE__, __FILE__);
    char* w = buf;
    for (int y = 0; y < p->ny; y++) {
      for (int x = 0; x < p->nx; x++) {
        const size_t c = (size_t)y * p->nx + x;
        w = fmt_uint(w, (unsigned)x); *w++ = ' ';
        w = fmt_uint(w, (unsigned)y); *w++ = ' ';
        w = fmt_e12(w, u_x[c]); *w++ = ' ';
        w = fmt_e12(w, u_y[c]); *w++ = ' ';
        w = fmt_e12(w, u[c]); *w++ = ' ';
        w = fmt_e12(w, pressure[c]); *w++ = ' ';
        w = fmt_uint(w, (unsigned)obstacles[c]); *w++ = '\n';
        if ((size_t)(w - buf) >= cap) { fwrite(buf, 1, (size_t)(w - buf), fp); w = buf; }
      }
    }
    fwrite(buf, 1, (size_t)(w - buf), fp);
    free(buf);
    fclose(fp);
  }
  fp = fopen(AVVELSFILE, "w");
  if (fp == NULL) die("could not open file output file", __LINE__, __FILE__);
  for (int t = 0; t < p->max_iters; t++) {
    char line[64];
    char* w = fmt_uint(line, (unsigned)t);
    *w++ = ':'; *w++ = '\t';
    w = fmt_e12(w, av_vels[t]);
    *w++ = '\n';
    fwrite(line, 1, (size_t)(w - line), fp);
  }
  fclose(fp);
}

/* av_velocity + calc_reynolds of the reference (d2q9-bgk.c:707-757, 1002-1008) on the final
 * macroscopic fields: sequential float accumulator fed through a double sqrt, then
 * av * reynolds_dim / viscosity. */
static float calc_reynolds(const deck_params* p, const int* obstacles, const float* u_x, const float* u_y,
                           float free_cells_inv)
{
  float tot_u = 0.0f;
  const size_t n = (size_t)p->nx * p->ny;
  for (size_t c = 0; c < n; c++)
    if (!obstacles[c]) tot_u += sqrt((u_x[c] * u_x[c]) + (u_y[c] * u_y[c]));
  const float av = tot_u * free_cells_inv;
  const float viscosity = 1.0f / 6.0f * (2.0f / p->omega - 1.0f);
  return av * p->reynolds_dim / viscosity;
}

int main(int argc, char* argv[])
{
  if (argc != 3) usage(argv[0]);
  const char* paramfile = argv[1];
  const char* obstaclefile = argv[2];

  deck_params p;
  read_params(paramfile, &p);
  int* obstacles = read_obstacles(obstaclefile, &p);
  const float free_cells_inv = lbm_b200_free_cells_inv(obstacles, (long)p.nx * p.ny);

  int n_slabs = 1;
  const char* env = getenv("LBM_GPUS");
  if (env && *env) n_slabs = strcmp(env, "all") == 0 ? lbm_b200_device_count() : atoi(env);
  if (n_slabs < 1) die("LBM_GPUS must be a positive integer or 'all'", __LINE__, __FILE__);
  int* devices = NULL;
  env = getenv("LBM_DEVICES");
  if (env && *env) {
    devices = (int*)malloc(sizeof(int) * (size_t)n_slabs);
    char* copy = strdup(env);
    int i = 0;
    for (char* tok = strtok(copy, ","); tok && i < n_slabs; tok = strtok(NULL, ",")) devices[i++] = atoi(tok);
    if (i != n_slabs) die("LBM_DEVICES must list one device per slab (LBM_GPUS)", __LINE__, __FILE__);
    free(copy);
  }

  /* device allocation + upload sit with initialise(), outside the timed region, exactly
   * where the reference mallocs and scatters (d2q9-bgk.c:208-209 precede tic at 278) */
  lbm_b200* sim = NULL;
  env = getenv("LBM_INPLACE");
  if (env && atoi(env) == 1)
    LBM_TRY(lbm_b200_create_inplace(&sim, p.nx, p.ny, p.density, p.accel, p.omega, obstacles, n_slabs, devices));
  else
    LBM_TRY(lbm_b200_create(&sim, p.nx, p.ny, p.density, p.accel, p.omega, obstacles, n_slabs, devices));
  float* av_vels = (float*)malloc(sizeof(float) * (size_t)(p.max_iters > 0 ? p.max_iters : 1));
  if (av_vels == NULL) die("cannot allocate memory for av_vels", __LINE__, __FILE__);

  struct timeval timstr;
  struct rusage ru;
  gettimeofday(&timstr, NULL);
  const double tic = timstr.tv_sec + (timstr.tv_usec / 1000000.0);

  LBM_TRY(lbm_b200_run(sim, p.max_iters, av_vels));   /* every step, the final sync and the av_vels gather */

  gettimeofday(&timstr, NULL);
  const double toc = timstr.tv_sec + (timstr.tv_usec / 1000000.0);
  getrusage(RUSAGE_SELF, &ru);
  const double usrtim = ru.ru_utime.tv_sec + (ru.ru_utime.tv_usec / 1000000.0);
  const double systim = ru.ru_stime.tv_sec + (ru.ru_stime.tv_usec / 1000000.0);

  const size_t n = (size_t)p.nx * p.ny;
  float* fields = (float*)malloc(sizeof(float) * 4 * n);
  if (fields == NULL) die("cannot allocate memory for the final state", __LINE__, __FILE__);
  float *u_x = fields, *u_y = fields + n, *u = fields + 2 * n, *pressure = fields + 3 * n;
  LBM_TRY(lbm_b200_get_final_state(sim, u_x, u_y, u, pressure));

  printf("==done==\n");
  printf("Reynolds number:\t\t%.12E\n", calc_reynolds(&p, obstacles, u_x, u_y, free_cells_inv));
  printf("Elapsed time:\t\t\t%.6lf (s)\n", toc - tic);
  printf("Elapsed user CPU time:\t\t%.6lf (s)\n", usrtim);
  printf("Elapsed system CPU time:\t%.6lf (s)\n", systim);
  fflush(stdout);

  env = getenv("LBM_VERBOSE");
  if (env && atoi(env)) {
    float ms = 0.f;
    lbm_b200_elapsed_ms(sim, &ms);
    const double lups = (double)n * p.max_iters;
    fprintf(stderr, "slabs: %d  wall: %.1f MLUPS  device: %.1f MLUPS, %.1f GB/s at 72 B/cell/step\n",
            n_slabs, lups / (toc - tic) / 1e6, lups / (ms * 1e-3) / 1e6, lups * 72.0 / (ms * 1e-3) / 1e9);
  }

  env = getenv("LBM_FINAL_STATE");
  write_values(&p, obstacles, u_x, u_y, u, pressure, av_vels, !(env && strcmp(env, "0") == 0));

  lbm_b200_destroy(sim);
  free(fields);
  free(av_vels);
  free(obstacles);
  free(devices);
  return EXIT_SUCCESS;
}
