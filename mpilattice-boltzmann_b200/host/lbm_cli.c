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

#include "deck_io.h"

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
