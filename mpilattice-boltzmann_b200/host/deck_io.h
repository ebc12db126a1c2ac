/*
 * deck_io.h -- what both host programs (d2q9-bgk and d2q9-bgk-mp) share: the reference's input formats, output
 * formats, error convention and Reynolds number, in plain C like the reference.
 */
#ifndef LBM_DECK_IO_H
#define LBM_DECK_IO_H

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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

/* (the launcher d2q9-bgk-mp has its own usage line with -np) */
static void __attribute__((unused)) usage(const char* exe)
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

#endif /* LBM_DECK_IO_H */
