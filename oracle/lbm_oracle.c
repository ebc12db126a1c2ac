/*
 * lbm_oracle.c -- CPU restatement of the reference D2Q9-BGK hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline legs may load this library, and only as the checker.  The product
 * (include/lbm_b200.h, mpilattice-boltzmann_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  Built with -ffp-contract=off this restatement is bit-identical
 * (every population of every cell, and every av_vels entry) to the unmodified reference
 * source compiled with the same strict-IEEE flag (oracle/_ref/d2q9-bgk.strict, one rank),
 * and both pass the reference's own check/check.py against the goldens in check/ -- see
 * tests/test_oracle_pinned.py and oracle/pin_oracle.py.
 *
 * Every function cites the lines of /root/reference/d2q9-bgk.c it restates.  The order of
 * every floating-point operation follows the reference exactly, because the parity bar for
 * the populations is bit-exactness; see DESIGN.md "Arithmetic contract".
 *
 * Layout is the reference's: array-of-structs cells, 9 floats per cell, row-major, with one
 * halo row below (row 0) and above (row rows+1) the slab, obstacles as one int per cell.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NSPEEDS 9

/* d2q9-bgk.c:834-862 -- rows per rank and first global row of each rank.  The last rank is
 * given at least 3 rows so the accelerated row (global ny-2) is never an exchanged edge row. */
void lbm_oracle_decompose(int ny, int size, int* ny_local, int* displs)
{
  int orig = ny / size;
  int left = ny % size;
  int one_for_last = 0, one_less_for_second_to_last = 0;
  if (orig < 3 && left) { left--; one_for_last = 1; }
  else if (orig < 3 && !left) { one_for_last = 1; one_less_for_second_to_last = 1; }
  for (int p = 0; p < size; p++) {
    if (p < size - 2) ny_local[p] = orig;
    else if (p == size - 2) ny_local[p] = orig - one_less_for_second_to_last;
    else ny_local[p] = orig + one_for_last;
    if (p < left) ny_local[p]++;
    displs[p] = (p == 0) ? 0 : displs[p - 1] + ny_local[p - 1];
  }
}

/* d2q9-bgk.c:880-902 -- uniform initial populations. */
void lbm_oracle_init(float* cells, long ncells, float density)
{
  float w0 = density * 4.0f / 9.0f;
  float w1 = density / 9.0f;
  float w2 = density / 36.0f;
  for (long c = 0; c < ncells; c++) {
    float* s = cells + c * NSPEEDS;
    s[0] = w0;
    s[1] = w1; s[2] = w1; s[3] = w1; s[4] = w1;
    s[5] = w2; s[6] = w2; s[7] = w2; s[8] = w2;
  }
}

/* d2q9-bgk.c:442-478 -- body force on ONE row (the caller passes global row ny-2). */
void lbm_oracle_accelerate_row(float* row_cells, const int* row_obstacles, int nx,
                               float density, float accel)
{
  float w1 = density * accel * 0.111111111111111111111111f;
  float w2 = density * accel * 0.0277777777777777777777778f;
  for (int x = 0; x < nx; x++) {
    float* s = row_cells + (long)x * NSPEEDS;
    if (!row_obstacles[x] && s[3] - w1 > 0.0f && s[6] - w2 > 0.0f && s[7] - w2 > 0.0f) {
      s[1] += w1; s[5] += w2; s[8] += w2;
      s[3] -= w1; s[6] -= w2; s[7] -= w2;
    }
  }
}

/* d2q9-bgk.c:520-698 for one cell: pull, moments, equilibrium, relax or bounce back.
 * Returns the cell's contribution sqrt(m^2)/rho as the double the reference forms at 667/684
 * (negative = blocked cell, contributes nothing). */
static inline double oracle_cell(const float* cells, float* tmp_cells, int blocked, int nx,
                                 long row, long row_s, long row_n, int x, float omega)
{
  const float ic_sq = 3.0f;
  const float w0 = 4.0f / 9.0f, w1 = 1.0f / 9.0f, w2 = 1.0f / 36.0f;
  int x_e = x + 1; if (x_e >= nx) x_e -= nx;                       /* 527-528 */
  int x_w = (x == 0) ? (nx - 1) : (x - 1);                         /* 529 */
  float f[NSPEEDS];
  f[0] = cells[(row   * nx + x  ) * NSPEEDS + 0];                  /* 530-538 */
  f[1] = cells[(row   * nx + x_w) * NSPEEDS + 1];
  f[2] = cells[(row_s * nx + x  ) * NSPEEDS + 2];
  f[3] = cells[(row   * nx + x_e) * NSPEEDS + 3];
  f[4] = cells[(row_n * nx + x  ) * NSPEEDS + 4];
  f[5] = cells[(row_s * nx + x_w) * NSPEEDS + 5];
  f[6] = cells[(row_s * nx + x_e) * NSPEEDS + 6];
  f[7] = cells[(row_n * nx + x_e) * NSPEEDS + 7];
  f[8] = cells[(row_n * nx + x_w) * NSPEEDS + 8];
  float* out = tmp_cells + (row * nx + x) * NSPEEDS;

  if (blocked) {                                                   /* 687-695 */
    out[0] = f[0]; out[3] = f[1]; out[4] = f[2]; out[1] = f[3]; out[2] = f[4];
    out[7] = f[5]; out[8] = f[6]; out[5] = f[7]; out[6] = f[8];
    return -1.0;
  }

  float dens = f[0];                                               /* 546-554 */
  dens += f[1]; dens += f[2]; dens += f[3]; dens += f[4];
  dens += f[5]; dens += f[6]; dens += f[7]; dens += f[8];
  float densinv = 1.0f / dens;                                     /* 561 */

  float u_x = f[1] + f[5];                                         /* 570-574 (momentum) */
  u_x += f[8]; u_x -= f[3]; u_x -= f[6]; u_x -= f[7];
  float u_y = f[2] + f[5];                                         /* 576-580 */
  u_y += f[6]; u_y -= f[4]; u_y -= f[7]; u_y -= f[8];
  float u_sq = u_x * u_x + u_y * u_y;                              /* 589 */

  float uvec[NSPEEDS], t3[NSPEEDS], t3sq[NSPEEDS], d_equ[NSPEEDS];
  uvec[1] = u_x;        uvec[2] = u_y;                             /* 596-603 */
  uvec[3] = -u_x;       uvec[4] = -u_y;
  uvec[5] = u_x + u_y;  uvec[6] = -u_x + u_y;
  uvec[7] = -u_x - u_y; uvec[8] = u_x - u_y;
  for (int k = 1; k < NSPEEDS; k++) {
    t3[k] = uvec[k] * ic_sq;                                       /* 610-617 */
    t3sq[k] = t3[k] * uvec[k];                                     /* 624-631 */
  }
  d_equ[0] = w0 * (dens - 0.5f * densinv * ic_sq * u_sq);          /* 638 */
  for (int k = 1; k < NSPEEDS; k++) {                              /* 639-646 */
    float w = (k < 5) ? w1 : w2;
    d_equ[k] = w * (dens + t3[k] + 0.5f * densinv * ic_sq * (t3sq[k] - u_sq));
  }
  for (int k = 0; k < NSPEEDS; k++)                                /* 658-666 / 675-683 */
    out[k] = f[k] + omega * (d_equ[k] - f[k]);
  return sqrt(u_sq) * densinv;                                     /* 667 / 684 */
}

/* d2q9-bgk.c:493-704 -- timestep(start, end) on a slab with halo rows; returns tot_u.
 * `scratch` holds one double per cell of the processed rows.  Cells are independent, so rows
 * are updated in parallel; the Sigma|u| accumulation is then replayed strictly in the
 * reference's sequential order (row-major, float accumulator through a double add) so the
 * result does not depend on the thread count. */
static double oracle_exact_sum;   /* fp64 sum of the last lbm_oracle_timestep_rows call (test aid) */

float lbm_oracle_timestep_rows(const float* cells, float* tmp_cells, const int* obstacles,
                               int nx, int start, int end, float omega, double* scratch)
{
  #pragma omp parallel for schedule(static)
  for (int ii = start; ii < end; ii++) {
    for (int x = 0; x < nx; x++) {
      scratch[(long)(ii - start) * nx + x] =
          oracle_cell(cells, tmp_cells, obstacles[(long)ii * nx + x], nx,
                      ii, ii - 1, ii + 1, x, omega);               /* 511-512: y_s, y_n */
    }
  }
  float tot_u = 0.0f;                                              /* 502 */
  double exact = 0.0;
  long n = (long)(end - start) * nx;
  for (long c = 0; c < n; c++)
    if (scratch[c] >= 0.0) {
      tot_u += scratch[c];                                         /* 667: float += double */
      exact += scratch[c];
    }
  oracle_exact_sum = exact;
  return tot_u;
}

/* d2q9-bgk.c:315-378 on ONE rank (size 1): the halo exchange degenerates to the periodic
 * wrap (295-303: halo above the last row <- first row, halo below the first row <- last
 * row), then accelerate (345-348), interior rows (350), the two edge rows (365-366), the
 * per-step average (367) and the buffer swap (376-378).
 * cells: ny*nx*9 floats (no halo), updated in place.  av_vels: iters floats.
 * av_exact (optional): the same per-cell terms summed in fp64 -- NOT what the reference
 * computes; it separates "the GPU's terms are right" from "the reference's sequential fp32
 * accumulation has rounding error of its own" in the av_vels comparison. */
int lbm_oracle_run(int nx, int ny, int iters, float density, float accel, float omega,
                   float free_cells_inv, const int* obstacles, float* cells, float* av_vels,
                   double* av_exact)
{
  long row = (long)nx * NSPEEDS;
  float* a = (float*)malloc(sizeof(float) * row * (ny + 2));
  float* b = (float*)malloc(sizeof(float) * row * (ny + 2));
  int* obst = (int*)calloc((size_t)(ny + 2) * nx, sizeof(int));
  double* scratch = (double*)malloc(sizeof(double) * (size_t)ny * nx);
  if (!a || !b || !obst || !scratch) { free(a); free(b); free(obst); free(scratch); return 1; }
  memcpy(a + row, cells, sizeof(float) * row * ny);
  memset(b, 0, sizeof(float) * row * (ny + 2));
  memcpy(obst + nx, obstacles, sizeof(int) * (size_t)ny * nx);

  for (int tt = 0; tt < iters; tt++) {
    memcpy(a + row * (ny + 1), a + row, sizeof(float) * row);      /* 295,300,327 */
    memcpy(a, a + row * ny, sizeof(float) * row);                  /* 297,302,327 */
    lbm_oracle_accelerate_row(a + row * (ny - 1), obst + (long)nx * (ny - 1), nx,
                              density, accel);                     /* 449: ii = ny_local-1 */
    float local = lbm_oracle_timestep_rows(a, b, obst, nx, 2, ny, omega, scratch);      /* 350 */
    double exact = oracle_exact_sum;
    local += lbm_oracle_timestep_rows(a, b, obst, nx, 1, 2, omega, scratch);            /* 365 */
    exact += oracle_exact_sum;
    local += lbm_oracle_timestep_rows(a, b, obst, nx, ny, ny + 1, omega, scratch);      /* 366 */
    exact += oracle_exact_sum;
    av_vels[tt] = local * free_cells_inv;                          /* 367 */
    if (av_exact) av_exact[tt] = exact * (double)free_cells_inv;
    float* t = a; a = b; b = t;                                    /* 376-378 */
  }
  memcpy(cells, a + row, sizeof(float) * row * ny);
  free(a); free(b); free(obst); free(scratch);
  return 0;
}

/* d2q9-bgk.c:707-757 (size 1) -- Sigma|u| over free cells of the final state, times 1/free. */
float lbm_oracle_av_velocity(const float* cells, const int* obstacles, int nx, int ny,
                             float free_cells_inv)
{
  float tot_u = 0.0f;
  for (long c = 0; c < (long)nx * ny; c++) {
    if (obstacles[c]) continue;
    const float* s = cells + c * NSPEEDS;
    float local_density = 0.0f;
    for (int k = 0; k < NSPEEDS; k++) local_density += s[k];
    float u_x = (s[1] + s[5] + s[8] - (s[3] + s[6] + s[7])) / local_density;
    float u_y = (s[2] + s[5] + s[6] - (s[4] + s[7] + s[8])) / local_density;
    tot_u += sqrt((u_x * u_x) + (u_y * u_y));
  }
  return tot_u * free_cells_inv;
}

/* d2q9-bgk.c:1002-1008 */
float lbm_oracle_reynolds(const float* cells, const int* obstacles, int nx, int ny,
                          float free_cells_inv, float omega, int reynolds_dim)
{
  const float viscosity = 1.0f / 6.0f * (2.0f / omega - 1.0f);
  return lbm_oracle_av_velocity(cells, obstacles, nx, ny, free_cells_inv) * reynolds_dim / viscosity;
}

/* d2q9-bgk.c:1071-1112 -- the four macroscopic columns of final_state.dat. */
void lbm_oracle_final_state(const float* cells, const int* obstacles, int nx, int ny,
                            float density, float* u_x, float* u_y, float* u, float* pressure)
{
  const float c_sq = 1.0f / 3.0f;
  for (long c = 0; c < (long)nx * ny; c++) {
    if (obstacles[c]) {
      u_x[c] = u_y[c] = u[c] = 0.0f;
      pressure[c] = density * c_sq;
      continue;
    }
    const float* s = cells + c * NSPEEDS;
    float local_density = 0.0f;
    for (int k = 0; k < NSPEEDS; k++) local_density += s[k];
    u_x[c] = (s[1] + s[5] + s[8] - (s[3] + s[6] + s[7])) / local_density;
    u_y[c] = (s[2] + s[5] + s[6] - (s[4] + s[7] + s[8])) / local_density;
    u[c] = sqrt((u_x[c] * u_x[c]) + (u_y[c] * u_y[c]));
    pressure[c] = local_density * c_sq;
  }
}

/* d2q9-bgk.c:1011-1032 -- the (disabled) conservation invariant, kept as a test property. */
double lbm_oracle_total_density(const float* cells, long ncells)
{
  double total = 0.0;
  for (long c = 0; c < ncells * NSPEEDS; c++) total += cells[c];
  return total;
}
