/*
 * lbm_b200.h -- C-ABI of the B200-native D2Q9-BGK timestep path (liblbm_b200.so).
 *
 * This is the drop-in boundary for the hot path of ag14774/MPILattice-Boltzmann: everything
 * the reference's main() does between `tic` and `toc` (d2q9-bgk.c:278-398) plus the state it
 * hands to calc_reynolds/write_values afterwards.  The reference has no plugin or FFI layer;
 * each entry point below cites the internal reference interface it replaces.  Plain pointers
 * and sizes only -- no CUDA, torch or C++ types cross this boundary.
 *
 * Data contract (same as the reference):
 *   cells      array-of-structs, 9 floats per cell in the reference's speed order
 *              (d2q9-bgk.c:7-13, 95-98), row-major, row 0 first: cells[(y*nx + x)*9 + k]
 *   obstacles  one int per cell, 0 = fluid, non-zero = blocked (d2q9-bgk.c:875, 905-911)
 *   av_vels    one float per timestep (d2q9-bgk.c:367, 396)
 * Inside the library the populations live as 9 fp32 planes (structure of arrays) and the
 * obstacle map as 1 bit per cell; see DESIGN.md.
 *
 * Error convention: every function returning int returns LBM_B200_OK (0) or a non-zero code;
 * lbm_b200_last_error() then describes the failure (thread-local).  The reference's
 * convention is die(message, __LINE__, __FILE__) (d2q9-bgk.c:1145-1151): the host program
 * passes lbm_b200_last_error() to its own die().  There is no CPU fallback: without a usable
 * CUDA device every compute entry point fails with LBM_B200_ERR_CUDA.
 *
 * Threading: a handle is driven by one host thread at a time.
 */
#ifndef LBM_B200_H
#define LBM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define LBM_B200_ABI_VERSION 2

enum {
  LBM_B200_OK = 0,
  LBM_B200_ERR_ARG = 1,      /* bad argument (sizes, NULL, unsupported shape) */
  LBM_B200_ERR_CUDA = 2,     /* CUDA runtime / driver failure, or no device */
  LBM_B200_ERR_ALLOC = 3,    /* host or device allocation failed */
  LBM_B200_ERR_STATE = 4     /* call not valid in the handle's current state */
};

typedef struct lbm_b200 lbm_b200;   /* opaque: owns device memory, streams, graphs */

/* ---- host-only helpers (usable without a GPU) ------------------------------------- */

int lbm_b200_abi_version(void);

/* Row-slab decomposition of ny rows over n_slabs, replacing d2q9-bgk.c:834-862: ny/n rows
 * each, the remainder one per slab from slab 0, and the last slab never thinner than 3 rows
 * (the accelerated row ny-2 must not be an exchanged edge row).  rows[] and first_row[] have
 * n_slabs entries.  Fails if any slab would get fewer than 3 rows. */
int lbm_b200_decompose(int ny, int n_slabs, int* rows, int* first_row);

/* Work decomposition of the two-timesteps-per-pass kernel (no counterpart in the reference; exposed so that it can
 * be checked without a GPU): a slab of `rows` rows is cut into `*bands` bands of `*rows_per_band` rows -- the last
 * one takes the remainder, and the first and the last hold at least two rows -- times ceil(nx / 120) column strips.
 * band_rows = 0 picks the height automatically for a device with `sms` multiprocessors. */
int lbm_b200_plan_bands(int rows, int nx, int band_rows, int sms, int* bands, int* rows_per_band);

/* The same for `steps` = 2..4 timesteps per pass in the default launch shape (2: kernel 5; 3, 4: kernel 7 with one
 * staging row and one warp per CTA) and, with ring != 0, for a slab of a multi-GPU ring: its first and last band then
 * hold at least the four rows kernel 7 pushes per direction from one work item (two for kernel 5). */
int lbm_b200_plan_bands_ex(int rows, int nx, int band_rows, int sms, int steps, int ring, int* bands, int* rows_per_band);

/* 1.0f / (number of unblocked cells), replacing d2q9-bgk.c:805, 945-950. */
float lbm_b200_free_cells_inv(const int* obstacles, long n_cells);

const char* lbm_b200_last_error(void);

/* Number of visible CUDA devices (0 if there is none or the driver is missing). */
int lbm_b200_device_count(void);

/* Device self-test of the arithmetic restatements the step kernels rely on for bit-identity with the reference
 * (csrc/lbm_cell.cuh): counts[0] / counts[1] = floats in [2^-101, 2^126) whose restated reciprocal / square root
 * differs from CUDA's correctly rounded __frcp_rn / __fsqrt_rn (all 1.9e9 bit patterns are tried), counts[2] =
 * packed fp32x2 add / sub / mul results that differ from the scalar IEEE operations over 2^28 operand pairs drawn
 * from all bit patterns.  All three must be 0.  No counterpart in the reference. */
int lbm_b200_selftest(int device, unsigned long long counts[3]);

/* ---- whole-domain solver in one process -------------------------------------------- */

/* Replaces the device-independent part of initialise() (d2q9-bgk.c:865-902, 966-970):
 * allocates the two population buffers, fills them with the uniform initial state
 * w0 = 4/9 rho, w1 = rho/9, w2 = rho/36, uploads the bit-packed obstacle map and splits the
 * grid into n_slabs row slabs (lbm_b200_decompose).  devices[i] is the CUDA device of slab i
 * (NULL: slab i on device i).  Slabs on different devices exchange halo rows with direct
 * NVLink stores; several slabs may share one device (then they run in lock step on one
 * stream -- used to test the exchange on a single GPU).
 * Requires nx >= 4, ny >= 3. */
int lbm_b200_create(lbm_b200** handle, int nx, int ny, float density, float accel, float omega,
                    const int* obstacles, int n_slabs, const int* devices);

/* As lbm_b200_create, but with ONE population buffer per slab instead of the reference's cells/tmp_cells pair
 * (d2q9-bgk.c:865-872, swapped at 376-378): the timestep streams in place (the "AA" access pattern: steps
 * alternate between a neighbour-access and a cell-local flavour, csrc/lbm_kernels.cuh kernel 4).  Half the
 * device memory per cell (36 B + 1 bit), the same 72 B/cell/step of traffic and bit-identical results; state
 * in/out is staged through a bounded buffer ("staging_bytes").  Requires nx % 4 == 0, nx >= 8. */
int lbm_b200_create_inplace(lbm_b200** handle, int nx, int ny, float density, float accel, float omega,
                            const int* obstacles, int n_slabs, const int* devices);

/* Obstacle map formats of the _ex constructors.  The reference keeps one int per cell (d2q9-bgk.c:875), which at
 * 16384 x 16384 is a 1.07 GB upload for a 33 MB mask; callers that build their map themselves can hand it over in a
 * denser form:
 *   LBM_B200_OBST_INT32  int per cell, non-zero = blocked (the reference's layout; what the plain constructors take)
 *   LBM_B200_OBST_UINT8  one byte per cell, non-zero = blocked
 *   LBM_B200_OBST_BITS   one bit per cell: rows padded to whole 32-bit words ((nx + 31) / 32 per row), cell x of a
 *                        row is bit (x & 31) of word (x >> 5); padding bits are ignored */
enum { LBM_B200_OBST_INT32 = 0, LBM_B200_OBST_UINT8 = 1, LBM_B200_OBST_BITS = 2 };

/* lbm_b200_create / lbm_b200_create_inplace (inplace != 0) with the obstacle map in any of the formats above. */
int lbm_b200_create_ex(lbm_b200** handle, int nx, int ny, float density, float accel, float omega,
                       const void* obstacles, int obstacles_format, int n_slabs, const int* devices, int inplace);

/* ---- one slab per process (one rank per GPU; ranks launched by torchrun or similar) --- */

/* As lbm_b200_create, for the slab [first_row, first_row + rows) of a ny_global-row grid owned
 * by `rank` of `n_ranks` (ring neighbours rank-1 and rank+1, as d2q9-bgk.c:244-247).
 * obstacles_slab holds this slab's rows only (the reference's MPI_Scatterv, 968-970) and
 * free_cells_inv is the global value (the reference's MPI_Bcast, 966). */
int lbm_b200_create_slab(lbm_b200** handle, int nx, int ny_global, int first_row, int rows,
                         int rank, int n_ranks, float density, float accel, float omega,
                         float free_cells_inv, const int* obstacles_slab, int device);

/* The in-place twin of lbm_b200_create_slab (see lbm_b200_create_inplace); every rank of a ring must use the
 * same kind.  After an odd number of timesteps the populations of a slab's edge rows live in the neighbours'
 * buffers: call the state getters only when every rank has finished its run (lbm_b200_sync + a barrier). */
int lbm_b200_create_slab_inplace(lbm_b200** handle, int nx, int ny_global, int first_row, int rows,
                                 int rank, int n_ranks, float density, float accel, float omega,
                                 float free_cells_inv, const int* obstacles_slab, int device);

/* lbm_b200_create_slab / lbm_b200_create_slab_inplace (inplace != 0) with the slab's obstacle rows in any format. */
int lbm_b200_create_slab_ex(lbm_b200** handle, int nx, int ny_global, int first_row, int rows,
                            int rank, int n_ranks, float density, float accel, float omega,
                            float free_cells_inv, const void* obstacles_slab, int obstacles_format, int device, int inplace);

/* Size in bytes of the blob written by lbm_b200_ipc_export. */
int lbm_b200_ipc_blob_bytes(void);

/* Writes this slab's CUDA IPC handles (population buffers + flag words) so that the two ring
 * neighbours can map them.  Replaces MPI_Recv_init/MPI_Send_init (d2q9-bgk.c:295-313). */
int lbm_b200_ipc_export(lbm_b200* handle, void* blob);

/* Maps the neighbours' buffers.  `south` is the blob of rank-1 (owner of the rows below this
 * slab), `north` that of rank+1; with 2 ranks both are the same rank's blob.  Collective in
 * spirit: every rank must call it before any rank runs. */
int lbm_b200_ipc_connect(lbm_b200* handle, const void* south_blob, const void* north_blob);

/* ---- the hot path --------------------------------------------------------------------- */

/* Advances the state by `iters` timesteps, replacing the loop d2q9-bgk.c:315-394 (halo
 * exchange, accelerate_flow, timestep on interior and edge rows, per-step average, buffer
 * swap) and the final reduction (396).  All steps are enqueued without host synchronisation;
 * one synchronisation and one device-to-host copy happen at the end.
 * av_vels (host, `iters` floats, may be NULL) receives, per step,
 *   (float)(Sigma over this handle's free cells of |m|/rho) * free_cells_inv
 * i.e. the reference's av_vels_local[tt] (367); for a whole-domain handle that is the final
 * av_vels[tt], for a slab handle the ranks' arrays are summed element-wise as in (396). */
int lbm_b200_run(lbm_b200* handle, int iters, float* av_vels);

/* The same work split for device-side timing: enqueue returns at once, sync blocks, and
 * elapsed_ms reports the CUDA-event time of the last enqueue's timestep kernels (maximum
 * over the handle's devices); fetch copies the per-step averages of the last enqueue. */
int lbm_b200_enqueue(lbm_b200* handle, int iters);
int lbm_b200_sync(lbm_b200* handle);
int lbm_b200_elapsed_ms(lbm_b200* handle, float* ms);
int lbm_b200_fetch_av_vels(lbm_b200* handle, int iters, float* av_vels);

/* ---- state in and out -------------------------------------------------------------- */

/* Rows owned by the handle: the whole grid, or the slab.  Any pointer may be NULL. */
int lbm_b200_shape(const lbm_b200* handle, int* nx, int* rows, int* first_row);

/* Current populations in the reference's AoS layout, rows*nx*9 floats (what main() hands to
 * calc_reynolds/write_values, d2q9-bgk.c:408, 420). */
int lbm_b200_get_cells(lbm_b200* handle, float* cells);

/* Overwrites the populations (rows*nx*9 floats, AoS).  Not in the reference; lets tests start
 * from arbitrary states. */
int lbm_b200_set_cells(lbm_b200* handle, const float* cells);

/* Macroscopic fields of the current state, computed on the device exactly as write_values
 * does (d2q9-bgk.c:1076-1111): four arrays of rows*nx floats.  Blocked cells give
 * u_x = u_y = u = 0 and pressure = density/3.  Any pointer may be NULL. */
int lbm_b200_get_final_state(lbm_b200* handle, float* u_x, float* u_y, float* u, float* pressure);

/* Tuning knobs, all optional.  Unknown keys fail with LBM_B200_ERR_ARG.
 *   "kernel"        0 = auto, 1 = one cell per thread, 2 = four cells per thread (128-bit); reads back 3
 *                   when the resident variant of kernel 2 is in use, 4 on an in-place handle, 5 when two
 *                   timesteps are fused per pass ("fused2"), 7 when three or four are ("fused_steps") and 6 when the
 *                   grid lives in a cluster's shared memory
 *   "fused2"        two timesteps per pass over HBM (kernel 5: the first step of a 120-column strip goes into a
 *                   shared-memory ring, the second comes out of it; half the DRAM traffic per step, bit-identical
 *                   results; ring slabs keep two halo rows per side and exchange once per pass).  1 = on where it
 *                   applies (ping-pong handle, nx % 4 == 0, nx >= 240, >= 4 rows per slab), 0 = off, -1 = automatic
 *                   (on from 2^22 cells per GPU).  Reads back whether it is in use; "kernel" then reads 5 or 7.
 *                   On a multi-process ring set it on every rank while the ring is idle.
 *   "fused_steps"   timesteps per pass over HBM where "fused2" is in use: 2 = kernel 5; 3 or 4 = kernel 7 (a chain
 *                   of shared-memory rings, one per intermediate step; a third / a quarter of the DRAM traffic per
 *                   step; ring slabs keep four halo rows per side and need >= 6 rows each -- thinner ones fall back to
 *                   kernel 5); 0 = automatic (3).  Reads back the number in use (1 without "fused2").
 *                   Runs whose length is not a multiple end with a shorter pass through the same kernel.
 *   "fused_ctas"    kernel 7: CTAs per SM its resident warps are grouped into (0 = automatic: one warp per CTA, so that
 *                   an SM slot is free again the moment a work item ends; other values are for A/B measurements)
 *   "fused_k7"      1 = kernel 7 also for "fused_steps" = 2 (default 0: kernel 5)
 *   "cluster"       kernel 6: the whole grid resident in the shared memory of ONE 16-CTA thread-block cluster for up
 *                   to 256 timesteps per launch, halo rows read from the neighbour CTA over distributed shared
 *                   memory, one hardware cluster barrier per step (for launch-latency-bound decks: 128 x 128 and
 *                   128 x 256 fit; single-GPU ping-pong handles with ny % 16 == 0).  1 = wherever it fits, 0 = never,
 *                   -1 = automatic (where it fits and no other kernel / launch mode was asked for).  Reads back
 *                   whether it is in use; "kernel" then reads 6.
 *   "cluster_rows"  kernel 6 on 128-cell-wide grids of up to 256 rows: 1 (default) = one warp per row for the whole
 *                   launch, packed-pair arithmetic, halo rows pushed into the neighbour CTA (kernel 6b: 1.6 us per
 *                   step on the 128 x 128 deck instead of 2.3), 0 = the general form.  Reads back which one runs.
 *   "fused_deep"    staging rows of kernels 5 and 7: 1 = two (the copy runs two rows ahead of the arithmetic; kernel 5:
 *                   3 CTAs x 4 warps per SM), 0 = one (kernel 5: 2 CTAs x 8 warps; kernel 7: more resident warps),
 *                   -1 (default) = automatic: two for kernel 5, one for kernel 7.  Reads back 0 or 1.
 *   "prefetch_rows" kernel 5: bulk L2 prefetch this many rows ahead of the copy (default 0 = off: measured slower)
 *   "spin_timeout_ms"  how long a kernel of a ring slab waits for a neighbour's halo flag before it gives up and
 *                   lbm_b200_sync reports LBM_B200_ERR_STATE (default 30000)
 *   "debug_skip_slab"  test hook: the step kernels of this slab of a whole-domain handle are not launched
 *   "band_rows"     rows per work item of kernels 5 and 7 (0 = automatic)
 *   "inplace"       read-only: 1 on a handle made by lbm_b200_create_inplace
 *   "staging_bytes" in-place handles: size of the device staging buffer that get_cells / set_cells /
 *                   get_final_state move the state through, in chunks of whole rows (default 256 MB)
 *   "resident"      1 = run up to 256 timesteps per cooperative launch with a grid-wide barrier between
 *                   steps (for launch-latency-bound grids), 0 = never, -1 = automatic (single-GPU grids of
 *                   up to 2^22 cells)
 *   "graph_steps"   timesteps per CUDA-graph launch (0 = plain launches, -1 = automatic: graphs of 256 steps
 *                   for grids of up to 2^22 cells per GPU).  Ring slabs replay one independent graph per device:
 *                   their kernels are ordered by the flag words they exchange, not by stream dependencies
 *   "ctas_per_sm"   persistent-grid size in CTAs per SM (0 = occupancy query)
 *   "min_ctas"      register budget of the 128-bit kernel: 2, 3 or 4 resident CTAs per SM
 *   "cache_hint"    0 = read-only loads + plain stores, 1 = streaming loads and stores,
 *                   2 = read-only loads + streaming stores, 4 = as 0 plus a bulk L2 prefetch of each
 *                   warp's next row segment.  In-place handles: 0 = L2-only loads (ld.global.cg) + plain
 *                   stores, 1 = streaming loads and stores, 2 = plain (L1-cached) loads + plain stores
 */
int lbm_b200_set_option(lbm_b200* handle, const char* key, long value);
int lbm_b200_get_option(const lbm_b200* handle, const char* key, long* value);

void lbm_b200_destroy(lbm_b200* handle);

#ifdef __cplusplus
}
#endif
#endif /* LBM_B200_H */
