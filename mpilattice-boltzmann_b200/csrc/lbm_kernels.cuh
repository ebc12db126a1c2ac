// lbm_kernels.cuh -- the fused timestep kernels (sm_100a) and their small helpers.
//
// One timestep = ONE pass over the slab: pull-stream (propagate), bounce-back (rebound), BGK
// collision, next step's accelerate_flow on row ny-2, and the block-level partial of
// Sigma |m|/rho -- replacing reference d2q9-bgk.c:345-367.  Populations are nine fp32 planes
// (structure of arrays) of (rows+8) x nx floats: padded row 0 and rows+1 are the halo rows next to the
// slab, rows+2 .. rows+7 the neighbours' rows further away (kernels 5 and 7 on a ring), rows 1..rows are owned.  The
// obstacle map is 1 bit per cell, 32 cells per word, rows padded to whole words.  Algorithmic
// traffic: 9 loads + 9 stores = 72 B per cell per step (+1 bit).
//
//   kernel 1  step_scalar     one cell per thread, any nx
//   kernel 2  step_vec4       one warp per 128-cell row segment, 128-bit accesses + shuffles (0.989 of copy peak)
//   kernel 3  steps_resident  kernel 2 in a cooperative many-steps-per-launch loop (launch-latency-bound grids)
//   kernel 4  step_inplace    ONE buffer, AA access pattern (two alternating flavours)
//   kernel 5  steps2_strip    TWO timesteps per pass over HBM through a shared-memory ring
//   kernel 6  steps_cluster   the grid resident in the shared memory of one 16-CTA cluster, halo rows over DSMEM
//             (6b, steps_cluster_rows in lbm_cluster.cuh: one warp per 128-cell row -- the two smallest shipped decks)
//   kernel 7  steps_strip     K = 1..4 timesteps per pass through K-1 rings (lbm_stepsk.cuh; K = 3 is the default for
//             big grids)
// Ring slabs (multi-GPU): halo rows are stored straight into the neighbours' buffers over NVLink by the
// step kernels themselves; flag words (one per 128-cell chunk / 120-column strip) order the exchange.
#pragma once
#include <cstdint>
#include <type_traits>
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include "lbm_cell.cuh"

namespace lbm {

constexpr int kSegCells = 128;   // cells per warp work item in the vec4 kernel (32 lanes x 4)

// What a step kernel needs.  Passed by value (lives in the constant bank).
struct StepArgs {
  const float* src;          // plane 0 of the source buffer
  float* dst;                // plane 0 of the destination buffer
  size_t plane;              // floats between planes = (rows+8)*nx (four halo rows per side)
  const uint32_t* mask;      // bit-packed obstacles of the owned rows, row r at (r-1)*mask_row_words
  int mask_row_words;
  int nx;
  int chunks;                // ceil(nx / kSegCells)
  int row_begin, row_count, row_stride;   // processed padded rows: row_begin + i*row_stride
  int row_first, row_last;   // first / last owned padded row (1, rows)
  int south_of_first;        // padded source row below row_first (0 = halo, or rows = y-wrap)
  int north_of_last;         // padded source row above row_last (rows+1 = halo, or 1 = y-wrap)
  int accel_row;             // padded row of global row ny-2 if the next step's force is to be folded in, else -1
  StepConst c;
  double* partials;          // one double per CTA: this launch's share of Sigma |m|/rho
  // ---- halo exchange over peer memory (used only by the <PEER> instantiations) ----
  float* north_dst; size_t north_plane; int north_row;   // planes 2,5,6 of row_last  -> neighbour row north_row
  float* south_dst; size_t south_plane; int south_row;   // planes 4,7,8 of row_first -> neighbour row south_row
  // local flag words the neighbours write / the neighbours' flag words (peer memory): one word per slab for the
  // scalar kernel's launch-level handshake, one word per 128-cell chunk for the 128-bit kernels
  const unsigned* wait_from_south; const unsigned* wait_from_north;
  unsigned* signal_north; unsigned* signal_south;
  unsigned* epoch;           // local: number of states this slab has published
  unsigned* done;            // local: CTA completion counter of this launch
  unsigned* error;           // local: 0, or what a wait on a neighbour's flag gave up on (see spin_until)
  unsigned long long timeout_ns;   // how long one wait may last before it gives up
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p)
{
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v)
{
  asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned long long global_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Error word of a slab: (index of the chunk / strip waited for) << 8 | direction (1 = from the south, 2 = from the
// north); 0 = no error.
constexpr unsigned kWaitFromSouth = 1u, kWaitFromNorth = 2u;

// Waits until *flag >= need (wrap-around compare).  Where the reference would hang for ever in MPI_Waitall
// (d2q9-bgk.c:364) when a neighbour rank died, this wait is BOUNDED: after a.timeout_ns it records what it waited
// for in the slab's error word and gives up; every later wait of the slab then returns at once, so that the work
// already enqueued drains (producing garbage) and lbm_b200_sync can report LBM_B200_ERR_STATE instead of the GPU
// spinning until it is reset.
__device__ __noinline__ void spin_slow(const unsigned* flag, unsigned need, const StepArgs& a, unsigned code)
{
  volatile unsigned* err = a.error;
  if (*err) return;
  const unsigned long long t0 = global_ns();
  for (;;) {
    for (int i = 0; i < 32; i++) {
      if ((int)(ld_acquire_sys(flag) - need) >= 0) return;
      __nanosleep(20);
    }
    if (*err) return;
    if (global_ns() - t0 > a.timeout_ns) { atomicCAS(a.error, 0u, code); return; }
  }
}
__device__ __forceinline__ void spin_until(const unsigned* flag, unsigned need, const StepArgs& a, unsigned code)
{
  if ((int)(ld_acquire_sys(flag) - need) < 0) spin_slow(flag, need, a, code);
}

// Blocks until both ring neighbours have published the halo rows of the state this slab is
// about to read (their flag >= this slab's epoch).  Replaces MPI_Waitall (d2q9-bgk.c:364).
__device__ __forceinline__ void peer_wait(const StepArgs& a)
{
  if (threadIdx.x == 0) {
    const unsigned need = *reinterpret_cast<volatile unsigned*>(a.epoch);
    spin_until(a.wait_from_south, need, a, kWaitFromSouth);
    spin_until(a.wait_from_north, need, a, kWaitFromNorth);
  }
  __syncthreads();
}

// After every CTA has stored its share of the outgoing halo rows into the neighbours'
// buffers, the last CTA to finish publishes the new epoch to both neighbours.  Replaces the
// completion of MPI_Startall's sends (d2q9-bgk.c:327).
__device__ __forceinline__ void peer_signal(const StepArgs& a)
{
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(a.done, 1u);
    if (prev == gridDim.x - 1) {
      *a.done = 0;
      __threadfence_system();
      const unsigned next = *reinterpret_cast<volatile unsigned*>(a.epoch) + 1u;
      st_release_sys(a.signal_north, next);
      st_release_sys(a.signal_south, next);
      *reinterpret_cast<volatile unsigned*>(a.epoch) = next;
    }
  }
}

// Segment-granular versions for the 128-bit kernels, where the edge-row segments are the first work items of
// the SAME launch as the interior.  Every 128-cell chunk of an edge row has its own flag word in the neighbour:
//   * the warp that owns chunk c of the first (last) row waits until the southern (northern) neighbour has
//     published chunks c-1, c, c+1 (periodic) of the state it is about to read -- the two extra chunks supply the
//     scalars that cross the segment ends, and the same three flags say that the neighbour is done reading the
//     halo chunk this warp is about to overwrite;
//   * after its stores one system-scope fence and ONE flag store publish the chunk -- no counter, no "last warp",
//     so the exchange latency per step is one fence plus one NVLink hop.
// The epoch (number of states this slab has published) advances once per launch, by the last CTA to leave.
__device__ __forceinline__ void warp_peer_wait(const StepArgs& a, bool first, int ch)
{
  const int lane = threadIdx.x & 31;
  if (lane < 3) {
    int c = ch - 1 + lane;
    if (c < 0) c = a.chunks - 1; else if (c >= a.chunks) c = 0;
    const unsigned* p = (first ? a.wait_from_south : a.wait_from_north) + c;
    const unsigned need = *reinterpret_cast<volatile unsigned*>(a.epoch);
    spin_until(p, need, a, ((unsigned)c << 8) | (first ? kWaitFromSouth : kWaitFromNorth));
  }
  __syncwarp();
}
__device__ __forceinline__ void warp_peer_signal(const StepArgs& a, bool first, int ch)
{
  __threadfence_system();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    const unsigned next = *reinterpret_cast<volatile unsigned*>(a.epoch) + 1u;
    st_release_sys((first ? a.signal_south : a.signal_north) + ch, next);
  }
}
// end of a ring launch: the last CTA to leave advances the epoch (device scope: only this slab reads it)
__device__ __forceinline__ void peer_advance_epoch(const StepArgs& a)
{
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(a.done, 1u);
    if (prev == gridDim.x - 1) {
      *a.done = 0;
      *reinterpret_cast<volatile unsigned*>(a.epoch) = *reinterpret_cast<volatile unsigned*>(a.epoch) + 1u;
    }
  }
}

// Global access flavours of the vec4 kernel (template parameter HINT):
//   0  ld.global.nc (read-only path) + plain st.global
//   1  ld.global.cs + st.global.cs   (streaming: evict-first in L1 and L2)
//   2  ld.global.nc + st.global.cs
//   3  ld.global.cg (L2 only, coherent within a launch) + plain st.global   [resident kernel]
//   4  as 0, plus a bulk L2 prefetch (cp.async.bulk.prefetch.L2) of the warp's NEXT segment
template <int HINT>
__device__ __forceinline__ float4 load4(const float* p)
{
  if (HINT == 1) return __ldcs(reinterpret_cast<const float4*>(p));
  if (HINT == 3) return __ldcg(reinterpret_cast<const float4*>(p));
  return __ldg(reinterpret_cast<const float4*>(p));
}
template <int HINT>
__device__ __forceinline__ float load1(const float* p)
{
  if (HINT == 1) return __ldcs(p);
  if (HINT == 3) return __ldcg(p);
  return __ldg(p);
}
template <int HINT>
__device__ __forceinline__ void store4(float* p, float4 v)
{
  if (HINT == 0 || HINT == 3 || HINT == 4) *reinterpret_cast<float4*>(p) = v;
  else __stcs(reinterpret_cast<float4*>(p), v);
}

// Sum of `v` over the CTA, written by thread 0 to *out.  Deterministic (fixed tree).
__device__ __forceinline__ void block_sum_to(double v, double* out)
{
  __shared__ double warp_sums[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_sums[warp] = v;
  __syncthreads();
  if (warp == 0) {
    const int nwarps = (blockDim.x + 31) >> 5;
    double s = (lane < nwarps) ? warp_sums[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) *out = s;
  }
}

// ---------------------------------------------------------------------------------------
// Kernel 2 ("vec4"): one warp per 128-cell row segment, four cells per lane.  Every global
// access is a 128-bit aligned, fully coalesced row load/store; the +-1 x-shifts of the six
// x-moving populations come from the neighbouring lane by warp shuffle, and only the two end
// lanes of a segment fetch one extra scalar across the segment (or the periodic) boundary.
// CTAs stride over the segments (grid sized by the host).  Requires nx % 4 == 0, nx >= 8.
// ---------------------------------------------------------------------------------------
// One pass of this CTA over its share of the row segments: src -> dst.  Returns the thread's share of
// Sigma |m|/rho.  `accel_row` = padded row that gets the next step's body force folded in (or -1).
template <bool PEER, int HINT>
__device__ __forceinline__ double vec4_pass(const StepArgs& a, const float* __restrict__ src, float* __restrict__ dst,
                                            const int accel_row)
{
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  const long nseg = (long)a.row_count * a.chunks;
  const size_t P = a.plane;
  double acc = 0.0;

  for (long seg = (long)blockIdx.x * warps + (threadIdx.x >> 5); seg < nseg; seg += (long)gridDim.x * warps) {
    int row, ch;
    bool edge = false;
    if (PEER) {
      // multi-GPU: the two edge rows are the first 2*chunks work items, so their halo stores leave over
      // NVLink at the start of the launch and the interior hides the exchange (d2q9-bgk.c:326-366)
      const long e2 = 2L * a.chunks;
      if (seg < e2) {
        edge = true;
        const bool first = seg < a.chunks;
        row = first ? a.row_first : a.row_last;
        ch = (int)(first ? seg : seg - a.chunks);
        warp_peer_wait(a, first, ch);
      } else {
        const long s2 = seg - e2;
        const int ri = (int)(s2 / a.chunks);
        ch = (int)(s2 - (long)ri * a.chunks);
        row = a.row_first + 1 + ri;
      }
    } else {
      const int ri = (int)(seg / a.chunks);
      ch = (int)(seg - (long)ri * a.chunks);
      row = a.row_begin + ri * a.row_stride;
    }
    const int rs = (row == a.row_first) ? a.south_of_first : row - 1;
    const int rn = (row == a.row_last) ? a.north_of_last : row + 1;
    const int x0 = ch * kSegCells + lane * 4;
    const bool active = x0 < a.nx;
    const bool west_edge = (lane == 0);
    const bool east_edge = (lane == 31) || (x0 + 4 >= a.nx);
    const int xw = (x0 == 0) ? a.nx - 1 : x0 - 1;
    const int xe = (x0 + 4 >= a.nx) ? 0 : x0 + 4;

    const size_t o_c = (size_t)row * a.nx, o_s = (size_t)rs * a.nx, o_n = (size_t)rn * a.nx;

    if (HINT == 4 && !PEER) {
      // pull the nine 512-byte row pieces of this warp's next segment into L2 while this one is processed:
      // lane k asks for plane k with one bulk-prefetch instruction
      const long nxt = seg + (long)gridDim.x * warps;
      if (nxt < nseg && lane < 9) {
        const int nri = (int)(nxt / a.chunks);
        const int nch = (int)(nxt - (long)nri * a.chunks);
        const int nrow = a.row_begin + nri * a.row_stride;
        const int nrs = (nrow == a.row_first) ? a.south_of_first : nrow - 1;
        const int nrn = (nrow == a.row_last) ? a.north_of_last : nrow + 1;
        // planes 0,1,3 come from the row itself, 2,5,6 from the row below, 4,7,8 from the row above
        const int from = (lane == 0 || lane == 1 || lane == 3) ? nrow : ((lane == 2 || lane == 5 || lane == 6) ? nrs : nrn);
        const float* p = src + (size_t)lane * P + (size_t)from * a.nx + (size_t)nch * kSegCells;
        const unsigned bytes = (unsigned)min(kSegCells, a.nx - nch * kSegCells) * 4u;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
      }
    }

    float4 c[9];
    float e_c = 0.f, e_s = 0.f, e_n = 0.f;
    uint32_t mw = 0;
    if (active) {
      c[0] = load4<HINT>(src + 0 * P + o_c + x0);
      c[1] = load4<HINT>(src + 1 * P + o_c + x0);
      c[2] = load4<HINT>(src + 2 * P + o_s + x0);
      c[3] = load4<HINT>(src + 3 * P + o_c + x0);
      c[4] = load4<HINT>(src + 4 * P + o_n + x0);
      c[5] = load4<HINT>(src + 5 * P + o_s + x0);
      c[6] = load4<HINT>(src + 6 * P + o_s + x0);
      c[7] = load4<HINT>(src + 7 * P + o_n + x0);
      c[8] = load4<HINT>(src + 8 * P + o_n + x0);
      mw = __ldg(a.mask + (size_t)(row - 1) * a.mask_row_words + (x0 >> 5));
      if (west_edge || east_edge) {
        // west end: populations 1, 5, 8 arrive from x-1; east end: 3, 6, 7 arrive from x+1
        e_c = load1<HINT>(west_edge ? src + 1 * P + o_c + xw : src + 3 * P + o_c + xe);
        e_s = load1<HINT>(west_edge ? src + 5 * P + o_s + xw : src + 6 * P + o_s + xe);
        e_n = load1<HINT>(west_edge ? src + 8 * P + o_n + xw : src + 7 * P + o_n + xe);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 9; k++) c[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    // neighbour lanes supply the value that crosses the 4-cell boundary
    const float up1 = __shfl_up_sync(0xffffffffu, c[1].w, 1);
    const float up5 = __shfl_up_sync(0xffffffffu, c[5].w, 1);
    const float up8 = __shfl_up_sync(0xffffffffu, c[8].w, 1);
    const float dn3 = __shfl_down_sync(0xffffffffu, c[3].x, 1);
    const float dn6 = __shfl_down_sync(0xffffffffu, c[6].x, 1);
    const float dn7 = __shfl_down_sync(0xffffffffu, c[7].x, 1);
    const float w1v = west_edge ? e_c : up1, w5v = west_edge ? e_s : up5, w8v = west_edge ? e_n : up8;
    const float e3v = east_edge ? e_c : dn3, e6v = east_edge ? e_s : dn6, e7v = east_edge ? e_n : dn7;

    const unsigned bits = active ? ((mw >> (x0 & 31)) & 0xFu) : 0u;
    const bool any_blocked = __any_sync(0xffffffffu, bits != 0u);   // warp-uniform: packed pairs where no lane is blocked
    if (active) {
      const bool fold_accel = (row == accel_row);
      float f[4][9];
      f[0][0] = c[0].x; f[1][0] = c[0].y; f[2][0] = c[0].z; f[3][0] = c[0].w;
      f[0][1] = w1v;    f[1][1] = c[1].x; f[2][1] = c[1].y; f[3][1] = c[1].z;
      f[0][2] = c[2].x; f[1][2] = c[2].y; f[2][2] = c[2].z; f[3][2] = c[2].w;
      f[0][3] = c[3].y; f[1][3] = c[3].z; f[2][3] = c[3].w; f[3][3] = e3v;
      f[0][4] = c[4].x; f[1][4] = c[4].y; f[2][4] = c[4].z; f[3][4] = c[4].w;
      f[0][5] = w5v;    f[1][5] = c[5].x; f[2][5] = c[5].y; f[3][5] = c[5].z;
      f[0][6] = c[6].y; f[1][6] = c[6].z; f[2][6] = c[6].w; f[3][6] = e6v;
      f[0][7] = c[7].y; f[1][7] = c[7].z; f[2][7] = c[7].w; f[3][7] = e7v;
      f[0][8] = w8v;    f[1][8] = c[8].x; f[2][8] = c[8].y; f[3][8] = c[8].z;

      acc += (double)collide4(f, bits, any_blocked, a.c, fold_accel);

#pragma unroll
      for (int k = 0; k < 9; k++)
        store4<HINT>(dst + k * P + o_c + x0, make_float4(f[0][k], f[1][k], f[2][k], f[3][k]));

      if (PEER) {
        // one-row halo exchange: NVLink stores straight into the neighbour's halo row
        // (replaces the MPI_Send_init/MPI_Recv_init pairs of d2q9-bgk.c:295-313; only the
        // three populations that cross the slab boundary travel, 12*nx B instead of 36*nx B)
        if (row == a.row_last) {
          const size_t o = (size_t)a.north_row * a.nx + x0;
          *reinterpret_cast<float4*>(a.north_dst + 2 * a.north_plane + o) = make_float4(f[0][2], f[1][2], f[2][2], f[3][2]);
          *reinterpret_cast<float4*>(a.north_dst + 5 * a.north_plane + o) = make_float4(f[0][5], f[1][5], f[2][5], f[3][5]);
          *reinterpret_cast<float4*>(a.north_dst + 6 * a.north_plane + o) = make_float4(f[0][6], f[1][6], f[2][6], f[3][6]);
        }
        if (row == a.row_first) {
          const size_t o = (size_t)a.south_row * a.nx + x0;
          *reinterpret_cast<float4*>(a.south_dst + 4 * a.south_plane + o) = make_float4(f[0][4], f[1][4], f[2][4], f[3][4]);
          *reinterpret_cast<float4*>(a.south_dst + 7 * a.south_plane + o) = make_float4(f[0][7], f[1][7], f[2][7], f[3][7]);
          *reinterpret_cast<float4*>(a.south_dst + 8 * a.south_plane + o) = make_float4(f[0][8], f[1][8], f[2][8], f[3][8]);
        }
      }
    }
    if (PEER && edge) warp_peer_signal(a, row == a.row_first, ch);
  }

  return acc;
}

// PEER = true: a slab of a multi-GPU ring -- ONE launch per timestep computes the whole slab, pushes the
// outgoing halo rows into the neighbours' buffers and does the flag handshake (loads are L2-coherent
// ld.global.cg, HINT 3, because halo rows are written by another GPU).
template <bool PEER, int MIN_CTAS, int HINT>
__global__ void __launch_bounds__(256, MIN_CTAS) step_vec4(const StepArgs a)
{
  const double acc = vec4_pass<PEER, PEER ? 3 : HINT>(a, a.src, a.dst, a.accel_row);
  block_sum_to(acc, a.partials + blockIdx.x);
  if (PEER) peer_advance_epoch(a);
}

// ---------------------------------------------------------------------------------------
// Kernel 4 ("in place", the AA access pattern): ONE population buffer instead of the ping-pong pair, same
// 72 B/cell/step.  Every cell reads and writes exactly the same nine memory locations in a step, so a step
// needs no second buffer and no ordering between cells.  Two alternating step flavours:
//   NEIGHBOUR (state layout L0 -> L1): pull population i from (slot i, x - c_i) exactly like step_vec4,
//       collide, and write the post-collision population opp(i) back to that SAME location.  L1 therefore
//       holds, at (slot i, y), the population that arrives at y travelling in direction opp(i).
//   LOCAL (L1 -> L0): population j of cell y is at (slot opp(j), y); collide; store population j at
//       (slot j, y).  All nine accesses are aligned 128-bit row accesses of the cell's own row.
// L0 is the canonical layout of the ping-pong kernels (post-collision, not yet streamed), so after an even
// number of steps the buffer is bit-identical to theirs; after an odd number the accessors below decode L1.
// The arithmetic is the same collide()/accelerate() as everywhere else.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int opposite(int k) { return k == 0 ? 0 : (k <= 4 ? ((k + 1) & 3) + 1 : ((k - 3) & 3) + 5); }

template <int HINT>
__device__ __forceinline__ float4 ld4_rw(const float* p)
{
  if (HINT == 1) return __ldcs(reinterpret_cast<const float4*>(p));
  if (HINT == 2) return *reinterpret_cast<const float4*>(p);
  return __ldcg(reinterpret_cast<const float4*>(p));
}
template <int HINT>
__device__ __forceinline__ float ld1_rw(const float* p)
{
  if (HINT == 1) return __ldcs(p);
  if (HINT == 2) return *p;
  return __ldcg(p);
}
template <int HINT>
__device__ __forceinline__ void st4_rw(float* p, float4 v)
{
  if (HINT == 1) __stcs(reinterpret_cast<float4*>(p), v);
  else *reinterpret_cast<float4*>(p) = v;
}
// the first / last three floats of an aligned group of four (the fourth belongs to the neighbouring segment)
__device__ __forceinline__ void st3_low(float* p, float4 v)
{
  *reinterpret_cast<float2*>(p) = make_float2(v.x, v.y);
  p[2] = v.z;
}
__device__ __forceinline__ void st3_high(float* p, float4 v)
{
  p[1] = v.y;
  *reinterpret_cast<float2*>(p + 2) = make_float2(v.z, v.w);
}

// One 128-cell row segment in place.  EDGE = true: an edge row of a ring slab (multi-GPU) -- what crosses the
// slab boundary:
//   LOCAL flavour      (-> L0) copies of planes 2,5,6 of the last row go to the northern neighbour's halo row 0 and
//                      planes 4,7,8 of the first row to the southern neighbour's halo row rows+1 (as step_vec4);
//   NEIGHBOUR flavour  (-> L1) pulls from the local halo rows, and the values it would write back into a halo row are
//                      stored straight into the neighbour's OWNED edge row instead (its row `rows` / row 1), which is
//                      where the neighbour's next LOCAL step reads them -- nobody else touches those slots.
template <bool NEIGHBOUR, bool EDGE, int HINT>
__device__ __forceinline__ void inplace_segment(const StepArgs& a, float* __restrict__ buf, const int row, const int ch,
                                                const int accel_row, double& acc)
{
  const int lane = threadIdx.x & 31;
  const size_t P = a.plane;
  const int x0 = ch * kSegCells + lane * 4;
  const bool active = x0 < a.nx;
  const size_t o_c = (size_t)row * a.nx;
  const unsigned bits = active ? ((__ldg(a.mask + (size_t)(row - 1) * a.mask_row_words + (x0 >> 5)) >> (x0 & 31)) & 0xFu) : 0u;
  const bool fold_accel = (row == accel_row);
  const bool any_blocked = __any_sync(0xffffffffu, bits != 0u);
  float f[4][9];

  if (!NEIGHBOUR) {
    if (active) {
#pragma unroll
      for (int k = 0; k < 9; k++) {
        const float4 v = ld4_rw<HINT>(buf + (size_t)opposite(k) * P + o_c + x0);
        f[0][k] = v.x; f[1][k] = v.y; f[2][k] = v.z; f[3][k] = v.w;
      }
      acc += (double)collide4(f, bits, any_blocked, a.c, fold_accel);
#pragma unroll
      for (int k = 0; k < 9; k++)
        st4_rw<HINT>(buf + (size_t)k * P + o_c + x0, make_float4(f[0][k], f[1][k], f[2][k], f[3][k]));
      if (EDGE) {
        if (row == a.row_last) {
          const size_t o = (size_t)a.north_row * a.nx + x0;
          *reinterpret_cast<float4*>(a.north_dst + 2 * a.north_plane + o) = make_float4(f[0][2], f[1][2], f[2][2], f[3][2]);
          *reinterpret_cast<float4*>(a.north_dst + 5 * a.north_plane + o) = make_float4(f[0][5], f[1][5], f[2][5], f[3][5]);
          *reinterpret_cast<float4*>(a.north_dst + 6 * a.north_plane + o) = make_float4(f[0][6], f[1][6], f[2][6], f[3][6]);
        }
        if (row == a.row_first) {
          const size_t o = (size_t)a.south_row * a.nx + x0;
          *reinterpret_cast<float4*>(a.south_dst + 4 * a.south_plane + o) = make_float4(f[0][4], f[1][4], f[2][4], f[3][4]);
          *reinterpret_cast<float4*>(a.south_dst + 7 * a.south_plane + o) = make_float4(f[0][7], f[1][7], f[2][7], f[3][7]);
          *reinterpret_cast<float4*>(a.south_dst + 8 * a.south_plane + o) = make_float4(f[0][8], f[1][8], f[2][8], f[3][8]);
        }
      }
    }
    return;
  }

  // ---- NEIGHBOUR flavour: the pull of step_vec4, then the push back into the very same locations ----
  const int rs = (row == a.row_first) ? a.south_of_first : row - 1;
  const int rn = (row == a.row_last) ? a.north_of_last : row + 1;
  const bool west_edge = (lane == 0);
  const bool east_edge = (lane == 31) || (x0 + 4 >= a.nx);
  const int xw = (x0 == 0) ? a.nx - 1 : x0 - 1;
  const int xe = (x0 + 4 >= a.nx) ? 0 : x0 + 4;
  const size_t o_s = (size_t)rs * a.nx, o_n = (size_t)rn * a.nx;

  float4 c[9];
  float w_c = 0.f, w_s = 0.f, w_n = 0.f, e_c = 0.f, e_s = 0.f, e_n = 0.f;
  if (active) {
    c[0] = ld4_rw<HINT>(buf + 0 * P + o_c + x0);
    c[1] = ld4_rw<HINT>(buf + 1 * P + o_c + x0);
    c[2] = ld4_rw<HINT>(buf + 2 * P + o_s + x0);
    c[3] = ld4_rw<HINT>(buf + 3 * P + o_c + x0);
    c[4] = ld4_rw<HINT>(buf + 4 * P + o_n + x0);
    c[5] = ld4_rw<HINT>(buf + 5 * P + o_s + x0);
    c[6] = ld4_rw<HINT>(buf + 6 * P + o_s + x0);
    c[7] = ld4_rw<HINT>(buf + 7 * P + o_n + x0);
    c[8] = ld4_rw<HINT>(buf + 8 * P + o_n + x0);
    if (west_edge) {
      w_c = ld1_rw<HINT>(buf + 1 * P + o_c + xw);
      w_s = ld1_rw<HINT>(buf + 5 * P + o_s + xw);
      w_n = ld1_rw<HINT>(buf + 8 * P + o_n + xw);
    }
    if (east_edge) {
      e_c = ld1_rw<HINT>(buf + 3 * P + o_c + xe);
      e_s = ld1_rw<HINT>(buf + 6 * P + o_s + xe);
      e_n = ld1_rw<HINT>(buf + 7 * P + o_n + xe);
    }
  } else {
#pragma unroll
    for (int k = 0; k < 9; k++) c[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float up1 = __shfl_up_sync(0xffffffffu, c[1].w, 1);
  const float up5 = __shfl_up_sync(0xffffffffu, c[5].w, 1);
  const float up8 = __shfl_up_sync(0xffffffffu, c[8].w, 1);
  const float dn3 = __shfl_down_sync(0xffffffffu, c[3].x, 1);
  const float dn6 = __shfl_down_sync(0xffffffffu, c[6].x, 1);
  const float dn7 = __shfl_down_sync(0xffffffffu, c[7].x, 1);

  f[0][0] = c[0].x; f[1][0] = c[0].y; f[2][0] = c[0].z; f[3][0] = c[0].w;
  f[0][1] = west_edge ? w_c : up1; f[1][1] = c[1].x; f[2][1] = c[1].y; f[3][1] = c[1].z;
  f[0][2] = c[2].x; f[1][2] = c[2].y; f[2][2] = c[2].z; f[3][2] = c[2].w;
  f[0][3] = c[3].y; f[1][3] = c[3].z; f[2][3] = c[3].w; f[3][3] = east_edge ? e_c : dn3;
  f[0][4] = c[4].x; f[1][4] = c[4].y; f[2][4] = c[4].z; f[3][4] = c[4].w;
  f[0][5] = west_edge ? w_s : up5; f[1][5] = c[5].x; f[2][5] = c[5].y; f[3][5] = c[5].z;
  f[0][6] = c[6].y; f[1][6] = c[6].z; f[2][6] = c[6].w; f[3][6] = east_edge ? e_s : dn6;
  f[0][7] = c[7].y; f[1][7] = c[7].z; f[2][7] = c[7].w; f[3][7] = east_edge ? e_n : dn7;
  f[0][8] = west_edge ? w_n : up8; f[1][8] = c[8].x; f[2][8] = c[8].y; f[3][8] = c[8].z;

  if (active) acc += (double)collide4(f, bits, any_blocked, a.c, fold_accel);
  // the value that belongs to the neighbouring lane's aligned group of four: (slot 1, x-1) <- f3(x) etc.
  const float g3 = __shfl_down_sync(0xffffffffu, f[0][3], 1);
  const float g7 = __shfl_down_sync(0xffffffffu, f[0][7], 1);
  const float g6 = __shfl_down_sync(0xffffffffu, f[0][6], 1);
  const float g1 = __shfl_up_sync(0xffffffffu, f[3][1], 1);
  const float g8 = __shfl_up_sync(0xffffffffu, f[3][8], 1);
  const float g5 = __shfl_up_sync(0xffffffffu, f[3][5], 1);
  if (active) {
    // where the rows below / above are written: the rows just read, or (slab edge rows of a ring) the
    // neighbour's owned edge row over NVLink
    float* bs = buf; float* bn = buf;
    size_t Ps = P, Pn = P, w_s_off = o_s, w_n_off = o_n;
    if (EDGE) {
      if (row == a.row_first) { bs = a.south_dst; Ps = a.south_plane; w_s_off = (size_t)a.south_row * a.nx; }
      if (row == a.row_last) { bn = a.north_dst; Pn = a.north_plane; w_n_off = (size_t)a.north_row * a.nx; }
    }
    st4_rw<HINT>(buf + 0 * P + o_c + x0, make_float4(f[0][0], f[1][0], f[2][0], f[3][0]));
    st4_rw<HINT>(bs + 2 * Ps + w_s_off + x0, make_float4(f[0][4], f[1][4], f[2][4], f[3][4]));
    st4_rw<HINT>(bn + 4 * Pn + w_n_off + x0, make_float4(f[0][2], f[1][2], f[2][2], f[3][2]));
    const float4 v1 = make_float4(f[1][3], f[2][3], f[3][3], g3);   // slot 1, this row
    const float4 v5 = make_float4(f[1][7], f[2][7], f[3][7], g7);   // slot 5, row below
    const float4 v8 = make_float4(f[1][6], f[2][6], f[3][6], g6);   // slot 8, row above
    const float4 v3 = make_float4(g1, f[0][1], f[1][1], f[2][1]);   // slot 3, this row
    const float4 v6 = make_float4(g8, f[0][8], f[1][8], f[2][8]);   // slot 6, row below
    const float4 v7 = make_float4(g5, f[0][5], f[1][5], f[2][5]);   // slot 7, row above
    if (!east_edge) {
      st4_rw<HINT>(buf + 1 * P + o_c + x0, v1);
      st4_rw<HINT>(bs + 5 * Ps + w_s_off + x0, v5);
      st4_rw<HINT>(bn + 8 * Pn + w_n_off + x0, v8);
    } else {
      // the fourth element is the next segment's west scalar; this lane's own east scalars go back too
      st3_low(buf + 1 * P + o_c + x0, v1);
      st3_low(bs + 5 * Ps + w_s_off + x0, v5);
      st3_low(bn + 8 * Pn + w_n_off + x0, v8);
      buf[3 * P + o_c + xe] = f[3][1];
      bs[6 * Ps + w_s_off + xe] = f[3][8];
      bn[7 * Pn + w_n_off + xe] = f[3][5];
    }
    if (!west_edge) {
      st4_rw<HINT>(buf + 3 * P + o_c + x0, v3);
      st4_rw<HINT>(bs + 6 * Ps + w_s_off + x0, v6);
      st4_rw<HINT>(bn + 7 * Pn + w_n_off + x0, v7);
    } else {
      st3_high(buf + 3 * P + o_c + x0, v3);
      st3_high(bs + 6 * Ps + w_s_off + x0, v6);
      st3_high(bn + 7 * Pn + w_n_off + x0, v7);
      buf[1 * P + o_c + xw] = f[0][3];
      bs[5 * Ps + w_s_off + xw] = f[0][7];
      bn[8 * Pn + w_n_off + xw] = f[0][6];
    }
  }
}

// PEER = true: a slab of a multi-GPU ring.  As in step_vec4<PEER> the two edge rows are the first work items of
// the launch and carry the flag handshake; the interior rows run the same code as a single-GPU launch.
template <bool NEIGHBOUR, bool PEER, int HINT>
__device__ __forceinline__ double inplace_pass(const StepArgs& a, float* __restrict__ buf, const int accel_row)
{
  const int warps = blockDim.x >> 5;
  const long nseg = (long)a.row_count * a.chunks;
  double acc = 0.0;

  for (long seg = (long)blockIdx.x * warps + (threadIdx.x >> 5); seg < nseg; seg += (long)gridDim.x * warps) {
    if (PEER) {
      const long e2 = 2L * a.chunks;
      if (seg < e2) {
        const bool first = seg < a.chunks;
        const int ch = (int)(first ? seg : seg - a.chunks);
        warp_peer_wait(a, first, ch);
        inplace_segment<NEIGHBOUR, true, HINT>(a, buf, first ? a.row_first : a.row_last, ch, accel_row, acc);
        warp_peer_signal(a, first, ch);
      } else {
        const long s2 = seg - e2;
        const int ri = (int)(s2 / a.chunks);
        inplace_segment<NEIGHBOUR, false, HINT>(a, buf, a.row_first + 1 + ri, (int)(s2 - (long)ri * a.chunks), accel_row, acc);
      }
    } else {
      const int ri = (int)(seg / a.chunks);
      inplace_segment<NEIGHBOUR, false, HINT>(a, buf, a.row_begin + ri * a.row_stride, (int)(seg - (long)ri * a.chunks),
                                              accel_row, acc);
    }
  }
  return acc;
}

// Loads of a ring slab are always L2-coherent (HINT 0): halo rows and edge-row slots are written by another GPU.
template <bool NEIGHBOUR, bool PEER, int HINT>
__global__ void __launch_bounds__(256, 2) step_inplace(const StepArgs a)
{
  const double acc = inplace_pass<NEIGHBOUR, PEER, PEER ? 0 : HINT>(a, a.dst, a.accel_row);
  block_sum_to(acc, a.partials + blockIdx.x);
  if (PEER) peer_advance_epoch(a);
}

// Where the canonical population k of the cell (x, padded row) lives: in place (odd = 0: layout L0), or
// (odd = 1: layout L1 of the in-place kernels) at the destination cell in the opposite slot -- periodic in x,
// and across the slab's first / last row in the ring neighbour's owned edge row (the slab itself when it is the
// whole grid).
struct Layout {
  size_t plane;
  int nx, rows, odd;
  const float* south; size_t south_plane; int south_rows;   // owner of the row below row 1, and its last owned row
  const float* north; size_t north_plane;                   // owner of the row above row `rows` (its row 1)
};
__device__ __forceinline__ const float* locate(const Layout& l, const float* buf, int k, int x, int row)
{
  if (!l.odd || k == 0) return buf + (size_t)k * l.plane + (size_t)row * l.nx + x;
  const int cx = (k == 1 || k == 5 || k == 8) ? 1 : ((k == 3 || k == 6 || k == 7) ? -1 : 0);
  const int cy = (k == 2 || k == 5 || k == 6) ? 1 : ((k == 4 || k == 7 || k == 8) ? -1 : 0);
  int xo = x + cx;
  const int ro = row + cy;
  if (xo < 0) xo = l.nx - 1; else if (xo >= l.nx) xo = 0;
  if (ro < 1) return l.south + (size_t)opposite(k) * l.south_plane + (size_t)l.south_rows * l.nx + xo;
  if (ro > l.rows) return l.north + (size_t)opposite(k) * l.north_plane + (size_t)l.nx + xo;
  return buf + (size_t)opposite(k) * l.plane + (size_t)ro * l.nx + xo;
}
__device__ __forceinline__ float* locate(const Layout& l, float* buf, int k, int x, int row)
{
  return const_cast<float*>(locate(l, const_cast<const float*>(buf), k, x, row));
}

// ---------------------------------------------------------------------------------------
// Kernel 5 ("fused2"): TWO timesteps per pass over HBM.  The streaming kernels above move 72 B per cell and step
// and sit at the HBM roofline; the only way past it is to touch DRAM less often.  Here a warp walks down a strip
// of 120 owned columns: the first step is computed for the 128 aligned columns around the strip (4 redundant on
// each side, so the second step finds its x-neighbours) and kept in a three-row ring in shared memory; the second
// step pulls from that ring and stores to the other buffer.  Per cell and PAIR of steps: 36 B read + 36 B written
// (+6.7 % redundant columns, +2/band_rows redundant rows) -- about half the DRAM traffic per step, at the price of
// 1.07x the arithmetic; the kernel is bound by instruction issue, not by HBM.
//   work item  = (band of `band_rows` rows, strip); one warp per item, no inter-warp synchronisation
//   step 1     row y needs planes 2,5,6 of row y-1, 0,1,3 of row y and 4,7,8 of row y+1: every (row, plane) is
//              loaded exactly once per strip, as in step_vec4 (128-bit loads + shuffles + end-lane scalars)
//   step 2     row y needs step-1 rows y-1, y, y+1: 128-bit shared-memory loads + shuffles; lanes 1..30 own output
// The arithmetic per cell and step is the same collide()/accelerate(): results are bit-identical to two launches
// of step_vec4.  Ring slabs keep two halo rows per side (template parameter PEER below).
// ---------------------------------------------------------------------------------------
constexpr int kStripOut = 120;    // owned columns per strip (30 lanes x 4)

struct FusedArgs {
  int band_rows;        // rows per work item
  int bands, strips;
  int accel_row;        // padded row of global row ny-2, or -1 (the first step always gets the second step's force)
  int fold_last;        // whether the launch's last step folds the following step's force in
  int partial_stride;   // doubles between the two steps' partials
  int south_rows, north_rows;   // owned rows of the ring neighbours (PEER)
  int prefetch_rows;    // L2 prefetch distance in rows (0 = off)
};

// (destination = a 32-bit shared-window address, converted once per kernel, not per copy)
__device__ __forceinline__ void cp_async16(unsigned smem_addr, const void* gmem)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_addr), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// Ring slabs: one flag word per 120-column strip and direction, same protocol as warp_peer_wait/warp_peer_signal.
__device__ __forceinline__ void strip_wait(const StepArgs& a, const unsigned* flags, unsigned dir, int strip, int strips, int lane)
{
  if (lane < 3) {
    int c = strip - 1 + lane;
    if (c < 0) c = strips - 1; else if (c >= strips) c = 0;
    const unsigned need = *reinterpret_cast<volatile unsigned*>(a.epoch);
    spin_until(flags + c, need, a, ((unsigned)c << 8) | dir);
  }
  __syncwarp();
}
__device__ __forceinline__ void strip_signal(const StepArgs& a, unsigned* flags, int strip, int lane)
{
  __threadfence_system();
  __syncwarp();
  if (lane == 0) st_release_sys(flags + strip, *reinterpret_cast<volatile unsigned*>(a.epoch) + 1u);
}

// Shared memory of one warp: a three-row ring of the six first-step planes the second step reads later (planes
// 4,7,8 of a row are consumed straight from registers) and one staging row filled by cp.async one row ahead of the
// arithmetic.  Every lane reads back only the 16-byte cells it wrote (or copied) itself -- what crosses lanes goes
// through warp shuffles -- so neither a warp barrier nor an mbarrier sits between the producers and consumers.
constexpr int kRingPlanes = 6;                                   // ring order: planes 0, 1, 3, 2, 5, 6
constexpr int kRingSlot = kRingPlanes * 32;                      // float4 per ring row
constexpr int kStageSlot = 9 * 32;                               // float4 per staging row
// DEEP: two staging rows (the copy runs two rows ahead of the arithmetic) at 3 CTAs x 4 warps per SM and up to 168
// registers, instead of one staging row at 2 CTAs x 8 warps and 128 registers.
__host__ __device__ constexpr int fused_warp_float4(bool deep) { return 3 * kRingSlot + (deep ? 2 : 1) * kStageSlot; }

// 128-bit accesses that deliver / take two packed pairs in 64-bit registers (no re-packing around the access)
__device__ __forceinline__ void lds2(const float4* p, f2& a, f2& b)
{
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
}
__device__ __forceinline__ void sts2(float4* p, f2 a, f2 b)
{
  asm volatile("st.shared.v2.u64 [%0], {%1, %2};" :: "r"((unsigned)__cvta_generic_to_shared(p)), "l"(a), "l"(b) : "memory");
}
template <int HINT>
__device__ __forceinline__ void stg2(float* p, f2 a, f2 b)
{
  if (HINT == 0 || HINT == 3 || HINT == 4) asm volatile("st.global.v2.u64 [%0], {%1, %2};" :: "l"(p), "l"(a), "l"(b) : "memory");
  else asm volatile("st.global.cs.v2.u64 [%0], {%1, %2};" :: "l"(p), "l"(a), "l"(b) : "memory");
}

// The rare path of collide_quad: a row segment with obstacles, or the driven row (the body force folded in).  Kept out
// of line -- with nothing but one array and scalars crossing the call: every other interface tried (the pairs
// themselves, a StepConst reference, results through references) cost the COMMON path 3..7 % through its register
// allocation (same-box A/B, profiles/r02_fused2.md).
__device__ __noinline__ float collide4_generic(float (&f)[4][9], unsigned bits, float omega, bool fold, float aw1, float aw2, float nz)
{
  StepConst c;
  c.omega = omega; c.aw1 = aw1; c.aw2 = aw2; c.negzero = nz;
  return collide4_masked(f, bits, c, fold);
}

// A lane's four cells as two packed pairs.  ROT = false: p = cells (0,1), q = cells (2,3); ROT = true: p = cells
// (1,2), q = cells (3,0).  Which pairing is free of register moves depends on where the operands come from: the
// six x-shifted populations of a row arrive as an aligned group of four plus one shuffled value, and that value
// lands in the register its dead neighbour vacated only under ONE of the two pairings (see step1 / step2 below).
// `any_blocked` and `fold` are warp-uniform.  Returns the lane's sum of |m|/rho in the reference's cell order.
template <bool ROT>
__device__ __forceinline__ float collide_quad_fast(f2 (&p)[9], f2 (&q)[9], const StepConst& c)
{
  float2 up, uq;
  collide_pairs(p, q, c.omega, c.negzero, up, uq);
  return ROT ? add(add(add(uq.y, up.x), up.y), uq.x) : add(add(add(up.x, up.y), uq.x), uq.y);
}
// bits = the four cells' obstacle bits in column order (bit j = cell j)
__host__ __device__ __forceinline__ unsigned pair_order(unsigned bits, bool rot)
{
  return rot ? (((bits >> 1) & 7u) | ((bits & 1u) << 3)) : bits;
}
template <bool ROT>
__device__ __forceinline__ float collide_quad_generic(f2 (&p)[9], f2 (&q)[9], unsigned bits, const StepConst& c, bool fold)
{
  float f[4][9];
#pragma unroll
  for (int k = 0; k < 9; k++) {
    const float2 a = unpack2(p[k]), b = unpack2(q[k]);
    if (ROT) { f[1][k] = a.x; f[2][k] = a.y; f[3][k] = b.x; f[0][k] = b.y; }
    else { f[0][k] = a.x; f[1][k] = a.y; f[2][k] = b.x; f[3][k] = b.y; }
  }
  const float u4 = collide4_generic(f, bits, c.omega, fold, c.aw1, c.aw2, c.negzero);
#pragma unroll
  for (int k = 0; k < 9; k++) {
    if (ROT) { p[k] = pack2(f[1][k], f[2][k]); q[k] = pack2(f[3][k], f[0][k]); }
    else { p[k] = pack2(f[0][k], f[1][k]); q[k] = pack2(f[2][k], f[3][k]); }
  }
  return u4;
}

// PEER   = a slab of a multi-GPU ring.  Padded rows: 0 / rows+1 are the halo rows next to the slab, rows+2 / rows+3
//          the second halo rows (the neighbours' rows one further away); obstacle words of the two adjacent
//          neighbour rows follow the slab's own.  The first step is ALSO computed for the neighbours' edge rows
//          (redundantly, from the two halo rows), so the second step of the slab's own edge rows needs no exchange
//          in the middle of a pass.  Once per pass a strip pushes what the neighbour's next pass pulls: planes
//          0,1,3 + 2,5,6 of its last row and 2,5,6 of the row below go north, 0,1,3 + 4,7,8 of its first row and
//          4,7,8 of the row above go south -- 9 (+2) plane rows per direction and pair of steps.
// SINGLE = one timestep only (the odd tail of a run on a ring): the first step's result is the output.
//
// Columns: a warp's 32 lanes hold 128 aligned columns, of which lanes 1..30 (120 columns) are the strip's own.  The
// first step is computed for all 128; its results for the outermost column on either side would need a 129th
// column and are simply wrong -- nobody reads them: the second step of the own columns needs first-step columns
// -1 .. 120 of the strip only.  (A shuffle without a source lane returns the lane's own value, so the unused cells
// stay finite.)
template <int HINT, bool PEER, bool SINGLE, bool DEEP = false>
__global__ void __launch_bounds__(DEEP ? 128 : 256, DEEP ? 3 : 2) steps2_strip(const StepArgs a, const FusedArgs g)
{
  extern __shared__ float4 fused_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  float4* const ring = fused_smem + (size_t)warp * fused_warp_float4(DEEP) + lane;   // this lane's cells of [3][6][32]
  const float4* const stage0 = ring + 3 * kRingSlot;                          // this lane's cells of [1 or 2][9][32]
  const unsigned stage0_s = (unsigned)__cvta_generic_to_shared(stage0);
  const float* __restrict__ src = a.src;
  float* __restrict__ dst = a.dst;
  const size_t P = a.plane;
  const int rows = a.row_last;                               // owned padded rows are 1..rows
  const int nx = a.nx;
  double acc1 = 0.0, acc2 = 0.0;

  const long nitems = (long)g.bands * g.strips;
  for (long item = (long)blockIdx.x * warps + warp; item < nitems; item += (long)gridDim.x * warps) {
    int band = (int)(item / g.strips);
    const int strip = (int)(item - (long)band * g.strips);
    // a ring slab does its two edge bands first: their halo rows are on the way to the neighbours (and published)
    // while the interior bands run, as the reference overlaps its halo exchange with the interior rows (326-366)
    if (PEER && g.bands > 2) band = (band == 0) ? 0 : (band == 1 ? g.bands - 1 : band - 1);

    // One work item.  EDGE (ring slabs only): the band touches the slab's first or last row -- halo rows, flag
    // handshake and pushes; every other band of a ring slab runs exactly the single-GPU code.
    auto run_item = [&](auto edge_tag) {
      constexpr bool EDGE = decltype(edge_tag)::value;
      // padded row of the 0-based row y in [-2, rows+1]: halo rows at a ring slab's edges, periodic otherwise
      auto prow = [&](const int y) -> int {
        if ((unsigned)y < (unsigned)rows) return y + 1;     // an owned row: the common case
        if (EDGE) return y == -1 ? 0 : (y == rows ? rows + 1 : (y == -2 ? rows + 2 : rows + 3));
        return (y < 0) ? y + rows + 1 : y - rows + 1;
      };
      const int yb = band * g.band_rows;                       // owned rows of the item, 0-based: [yb, ye)
      const int ye = (band == g.bands - 1) ? rows : yb + g.band_rows;
      // this lane's aligned group of four columns (periodic): the strip's 120 owned columns are lanes 1..30
      int gx = strip * kStripOut - 4 + 4 * lane;
      if (gx < 0) gx += nx;
      while (gx >= nx) gx -= nx;
      const bool owned = lane >= 1 && lane <= 30 && strip * kStripOut + 4 * (lane - 1) < nx;
      const uint32_t* const mask_x = a.mask + (gx >> 5);
      const int mask_shift = gx & 31;

      if (EDGE) {
        // the halo rows this item pulls (and the neighbour's halo columns it overwrites) are ordered by the flags
        if (yb == 0) strip_wait(a, a.wait_from_south, kWaitFromSouth, strip, g.strips, lane);
        if (ye == rows) strip_wait(a, a.wait_from_north, kWaitFromNorth, strip, g.strips, lane);
      }

      // ---- asynchronous copy of what the first step of row y pulls, into the staging row; returns the row's
      //      obstacle word of this lane's columns (its bits are >> mask_shift) ----
      // (q_y = the row the next copy is for; q_s, q_c, q_n = the padded rows it pulls from, rolled forward per copy;
      //  offsets inside a plane are 32-bit: the host refuses slabs of 2^32 cells)
      int q_y = SINGLE ? yb : yb - 1;
      int q_s = prow(q_y - 1), q_c = prow(q_y), q_n = prow(q_y + 1);
      // L2 prefetch kPrefetchRows rows ahead of the copy: lane k < 9 asks for plane k's 512-byte row piece with one
      // bulk-prefetch instruction, so that the copy into shared memory finds its data in L2 (with the copy only one
      // row ahead of the arithmetic, 19 % of all stall samples sat on the first read of the staging row).
      // Every (row, plane) piece of the band is requested once: plane k is pulled from row y + pf_dy of the row y
      // it is used for.
      const int pf_dy = (lane == 0 || lane == 1 || lane == 3) ? 0 : ((lane == 2 || lane == 5 || lane == 6) ? -1 : 1);
      int pf_x = strip * kStripOut - 4;                       // a contiguous piece even where the strip wraps around
      pf_x = max(0, min(pf_x, nx - 128));
      const float* const pf_base = src + (size_t)min(lane, 8) * P + pf_x;
      const int pf_last = SINGLE ? ye - 1 : ye;               // last row whose first step this item computes
      auto prefetch_row = [&](const int y) {                  // what the first step of row y pulls
        if (lane < 9 && y <= pf_last) {
          const float* p = pf_base + (unsigned)prow(y + pf_dy) * (unsigned)nx;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(512u) : "memory");
        }
      };
      for (int d = 1; d < g.prefetch_rows; d++) prefetch_row(q_y + d);
      unsigned stage_s = stage0_s;                            // where the next copy goes / the next first step reads
      const float4* stage = stage0;
      auto issue = [&]() -> unsigned {
        if (g.prefetch_rows > 0) prefetch_row(q_y + g.prefetch_rows);
        const float* pc = src + ((unsigned)q_c * (unsigned)nx + (unsigned)gx);
        const float* ps = src + ((unsigned)q_s * (unsigned)nx + (unsigned)gx);
        const float* pn = src + ((unsigned)q_n * (unsigned)nx + (unsigned)gx);
        cp_async16(stage_s + 0 * 512, pc + 0 * P);
        cp_async16(stage_s + 1 * 512, pc + 1 * P);
        cp_async16(stage_s + 2 * 512, ps + 2 * P);
        cp_async16(stage_s + 3 * 512, pc + 3 * P);
        cp_async16(stage_s + 4 * 512, pn + 4 * P);
        cp_async16(stage_s + 5 * 512, ps + 5 * P);
        cp_async16(stage_s + 6 * 512, ps + 6 * P);
        cp_async16(stage_s + 7 * 512, pn + 7 * P);
        cp_async16(stage_s + 8 * 512, pn + 8 * P);
        cp_async_commit();
        if (DEEP) stage_s ^= (stage0_s ^ (stage0_s + kStageSlot * 16));      // the other staging row next time
        // obstacle words: the slab's own rows, then (ring) the southern and the northern neighbour's adjacent row
        const int mrow = EDGE ? ((q_y < 0) ? rows : (q_y >= rows ? rows + 1 : q_y)) : q_c - 1;
        const unsigned word = __ldg(mask_x + (unsigned)mrow * (unsigned)a.mask_row_words);   // used a whole row later: no stall here
        q_y++;
        q_s = q_c; q_c = q_n; q_n = prow(q_y + 1);
        return word;
      };

      // ---- a finished (owned) row y (o[k] = plane k of the lane's four columns): to the destination buffer and, on
      //      a ring, into the neighbours' halo rows ----
      auto emit = [&](const int y, const f2 (&p)[9], const f2 (&q)[9]) {
        float* const d = dst + ((unsigned)(y + 1) * (unsigned)nx + (unsigned)gx);
#pragma unroll
        for (int k = 0; k < 9; k++) stg2<HINT>(d + k * P, p[k], q[k]);
        if (EDGE && (y <= 1 || y >= rows - 2)) {
          auto put = [&](float* base, size_t plane, int row, int k) {
            stg2<0>(base + k * plane + (size_t)row * nx + gx, p[k], q[k]);
          };
          if (y == 0) {                                        // the southern slab's halo row rows+1
            const int r = g.south_rows + 1;
            put(a.south_dst, a.south_plane, r, 0); put(a.south_dst, a.south_plane, r, 1); put(a.south_dst, a.south_plane, r, 3);
            put(a.south_dst, a.south_plane, r, 4); put(a.south_dst, a.south_plane, r, 7); put(a.south_dst, a.south_plane, r, 8);
          }
          if (y == 1) {                                        // its second halo row
            const int r = g.south_rows + 3;
            put(a.south_dst, a.south_plane, r, 4); put(a.south_dst, a.south_plane, r, 7); put(a.south_dst, a.south_plane, r, 8);
          }
          if (y == rows - 1) {                                 // the northern slab's halo row 0
            put(a.north_dst, a.north_plane, 0, 0); put(a.north_dst, a.north_plane, 0, 1); put(a.north_dst, a.north_plane, 0, 3);
            put(a.north_dst, a.north_plane, 0, 2); put(a.north_dst, a.north_plane, 0, 5); put(a.north_dst, a.north_plane, 0, 6);
          }
          if (y == rows - 2) {                                 // its second halo row (planes 3 and 7 ride along: if this
            const int r = g.north_rows + 2;                    // is the driven row, the neighbour's copy gets its own pre-pass)
            put(a.north_dst, a.north_plane, r, 2); put(a.north_dst, a.north_plane, r, 5); put(a.north_dst, a.north_plane, r, 6);
            put(a.north_dst, a.north_plane, r, 3); put(a.north_dst, a.north_plane, r, 7);
          }
        }
      };
      // after the rows a neighbour waits for have been pushed: publish this strip (whole warp)
      auto publish = [&](const int y) {
        if (EDGE && (y == 1 || y == rows - 1)) {
          if (y == 1) strip_signal(a, a.signal_south, strip, lane);
          if (y == rows - 1) strip_signal(a, a.signal_north, strip, lane);
        }
      };

      // ---- first step of row y out of the staging row (mword = its obstacle word).  The lane's four cells run as
      //      the pairs (1,2) and (3,0): a west-moving population's cells 1,2 are the first half of its aligned group
      //      and cell 3 + the value shuffled in for cell 0 fill the second half (mirrored for east-moving ones), so
      //      only the three unshifted planes need register moves.  Results keep that rotated order (columns 1,2,3,0
      //      of the group): six planes go into ring row `slot`, planes 4,7,8 come back in registers.  With `ahead`
      //      the next copy (one row ahead; DEEP: two) is issued as soon as the staging row has been read; its obstacle
      //      word is returned.
      auto step1 = [&](const int y, const unsigned mword, float4* const slot, const bool ahead, const bool more_copies_pending,
                       const bool count, f2 (&kp)[3], f2 (&kq)[3]) -> unsigned {
        // (DEEP: the copy for the row after this one is in flight as well and may stay so)
        if (DEEP && more_copies_pending) cp_async_wait_but_one(); else cp_async_wait_all();
        f2 lo[9], hi[9];                                         // lo = columns 0,1 of the lane's group, hi = columns 2,3
#pragma unroll
        for (int k = 0; k < 9; k++) lds2(stage + k * 32, lo[k], hi[k]);
        if (DEEP) stage = (stage == stage0) ? stage0 + kStageSlot : stage0;
        const float up1 = __shfl_up_sync(0xffffffffu, hi2(hi[1]), 1);
        const float up5 = __shfl_up_sync(0xffffffffu, hi2(hi[5]), 1);
        const float up8 = __shfl_up_sync(0xffffffffu, hi2(hi[8]), 1);
        const float dn3 = __shfl_down_sync(0xffffffffu, lo2(lo[3]), 1);
        const float dn6 = __shfl_down_sync(0xffffffffu, lo2(lo[6]), 1);
        const float dn7 = __shfl_down_sync(0xffffffffu, lo2(lo[7]), 1);
        // every lane has read its own staging cells (the shuffles consumed them): the next row's copy may start and
        // flies during both steps' arithmetic
        const unsigned word_next = ahead ? issue() : 0u;
        f2 p[9], q[9];                                         // p = cells (1,2), q = cells (3,0)
        p[0] = pack2(hi2(lo[0]), lo2(hi[0])); q[0] = pack2(hi2(hi[0]), lo2(lo[0]));
        p[2] = pack2(hi2(lo[2]), lo2(hi[2])); q[2] = pack2(hi2(hi[2]), lo2(lo[2]));
        p[4] = pack2(hi2(lo[4]), lo2(hi[4])); q[4] = pack2(hi2(hi[4]), lo2(lo[4]));
        p[1] = lo[1]; q[1] = pack2(lo2(hi[1]), up1);
        p[5] = lo[5]; q[5] = pack2(lo2(hi[5]), up5);
        p[8] = lo[8]; q[8] = pack2(lo2(hi[8]), up8);
        p[3] = hi[3]; q[3] = pack2(dn3, hi2(lo[3]));
        p[6] = hi[6]; q[6] = pack2(dn6, hi2(lo[6]));
        p[7] = hi[7]; q[7] = pack2(dn7, hi2(lo[7]));
        const unsigned bits = (mword >> mask_shift) & 0xFu;
        const bool fold = (SINGLE ? g.fold_last != 0 : true) && (y + 1 == g.accel_row);   // the driven row is an owned row
        const bool any_blocked = __any_sync(0xffffffffu, bits != 0u);
        // (each branch stores its own results: a join would pin 36 registers to common locations)
        auto finish = [&](const float u4) {
          acc1 += (double)(count ? u4 : 0.0f);
          if (SINGLE) {
            if (owned) {
              f2 op[9], oq[9];                                 // back to column order: cells (0,1), (2,3)
#pragma unroll
              for (int k = 0; k < 9; k++) { op[k] = pack2(hi2(q[k]), lo2(p[k])); oq[k] = pack2(hi2(p[k]), lo2(q[k])); }
              emit(y, op, oq);
            }
            return;
          }
          sts2(slot + 0 * 32, p[0], q[0]);
          sts2(slot + 1 * 32, p[1], q[1]);
          sts2(slot + 2 * 32, p[3], q[3]);
          sts2(slot + 3 * 32, p[2], q[2]);
          sts2(slot + 4 * 32, p[5], q[5]);
          sts2(slot + 5 * 32, p[6], q[6]);
          kp[0] = p[4]; kq[0] = q[4]; kp[1] = p[7]; kq[1] = q[7]; kp[2] = p[8]; kq[2] = q[8];
        };
        if (!any_blocked && !fold) finish(collide_quad_fast<true>(p, q, a.c));
        else finish(collide_quad_generic<true>(p, q, bits, a.c, fold));
        if (SINGLE) publish(y);
        return word_next;
      };

      // ---- second step of row y (0-based, owned): planes 2,5,6 of first-step row y-1 (ring row s_s), planes 0,1,3
      //      of row y (ring row s_c), planes 4,7,8 of row y+1 from the registers of the first step that just ran.
      //      All of them hold columns 1,2,3,0 of the lane's group; here the cells run as the pairs (0,1) and (2,3),
      //      which again leaves only the unshifted planes with register moves and yields the output in column order.
      auto step2 = [&](const int y, const unsigned mword, const float4* const s_s, const float4* const s_c,
                       const f2 (&kp)[3], const f2 (&kq)[3]) {
        f2 lo[9], hi[9];                                         // lo = columns 1,2 of the lane's group, hi = columns 3,0
        lds2(s_c + 0 * 32, lo[0], hi[0]); lds2(s_c + 1 * 32, lo[1], hi[1]); lds2(s_c + 2 * 32, lo[3], hi[3]);
        lds2(s_s + 3 * 32, lo[2], hi[2]); lds2(s_s + 4 * 32, lo[5], hi[5]); lds2(s_s + 5 * 32, lo[6], hi[6]);
        lo[4] = kp[0]; hi[4] = kq[0]; lo[7] = kp[1]; hi[7] = kq[1]; lo[8] = kp[2]; hi[8] = kq[2];
        // cell 0 of an east-moving population comes from the previous lane's column 3, cell 3 of a west-moving one
        // from the next lane's column 0
        const float up1 = __shfl_up_sync(0xffffffffu, lo2(hi[1]), 1);
        const float up5 = __shfl_up_sync(0xffffffffu, lo2(hi[5]), 1);
        const float up8 = __shfl_up_sync(0xffffffffu, lo2(hi[8]), 1);
        const float dn3 = __shfl_down_sync(0xffffffffu, hi2(hi[3]), 1);
        const float dn6 = __shfl_down_sync(0xffffffffu, hi2(hi[6]), 1);
        const float dn7 = __shfl_down_sync(0xffffffffu, hi2(hi[7]), 1);
        const unsigned bits = (mword >> mask_shift) & 0xFu;
        const bool any_blocked = __any_sync(0xffffffffu, bits != 0u);
        if (!owned) return;
        f2 p[9], q[9];                                         // p = cells (0,1), q = cells (2,3)
        p[0] = pack2(hi2(hi[0]), lo2(lo[0])); q[0] = pack2(hi2(lo[0]), lo2(hi[0]));
        p[2] = pack2(hi2(hi[2]), lo2(lo[2])); q[2] = pack2(hi2(lo[2]), lo2(hi[2]));
        p[4] = pack2(hi2(hi[4]), lo2(lo[4])); q[4] = pack2(hi2(lo[4]), lo2(hi[4]));
        p[1] = pack2(up1, hi2(hi[1])); q[1] = lo[1];
        p[5] = pack2(up5, hi2(hi[5])); q[5] = lo[5];
        p[8] = pack2(up8, hi2(hi[8])); q[8] = lo[8];
        p[3] = lo[3]; q[3] = pack2(lo2(hi[3]), dn3);
        p[6] = lo[6]; q[6] = pack2(lo2(hi[6]), dn6);
        p[7] = lo[7]; q[7] = pack2(lo2(hi[7]), dn7);
        const bool fold = g.fold_last && (y + 1 == g.accel_row);
        if (!any_blocked && !fold) {
          acc2 += (double)collide_quad_fast<false>(p, q, a.c);
          emit(y, p, q);
        } else {
          acc2 += (double)collide_quad_generic<false>(p, q, bits, a.c, fold);
          emit(y, p, q);
        }
      };

      f2 kp[3], kq[3];
      // Rows r0 .. r_last get their first step (yb-1 .. ye; SINGLE: yb .. ye-1); the copies run D rows ahead.
      // (w_c, w_n, w_nn = obstacle words of the rows of the next second step, the next first step and -- DEEP -- the
      // first step after that.)
      constexpr int D = DEEP ? 2 : 1;
      const int r0 = SINGLE ? yb : yb - 1, r_last = SINGLE ? ye - 1 : ye;
      unsigned w_c = 0u, w_n = issue(), w_nn = 0u;
      if (DEEP && r0 + 1 <= r_last) w_nn = issue();
      float4 *s_s = ring, *s_c = ring + kRingSlot, *s_n = ring + 2 * kRingSlot;
#pragma unroll 1
      for (int r = r0; r <= r_last; r++) {
        const unsigned w_new = step1(r, w_n, s_n, r + D <= r_last, DEEP && r + 1 <= r_last, owned && r >= yb && r < ye, kp, kq);
        if (!SINGLE) {
          // every owned row gets its second step as soon as the first step of the row above it has run
          if (r > yb) {
            step2(r - 1, w_c, s_s, s_c, kp, kq);
            publish(r - 1);
          }
          float4* const t = s_s; s_s = s_c; s_c = s_n; s_n = t;
        }
        w_c = w_n;
        if (DEEP) { w_n = w_nn; w_nn = w_new; } else w_n = w_new;
      }
    };
    if (PEER && (band == 0 || band == g.bands - 1)) run_item(std::true_type{});
    else run_item(std::false_type{});
  }

  block_sum_to(acc1, a.partials + blockIdx.x);
  if (!SINGLE) {
    __syncthreads();
    block_sum_to(acc2, a.partials + g.partial_stride + blockIdx.x);
  }
  if (PEER) peer_advance_epoch(a);
}

// ---------------------------------------------------------------------------------------
// Kernel 3 ("resident"): many timesteps in ONE cooperative launch for grids that are launch-latency
// bound (the shipped 128..1024-wide decks: a step is a few microseconds of work).  The same pass as
// step_vec4 runs `steps` times with a grid-wide barrier in between and the two buffers swapping roles;
// loads bypass L1 (ld.global.cg) so every step sees what other SMs stored in the previous one.
// ---------------------------------------------------------------------------------------
struct ResidentArgs {
  float* buf0;       // holds the current state at launch
  float* buf1;
  int steps;         // timesteps in this launch (<= the partials capacity)
  int fold_last;     // whether the last of them folds the next step's body force in
  int partial_stride;  // doubles between consecutive steps' partials (= grid size)
};

template <int MIN_CTAS>
__global__ void __launch_bounds__(256, MIN_CTAS) steps_resident(const StepArgs a, const ResidentArgs r)
{
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  for (int t = 0; t < r.steps; t++) {
    const float* src = (t & 1) ? r.buf1 : r.buf0;
    float* dst = (t & 1) ? r.buf0 : r.buf1;
    const int accel_row = (t + 1 < r.steps || r.fold_last) ? a.accel_row : -1;
    const double acc = vec4_pass<false, 3>(a, src, dst, accel_row);
    block_sum_to(acc, a.partials + (size_t)t * r.partial_stride + blockIdx.x);
    grid.sync();
  }
}

// ---------------------------------------------------------------------------------------
// Kernel 6 ("cluster"): the whole grid RESIDENT IN SHARED MEMORY of one thread-block cluster, many timesteps per
// launch (first option).  For the shipped 128-wide decks a timestep is ~16 K cell updates: far too
// little work to hide a kernel launch (2.7 us per step from CUDA graphs) or a grid-wide barrier through L2.  Here
// kClusterCtas CTAs of one cluster (16: the non-portable maximum) each keep ny/16 rows of all nine planes in
// their shared memory, twice (ping-pong); a step is one pass of every thread over its cell(s) -- nine loads from
// shared memory, the three planes of the row above / below an edge row straight out of the NEIGHBOUR CTA's shared
// memory (distributed shared memory: the halo exchange of d2q9-bgk.c:326-364 becomes a remote load), collide(),
// nine stores into the other buffer -- and ONE hardware cluster barrier.  Global memory is touched at the start
// and the end of the launch and for one double per warp and step (the Sigma |m|/rho partials).
// Per-cell arithmetic: collide()/accelerate() as everywhere, so the populations are bit-identical.
// ---------------------------------------------------------------------------------------
constexpr int kClusterCtas = 16;

struct ClusterArgs {
  const float* in;       // plane 0 of the buffer that holds the state at launch
  float* out;            // plane 0 of the buffer that receives the state after `steps` steps
  size_t plane;          // floats between planes of in / out
  const uint32_t* mask;
  int mask_row_words;
  int nx, rows_per_cta;  // ny = kClusterCtas * rows_per_cta
  int steps, fold_last;
  int accel_row;         // 0-based global row ny-2
  StepConst c;
  double* partials;      // [steps][partial_stride]: one double per warp of the cluster
  int partial_stride;
};

__global__ void __launch_bounds__(1024, 1) steps_cluster(const ClusterArgs a)
{
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float cluster_smem[];
  const int rank = (int)cluster.block_rank();
  const int nx = a.nx, rows = a.rows_per_cta, ncell = rows * nx;   // this CTA's cells
  float* const buf0 = cluster_smem;                        // [9][ncell]
  float* const buf1 = cluster_smem + 9 * (size_t)ncell;
  const int first_row = rank * rows;                       // 0-based global row of local row 0
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;

  for (int c = threadIdx.x; c < ncell; c += blockDim.x) {
    const size_t g = (size_t)(first_row + 1) * nx + c;     // padded row = global row + 1; rows are contiguous
#pragma unroll
    for (int k = 0; k < 9; k++) buf0[k * ncell + c] = __ldg(a.in + k * a.plane + g);
  }
  // the rows below local row 0 / above local row rows-1 live in the ring neighbours' shared memory (periodic in y)
  const int r_s = (rank + kClusterCtas - 1) % kClusterCtas, r_n = (rank + 1) % kClusterCtas;
  const float* const south0 = cluster.map_shared_rank(buf0, r_s);
  const float* const south1 = cluster.map_shared_rank(buf1, r_s);
  const float* const north0 = cluster.map_shared_rank(buf0, r_n);
  const float* const north1 = cluster.map_shared_rank(buf1, r_n);
  cluster.sync();

  for (int t = 0; t < a.steps; t++) {
    const float* const cur = (t & 1) ? buf1 : buf0;
    float* const nxt = (t & 1) ? buf0 : buf1;
    const float* const south = (t & 1) ? south1 : south0;
    const float* const north = (t & 1) ? north1 : north0;
    const bool fold = (t + 1 < a.steps) || a.fold_last;
    double acc = 0.0;
    for (int c = threadIdx.x; c < ncell; c += blockDim.x) {
      const int row = c / nx, x = c - row * nx;
      const int xw = (x == 0) ? nx - 1 : x - 1, xe = (x + 1 == nx) ? 0 : x + 1;
      const float* const below = (row > 0) ? cur + (row - 1) * nx : south + (rows - 1) * nx;
      const float* const above = (row + 1 < rows) ? cur + (row + 1) * nx : north;
      const float* const here = cur + row * nx;
      float f[9];
      f[0] = here[0 * ncell + x];
      f[1] = here[1 * ncell + xw];
      f[2] = below[2 * ncell + x];
      f[3] = here[3 * ncell + xe];
      f[4] = above[4 * ncell + x];
      f[5] = below[5 * ncell + xw];
      f[6] = below[6 * ncell + xe];
      f[7] = above[7 * ncell + xe];
      f[8] = above[8 * ncell + xw];
      const int grow = first_row + row;
      const bool blocked = (__ldg(a.mask + (size_t)grow * a.mask_row_words + (x >> 5)) >> (x & 31)) & 1u;
      acc += (double)collide(f, blocked, a.c.omega);
      if (fold && grow == a.accel_row) accelerate(f, blocked, a.c.aw1, a.c.aw2);
#pragma unroll
      for (int k = 0; k < 9; k++) nxt[k * ncell + c] = f[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) a.partials[(size_t)t * a.partial_stride + rank * warps + warp] = acc;
    cluster.sync();                                        // everybody's new rows are visible cluster-wide
  }

  const float* const fin = (a.steps & 1) ? buf1 : buf0;
  for (int c = threadIdx.x; c < ncell; c += blockDim.x) {
    const size_t g = (size_t)(first_row + 1) * nx + c;
#pragma unroll
    for (int k = 0; k < 9; k++) a.out[k * a.plane + g] = fin[k * ncell + c];
  }
}

// ---------------------------------------------------------------------------------------
// Kernel 1 ("scalar"): one cell per thread, any nx.  The generic path (nx not a multiple of
// 4) and the baseline the vec4 kernel is measured against.
// ---------------------------------------------------------------------------------------
template <bool PEER>
__global__ void __launch_bounds__(256) step_scalar(const StepArgs a)
{
  if (PEER) peer_wait(a);

  const float* __restrict__ src = a.src;
  float* __restrict__ dst = a.dst;
  const size_t P = a.plane;
  const long ncell = (long)a.row_count * a.nx;
  double acc = 0.0;

  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < ncell; i += (long)gridDim.x * blockDim.x) {
    const int ri = (int)(i / a.nx);
    const int x = (int)(i - (long)ri * a.nx);
    const int row = a.row_begin + ri * a.row_stride;
    const int rs = (row == a.row_first) ? a.south_of_first : row - 1;
    const int rn = (row == a.row_last) ? a.north_of_last : row + 1;
    const int xw = (x == 0) ? a.nx - 1 : x - 1;
    const int xe = (x + 1 >= a.nx) ? 0 : x + 1;
    const size_t o_c = (size_t)row * a.nx, o_s = (size_t)rs * a.nx, o_n = (size_t)rn * a.nx;

    float f[9];
    f[0] = __ldg(src + 0 * P + o_c + x);
    f[1] = __ldg(src + 1 * P + o_c + xw);
    f[2] = __ldg(src + 2 * P + o_s + x);
    f[3] = __ldg(src + 3 * P + o_c + xe);
    f[4] = __ldg(src + 4 * P + o_n + x);
    f[5] = __ldg(src + 5 * P + o_s + xw);
    f[6] = __ldg(src + 6 * P + o_s + xe);
    f[7] = __ldg(src + 7 * P + o_n + xe);
    f[8] = __ldg(src + 8 * P + o_n + xw);
    const uint32_t mw = __ldg(a.mask + (size_t)(row - 1) * a.mask_row_words + (x >> 5));
    const bool blocked = (mw >> (x & 31)) & 1u;

    acc += (double)collide(f, blocked, a.c.omega);
    if (row == a.accel_row) accelerate(f, blocked, a.c.aw1, a.c.aw2);

#pragma unroll
    for (int k = 0; k < 9; k++) dst[k * P + o_c + x] = f[k];

    if (PEER) {
      if (row == a.row_last) {
        const size_t o = (size_t)a.north_row * a.nx + x;
        a.north_dst[2 * a.north_plane + o] = f[2];
        a.north_dst[5 * a.north_plane + o] = f[5];
        a.north_dst[6 * a.north_plane + o] = f[6];
      }
      if (row == a.row_first) {
        const size_t o = (size_t)a.south_row * a.nx + x;
        a.south_dst[4 * a.south_plane + o] = f[4];
        a.south_dst[7 * a.south_plane + o] = f[7];
        a.south_dst[8 * a.south_plane + o] = f[8];
      }
    }
  }

  block_sum_to(acc, a.partials + blockIdx.x);
  if (PEER) peer_signal(a);
}

// ---------------------------------------------------------------------------------------
// Small kernels around the step.
// ---------------------------------------------------------------------------------------

// accelerate_flow (d2q9-bgk.c:442-478) as a pre-pass on one row: needed once per run, before
// the first step; all later steps get their force folded into the previous step's store.
// `l` says where the row's populations live (in place, or decoded from the in-place kernel's L1 layout).
__global__ void accelerate_row(float* buf, Layout l, const uint32_t* mask_row, int row, float aw1, float aw2)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= l.nx) return;
  const bool blocked = (mask_row[x >> 5] >> (x & 31)) & 1u;
  float* p1 = locate(l, buf, 1, x, row); float* p3 = locate(l, buf, 3, x, row);
  float* p5 = locate(l, buf, 5, x, row); float* p6 = locate(l, buf, 6, x, row);
  float* p7 = locate(l, buf, 7, x, row); float* p8 = locate(l, buf, 8, x, row);
  const float f3 = *p3, f6 = *p6, f7 = *p7;
  if (!blocked && sub(f3, aw1) > 0.0f && sub(f6, aw2) > 0.0f && sub(f7, aw2) > 0.0f) {
    *p1 = add(*p1, aw1);
    *p5 = add(*p5, aw2);
    *p8 = add(*p8, aw2);
    *p3 = sub(f3, aw1);
    *p6 = sub(f6, aw2);
    *p7 = sub(f7, aw2);
  }
}

// The same pre-pass for the copy of the driven row that the slab NORTH of its owner keeps in its second halo row
// (kernel 5 on a ring).  That copy is pushed by the owner during the last pass of the previous run, and nothing but
// the strip flags orders that push against this kernel (the owner may be another process on another GPU whose
// stream is still draining): every thread first waits until the strip that covers its column has been published
// (flag >= this slab's epoch, exactly the condition the next pass's edge items wait for).
__global__ void accelerate_halo_row(float* buf, Layout l, const uint32_t* mask_row, int row, float aw1, float aw2,
                                    const StepArgs a, int strips)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= l.nx) return;
  const int c = min(x / kStripOut, strips - 1);
  const unsigned need = *reinterpret_cast<volatile unsigned*>(a.epoch);
  spin_until(a.wait_from_south + c, need, a, ((unsigned)c << 8) | kWaitFromSouth);
  const bool blocked = (mask_row[x >> 5] >> (x & 31)) & 1u;
  float* p = buf + (size_t)row * l.nx + x;          // always the canonical layout (ping-pong handles only)
  const size_t P = l.plane;
  const float f3 = __ldcg(p + 3 * P), f6 = __ldcg(p + 6 * P), f7 = __ldcg(p + 7 * P);
  if (!blocked && sub(f3, aw1) > 0.0f && sub(f6, aw2) > 0.0f && sub(f7, aw2) > 0.0f) {
    p[1 * P] = add(__ldcg(p + 1 * P), aw1);
    p[5 * P] = add(__ldcg(p + 5 * P), aw2);
    p[8 * P] = add(__ldcg(p + 8 * P), aw2);
    p[3 * P] = sub(f3, aw1);
    p[6 * P] = sub(f6, aw2);
    p[7 * P] = sub(f7, aw2);
  }
}

// av[base + t] = (float)(sum of the step's CTA partials, fixed order, fp64) * free_cells_inv
// for t in [0, steps): d2q9-bgk.c:367.  One warp per step.  `cursor` (device) holds base and
// is advanced by `steps` so that graph replays append.
__global__ void reduce_partials(const double* partials, int per_step, int steps, float free_cells_inv,
                                float* av, unsigned* cursor)
{
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned base = *cursor;
  if (warp < steps) {
    const double* p = partials + (size_t)warp * per_step;
    double s = 0.0;
    for (int i = lane; i < per_step; i += 32) s += p[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) av[base + warp] = mul((float)s, free_cells_inv);
  }
}
__global__ void advance_cursor(unsigned* cursor, unsigned by) { *cursor += by; }

// int-per-cell obstacle rows (the reference's layout, d2q9-bgk.c:875) -> 1 bit per cell.  One warp packs
// 32 words of one row at a time: coalesced 128 B reads, one ballot per word.
__global__ void pack_mask(const int* obstacles, int nx, int rows, int row_words, uint32_t* mask,
                          unsigned long long* blocked_cells)
{
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int groups = (row_words + 31) / 32;                 // 32-word groups per row
  if (warp >= (long)rows * groups) return;
  const int r = (int)(warp / groups);
  const int w0 = (int)(warp - (long)r * groups) * 32;
  const int* row = obstacles + (size_t)r * nx;
  uint32_t mine = 0;
  for (int i = 0; i < 32; i++) {
    const int x = (w0 + i) * 32 + lane;
    const unsigned word = __ballot_sync(0xffffffffu, x < nx && row[x] != 0);
    if (i == lane) mine = word;
  }
  if (w0 + lane < row_words) mask[(size_t)r * row_words + w0 + lane] = mine;
  // number of blocked cells (for free_cells_inv, d2q9-bgk.c:945-950): one atomic per warp
  unsigned n = __popc(mine);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if (lane == 0 && n && blocked_cells) atomicAdd(blocked_cells, (unsigned long long)n);
}

// The same from one byte per cell (LBM_B200_OBST_UINT8): a quarter of the upload.
__global__ void pack_mask_u8(const unsigned char* obstacles, int nx, int rows, int row_words, uint32_t* mask,
                             unsigned long long* blocked_cells)
{
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int groups = (row_words + 31) / 32;
  if (warp >= (long)rows * groups) return;
  const int r = (int)(warp / groups);
  const int w0 = (int)(warp - (long)r * groups) * 32;
  const unsigned char* row = obstacles + (size_t)r * nx;
  uint32_t mine = 0;
  for (int i = 0; i < 32; i++) {
    const int x = (w0 + i) * 32 + lane;
    const unsigned word = __ballot_sync(0xffffffffu, x < nx && row[x] != 0);
    if (i == lane) mine = word;
  }
  if (w0 + lane < row_words) mask[(size_t)r * row_words + w0 + lane] = mine;
  unsigned n = __popc(mine);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if (lane == 0 && n && blocked_cells) atomicAdd(blocked_cells, (unsigned long long)n);
}

// Already bit-packed rows (LBM_B200_OBST_BITS, the device layout): clear the padding bits past nx in the last word
// of every row and count the blocked cells.
__global__ void adopt_mask_bits(uint32_t* mask, int nx, int rows, int row_words, unsigned long long* blocked_cells)
{
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned n = 0;
  if (i < (long)rows * row_words) {
    const int w = (int)(i % row_words);
    uint32_t v = mask[i];
    const int valid = nx - w * 32;                            // cells this word covers
    if (valid < 32) { v &= (1u << valid) - 1u; mask[i] = v; }
    n = __popc(v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if ((threadIdx.x & 31) == 0 && n && blocked_cells) atomicAdd(blocked_cells, (unsigned long long)n);
}

// ---------------------------------------------------------------------------------------
// Self-test of the arithmetic restatements (lbm_b200_selftest): counts[0] = floats in the fast range whose rcp_fast
// differs from __frcp_rn, counts[1] = the same for sqrt_fast / __fsqrt_rn (every bit pattern in [kFastLo, kFastHi] is
// tried), counts[2] = packed add/sub/mul results that differ from the scalar intrinsics over `pairs` pseudo-random
// operand pairs drawn from ALL bit patterns (NaN results compare equal to NaN results).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ bool same_float(float a, float b)
{
  return __float_as_uint(a) == __float_as_uint(b) || (a != a && b != b);
}
__global__ void selftest_math(unsigned long long* counts, unsigned long long pairs, float nz)
{
  const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long nthreads = (unsigned long long)gridDim.x * blockDim.x;
  unsigned bad_rcp = 0, bad_sqrt = 0, bad_packed = 0;
  for (unsigned long long b = kFastLo + tid; b <= kFastHi; b += nthreads) {
    const float x = __uint_as_float((unsigned)b);
    if (__float_as_uint(rcp_fast(x)) != __float_as_uint(__frcp_rn(x))) bad_rcp++;
    if (__float_as_uint(sqrt_fast(x)) != __float_as_uint(__fsqrt_rn(x))) bad_sqrt++;
  }
  for (unsigned long long i = tid; i < pairs; i += nthreads) {
    // four 32-bit patterns from a counter hash (splitmix64); every exponent, both signs, NaNs and infinities occur
    unsigned long long z = i * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull, w[2];
    for (int j = 0; j < 2; j++) {
      z += 0x9E3779B97F4A7C15ull;
      unsigned long long t = z;
      t = (t ^ (t >> 30)) * 0xBF58476D1CE4E5B9ull;
      t = (t ^ (t >> 27)) * 0x94D049BB133111EBull;
      w[j] = t ^ (t >> 31);
    }
    const float a0 = __uint_as_float((unsigned)w[0]), a1 = __uint_as_float((unsigned)(w[0] >> 32));
    const float b0 = __uint_as_float((unsigned)w[1]), b1 = __uint_as_float((unsigned)(w[1] >> 32));
    const f2 a = pack2(a0, a1), b = pack2(b0, b1);
    const float2 s = unpack2(add2(a, b)), d = unpack2(sub2(a, b)), m = unpack2(mul2(a, b, nz)), ms = unpack2(mul2(a, b0, nz));
    // (the multiply followed by an add: the pattern ptxas would contract into one FFMA2 if it saw a plain product)
    const float2 ma = unpack2(add2(mul2(a, b, nz), a));
    bad_packed += !same_float(s.x, add(a0, b0)) + !same_float(s.y, add(a1, b1)) + !same_float(d.x, sub(a0, b0)) +
                  !same_float(d.y, sub(a1, b1)) + !same_float(m.x, mul(a0, b0)) + !same_float(m.y, mul(a1, b1)) +
                  !same_float(ms.x, mul(a0, b0)) + !same_float(ms.y, mul(a1, b0)) +
                  !same_float(ma.x, add(mul(a0, b0), a0)) + !same_float(ma.y, add(mul(a1, b1), a1));
  }
  if (bad_rcp) atomicAdd(counts + 0, (unsigned long long)bad_rcp);
  if (bad_sqrt) atomicAdd(counts + 1, (unsigned long long)bad_sqrt);
  if (bad_packed) atomicAdd(counts + 2, (unsigned long long)bad_packed);
}

// uniform initial state, every padded row (d2q9-bgk.c:880-902)
__global__ void fill_planes(float* buf, size_t plane, float w0, float w1, float w2)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= plane) return;
  buf[i] = w0;
#pragma unroll
  for (int k = 1; k < 5; k++) buf[k * plane + i] = w1;
#pragma unroll
  for (int k = 5; k < 9; k++) buf[k * plane + i] = w2;
}

// planes -> array of structs for `ncell` cells starting at padded row `row0`, and back
__global__ void soa_to_aos(const float* buf, Layout l, int row0, size_t ncell, float* aos)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncell * 9) return;
  const size_t cell = i / 9;
  const int k = (int)(i - cell * 9);
  const size_t r = cell / l.nx;
  aos[i] = *locate(l, buf, k, (int)(cell - r * l.nx), row0 + (int)r);
}
// always writes the canonical layout (l.odd is ignored)
__global__ void aos_to_soa(const float* aos, size_t plane, size_t first, size_t ncell, float* buf)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncell * 9) return;
  const size_t cell = i / 9;
  const int k = (int)(i - cell * 9);
  buf[k * plane + first + cell] = aos[i];
}

// Macroscopic fields exactly as write_values computes them (d2q9-bgk.c:1076-1111) for `ncell` cells
// starting at padded row `row0`; out holds four planes of `ncell` floats: u_x, u_y, |u|, pressure.
__global__ void final_state(const float* buf, Layout l, int row0, const uint32_t* mask, int mask_row_words,
                            size_t ncell, float density, float* out)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncell) return;
  const size_t r = i / l.nx;
  const int x = (int)(i - r * l.nx);
  const int row = row0 + (int)r;
  const bool blocked = (mask[(size_t)(row - 1) * mask_row_words + (x >> 5)] >> (x & 31)) & 1u;
  constexpr float c_sq = 1.0f / 3.0f;
  float ux = 0.f, uy = 0.f, u = 0.f, pr = mul(density, c_sq);
  if (!blocked) {
    float f[9];
#pragma unroll
    for (int k = 0; k < 9; k++) f[k] = *locate(l, buf, k, x, row);
    float rho = add(0.0f, f[0]);
#pragma unroll
    for (int k = 1; k < 9; k++) rho = add(rho, f[k]);
    ux = __fdiv_rn(sub(add(add(f[1], f[5]), f[8]), add(add(f[3], f[6]), f[7])), rho);
    uy = __fdiv_rn(sub(add(add(f[2], f[5]), f[6]), add(add(f[4], f[7]), f[8])), rho);
    u = __fsqrt_rn(add(mul(ux, ux), mul(uy, uy)));
    pr = mul(rho, c_sq);
  }
  out[i] = ux; out[ncell + i] = uy; out[2 * ncell + i] = u; out[3 * ncell + i] = pr;
}

}  // namespace lbm
