// lbm_stepsk.cuh -- kernel 7 ("fusedk"): K timesteps per pass over HBM, K = 1..4.
//
// Kernel 5 (steps2_strip, lbm_kernels.cuh) halves the DRAM traffic of a timestep and ends up at 0.91 of the HBM copy
// bandwidth again (profiles/r02_summary.md): the only way further is to touch DRAM even less often.  This kernel is
// the same scheme with a chain of K steps: a warp walks down a strip of 128 aligned columns (120 owned, 4 halo
// columns per side -- exactly what FOUR steps consume, one column per side and step), and row r of the walk runs
//   step 1 of row r        out of the staging row that cp.async filled D = 1 or 2 rows ahead,
//   step 2 of row r-1      out of ring 1 (the step-1 results the second step still needs, in shared memory),
//   ...
//   step K of row r-K+1    out of ring K-1, stored to the other buffer (lanes 1..30).
// Per cell and K steps 36 B are read and 36 B written (+6.7 % redundant columns, +2(K-1)/band_rows redundant rows).
// A step's results go into its ring as six planes (0,1,3,2,5,6); planes 4,7,8 of a row are consumed by the next
// step of the row below in the same iteration and stay in registers.  Odd steps take their operands in column
// order and leave them rotated by one column (cells 1,2,3,0 of a lane's group), even steps take rotated operands
// and leave column order (see collide_quad in lbm_kernels.cuh): a chain needs no register shuffling in between.
// The arithmetic per cell and step is collide()/accelerate() of lbm_cell.cuh: bit-identical to K launches of
// step_vec4 and to the oracle.
//
// Launch shape: ONE warp per CTA, as many CTAs per SM as its shared memory holds (K = 3, D = 1: eleven) -- an SM slot
// is free again the moment a work item ends; the per-step sums leave the kernel as one fp64 partial per warp.  K = 3
// with one staging row is the automatic choice wherever a fused kernel runs (183-190 GLUPS at 16384^2 on one B200,
// DRAM at 0.72 of the copy peak, bound by the FMA pipe; profiles/r02_fused2.md, profiles/r02_summary.md).
//
// Ring slabs (PEER): every slab keeps kHalo = 4 halo rows per side, all nine planes (padded rows: 0 and rows+1 next
// to the slab, rows+2(d-1) / rows+2(d-1)+1 the southern / northern neighbour's row d rows away, d = 2..4).  The
// steps before the last are ALSO computed for the neighbours' rows they need (from the halo rows), so a pass needs no
// exchange in the middle.  Once per pass the items of the two edge bands copy their final rows 0..3 and rows-4..
// rows-1 into the neighbours' halo rows over NVLink and publish one flag word per strip and direction, after having
// waited for the neighbours' words of strips c-1, c, c+1 -- the protocol of kernel 5, with a deeper halo.
#pragma once
#include "lbm_kernels.cuh"

namespace lbm {

constexpr int kHalo = 4;          // halo rows per side a ring slab keeps for this kernel (= the largest K)

struct StepsKArgs {
  int band_rows;        // rows per work item
  int bands, strips;
  int accel_y;          // 0-based row of global row ny-2 as this slab addresses it (-2 on the slab north of its owner
                        // on a ring), or a value no row has
  int fold_last;        // whether the pass's last step folds the following step's force in
  int partial_stride;   // doubles between consecutive steps' partials
  int south_rows, north_rows;   // owned rows of the ring neighbours (PEER)
};

// Shared memory of one warp (float4 per lane x 32 lanes): K-1 rings and D staging rows of nine planes.  A ring holds
// what the next step still has to read of a step's results: planes 0,1,3 of two rows (the row written in this
// iteration and the one written in the previous iteration, which the next step reads now) and planes 2,5,6 of three
// rows (the next step reads them one iteration later still, as its row's southern neighbour).
constexpr int kRingA = 3 * 32;                                     // one row of planes 0,1,3
constexpr int kRingB = 3 * 32;                                     // one row of planes 2,5,6
constexpr int kRingK = 2 * kRingA + 3 * kRingB;                    // one ring: 7680 bytes
__host__ __device__ constexpr int stepsk_warp_float4(int k, int d) { return (k - 1) * kRingK + d * kStageSlot; }
__host__ __device__ constexpr size_t stepsk_warp_bytes(int k, int d) { return (size_t)stepsk_warp_float4(k, d) * 16; }
// resident warps per SM: what 227 KB of shared memory hold in CTAs of one warp (a CTA costs 1 KB more), at most 12 --
// the register file then allows 170 registers per thread
__host__ __device__ constexpr int stepsk_max_warps(int k, int d)
{
  return (int)(232448 / (stepsk_warp_bytes(k, d) + 1024)) > 12 ? 12 : (int)(232448 / (stepsk_warp_bytes(k, d) + 1024));
}

// The ring protocol's two calls, out of line: they run twice per EDGE item, and every instruction they would add to the
// walk's body costs all items (unrolling the nine plane copies of a push inside the body: -1.7 % on every slab).
__device__ __noinline__ void strip_wait_cold(const StepArgs& a, const unsigned* flags, unsigned dir, int strip, int strips, int lane)
{
  strip_wait(a, flags, dir, strip, strips, lane);
}
__device__ __noinline__ void strip_signal_cold(const StepArgs& a, unsigned* flags, int strip, int lane)
{
  strip_signal(a, flags, strip, lane);
}

// One work item: the walk of one warp down one strip of one band.  `edge` (ring slabs only, warp-uniform): the band
// touches the slab's first or last row -- halo rows, flag handshake and pushes.  It is a RUN-TIME flag on one copy of
// the walk: with the edge form as a second copy of the code (inlined, or out of line) the items of a pass's first
// wave -- edge and interior items side by side on every SM -- fought over the instruction cache: same instruction
// count, 12 % longer passes on 2048-row slabs (3.6 waves per pass), 2.8 % on 16384-row ones (ncu, profiles/
// r02_fused2.md).  acc[s] += the item's share of step s+1's Sigma |m|/rho.
template <int K, int D, int HINT, bool PEER>
__device__ __forceinline__ void stepsk_item(const StepArgs& a, const StepsKArgs& g, const int band, const int strip, const bool edge,
                                            float4* const ring0, const float4* const stage0, const int lane, double (&acc)[K])
{
  const bool EDGE = PEER && edge;
  const unsigned stage0_s = (unsigned)__cvta_generic_to_shared(stage0);
  // (everything the walk needs of the two argument structs, by value)
  const float* __restrict__ src = a.src;
  float* __restrict__ dst = a.dst;
  const size_t P = a.plane;
  const int rows = a.row_last;                               // owned padded rows are 1..rows
  const int nx = a.nx;
  const StepConst c = a.c;
  const int mask_row_words = a.mask_row_words;
  const int band_rows = g.band_rows, bands = g.bands, strips = g.strips, accel_y = g.accel_y, fold_last = g.fold_last;
  const int south_rows = g.south_rows, north_rows = g.north_rows;
  float* const south_dst = a.south_dst;
  float* const north_dst = a.north_dst;
  const size_t south_plane = a.south_plane, north_plane = a.north_plane;

  // padded row of the 0-based row y in [-kHalo, rows+kHalo): halo rows at a ring slab's edges, periodic otherwise
  auto prow = [&](const int y) -> int {
    if ((unsigned)y < (unsigned)rows) return y + 1;     // an owned row: the common case
    if (EDGE) {
      if (y < 0) return y == -1 ? 0 : rows - 2 * (y + 1);            // d = -y: rows + 2(d-1)
      return y == rows ? rows + 1 : rows + 2 * (y - rows) + 1;        // d = y-rows+1: rows + 2(d-1) + 1
    }
    return (y < 0) ? y + rows + 1 : y - rows + 1;
  };
  // row of obstacle words: the slab's own rows, then (ring) the neighbours' rows in the order of the halo rows
  auto mrow = [&](const int y) -> int {
    if ((unsigned)y < (unsigned)rows) return y;
    if (EDGE) return (y < 0) ? rows - 2 * (y + 1) : rows + 2 * (y - rows) + 1;
    return (y < 0) ? y + rows : y - rows;
  };
  // the row a 0-based y stands for when it is compared with accel_y (periodic on one GPU)
  auto ident = [&](const int y) -> int {
    if (EDGE || (unsigned)y < (unsigned)rows) return y;
    return (y < 0) ? y + rows : y - rows;
  };
  const int yb = band * band_rows;                       // owned rows of the item, 0-based: [yb, ye)
  const int ye = (band == bands - 1) ? rows : yb + band_rows;
  // this lane's aligned group of four columns (periodic): the strip's 120 owned columns are lanes 1..30
  int gx = strip * kStripOut - 4 + 4 * lane;
  if (gx < 0) gx += nx;
  while (gx >= nx) gx -= nx;
  const bool owned = lane >= 1 && lane <= 30 && strip * kStripOut + 4 * (lane - 1) < nx;
  const uint32_t* const mask_x = a.mask + (gx >> 5);
  const int mask_shift = gx & 31;

  if (EDGE) {
    // the halo rows this item pulls (and the neighbour's halo columns it overwrites) are ordered by the flags
    if (yb == 0) strip_wait_cold(a, a.wait_from_south, kWaitFromSouth, strip, strips, lane);
    if (ye == rows) strip_wait_cold(a, a.wait_from_north, kWaitFromNorth, strip, strips, lane);
  }

  // ---- asynchronous copy of what the first step of row q_y pulls, into a staging row; returns the row's
  //      obstacle word of this lane's columns (its bits are >> mask_shift) ----
  int q_y = yb - (K - 1);
  int q_s = prow(q_y - 1), q_c = prow(q_y), q_n = prow(q_y + 1);
  unsigned stage_s = stage0_s;                            // where the next copy goes
  const float4* stage = stage0;                           // what the next first step reads
  auto issue = [&]() -> unsigned {
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
    if (D == 2) stage_s ^= (stage0_s ^ (stage0_s + kStageSlot * 16));   // the other staging row next time
    const unsigned word = __ldg(mask_x + (unsigned)mrow(q_y) * (unsigned)mask_row_words);   // used rows later
    q_y++;
    q_s = q_c; q_c = q_n; q_n = prow(q_y + 1);
    return word;
  };

  // ---- a finished (owned) row y in column order: to the destination buffer ----
  auto emit = [&](const int y, const f2 (&p)[9], const f2 (&q)[9]) {
    float* const d = dst + ((unsigned)(y + 1) * (unsigned)nx + (unsigned)gx);
#pragma unroll
    for (int k = 0; k < 9; k++) stg2<HINT>(d + k * P, p[k], q[k]);
  };
  // ---- ring slabs: once the kHalo rows next to an edge are in the destination buffer, this lane copies its columns
  //      of them (all nine planes) into the neighbour's halo rows over NVLink, and the warp publishes the strip.  A
  //      rolled loop of loads (the lane's own stores of a moment ago) and stores: nothing of it sits in the walk's
  //      straight-line code, whose instruction-cache footprint decides the speed of every item (see stepsk_item). ----
  auto push_rows = [&](const int y0, const bool south) {
    if (!owned) return;
    float* const base = south ? south_dst : north_dst;
    const size_t plane = south ? south_plane : north_plane;
#pragma unroll 1
    for (int y = y0; y < y0 + kHalo; y++) {
      // row y of this slab is row south_rows + y of the southern slab / row y - rows of the northern one
      const int r = south ? south_rows + 2 * y + 1 : (rows - y == 1 ? 0 : north_rows + 2 * (rows - y - 1));
      const float* const from = dst + ((unsigned)(y + 1) * (unsigned)nx + (unsigned)gx);
      float* const to = base + ((size_t)r * nx + gx);
#pragma unroll 1
      for (int k = 0; k < 9; k++) *reinterpret_cast<float4*>(to + k * plane) = __ldcg(reinterpret_cast<const float4*>(from + k * P));
    }
  };
  auto publish = [&](const int y) {
    if (EDGE) {
      if (y == kHalo - 1) { push_rows(0, true); strip_signal_cold(a, a.signal_south, strip, lane); }
      if (y == rows - 1) { push_rows(rows - kHalo, false); strip_signal_cold(a, a.signal_north, strip, lane); }
    }
  };

  f2 kp[3], kq[3];                                         // planes 4,7,8 of the row the previous step just finished
  // ring rows (float4 offsets inside a ring): planes 0,1,3 of the row being written / the row read now, planes
  // 2,5,6 of the row read now (y-1), the row in between, the row being written
  int a_w = 0, a_r = kRingA;
  int o_s = 2 * kRingA, o_c = 2 * kRingA + kRingB, o_n = 2 * kRingA + 2 * kRingB;

  // ---- step S (1..K) of row y.  Odd steps: operands in column order (lo = columns 0,1 of the lane's group, hi
  //      = columns 2,3), cells run as the pairs (1,2) and (3,0), results rotated.  Even steps: operands rotated
  //      (lo = columns 1,2, hi = columns 3,0), cells run as the pairs (0,1) and (2,3), results in column order.
  //      Either way only the three unshifted planes need register moves (lbm_kernels.cuh, collide_quad). ----
  auto step = [&](auto s_tag, const int y, const unsigned mword, const bool ahead, const bool more_pending) -> unsigned {
    constexpr int S = decltype(s_tag)::value;
    constexpr bool ODD = (S & 1) != 0, LAST = (S == K);
    f2 lo[9], hi[9];
    if (S == 1) {
      // (D = 2: the copy for the row after this one is in flight as well and may stay so)
      if (D == 2 && more_pending) cp_async_wait_but_one(); else cp_async_wait_all();
#pragma unroll
      for (int k = 0; k < 9; k++) lds2(stage + k * 32, lo[k], hi[k]);
      if (D == 2) stage = (stage == stage0) ? stage0 + kStageSlot : stage0;
    } else {
      const float4* const rg = ring0 + (S - 2) * kRingK;
      const float4* const s_c = rg + a_r;
      const float4* const s_s = rg + o_s;
      lds2(s_c + 0 * 32, lo[0], hi[0]); lds2(s_c + 1 * 32, lo[1], hi[1]); lds2(s_c + 2 * 32, lo[3], hi[3]);
      lds2(s_s + 0 * 32, lo[2], hi[2]); lds2(s_s + 1 * 32, lo[5], hi[5]); lds2(s_s + 2 * 32, lo[6], hi[6]);
      lo[4] = kp[0]; hi[4] = kq[0]; lo[7] = kp[1]; hi[7] = kq[1]; lo[8] = kp[2]; hi[8] = kq[2];
    }
    // what crosses the lanes: the east-moving populations' column 3 goes to the next lane (its cell 0 pulls it),
    // the west-moving ones' column 0 to the previous lane
    const float up1 = __shfl_up_sync(0xffffffffu, ODD ? hi2(hi[1]) : lo2(hi[1]), 1);
    const float up5 = __shfl_up_sync(0xffffffffu, ODD ? hi2(hi[5]) : lo2(hi[5]), 1);
    const float up8 = __shfl_up_sync(0xffffffffu, ODD ? hi2(hi[8]) : lo2(hi[8]), 1);
    const float dn3 = __shfl_down_sync(0xffffffffu, ODD ? lo2(lo[3]) : hi2(hi[3]), 1);
    const float dn6 = __shfl_down_sync(0xffffffffu, ODD ? lo2(lo[6]) : hi2(hi[6]), 1);
    const float dn7 = __shfl_down_sync(0xffffffffu, ODD ? lo2(lo[7]) : hi2(hi[7]), 1);
    // every lane has read its own staging cells (the shuffles consumed them): the copy D rows ahead may start
    unsigned word_next = 0u;
    if (S == 1 && ahead) word_next = issue();
    const unsigned bits = (mword >> mask_shift) & 0xFu;
    const bool any_blocked = __any_sync(0xffffffffu, bits != 0u);
    if (LAST && !ODD && !owned) return word_next;
    const bool fold = (LAST ? fold_last != 0 : true) && (ident(y) == accel_y);
    f2 p[9], q[9];
    if (ODD) {                                             // p = cells (1,2), q = cells (3,0)
      p[0] = pack2(hi2(lo[0]), lo2(hi[0])); q[0] = pack2(hi2(hi[0]), lo2(lo[0]));
      p[2] = pack2(hi2(lo[2]), lo2(hi[2])); q[2] = pack2(hi2(hi[2]), lo2(lo[2]));
      p[4] = pack2(hi2(lo[4]), lo2(hi[4])); q[4] = pack2(hi2(hi[4]), lo2(lo[4]));
      p[1] = lo[1]; q[1] = pack2(lo2(hi[1]), up1);
      p[5] = lo[5]; q[5] = pack2(lo2(hi[5]), up5);
      p[8] = lo[8]; q[8] = pack2(lo2(hi[8]), up8);
      p[3] = hi[3]; q[3] = pack2(dn3, hi2(lo[3]));
      p[6] = hi[6]; q[6] = pack2(dn6, hi2(lo[6]));
      p[7] = hi[7]; q[7] = pack2(dn7, hi2(lo[7]));
    } else {                                               // p = cells (0,1), q = cells (2,3)
      p[0] = pack2(hi2(hi[0]), lo2(lo[0])); q[0] = pack2(hi2(lo[0]), lo2(hi[0]));
      p[2] = pack2(hi2(hi[2]), lo2(lo[2])); q[2] = pack2(hi2(lo[2]), lo2(hi[2]));
      p[4] = pack2(hi2(hi[4]), lo2(lo[4])); q[4] = pack2(hi2(lo[4]), lo2(hi[4]));
      p[1] = pack2(up1, hi2(hi[1])); q[1] = lo[1];
      p[5] = pack2(up5, hi2(hi[5])); q[5] = lo[5];
      p[8] = pack2(up8, hi2(hi[8])); q[8] = lo[8];
      p[3] = lo[3]; q[3] = pack2(lo2(hi[3]), dn3);
      p[6] = lo[6]; q[6] = pack2(lo2(hi[6]), dn6);
      p[7] = lo[7]; q[7] = pack2(lo2(hi[7]), dn7);
    }
    const bool count = owned && y >= yb && y < ye;
    // (each branch stores its own results: a join would pin 36 registers to common locations)
    auto finish = [&](const float u4) {
      acc[S - 1] += (double)(count ? u4 : 0.0f);
      if (LAST) {
        if (!ODD) {
          emit(y, p, q);
        } else if (owned) {
          f2 op[9], oq[9];                                 // back to column order: cells (0,1), (2,3)
#pragma unroll
          for (int k = 0; k < 9; k++) { op[k] = pack2(hi2(q[k]), lo2(p[k])); oq[k] = pack2(hi2(p[k]), lo2(q[k])); }
          emit(y, op, oq);
        }
        return;
      }
      float4* const slot_a = ring0 + (S - 1) * kRingK + a_w;
      float4* const slot_b = ring0 + (S - 1) * kRingK + o_n;
      sts2(slot_a + 0 * 32, p[0], q[0]);
      sts2(slot_a + 1 * 32, p[1], q[1]);
      sts2(slot_a + 2 * 32, p[3], q[3]);
      sts2(slot_b + 0 * 32, p[2], q[2]);
      sts2(slot_b + 1 * 32, p[5], q[5]);
      sts2(slot_b + 2 * 32, p[6], q[6]);
      kp[0] = p[4]; kq[0] = q[4]; kp[1] = p[7]; kq[1] = q[7]; kp[2] = p[8]; kq[2] = q[8];
    };
    if (!any_blocked && !fold) finish(collide_quad_fast<ODD>(p, q, c));
    else finish(collide_quad_generic<ODD>(p, q, bits, c, fold));
    return word_next;
  };

  // Walk: iteration r runs step s on row r-(s-1), s = 1..K, as far as that row belongs to the step's range
  // [yb-(K-s), ye+(K-s)); every range ends exactly with the last iteration.  The copies run D rows ahead;
  // W[j] = obstacle word of row r+D-1-j, so step s reads W[D-2+s].
  const int r0 = yb - (K - 1), r_last = ye + K - 2;
  constexpr int NW = K + D - 1;
  unsigned W[NW];
#pragma unroll
  for (int j = 0; j < NW; j++) W[j] = 0u;
  W[D - 1] = issue();
  if (D == 2 && r0 + 1 <= r_last) W[0] = issue();
#pragma unroll 1
  for (int r = r0; r <= r_last; r++) {
    const unsigned w_new = step(std::integral_constant<int, 1>{}, r, W[D - 1], r + D <= r_last, r + 1 <= r_last);
    if constexpr (K >= 2) { if (r >= yb - K + 3) step(std::integral_constant<int, 2>{}, r - 1, W[D], false, false); }
    if constexpr (K >= 3) { if (r >= yb - K + 5) step(std::integral_constant<int, 3>{}, r - 2, W[D + 1], false, false); }
    if constexpr (K >= 4) { if (r >= yb - K + 7) step(std::integral_constant<int, 4>{}, r - 3, W[D + 2], false, false); }
    // (here, not inside the last step: its two halo lanes leave early, and the flag store is lane 0's)
    if (r - (K - 1) >= yb) publish(r - (K - 1));
    const int t = o_s; o_s = o_c; o_c = o_n; o_n = t;
    const int u = a_w; a_w = a_r; a_r = u;
#pragma unroll
    for (int j = NW - 1; j >= 1; j--) W[j] = W[j - 1];
    W[0] = w_new;
  }
}

template <int K, int D, int HINT, bool PEER>
__global__ void __launch_bounds__(stepsk_max_warps(K, D) * 32, 1) steps_strip(const StepArgs a, const StepsKArgs g)
{
  static_assert(K >= 1 && K <= kHalo, "the strip's four halo columns per side carry at most four steps");
  static_assert(D == 1 || D == 2, "one or two staging rows");
  extern __shared__ float4 fused_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  float4* const ring0 = fused_smem + (size_t)warp * stepsk_warp_float4(K, D) + lane;   // this lane's cells of the K-1 rings
  const float4* const stage0 = ring0 + (K - 1) * kRingK;                               // ... and of [D][9][32]
  double acc[K];
#pragma unroll
  for (int s = 0; s < K; s++) acc[s] = 0.0;

  const long nitems = (long)g.bands * g.strips;
  for (long item = (long)blockIdx.x * warps + warp; item < nitems; item += (long)gridDim.x * warps) {
    int band = (int)(item / g.strips);
    const int strip = (int)(item - (long)band * g.strips);
    // a ring slab does its two edge bands first: their rows are on the way to the neighbours (and published) while
    // the interior bands run, as the reference overlaps its halo exchange with the interior rows (326-366)
    if (PEER && g.bands > 2) band = (band == 0) ? 0 : (band == 1 ? g.bands - 1 : band - 1);
    stepsk_item<K, D, HINT, PEER>(a, g, band, strip, PEER && (band == 0 || band == g.bands - 1), ring0, stage0, lane, acc);
  }

  // one partial per WARP and step (fixed tree, no shared memory: the 256 static bytes of block_sum_to would cost K = 3
  // its eleventh resident CTA)
#pragma unroll
  for (int s = 0; s < K; s++) {
    double v = acc[s];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) a.partials[(size_t)s * g.partial_stride + (size_t)blockIdx.x * warps + warp] = v;
  }
  // End of a ring launch: the last CTA to leave advances the slab's epoch.  No fence: the epoch is read by this slab's
  // NEXT launch only (stream order), and the last CTA knows that every other one has left (with one-warp CTAs the
  // fence of peer_advance_epoch sat at the end of every work item).
  if (PEER && threadIdx.x == 0) {
    const unsigned prev = atomicAdd(a.done, 1u);
    if (prev == gridDim.x - 1) {
      *a.done = 0;
      *reinterpret_cast<volatile unsigned*>(a.epoch) = *reinterpret_cast<volatile unsigned*>(a.epoch) + 1u;
    }
  }
}

}  // namespace lbm
