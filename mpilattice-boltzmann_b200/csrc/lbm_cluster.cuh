// lbm_cluster.cuh -- kernel 6b: the 128-cell-wide decks resident in the shared memory of one thread-block cluster.
// (Kernel 6, the general form for any nx, is steps_cluster in lbm_kernels.cuh.)
#pragma once
#include "lbm_kernels.cuh"

namespace lbm {

// ---------------------------------------------------------------------------------------
// Kernel 6b ("cluster, one warp per row"): the fast form of steps_cluster (lbm_kernels.cuh) for nx = 128, the width of
// the two smallest shipped decks.  The grid lives in the shared memory of one 16-CTA cluster for up to 256 timesteps
// per launch; a warp owns ONE row of 128 cells for the whole launch (four cells per lane), so everything but the
// populations -- addresses, obstacle bits, whether the row is the driven one -- is loop-invariant.  A timestep is:
// nine 128-bit shared-memory loads (the rows below / above come from the CTA's two halo rows), six shuffles (the
// periodic x-wrap is the shuffle's wrap-around: the row is exactly one warp wide), the packed-pair collision of
// kernel 5, nine 128-bit stores into the other buffer, for a CTA's first / last row three more stores straight into
// the neighbour CTA's halo row (distributed shared memory), and ONE hardware cluster barrier.  Odd steps leave the
// lane's four cells rotated by one column, even steps restore the order (collide_quad), as in kernels 5 and 7.
// Global memory is touched at the start and the end of the launch and for one double per warp and step.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned map_to_cta(unsigned smem_addr, unsigned rank)
{
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void sts2_cluster(unsigned cluster_addr, f2 a, f2 b)
{
  asm volatile("st.shared::cluster.v2.u64 [%0], {%1, %2};" :: "r"(cluster_addr), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

constexpr int kClusterRowCells = 128;      // nx of kernel 6b
constexpr int kClusterMaxRows = 16;        // rows per CTA: 16 warps x 128 registers fill the register file

// MAXROWS = 8 (255 registers per thread) or 16 (128 registers).  The loop body holds ONE copy of the collision: the two
// parities differ in how the operands are assembled and in the order of the four |m|/rho terms only, so that the
// loop (two assemblies, the packed collision, its masked form for rows with obstacles and the driven row) stays
// small enough for the instruction cache -- the first version called the masked form out of line from two places and
// spent 2 us per step in it.
template <int MAXROWS>
__global__ void __launch_bounds__(MAXROWS * 32, 1) steps_cluster_rows(const ClusterArgs a)
{
  extern __shared__ float4 cluster_rows_smem[];
  unsigned rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int lane = threadIdx.x & 31, row = threadIdx.x >> 5, rows = a.rows_per_cta;   // one warp per row
  // buffer b, plane k, padded row r (0 and rows+1 = halo rows): float4 index ((b*9 + k)*(rows+2) + r)*32 + lane
  const int prow = rows + 2;
  // (behind the two buffers: one double per step and row -- the step's share of Sigma |m|/rho stays on chip until the
  // launch ends, so that the barrier's release fence never waits for a store to global memory)
  double* const sums = reinterpret_cast<double*>(cluster_rows_smem + 2 * 9 * prow * 32);
  float4* const mine = cluster_rows_smem + lane;
  auto at = [&](int b, int k, int r) -> float4* { return mine + ((b * 9 + k) * prow + r) * 32; };
  const unsigned mine_s = (unsigned)__cvta_generic_to_shared(mine);
  const unsigned south = (rank + kClusterCtas - 1) % kClusterCtas, north = (rank + 1) % kClusterCtas;
  const int grow = (int)rank * rows + row;                      // 0-based global row
  const int pr = row + 1;                                       // this row's padded row

  // state at launch -> buffer 0 (column order); the halo rows come from the neighbours' edge rows in global memory
  {
    const size_t g = (size_t)(grow + 1) * kClusterRowCells + 4 * lane;
#pragma unroll
    for (int k = 0; k < 9; k++) *at(0, k, pr) = __ldg(reinterpret_cast<const float4*>(a.in + k * a.plane + g));
    const int ny = kClusterCtas * rows;
    if (row == 0 || row == rows - 1) {
      const int gy = (row == 0) ? (grow + ny - 1) % ny : (grow + 1) % ny;
      const size_t gh = (size_t)(gy + 1) * kClusterRowCells + 4 * lane;
      const int hr = (row == 0) ? 0 : rows + 1;
#pragma unroll
      for (int k = 0; k < 9; k++) *at(0, k, hr) = __ldg(reinterpret_cast<const float4*>(a.in + k * a.plane + gh));
      if (rows == 1) {                                          // a one-row CTA has both halo rows on one warp
        const size_t gn = (size_t)((grow + 1) % ny + 1) * kClusterRowCells + 4 * lane;
#pragma unroll
        for (int k = 0; k < 9; k++) *at(0, k, rows + 1) = __ldg(reinterpret_cast<const float4*>(a.in + k * a.plane + gn));
      }
    }
  }
  const unsigned bits = (__ldg(a.mask + (size_t)grow * a.mask_row_words + (lane >> 3)) >> ((4 * lane) & 31)) & 0xFu;
  const bool any_blocked = __any_sync(0xffffffffu, bits != 0u);
  const bool driven = (grow == a.accel_row);
  const unsigned blocked_a = pair_order(bits, true), blocked_b = pair_order(bits, false);
  const int west = (lane + 31) & 31, east = (lane + 1) & 31;
  // where this row's planes go in the neighbour CTAs' halo rows (first row: 4,7,8 south; last row: 2,5,6 north)
  const unsigned to_south = map_to_cta(mine_s, south) + (unsigned)((rows + 1) * 512);
  const unsigned to_north = map_to_cta(mine_s, north);
  __syncthreads();
  cluster_arrive();                                             // every CTA of the cluster runs before anybody pushes
  cluster_wait();

#pragma unroll 1
  for (int t = 0; t < a.steps; t++) {
    const int cur = t & 1, nxt = cur ^ 1;
    const bool rot = (cur == 0);                                // column order in, rotated out (and back in the next step)
    const bool fold = driven && ((t + 1 < a.steps) || a.fold_last);
    f2 lo[9], hi[9];
    lds2(at(cur, 0, pr), lo[0], hi[0]); lds2(at(cur, 1, pr), lo[1], hi[1]); lds2(at(cur, 3, pr), lo[3], hi[3]);
    lds2(at(cur, 2, pr - 1), lo[2], hi[2]); lds2(at(cur, 5, pr - 1), lo[5], hi[5]); lds2(at(cur, 6, pr - 1), lo[6], hi[6]);
    lds2(at(cur, 4, pr + 1), lo[4], hi[4]); lds2(at(cur, 7, pr + 1), lo[7], hi[7]); lds2(at(cur, 8, pr + 1), lo[8], hi[8]);
    f2 p[9], q[9];
    if (rot) {
      // operands in column order (lo = columns 0,1 of the lane's group, hi = columns 2,3): pairs (1,2) and (3,0)
      const float up1 = __shfl_sync(0xffffffffu, hi2(hi[1]), west), up5 = __shfl_sync(0xffffffffu, hi2(hi[5]), west);
      const float up8 = __shfl_sync(0xffffffffu, hi2(hi[8]), west);
      const float dn3 = __shfl_sync(0xffffffffu, lo2(lo[3]), east), dn6 = __shfl_sync(0xffffffffu, lo2(lo[6]), east);
      const float dn7 = __shfl_sync(0xffffffffu, lo2(lo[7]), east);
      p[0] = pack2(hi2(lo[0]), lo2(hi[0])); q[0] = pack2(hi2(hi[0]), lo2(lo[0]));
      p[2] = pack2(hi2(lo[2]), lo2(hi[2])); q[2] = pack2(hi2(hi[2]), lo2(lo[2]));
      p[4] = pack2(hi2(lo[4]), lo2(hi[4])); q[4] = pack2(hi2(hi[4]), lo2(lo[4]));
      p[1] = lo[1]; q[1] = pack2(lo2(hi[1]), up1);
      p[5] = lo[5]; q[5] = pack2(lo2(hi[5]), up5);
      p[8] = lo[8]; q[8] = pack2(lo2(hi[8]), up8);
      p[3] = hi[3]; q[3] = pack2(dn3, hi2(lo[3]));
      p[6] = hi[6]; q[6] = pack2(dn6, hi2(lo[6]));
      p[7] = hi[7]; q[7] = pack2(dn7, hi2(lo[7]));
    } else {
      // operands rotated (lo = columns 1,2, hi = columns 3,0): pairs (0,1) and (2,3), results in column order
      const float up1 = __shfl_sync(0xffffffffu, lo2(hi[1]), west), up5 = __shfl_sync(0xffffffffu, lo2(hi[5]), west);
      const float up8 = __shfl_sync(0xffffffffu, lo2(hi[8]), west);
      const float dn3 = __shfl_sync(0xffffffffu, hi2(hi[3]), east), dn6 = __shfl_sync(0xffffffffu, hi2(hi[6]), east);
      const float dn7 = __shfl_sync(0xffffffffu, hi2(hi[7]), east);
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
    float2 up, uq;
    if (!any_blocked && !fold) collide_pairs(p, q, a.c.omega, a.c.negzero, up, uq);
    else collide_pairs_masked(p, q, rot ? blocked_a : blocked_b, a.c, fold, up, uq);
    // the lane's four terms in the reference's cell order
    const float u4 = rot ? add(add(add(uq.y, up.x), up.y), uq.x) : add(add(add(up.x, up.y), uq.x), uq.y);
#pragma unroll
    for (int k = 0; k < 9; k++) sts2(at(nxt, k, pr), p[k], q[k]);
    if (row == 0) {                                             // planes 4,7,8 -> the southern CTA's halo row rows+1
      sts2_cluster(to_south + (unsigned)((nxt * 9 + 4) * prow * 512), p[4], q[4]);
      sts2_cluster(to_south + (unsigned)((nxt * 9 + 7) * prow * 512), p[7], q[7]);
      sts2_cluster(to_south + (unsigned)((nxt * 9 + 8) * prow * 512), p[8], q[8]);
    }
    if (row == rows - 1) {                                      // planes 2,5,6 -> the northern CTA's halo row 0
      sts2_cluster(to_north + (unsigned)((nxt * 9 + 2) * prow * 512), p[2], q[2]);
      sts2_cluster(to_north + (unsigned)((nxt * 9 + 5) * prow * 512), p[5], q[5]);
      sts2_cluster(to_north + (unsigned)((nxt * 9 + 6) * prow * 512), p[6], q[6]);
    }
    cluster_arrive();                                           // this warp's new row (and halo rows) are on their way
    double acc = (double)u4;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) sums[t * rows + row] = acc;
    cluster_wait();                                             // ... and everybody's are visible
  }
  __syncthreads();
  for (int i = threadIdx.x; i < a.steps * rows; i += blockDim.x)
    a.partials[(size_t)(i / rows) * a.partial_stride + rank * rows + (i % rows)] = sums[i];

  // the state after `steps` steps -> global memory in column order (an odd count leaves the cells rotated)
  {
    const int fin = a.steps & 1;
    const size_t g = (size_t)(grow + 1) * kClusterRowCells + 4 * lane;
#pragma unroll
    for (int k = 0; k < 9; k++) {
      const float4 v = *at(fin, k, pr);
      *reinterpret_cast<float4*>(a.out + k * a.plane + g) = fin ? make_float4(v.w, v.x, v.y, v.z) : v;
    }
  }
}

}  // namespace lbm
