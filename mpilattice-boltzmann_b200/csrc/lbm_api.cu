// lbm_api.cu -- implementation of include/lbm_b200.h: slab bookkeeping, launches, halo wiring.
//
// Host-side structure of the reference's timed region (d2q9-bgk.c:278-398), re-designed for
// B200: no host synchronisation inside the step loop, one persistent-grid kernel launch per
// slab region per step, halo rows pushed over NVLink by the edge-row kernel itself, per-step
// averages reduced on the device and copied back once.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <time.h>
#include <unistd.h>

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include "../../include/lbm_b200.h"
#include "lbm_kernels.cuh"
#include "lbm_stepsk.cuh"
#include "lbm_cluster.cuh"

namespace {

thread_local std::string g_error;

int fail(int code, const char* fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_error = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                        \
  do {                                                                                        \
    cudaError_t err__ = (expr);                                                               \
    if (err__ != cudaSuccess)                                                                 \
      return fail(LBM_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), \
                  __FILE__, __LINE__);                                                        \
  } while (0)

enum FlagWord { kFromSouth = 0, kFromNorth = 1, kEpoch = 2, kDone = 3, kError = 4, kFlagWords = 32 };

struct Neighbour {
  float* buf[2] = {nullptr, nullptr};   // the neighbour's two population buffers (peer or IPC mapped)
  uint32_t* mask = nullptr;             // the neighbour's obstacle words (read while connecting; stays mapped until destroy)
  unsigned* flags = nullptr;            // the neighbour's flag words
  size_t plane = 0;
  int rows = 0;
  bool ipc = false;                     // mapped with cudaIpcOpenMemHandle (must be closed)
};

struct Slab {
  int device = 0;
  int rows = 0, first_row = 0;
  size_t plane = 0;
  float* buf[2] = {nullptr, nullptr};
  uint32_t* mask = nullptr;
  unsigned* flags = nullptr;
  double* partials = nullptr;
  float* av_dev = nullptr;
  size_t av_cap = 0;
  unsigned* cursor = nullptr;
  unsigned long long* blocked_dev = nullptr;   // blocked cells of this slab, counted while packing the mask
  float* bounce = nullptr;              // in-place handles: bounded staging buffer for state in/out
  size_t bounce_bytes = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t copy_stream = nullptr;   // device-to-host copies of get_final_state, overlapped with its kernels
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
  Neighbour south, north;
  int accel_row = -1;                   // padded row of global row ny-2, or -1 if another slab owns it
  // launch geometry
  int threads_full = 256, grid_full = 1;
  int threads_edge = 256, grid_edge = 0;
  int threads_int = 256, grid_int = 0;
  int per_step = 1;
  int fused_grid = 0, fused_bands = 0, fused_band_rows = 0;   // two-steps-per-pass kernel (kernel 5)
  int fused_partials = 0;                // partial sums per step the fused kernel in use writes
  unsigned fusedk_attr = 0;              // kernel 7: bit 2K+D set once steps_strip<K, D> has its shared-memory opt-in on this device
  bool fused_attr = false;              // its dynamic shared memory opt-in has been made on this slab's device
  std::vector<cudaGraphExec_t> graphs;  // [parity]
};

struct IpcBlob {
  cudaIpcMemHandle_t buf[2];
  cudaIpcMemHandle_t flags;
  cudaIpcMemHandle_t mask;
  unsigned long long plane;
  int rows;
  int device;
  int pid;
  int inplace;
};

constexpr int kMaxCtasPerSm = 128;      // upper bound of the oversubscribed grid, in CTAs per SM
constexpr long kResidentAutoMinCells = 1L << 18;
constexpr long kGraphAutoCells = 1L << 22;  // up to 2048^2: step kernel <= ~60 us, launch gaps matter
constexpr long kFusedAutoMinCells = 1L << 22;   // per GPU: from 2048^2 the two-steps-per-pass kernel wins
constexpr int kChunkSteps = 256;        // steps whose CTA partials are kept before one reduce launch
constexpr size_t kBounceBytes = 256u << 20;   // staging buffer of in-place handles (state in/out goes through it in row chunks)

}  // namespace

struct lbm_b200 {
  int nx = 0, ny = 0;
  float density = 0, accel = 0, omega = 0, inv = 0;
  lbm::StepConst sc{};
  int mask_row_words = 0;
  int n_ranks = 1;                      // slabs in the ring (all processes)
  int rank0 = 0;                        // ring index of slabs[0]
  bool multi_process = false;
  bool connected = true;
  bool inplace = false;                 // ONE population buffer, AA access pattern (csrc/lbm_kernels.cuh, kernel 4)
  int cur = 0;                          // index of the buffer holding the current state; in-place handles: the
                                        // layout of buf[0] (0 = canonical L0, 1 = L1 after an odd number of steps)
  std::vector<Slab> slabs;
  // options
  long opt_staging_bytes = (long)kBounceBytes;
  long opt_kernel = 0, opt_graph_steps = -1, opt_ctas_per_sm = 0, opt_min_ctas = 2, opt_cache_hint = 0, opt_resident = -1;
  long opt_fused2 = -1, opt_band_rows = 0;   // -1 / 0 = automatic
  long opt_fused_k7 = 0;                // 1 = kernel 7 also for two timesteps per pass (instead of kernel 5)
  long opt_fused_ctas = 0;              // kernel 7: CTAs per SM its resident warps are split into (0 = automatic)
  long opt_cluster = -1;                // kernel 6: -1 automatic, 0 never, 1 wherever it fits
  long opt_fused_steps = 0;             // timesteps per pass over HBM of the fused kernel: 2 = kernel 5, 3 or 4 = kernel 7,
                                        // 0 = automatic (fused_steps_wanted)
  long opt_fused_deep = -1;             // staging rows of the fused kernels: 1 = two (the copy runs two rows ahead), 0 = one,
                                        // -1 = automatic: two for kernel 5 (3 CTAs x 4 warps per SM, +3 %), one for kernel 7
                                        // (more resident warps instead, +5..25 %; profiles/r02_fused2.md)
  long opt_prefetch_rows = 0;           // kernel 5: L2 prefetch distance in rows (0 = off: measured slower, profiles/r02_fused2.md)
  long opt_spin_timeout_ms = 30000;     // how long a kernel waits for a ring neighbour's flag before it gives up
  long opt_debug_skip_slab = -1;        // test hook: this slab's step kernels are not launched (its neighbours time out)
  bool failed = false;                  // a wait timed out: the state is garbage, only destroy is valid
  bool fused2 = false;                  // two (or more) timesteps per pass over HBM (kernel 5 / 7) are in use
  int fusedk = 0;                       // kernel 7 is in use with this many timesteps per pass (0 = kernel 5)
  int fused_strips = 0;
  bool resident = false;                // the cooperative many-steps-per-launch kernel is in use
  bool cluster = false;                 // the grid lives in the shared memory of one 16-CTA cluster (kernel 6)
  bool cluster_rows = false;            // ... in its one-warp-per-row form (nx = 128, kernel 6b)
  long opt_cluster_rows = 1;            // 0 = always the general form of kernel 6 (tests, A/B)
  int cluster_threads = 0;
  int last_iters = 0;
  int graph_len = 0;
  long launches = 0;                    // kernels launched by the last enqueue (all slabs)
};

namespace {

using lbm::StepArgs;

bool use_vec4(const lbm_b200* h)
{
  const bool ok = (h->nx % 4 == 0) && h->nx >= 8;
  if (h->opt_kernel == 1) return false;
  return ok;
}

template <typename K>
int occupancy(K kernel, int threads)
{
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, 0) != cudaSuccess || n < 1) n = 1;
  return n;
}

// persistent-grid geometry for `rows` rows of the slab
void plan_region(const lbm_b200* h, int device, int rows, int* threads, int* grid, bool resident = false)
{
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (rows <= 0) { *threads = 256; *grid = 0; return; }
  if (use_vec4(h)) {
    const long nseg = (long)rows * ((h->nx + lbm::kSegCells - 1) / lbm::kSegCells);
    long warps = (nseg + sms - 1) / sms;             // spread small grids over the SMs
    warps = std::max(1L, std::min(8L, warps));
    *threads = (int)warps * 32;
    int per_sm = (int)h->opt_ctas_per_sm;
    if (resident) {
      // every CTA must be co-resident for the grid-wide barrier
      const int fit = occupancy(lbm::steps_resident<2>, *threads);
      per_sm = per_sm > 0 ? std::min(per_sm, fit) : fit;
    } else if (per_sm <= 0) {
      per_sm = h->inplace               ? occupancy(lbm::step_inplace<true, false, 0>, *threads)
               : (h->opt_min_ctas >= 4) ? occupancy(lbm::step_vec4<false, 4, 0>, *threads)
               : (h->opt_min_ctas == 3) ? occupancy(lbm::step_vec4<false, 3, 0>, *threads)
                                        : occupancy(lbm::step_vec4<false, 2, 0>, *threads);
      // Oversubscribe the resident slots so the hardware CTA scheduler balances the two dies, but keep
      // enough segments per warp to amortise the CTA prologue and block reduction.  Measured optimum
      // (profiles/r01_sweep.md): 128 CTAs/SM at 16384 rows x 16384, 16-32 at 4096 rows, 2-16 at 2048 rows
      // -> CTAs per SM = segments / 16384, clamped to [resident slots, 128].
      per_sm = (int)std::max<long>(per_sm, std::min<long>(kMaxCtasPerSm, nseg >> 14));
    }
    const long want = (nseg + warps - 1) / warps;
    *grid = (int)std::max(1L, std::min(want, (long)sms * per_sm));
  } else {
    const long ncell = (long)rows * h->nx;
    *threads = 256;
    int per_sm = (int)h->opt_ctas_per_sm;
    if (per_sm <= 0) per_sm = occupancy(lbm::step_scalar<false>, 256);
    const long want = (ncell + 255) / 256;
    *grid = (int)std::max(1L, std::min(want, (long)sms * per_sm));
  }
}

// The resident kernel pays off where a step is only a few microseconds of work (launch latency bound).
bool want_resident(const lbm_b200* h)
{
  if (h->opt_resident == 0 || h->n_ranks != 1 || h->slabs.size() != 1 || !use_vec4(h) || h->inplace) return false;
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->slabs[0].device);
  if (!coop) return false;
  // measured (profiles/r01_summary.md): below ~512^2 a CUDA graph of plain launches is as fast or faster
  // (2.7 vs 2.9 us per step); from 1024^2 to 2048^2 the resident kernel wins by 16-25 %
  const long cells = (long)h->nx * h->ny;
  return h->opt_resident == 1 || (cells >= kResidentAutoMinCells && cells <= kGraphAutoCells);
}

// Kernel 6: single-GPU ping-pong handles whose grid fits twice into the shared memory of one 16-CTA cluster.
// (the one-warp-per-row form, kernel 6b, keeps two halo rows per CTA as well)
size_t cluster_smem_bytes(const lbm_b200* h)
{
  const size_t rows = (size_t)(h->ny / lbm::kClusterCtas);
  if (!h->cluster_rows) return 2 * 9 * sizeof(float) * rows * h->nx;
  return 2 * 9 * sizeof(float) * (rows + 2) * h->nx + (size_t)kChunkSteps * rows * sizeof(double);   // + the steps' sums
}

bool want_cluster(lbm_b200* h)
{
  if (h->opt_cluster == 0 || h->n_ranks != 1 || h->slabs.size() != 1 || h->inplace) return false;
  // automatic: only where the caller has not asked for a particular kernel / launch mode
  if (h->opt_cluster < 0 && (h->opt_kernel != 0 || h->opt_resident == 1 || h->opt_graph_steps >= 0 || h->opt_fused2 == 1)) return false;
  if (h->ny % lbm::kClusterCtas != 0 || h->ny / lbm::kClusterCtas < 1) return false;
  // 128 cells wide (the two smallest shipped decks): one warp per row for the whole launch
  h->cluster_rows = h->nx == lbm::kClusterRowCells && h->ny / lbm::kClusterCtas <= lbm::kClusterMaxRows && h->opt_cluster_rows != 0;
  const size_t bytes = cluster_smem_bytes(h);
  int max_smem = 0, cluster_ok = 0;
  const int dev = h->slabs[0].device;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&cluster_ok, cudaDevAttrClusterLaunch, dev);
  if (!cluster_ok || bytes > (size_t)max_smem) return false;
  const int ncell = (h->ny / lbm::kClusterCtas) * h->nx;
  h->cluster_threads = h->cluster_rows ? 32 * (h->ny / lbm::kClusterCtas) : std::min(1024, (ncell + 31) / 32 * 32);
  // can one cluster of 16 such CTAs be resident?  (non-portable cluster size: opt in first)
  cudaSetDevice(dev);
  void (*const kernel)(lbm::ClusterArgs) = !h->cluster_rows ? lbm::steps_cluster
      : (h->ny / lbm::kClusterCtas <= 8 ? lbm::steps_cluster_rows<8> : lbm::steps_cluster_rows<lbm::kClusterMaxRows>);
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(lbm::kClusterCtas); cfg.blockDim = dim3(h->cluster_threads); cfg.dynamicSmemBytes = bytes;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = lbm::kClusterCtas; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  int clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&clusters, kernel, &cfg) != cudaSuccess || clusters < 1) {
    cudaGetLastError();
    return false;
  }
  return true;
}

// Two timesteps per pass (kernel 5): ping-pong handles whose rows are at least two strips wide; on a ring every
// slab needs four rows (the driven row ny-2 must not be a row a neighbour recomputes).
bool want_fused2(const lbm_b200* h)
{
  if (h->opt_fused2 == 0 || h->inplace || !use_vec4(h) || h->nx < 2 * lbm::kStripOut || h->ny / h->n_ranks < 4) return false;
  // every slab of a ring needs 4 rows.  A whole-domain handle sees all of them; a one-slab-per-process handle sees
  // its own and (once connected) its neighbours' -- lbm_b200_enqueue refuses to run the kernel on thinner slabs.
  if (h->n_ranks > 1)
    for (const Slab& s : h->slabs)
      if (s.rows < 4) return false;
  // measured (profiles/r01_fused2.md): 1.46x at 2048^2, 1.55x at 4096^2, 1.69x at 16384^2; slower at 1024^2, where a
  // pass has too few strips x bands to fill the GPU
  return h->opt_fused2 == 1 || (long)h->nx * h->ny / h->n_ranks >= kFusedAutoMinCells;
}

// kernel 5 launch shapes: 8 warps per CTA, 2 CTAs per SM, one staging row -- or ("deep", single-GPU handles) 4 warps
// per CTA, 3 CTAs per SM, two staging rows
constexpr int fused_warps(bool deep) { return deep ? 4 : 8; }
constexpr size_t fused_smem(bool deep) { return (size_t)fused_warps(deep) * lbm::fused_warp_float4(deep) * sizeof(float4); }

// staging rows of the fused kernel in use (see opt_fused_deep)
int stage_rows(const lbm_b200* h)
{
  if (h->opt_fused_deep >= 0) return h->opt_fused_deep != 0 ? 2 : 1;
  return h->fusedk ? 1 : 2;
}

// Launch shape of kernel 7 for k steps per pass and d staging rows.  Default: ONE warp per CTA, as many CTAs per SM
// as fit into its shared memory (a CTA costs 1 KB on top of its warp's rings and staging rows) and its register file
// (12 x 32 x 168) -- an SM slot is free again the moment a work item ends.  Against CTAs of 4 or 11 warps: +3.5 % at
// 16384^2 and +8.5 % on a 16384 x 2048 slab (K = 3, profiles/r02_fused2.md).  want_ctas > 0 (option "fused_ctas")
// splits those warps into that many CTAs instead.
void stepsk_shape(int k, int d, int* warps_per_cta, int* ctas_per_sm, int want_ctas = 0)
{
  auto max_warps = [&]() {
    switch (k * 10 + d) {
      case 11: return lbm::stepsk_max_warps(1, 1); case 12: return lbm::stepsk_max_warps(1, 2);
      case 21: return lbm::stepsk_max_warps(2, 1); case 22: return lbm::stepsk_max_warps(2, 2);
      case 31: return lbm::stepsk_max_warps(3, 1); case 32: return lbm::stepsk_max_warps(3, 2);
      case 41: return lbm::stepsk_max_warps(4, 1); default: return lbm::stepsk_max_warps(4, 2);
    }
  };
  const int w = max_warps();
  const bool split = want_ctas > 0 && w % want_ctas == 0;
  *ctas_per_sm = split ? want_ctas : w;
  *warps_per_cta = split ? w / want_ctas : 1;
}

// Padded rows of a plane beyond the slab's own: lbm::kHalo halo rows per side.  Rows 0 and rows+1 are the ones next to
// the slab (all ring kernels), rows+2(d-1) / rows+2(d-1)+1 the southern / northern neighbour's row d rows away (d = 2
// for kernel 5, d = 2..4 for kernel 7).  The obstacle words of the neighbours' rows follow the slab's own in the same
// order: word row rows+2(d-1) = the southern neighbour's d-th row from its end, rows+2(d-1)+1 = the northern
// neighbour's d-th row.
constexpr int kPad = 2 * lbm::kHalo;
// 0-based row (negative / >= rows: a neighbour's) that padded row r holds
int y_of_padded(int rows, int r)
{
  if (r <= rows + 1) return r - 1;
  const int d = (r - rows) / 2 + 1;
  return ((r - rows) % 2 == 0) ? -d : rows - 1 + d;
}

// Kernel 7 (K = 3 or 4 timesteps per pass) where kernel 5 applies, K was asked for and the rows allow it.
// Timesteps per pass of the fused kernel.  Automatic: 3 (kernel 7 with one staging row, ten one-warp CTAs per SM)
// wherever a fused kernel runs -- one B200, GLUPS for K = 2 (kernel 5) / 3 / 4: 2048^2 120.7 / 124.7 / 95.7, 4096^2
// 143.2 / 145.7 / 105.7, 8192^2 157.8 / 172.0 / 167.5, 16384 x 2048 151.4 / 165.7 / 152.8, 16384^2 157.9 / 180.2 /
// 176.6 (profiles/r02_fused2.md).
int fused_steps_wanted(const lbm_b200* h)
{
  return h->opt_fused_steps != 0 ? (int)h->opt_fused_steps : 3;
}

bool want_fusedk(const lbm_b200* h)
{
  if (fused_steps_wanted(h) < 3 && !h->opt_fused_k7) return false;
  if (h->n_ranks == 1) return h->ny >= 2 * lbm::kHalo;
  // a ring: every slab holds its neighbours' kHalo rows, and the driven row ny-2 must not be a row a slab other than
  // the one north of its owner recomputes (a whole-domain handle sees all slabs, a one-slab-per-process handle its
  // own and -- once connected -- its neighbours': lbm_b200_enqueue refuses to run on thinner ones)
  for (const Slab& s : h->slabs)
    if (s.rows < lbm::kHalo + 2) return false;
  return h->ny / h->n_ranks >= lbm::kHalo + 2;
}

// Band plan of the fused kernels for k timesteps per pass (lbm_b200_plan_bands is the k = 2 case): every work item
// recomputes 2(k-1) rows (+ half a row of start-up), one item per resident warp at a time.
// `warps_per_sm` = resident warps (= work items in flight) per multiprocessor of the kernel variant in use.
void plan_bands_k(int rows, int nx, int band_rows, int sms, int k, int warps_per_sm, bool ring, int* bands_out, int* rows_per_band)
{
  const int strips = (nx + lbm::kStripOut - 1) / lbm::kStripOut;
  const int min_edge = (k <= 2) ? 2 : (ring ? lbm::kHalo : 1);     // rows the first and the last band must hold
  int want = band_rows;
  if (want <= 0) {
    const long slots = (long)sms * warps_per_sm;
    double best = 0;
    for (int b : {8, 12, 16, 24, 32, 48, 64, 96, 128, 192, 256}) {
      if (k <= 2 && b > 128) break;
      const int nb = (rows + b - 1) / b;
      const int per_b = (rows + nb - 1) / nb;
      // Fitted to the band sweeps of profiles/r02_fused2.md (2048^2, 4096^2, 16384 x 2048, 16384^2): the items run
      // `fill` deep on every resident warp; CTAs are handed out as others finish, so a launch costs its work plus a
      // ragged end of ~0.6 item; a launch whose items all fit at once runs every warp in the same phase (1.37 x
      // slower per row than the steady state) and no faster than a warp alone can go (0.63 of a full SM's pace).
      const double fill = (double)nb * strips / (double)slots;
      // (k >= 3: + the walk's start-up and drain -- 192-row bands measured 1.7 % faster than 128 at 16384^2, k = 4)
      const double over = 2.0 * (k - 1) + (k >= 3 ? 4.0 : 0.5);
      // (kernel 7 pays more for a launch whose items all start at once: 4096^2, K = 3: 128-row bands -- 0.69 of a
      // wave -- 147.3 GLUPS, 24..32-row bands 157.5)
      const double at_once = (k >= 3) ? 1.55 : 1.37;
      const double cost = (per_b + over) * (fill > 1.0 ? fill + 0.62 : at_once * std::max(0.63, fill));
      if (want <= 0 || cost < best) { best = cost; want = b; }
    }
  }
  int bands = std::max(1, (rows + want - 1) / want);
  int per = (rows + bands - 1) / bands;
  while (bands > 1 && (per < min_edge || rows - (bands - 1) * per < min_edge)) {
    bands--;
    per = (rows + bands - 1) / bands;
  }
  *bands_out = bands;
  *rows_per_band = per;
}

void plan(lbm_b200* h)
{
  h->resident = want_resident(h);
  h->fused2 = want_fused2(h);
  h->cluster = !(h->fused2 && h->opt_fused2 == 1) && want_cluster(h);
  if (h->cluster) h->resident = h->fused2 = false;
  h->fusedk = (h->fused2 && want_fusedk(h)) ? fused_steps_wanted(h) : 0;
  if (h->fused2) {
    h->resident = false;
    h->fused_strips = (h->nx + lbm::kStripOut - 1) / lbm::kStripOut;
    for (Slab& s : h->slabs) {
      int sms = 148;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s.device);
      int bands = 1, per = s.rows;
      int k7_wpc = 0, k7_ctas = 0;
      if (h->fusedk) stepsk_shape(h->fusedk, stage_rows(h), &k7_wpc, &k7_ctas, (int)h->opt_fused_ctas);
      const int resident_warps = h->fusedk ? k7_wpc * k7_ctas : (stage_rows(h) == 2 ? 12 : 16);
      plan_bands_k(s.rows, h->nx, (int)h->opt_band_rows, sms, h->fusedk ? h->fusedk : 2, resident_warps, h->n_ranks > 1, &bands, &per);
      s.fused_bands = bands;
      s.fused_band_rows = per;
      const long items = (long)h->fused_strips * bands;
      const int wpc = h->fusedk ? k7_wpc : fused_warps(stage_rows(h) == 2);
      s.fused_grid = (int)std::min<long>((items + wpc - 1) / wpc, 1L << 30);
      s.fused_partials = h->fusedk ? (int)std::min<long>((long)s.fused_grid * wpc, 1L << 30) : s.fused_grid;   // kernel 7: one per warp
      if (h->opt_ctas_per_sm > 0) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s.device);
        s.fused_grid = (int)std::min<long>(s.fused_grid, (long)sms * h->opt_ctas_per_sm);
      }
    }
  }
  for (Slab& s : h->slabs) {
    if (h->n_ranks == 1) {
      plan_region(h, s.device, s.rows, &s.threads_full, &s.grid_full, h->resident);
      s.grid_edge = s.grid_int = 0;
      s.per_step = std::max(s.grid_full, h->fused2 ? s.fused_partials : 0);
      if (h->cluster) s.per_step = std::max(s.per_step, lbm::kClusterCtas * (h->cluster_threads / 32));
    } else if (use_vec4(h)) {
      // one launch per step and slab: edge rows first, then the interior (csrc/lbm_kernels.cuh)
      plan_region(h, s.device, s.rows, &s.threads_full, &s.grid_full);
      s.grid_edge = s.grid_int = 0;
      s.per_step = std::max(s.grid_full, h->fused2 ? s.fused_partials : 0);
    } else {
      plan_region(h, s.device, 2, &s.threads_edge, &s.grid_edge);
      plan_region(h, s.device, s.rows - 2, &s.threads_int, &s.grid_int);
      s.grid_full = 0;
      s.per_step = s.grid_edge + s.grid_int;
    }
  }
}

void destroy_graphs(lbm_b200* h)
{
  for (Slab& s : h->slabs) {
    for (cudaGraphExec_t g : s.graphs)
      if (g) cudaGraphExecDestroy(g);
    s.graphs.clear();
  }
  h->graph_len = 0;
}

// NVTX range over the enqueue of the timestep loop: what the reference brackets with
// MPI_Pcontrol(1, "mainloop") ... MPI_Pcontrol(-1, "mainloop") (d2q9-bgk.c:276, 405) for its ITAC/TAU traces.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

double now_s()
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + ts.tv_nsec * 1e-9;
}

// kFlagWords for the slab-level handshake, then per direction one word per 128-cell chunk (kernels 2 and 4) or per
// 120-column strip (kernel 5), whichever is more
size_t flag_word_count(int nx)
{
  const size_t chunks = (size_t)((nx + lbm::kSegCells - 1) / lbm::kSegCells);
  const size_t strips = (size_t)((nx + lbm::kStripOut - 1) / lbm::kStripOut);
  return kFlagWords + 2 * std::max(chunks, strips);
}

// bytes of `rows` obstacle rows in the caller's format
size_t obstacle_row_bytes(const lbm_b200* h, int format)
{
  return format == LBM_B200_OBST_INT32 ? (size_t)h->nx * sizeof(int)
         : format == LBM_B200_OBST_UINT8 ? (size_t)h->nx
                                         : (size_t)h->mask_row_words * sizeof(uint32_t);
}

// Uploads `rows` obstacle rows (host, any of the three formats) and leaves them bit-packed at `mask_dst`; `staged`
// is device scratch of at least rows * obstacle_row_bytes.  Counts the blocked cells into *blocked (may be NULL).
int upload_mask(lbm_b200* h, Slab& s, const void* obstacles_rows, int format, int rows, void* staged, uint32_t* mask_dst,
                unsigned long long* blocked)
{
  const size_t bytes = (size_t)rows * obstacle_row_bytes(h, format);
  const long warps = (long)rows * ((h->mask_row_words + 31) / 32);
  const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
  if (format == LBM_B200_OBST_BITS) {
    CUDA_TRY(cudaMemcpyAsync(mask_dst, obstacles_rows, bytes, cudaMemcpyHostToDevice, s.stream));
    const long words = (long)rows * h->mask_row_words;
    lbm::adopt_mask_bits<<<(unsigned)((words + 255) / 256), 256, 0, s.stream>>>(mask_dst, h->nx, rows, h->mask_row_words, blocked);
  } else {
    CUDA_TRY(cudaMemcpyAsync(staged, obstacles_rows, bytes, cudaMemcpyHostToDevice, s.stream));
    if (format == LBM_B200_OBST_INT32)
      lbm::pack_mask<<<blocks, 256, 0, s.stream>>>(static_cast<const int*>(staged), h->nx, rows, h->mask_row_words, mask_dst, blocked);
    else
      lbm::pack_mask_u8<<<blocks, 256, 0, s.stream>>>(static_cast<const unsigned char*>(staged), h->nx, rows, h->mask_row_words, mask_dst, blocked);
  }
  CUDA_TRY(cudaGetLastError());
  return LBM_B200_OK;
}

int alloc_slab(lbm_b200* h, Slab& s, const void* obstacles_rows, int format)
{
  const bool trace = getenv("LBM_B200_TRACE") != nullptr;
  const double t0 = now_s();
  CUDA_TRY(cudaSetDevice(s.device));
  if ((unsigned long long)(s.rows + kPad) * (unsigned long long)h->nx >= (1ull << 32))
    return fail(LBM_B200_ERR_ARG, "a slab of %d x %d cells is too large: offsets inside a plane are 32-bit (use more slabs)", h->nx, s.rows);
  // padded rows: 0 and rows+1 are the halo rows next to the slab, the rest the neighbours' rows further away that
  // the fused kernels keep on a ring (see kPad)
  s.plane = (size_t)(s.rows + kPad) * h->nx;
  const size_t bytes = 9 * s.plane * sizeof(float);
  for (int b = 0; b < (h->inplace ? 1 : 2); b++) {
    cudaError_t e = cudaMalloc(&s.buf[b], bytes);
    if (e != cudaSuccess)
      return fail(LBM_B200_ERR_ALLOC, "cudaMalloc of %zu bytes for populations failed: %s", bytes, cudaGetErrorString(e));
  }
  // flag words: kFlagWords for the slab-level handshake, then one per 128-cell chunk from the south and one per
  // chunk from the north (csrc/lbm_kernels.cuh, warp_peer_wait)
  const size_t flag_words = flag_word_count(h->nx);
  CUDA_TRY(cudaMalloc(&s.flags, flag_words * sizeof(unsigned)));
  CUDA_TRY(cudaMemset(s.flags, 0, flag_words * sizeof(unsigned)));
  CUDA_TRY(cudaMalloc(&s.cursor, sizeof(unsigned)));
  CUDA_TRY(cudaMemset(s.cursor, 0, sizeof(unsigned)));
  CUDA_TRY(cudaEventCreate(&s.ev_start));
  CUDA_TRY(cudaEventCreate(&s.ev_stop));

  const double t1 = now_s();
  // obstacle rows: the reference's int-per-cell array (or one byte per cell) is uploaded into the (still unused)
  // second population buffer and bit-packed on the device -- 32 cells per word, rows padded to whole words; rows
  // that arrive bit-packed go straight into place
  // (the words of the neighbours' rows follow the slab's own, see kPad and set_halo_mask)
  const size_t words = (size_t)(s.rows + kPad) * h->mask_row_words;
  CUDA_TRY(cudaMalloc(&s.mask, std::max<size_t>(words, 1) * sizeof(uint32_t)));
  CUDA_TRY(cudaMemsetAsync(s.mask, 0, std::max<size_t>(words, 1) * sizeof(uint32_t), s.stream));
  CUDA_TRY(cudaMalloc(&s.blocked_dev, sizeof(unsigned long long)));
  CUDA_TRY(cudaMemsetAsync(s.blocked_dev, 0, sizeof(unsigned long long), s.stream));
  {
    void* staged = s.buf[h->inplace ? 0 : 1];        // at most 4 of the buffer's 36 bytes per cell
    int rc = upload_mask(h, s, obstacles_rows, format, s.rows, staged, s.mask, s.blocked_dev);
    if (rc) return rc;
  }

  if (trace) cudaStreamSynchronize(s.stream);
  const double t2 = now_s();
  // uniform initial state in both buffers, halo rows included (d2q9-bgk.c:880-902)
  const float w0 = h->density * 4.0f / 9.0f, w1 = h->density / 9.0f, w2 = h->density / 36.0f;
  const unsigned blocks = (unsigned)((s.plane + 255) / 256);
  for (int b = 0; b < (h->inplace ? 1 : 2); b++) lbm::fill_planes<<<blocks, 256, 0, s.stream>>>(s.buf[b], s.plane, w0, w1, w2);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(s.stream));

  if (trace)
    fprintf(stderr, "lbm_b200 trace: slab rows=%d alloc %.3f s, obstacle upload+pack %.3f s, fill %.3f s\n", s.rows,
            t1 - t0, t2 - t1, now_s() - t2);
  const int accel_global = h->ny - 2;
  s.accel_row = (accel_global >= s.first_row && accel_global < s.first_row + s.rows) ? accel_global - s.first_row + 1 : -1;
  return LBM_B200_OK;
}

// Obstacle words of the neighbours' rows a ring slab computes redundantly in the fused kernels (and of row -2 for
// the body-force pre-pass): rows_src[j] = the obstacle row that goes into word row rows + j (see kPad).
int set_halo_mask(lbm_b200* h, Slab& s, const void* const (&rows_src)[kPad], int format)
{
  CUDA_TRY(cudaSetDevice(s.device));
  const size_t row_bytes = obstacle_row_bytes(h, format);
  char* staged = nullptr;
  CUDA_TRY(cudaMalloc(&staged, kPad * row_bytes));
  int rc = LBM_B200_OK;
  for (int i = 0; i < kPad && rc == LBM_B200_OK; i++)
    rc = upload_mask(h, s, rows_src[i], format, 1, staged + (size_t)i * row_bytes, s.mask + (size_t)(s.rows + i) * h->mask_row_words, nullptr);
  cudaError_t e = cudaStreamSynchronize(s.stream);
  cudaFree(staged);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(LBM_B200_ERR_CUDA, "uploading the neighbours' obstacle rows failed: %s", cudaGetErrorString(e));
  return LBM_B200_OK;
}

// *moved = true when the array was reallocated: CUDA graphs captured reduce_partials with the old pointer and
// must be rebuilt (the caller destroys them).
int ensure_av_capacity(Slab& s, size_t iters, bool* moved)
{
  const size_t need = std::max<size_t>(iters, 1);
  if (s.av_cap >= need) return LBM_B200_OK;
  CUDA_TRY(cudaSetDevice(s.device));
  CUDA_TRY(cudaStreamSynchronize(s.stream));        // nothing in flight may still write the old array
  if (s.av_dev) CUDA_TRY(cudaFree(s.av_dev));
  s.av_dev = nullptr; s.av_cap = 0;
  const size_t cap = std::max(need, std::min<size_t>(2 * need, 1u << 20));   // room to grow without another move
  CUDA_TRY(cudaMalloc(&s.av_dev, cap * sizeof(float)));
  s.av_cap = cap;
  *moved = true;
  return LBM_B200_OK;
}

int ensure_partials(lbm_b200* h)
{
  for (Slab& s : h->slabs) {
    CUDA_TRY(cudaSetDevice(s.device));
    if (s.partials) CUDA_TRY(cudaFree(s.partials));
    s.partials = nullptr;
    CUDA_TRY(cudaMalloc(&s.partials, (size_t)kChunkSteps * s.per_step * sizeof(double)));
  }
  return LBM_B200_OK;
}

// where the canonical populations of the current state live (see lbm::locate)
lbm::Layout layout_of(const lbm_b200* h, const Slab& s)
{
  lbm::Layout l{};
  l.plane = s.plane; l.nx = h->nx; l.rows = s.rows; l.odd = h->inplace ? h->cur : 0;
  const bool ring = h->n_ranks > 1;
  l.south = ring ? s.south.buf[0] : s.buf[0]; l.south_plane = ring ? s.south.plane : s.plane; l.south_rows = ring ? s.south.rows : s.rows;
  l.north = ring ? s.north.buf[0] : s.buf[0]; l.north_plane = ring ? s.north.plane : s.plane;
  return l;
}

// Device staging for state in/out: the idle ping-pong buffer, or (in-place handles, which have none) a
// bounded bounce buffer.  *rows = how many grid rows of `row_bytes` fit.
int staging(lbm_b200* h, Slab& s, size_t row_bytes, int rows_wanted, float** ptr, int* rows)
{
  if (!h->inplace) {
    *ptr = s.buf[h->cur ^ 1];
    *rows = rows_wanted;                 // 9 planes of rows+2 rows always hold what the callers ask for
    return LBM_B200_OK;
  }
  const size_t want = std::min(std::max((size_t)h->opt_staging_bytes, row_bytes), row_bytes * (size_t)rows_wanted);
  if (s.bounce_bytes < want || s.bounce_bytes > std::max(want, (size_t)h->opt_staging_bytes)) {
    if (s.bounce) CUDA_TRY(cudaFree(s.bounce));
    s.bounce = nullptr; s.bounce_bytes = 0;
    cudaError_t e = cudaMalloc(&s.bounce, want);
    if (e != cudaSuccess) return fail(LBM_B200_ERR_ALLOC, "cudaMalloc of %zu bytes of staging failed: %s", want, cudaGetErrorString(e));
    s.bounce_bytes = want;
  }
  *ptr = s.bounce;
  *rows = (int)std::min<size_t>((size_t)rows_wanted, s.bounce_bytes / row_bytes);
  return LBM_B200_OK;
}

StepArgs base_args(const lbm_b200* h, const Slab& s, int slot, bool fold_accel)
{
  StepArgs a{};
  a.src = s.buf[h->inplace ? 0 : h->cur];
  a.dst = s.buf[h->inplace ? 0 : h->cur ^ 1];
  a.plane = s.plane;
  a.mask = s.mask;
  a.mask_row_words = h->mask_row_words;
  a.nx = h->nx;
  a.chunks = (h->nx + lbm::kSegCells - 1) / lbm::kSegCells;
  a.row_first = 1;
  a.row_last = s.rows;
  a.accel_row = fold_accel ? s.accel_row : -1;
  a.c = h->sc;
  a.partials = s.partials + (size_t)slot * s.per_step;
  a.error = s.flags + kError;
  a.timeout_ns = (unsigned long long)h->opt_spin_timeout_ms * 1000000ull;
  return a;
}

// halo wiring of a slab's step: where the outgoing rows go and which flag words order the exchange
void peer_args(const lbm_b200* h, const Slab& s, StepArgs& a)
{
  a.south_of_first = 0;
  a.north_of_last = s.rows + 1;
  a.north_dst = s.north.buf[h->cur ^ 1]; a.north_plane = s.north.plane; a.north_row = 0;
  a.south_dst = s.south.buf[h->cur ^ 1]; a.south_plane = s.south.plane; a.south_row = s.south.rows + 1;
  if (h->inplace) {
    // one buffer per slab; the NEIGHBOUR flavour (layout L0 -> L1) writes into the neighbours' owned edge rows,
    // the LOCAL flavour pushes copies into their halo rows like the ping-pong kernel
    a.north_dst = s.north.buf[0];
    a.south_dst = s.south.buf[0];
    if (h->cur == 0) { a.north_row = 1; a.south_row = s.south.rows; }
  }
  if (use_vec4(h)) {                               // one flag word per 128-cell chunk of the edge rows
    a.wait_from_south = s.flags + kFlagWords;
    a.wait_from_north = s.flags + kFlagWords + a.chunks;
    a.signal_north = s.north.flags + kFlagWords;   // I am my northern neighbour's south
    a.signal_south = s.south.flags + kFlagWords + a.chunks;
  } else {
    a.wait_from_south = s.flags + kFromSouth;
    a.wait_from_north = s.flags + kFromNorth;
    a.signal_north = s.north.flags + kFromSouth;
    a.signal_south = s.south.flags + kFromNorth;
  }
  a.epoch = s.flags + kEpoch;
  a.done = s.flags + kDone;
}

template <bool PEER>
int launch_step(lbm_b200* h, const Slab& s, const StepArgs& a, int grid, int threads)
{
  if (grid <= 0) return LBM_B200_OK;
  h->launches++;
  if (use_vec4(h)) {
#define LBM_LAUNCH_VEC4(M, H) lbm::step_vec4<PEER, M, H><<<grid, threads, 0, s.stream>>>(a)
#define LBM_LAUNCH_HINT(M)                                                                     \
  do {                                                                                         \
    if (h->opt_cache_hint == 1) LBM_LAUNCH_VEC4(M, 1);                                         \
    else if (h->opt_cache_hint == 2) LBM_LAUNCH_VEC4(M, 2);                                    \
    else if (h->opt_cache_hint == 4) LBM_LAUNCH_VEC4(M, 4);                                    \
    else LBM_LAUNCH_VEC4(M, 0);                                                                \
  } while (0)
    if (h->opt_min_ctas >= 4) LBM_LAUNCH_HINT(4);
    else if (h->opt_min_ctas == 3) LBM_LAUNCH_HINT(3);
    else LBM_LAUNCH_HINT(2);
#undef LBM_LAUNCH_HINT
#undef LBM_LAUNCH_VEC4
  } else {
    lbm::step_scalar<PEER><<<grid, threads, 0, s.stream>>>(a);
  }
  CUDA_TRY(cudaGetLastError());
  return LBM_B200_OK;
}

// One timestep on every slab of this handle: d2q9-bgk.c:326-378.
int enqueue_step(lbm_b200* h, int slot, bool fold_accel)
{
  if (h->inplace) {
    // one buffer: the NEIGHBOUR flavour takes layout L0 to L1, the LOCAL flavour takes it back
    for (Slab& s : h->slabs) {
      CUDA_TRY(cudaSetDevice(s.device));
      StepArgs a = base_args(h, s, slot, fold_accel);
      a.row_begin = 1; a.row_count = s.rows; a.row_stride = 1;
      h->launches++;
      if (h->n_ranks == 1) {
        a.south_of_first = s.rows;
        a.north_of_last = 1;
#define LBM_LAUNCH_INPLACE(H)                                                                             \
  do {                                                                                                    \
    if (h->cur == 0) lbm::step_inplace<true, false, H><<<s.grid_full, s.threads_full, 0, s.stream>>>(a);  \
    else lbm::step_inplace<false, false, H><<<s.grid_full, s.threads_full, 0, s.stream>>>(a);             \
  } while (0)
        if (h->opt_cache_hint == 1) LBM_LAUNCH_INPLACE(1);
        else if (h->opt_cache_hint == 2) LBM_LAUNCH_INPLACE(2);
        else LBM_LAUNCH_INPLACE(0);
#undef LBM_LAUNCH_INPLACE
      } else {
        peer_args(h, s, a);
        if (h->cur == 0) lbm::step_inplace<true, true, 0><<<s.grid_full, s.threads_full, 0, s.stream>>>(a);
        else lbm::step_inplace<false, true, 0><<<s.grid_full, s.threads_full, 0, s.stream>>>(a);
      }
      CUDA_TRY(cudaGetLastError());
    }
  } else if (h->n_ranks == 1) {
    Slab& s = h->slabs[0];
    CUDA_TRY(cudaSetDevice(s.device));
    StepArgs a = base_args(h, s, slot, fold_accel);
    a.row_begin = 1; a.row_count = s.rows; a.row_stride = 1;
    a.south_of_first = s.rows;          // periodic wrap in y without halo copies
    a.north_of_last = 1;
    int rc = launch_step<false>(h, s, a, s.grid_full, s.threads_full);
    if (rc) return rc;
  } else if (use_vec4(h)) {
    // ONE launch per slab: its first work items are the two edge rows, which wait for the neighbours'
    // previous halo rows, push this state's halo rows over NVLink and signal; the interior follows in the
    // same launch while the neighbours consume (replaces MPI_Startall ... interior ... MPI_Waitall, 326-366)
    for (Slab& s : h->slabs) {
      if (&s - h->slabs.data() == h->opt_debug_skip_slab) continue;
      CUDA_TRY(cudaSetDevice(s.device));
      StepArgs a = base_args(h, s, slot, fold_accel);
      a.row_begin = 1; a.row_count = s.rows; a.row_stride = 1;
      peer_args(h, s, a);
      int rc = launch_step<true>(h, s, a, s.grid_full, s.threads_full);
      if (rc) return rc;
    }
  } else {
    // scalar kernel: edge rows in their own launch (CTA-level handshake), then the interior
    for (Slab& s : h->slabs) {
      CUDA_TRY(cudaSetDevice(s.device));
      StepArgs a = base_args(h, s, slot, fold_accel);
      a.row_begin = 1; a.row_count = 2; a.row_stride = s.rows - 1;
      peer_args(h, s, a);
      int rc = launch_step<true>(h, s, a, s.grid_edge, s.threads_edge);
      if (rc) return rc;
    }
    for (Slab& s : h->slabs) {
      CUDA_TRY(cudaSetDevice(s.device));
      StepArgs a = base_args(h, s, slot, fold_accel);
      a.row_begin = 2; a.row_count = s.rows - 2; a.row_stride = 1;
      a.south_of_first = 0;
      a.north_of_last = s.rows + 1;
      a.partials += s.grid_edge;
      int rc = launch_step<false>(h, s, a, s.grid_int, s.threads_int);
      if (rc) return rc;
    }
  }
  h->cur ^= 1;
  return LBM_B200_OK;
}

// Two timesteps in one pass over HBM (kernel 5): partial slots `slot` and `slot`+1 -- or, with `single`, the odd
// last step of a run on a ring through the same strips (so that the strip-level handshake stays the only protocol
// in use and the neighbours' halo rows are refreshed).
// The kernel needs more dynamic shared memory than the default limit: the opt-in is per device (context) and is
// made once per slab when the kernel is first launched for it (Slab::fused_attr), never from process-global state.
int allow_fused_smem(Slab& s)
{
  if (s.fused_attr) return LBM_B200_OK;
  CUDA_TRY(cudaSetDevice(s.device));
  const int bytes = (int)fused_smem(false);
  CUDA_TRY(cudaFuncSetAttribute(lbm::steps2_strip<0, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem(true)));
  CUDA_TRY(cudaFuncSetAttribute(lbm::steps2_strip<3, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem(true)));
  CUDA_TRY(cudaFuncSetAttribute(lbm::steps2_strip<3, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem(true)));
  CUDA_TRY(cudaFuncSetAttribute(lbm::steps2_strip<0, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CUDA_TRY(cudaFuncSetAttribute(lbm::steps2_strip<1, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CUDA_TRY(cudaFuncSetAttribute(lbm::steps2_strip<2, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CUDA_TRY(cudaFuncSetAttribute(lbm::steps2_strip<3, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CUDA_TRY(cudaFuncSetAttribute(lbm::steps2_strip<3, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  s.fused_attr = true;
  return LBM_B200_OK;
}

int enqueue_fused2(lbm_b200* h, int slot, bool fold_last, bool single)
{
  for (Slab& s : h->slabs) {
    if (&s - h->slabs.data() == h->opt_debug_skip_slab) continue;
    CUDA_TRY(cudaSetDevice(s.device));
    if (int rc = allow_fused_smem(s)) return rc;
    StepArgs a = base_args(h, s, slot, false);
    a.row_begin = 1; a.row_count = s.rows; a.row_stride = 1;
    a.south_of_first = s.rows;
    a.north_of_last = 1;
    lbm::FusedArgs g{};
    g.band_rows = s.fused_band_rows;
    g.bands = s.fused_bands;
    g.strips = h->fused_strips;
    g.accel_row = s.accel_row;
    g.fold_last = fold_last ? 1 : 0;
    g.partial_stride = s.per_step;
    g.prefetch_rows = (int)h->opt_prefetch_rows;
    const bool deep = stage_rows(h) == 2;
    const size_t kFusedSmem = fused_smem(deep);
    const dim3 grid(s.fused_grid), block(fused_warps(deep) * 32);
    if (h->n_ranks == 1) {
      if (deep) lbm::steps2_strip<0, false, false, true><<<grid, block, kFusedSmem, s.stream>>>(a, g);
      else if (h->opt_cache_hint == 1) lbm::steps2_strip<1, false, false><<<grid, block, kFusedSmem, s.stream>>>(a, g);
      else if (h->opt_cache_hint == 2) lbm::steps2_strip<2, false, false><<<grid, block, kFusedSmem, s.stream>>>(a, g);
      else lbm::steps2_strip<0, false, false><<<grid, block, kFusedSmem, s.stream>>>(a, g);
    } else {
      peer_args(h, s, a);
      const int strips = h->fused_strips;                    // one flag word per strip and direction
      a.wait_from_south = s.flags + kFlagWords;
      a.wait_from_north = s.flags + kFlagWords + strips;
      a.signal_north = s.north.flags + kFlagWords;
      a.signal_south = s.south.flags + kFlagWords + strips;
      g.south_rows = s.south.rows;
      g.north_rows = s.north.rows;
      // (halo rows are written by the neighbours: L2-coherent loads, HINT 3 -- irrelevant for cp.async.cg)
      if (deep) {
        if (single) lbm::steps2_strip<3, true, true, true><<<grid, block, kFusedSmem, s.stream>>>(a, g);
        else lbm::steps2_strip<3, true, false, true><<<grid, block, kFusedSmem, s.stream>>>(a, g);
      } else {
        if (single) lbm::steps2_strip<3, true, true><<<grid, block, kFusedSmem, s.stream>>>(a, g);
        else lbm::steps2_strip<3, true, false><<<grid, block, kFusedSmem, s.stream>>>(a, g);
      }
    }
    CUDA_TRY(cudaGetLastError());
    h->launches++;
  }
  h->cur ^= 1;
  return LBM_B200_OK;
}

// K' <= K timesteps in one pass over HBM (kernel 7): partial slots `slot` .. `slot`+K'-1.  Shorter passes finish a
// chunk or a run through the same strips (and, on a ring, the same handshake and halo depth).
template <int K, int D, bool PEER>
int launch_stepsk(lbm_b200* h, Slab& s, const StepArgs& a, const lbm::StepsKArgs& g)
{
  int wpc = 0, ctas = 0;
  stepsk_shape(K, D, &wpc, &ctas, K == h->fusedk ? (int)h->opt_fused_ctas : 0);
  const size_t smem = (size_t)wpc * lbm::stepsk_warp_bytes(K, D);
  auto kernel = lbm::steps_strip<K, D, 0, PEER>;
  const unsigned bit = 1u << (2 * K + D + (PEER ? 16 : 0));
  if (!(s.fusedk_attr & bit)) {
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    s.fusedk_attr |= bit;
  }
  const long items = (long)g.bands * g.strips;
  long grid = (items + wpc - 1) / wpc;
  if (h->opt_ctas_per_sm > 0) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s.device);
    grid = std::min<long>(grid, (long)sms * h->opt_ctas_per_sm);
  }
  grid = std::min<long>(grid, s.per_step / wpc);             // one partial per warp and step
  kernel<<<dim3((unsigned)grid), dim3(wpc * 32), smem, s.stream>>>(a, g);
  CUDA_TRY(cudaGetLastError());
  return LBM_B200_OK;
}

int enqueue_fusedk(lbm_b200* h, int slot, int k, bool fold_last)
{
  for (Slab& s : h->slabs) {
    if (&s - h->slabs.data() == h->opt_debug_skip_slab) continue;
    CUDA_TRY(cudaSetDevice(s.device));
    StepArgs a = base_args(h, s, slot, false);
    a.row_begin = 1; a.row_count = s.rows; a.row_stride = 1;
    a.south_of_first = s.rows;
    a.north_of_last = 1;
    lbm::StepsKArgs g{};
    g.band_rows = s.fused_band_rows;
    g.bands = s.fused_bands;
    g.strips = h->fused_strips;
    g.accel_y = s.accel_row >= 1 ? s.accel_row - 1 : -1000;
    g.fold_last = fold_last ? 1 : 0;
    g.partial_stride = s.per_step;
    const bool peer = h->n_ranks > 1;
    if (peer) {
      peer_args(h, s, a);
      const int strips = h->fused_strips;                    // one flag word per strip and direction, as kernel 5
      a.wait_from_south = s.flags + kFlagWords;
      a.wait_from_north = s.flags + kFlagWords + strips;
      a.signal_north = s.north.flags + kFlagWords;
      a.signal_south = s.south.flags + kFlagWords + strips;
      g.south_rows = s.south.rows;
      g.north_rows = s.north.rows;
      // the slab north of the driven row's owner recomputes that row (its row -2) in the steps before the last
      if (s.first_row == 0 && s.accel_row < 0) g.accel_y = -2;
    }
    int rc;
    switch ((peer ? 100 : 0) + k * 10 + stage_rows(h)) {
      case 11: rc = launch_stepsk<1, 1, false>(h, s, a, g); break;
      case 12: rc = launch_stepsk<1, 2, false>(h, s, a, g); break;
      case 21: rc = launch_stepsk<2, 1, false>(h, s, a, g); break;
      case 22: rc = launch_stepsk<2, 2, false>(h, s, a, g); break;
      case 31: rc = launch_stepsk<3, 1, false>(h, s, a, g); break;
      case 32: rc = launch_stepsk<3, 2, false>(h, s, a, g); break;
      case 41: rc = launch_stepsk<4, 1, false>(h, s, a, g); break;
      case 42: rc = launch_stepsk<4, 2, false>(h, s, a, g); break;
      case 111: rc = launch_stepsk<1, 1, true>(h, s, a, g); break;
      case 112: rc = launch_stepsk<1, 2, true>(h, s, a, g); break;
      case 121: rc = launch_stepsk<2, 1, true>(h, s, a, g); break;
      case 122: rc = launch_stepsk<2, 2, true>(h, s, a, g); break;
      case 131: rc = launch_stepsk<3, 1, true>(h, s, a, g); break;
      case 132: rc = launch_stepsk<3, 2, true>(h, s, a, g); break;
      case 141: rc = launch_stepsk<4, 1, true>(h, s, a, g); break;
      default: rc = launch_stepsk<4, 2, true>(h, s, a, g); break;
    }
    if (rc) return rc;
    h->launches++;
  }
  h->cur ^= 1;
  return LBM_B200_OK;
}

int enqueue_reduce(lbm_b200* h, int steps)
{
  for (Slab& s : h->slabs) {
    CUDA_TRY(cudaSetDevice(s.device));
    const int blocks = (steps * 32 + 255) / 256;
    lbm::reduce_partials<<<blocks, 256, 0, s.stream>>>(s.partials, s.per_step, steps, h->inv, s.av_dev, s.cursor);
    lbm::advance_cursor<<<1, 1, 0, s.stream>>>(s.cursor, (unsigned)steps);
    CUDA_TRY(cudaGetLastError());
    h->launches += 2;
  }
  return LBM_B200_OK;
}

// Captures `len` steps (+ their reduce) into one CUDA graph per stream, for both buffer parities.  Slabs of a
// ring are ordered by the flag words their kernels exchange, not by stream dependencies, so every stream
// (device) gets its own independent graph and the host launches one graph per device and `len` steps.
int build_graphs(lbm_b200* h, int len)
{
  destroy_graphs(h);
  const int cur0 = h->cur;
  for (Slab& s : h->slabs) s.graphs.assign(2, nullptr);
  for (int parity = 0; parity < 2; parity++) {
    h->cur = parity;
    int rc = LBM_B200_OK;
    size_t begun = 0;
    for (; begun < h->slabs.size() && rc == LBM_B200_OK; begun++) {
      Slab& s = h->slabs[begun];
      if (!s.own_stream) continue;
      if (cudaSetDevice(s.device) != cudaSuccess || cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
        rc = fail(LBM_B200_ERR_CUDA, "graph capture could not begin: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (rc != LBM_B200_OK) begun--;                  // the slab that failed has no capture to end
    for (int t = 0; t < len && rc == LBM_B200_OK; t++) rc = enqueue_step(h, t, true);
    if (rc == LBM_B200_OK) rc = enqueue_reduce(h, len);
    // end every capture first: nothing may be instantiated while a stream of this thread still captures
    std::vector<cudaGraph_t> captured(h->slabs.size(), nullptr);
    for (size_t i = 0; i < begun; i++) {
      Slab& s = h->slabs[i];
      if (!s.own_stream) continue;
      cudaSetDevice(s.device);
      cudaError_t e = cudaStreamEndCapture(s.stream, &captured[i]);
      if (e != cudaSuccess && rc == LBM_B200_OK) rc = fail(LBM_B200_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
    }
    for (size_t i = 0; i < begun; i++) {
      Slab& s = h->slabs[i];
      if (!captured[i]) continue;
      cudaSetDevice(s.device);
      if (rc == LBM_B200_OK) {
        cudaError_t e = cudaGraphInstantiate(&s.graphs[parity], captured[i], 0);
        if (e != cudaSuccess) rc = fail(LBM_B200_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
      }
      cudaGraphDestroy(captured[i]);
    }
    cudaGetLastError();
    if (rc != LBM_B200_OK) { h->cur = cur0; destroy_graphs(h); return rc; }
  }
  h->cur = cur0;
  h->graph_len = len;
  return LBM_B200_OK;
}

// The slab-level (scalar kernel) and chunk-level (128-bit kernels) handshakes keep separate flag words but share
// the epoch: after switching kernels on a ring, every flag word is brought up to the slab's epoch.  The caller
// has drained the streams.
int resync_flags(lbm_b200* h)
{
  if (h->n_ranks == 1) return LBM_B200_OK;
  const size_t words = flag_word_count(h->nx);
  std::vector<unsigned> host(words);
  for (Slab& s : h->slabs) {
    CUDA_TRY(cudaSetDevice(s.device));
    unsigned epoch = 0;
    CUDA_TRY(cudaMemcpy(&epoch, s.flags + kEpoch, sizeof epoch, cudaMemcpyDeviceToHost));
    std::fill(host.begin(), host.end(), epoch);
    host[kDone] = 0;
    for (int i = kDone + 1; i < kFlagWords; i++) host[i] = 0;
    CUDA_TRY(cudaMemcpy(s.flags, host.data(), words * sizeof(unsigned), cudaMemcpyHostToDevice));
  }
  return LBM_B200_OK;
}

// The one-step ring kernels keep only three planes of one halo row per side up to date; the two-steps-per-pass
// kernels pull from two (kernel 5) or up to four (kernel 7) full halo rows per side.  When one is switched on for a
// live ring, every slab fetches those rows from its neighbours' current buffers (all ranks idle: the caller's
// responsibility, as for set_cells).
int pull_halos(lbm_b200* h)
{
  if (h->n_ranks == 1 || !h->connected || h->inplace) return LBM_B200_OK;
  for (Slab& s : h->slabs) {
    CUDA_TRY(cudaSetDevice(s.device));
    const size_t row_bytes = (size_t)h->nx * sizeof(float);
    const float* so = s.south.buf[h->cur];
    const float* no = s.north.buf[h->cur];
    float* me = s.buf[h->cur];
    for (int k = 0; k < 9; k++) {
      const size_t mk = (size_t)k * s.plane, sk = (size_t)k * s.south.plane, nk = (size_t)k * s.north.plane;
      for (int d = 1; d <= lbm::kHalo; d++) {            // the neighbours' rows d rows away (padded rows: see kPad)
        const int to_s = (d == 1) ? 0 : s.rows + 2 * (d - 1), to_n = (d == 1) ? s.rows + 1 : s.rows + 2 * (d - 1) + 1;
        if (s.south.rows >= d)
          CUDA_TRY(cudaMemcpyAsync(me + mk + (size_t)to_s * h->nx, so + sk + (size_t)(s.south.rows + 1 - d) * h->nx, row_bytes, cudaMemcpyDefault, s.stream));
        if (s.north.rows >= d)
          CUDA_TRY(cudaMemcpyAsync(me + mk + (size_t)to_n * h->nx, no + nk + (size_t)d * h->nx, row_bytes, cudaMemcpyDefault, s.stream));
      }
    }
    CUDA_TRY(cudaStreamSynchronize(s.stream));
  }
  return LBM_B200_OK;
}

int finish_create(lbm_b200* h)
{
  plan(h);
  return ensure_partials(h);
}

// argument checks come before the device check, so that they can be exercised on a box without a GPU
int check_common(int nx, int ny, float omega, const void* obstacles, lbm_b200** handle, bool inplace)
{
  if (!handle) return fail(LBM_B200_ERR_ARG, "handle pointer is NULL");
  *handle = nullptr;
  if (!obstacles) return fail(LBM_B200_ERR_ARG, "obstacles pointer is NULL");
  if (nx < 4 || ny < 3) return fail(LBM_B200_ERR_ARG, "grid %dx%d too small (need nx >= 4, ny >= 3)", nx, ny);
  if (!(omega > 0.0f)) return fail(LBM_B200_ERR_ARG, "omega must be positive");
  if (inplace && !(nx % 4 == 0 && nx >= 8))
    return fail(LBM_B200_ERR_ARG, "in-place streaming needs nx %% 4 == 0 and nx >= 8 (got %d)", nx);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
    return fail(LBM_B200_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
  return LBM_B200_OK;
}

void init_common(lbm_b200* h, int nx, int ny, float density, float accel, float omega, float inv)
{
  h->nx = nx; h->ny = ny;
  h->density = density; h->accel = accel; h->omega = omega; h->inv = inv;
  h->sc.omega = omega;
  h->sc.aw1 = density * accel * 0.111111111111111111111111f;    // d2q9-bgk.c:445
  h->sc.aw2 = density * accel * 0.0277777777777777777777778f;   // d2q9-bgk.c:446
  h->sc.negzero = -0.0f;                                         // see lbm::mul2 (csrc/lbm_cell.cuh)
  h->mask_row_words = (nx + 31) / 32;
  if (const char* e = getenv("LBM_B200_KERNEL")) h->opt_kernel = atol(e);
  if (const char* e = getenv("LBM_B200_GRAPH_STEPS")) h->opt_graph_steps = atol(e);
  if (const char* e = getenv("LBM_B200_CTAS_PER_SM")) h->opt_ctas_per_sm = atol(e);
  if (const char* e = getenv("LBM_B200_MIN_CTAS")) h->opt_min_ctas = atol(e);
  if (const char* e = getenv("LBM_B200_CACHE_HINT")) h->opt_cache_hint = atol(e);
  if (const char* e = getenv("LBM_B200_RESIDENT")) h->opt_resident = atol(e);
  if (const char* e = getenv("LBM_B200_FUSED2")) h->opt_fused2 = std::max(-1L, std::min(1L, atol(e)));
  if (const char* e = getenv("LBM_B200_BAND_ROWS")) h->opt_band_rows = std::max(0L, atol(e));
  if (const char* e = getenv("LBM_B200_CLUSTER")) h->opt_cluster = std::max(-1L, std::min(1L, atol(e)));
  if (const char* e = getenv("LBM_B200_CLUSTER_ROWS")) h->opt_cluster_rows = atol(e) != 0;
  if (const char* e = getenv("LBM_B200_FUSED_STEPS")) h->opt_fused_steps = std::max(0L, std::min((long)lbm::kHalo, atol(e)));
  if (const char* e = getenv("LBM_B200_FUSED_DEEP")) h->opt_fused_deep = std::max(-1L, std::min(1L, atol(e)));
  if (const char* e = getenv("LBM_B200_PREFETCH_ROWS")) h->opt_prefetch_rows = std::max(0L, std::min(16L, atol(e)));
}

}  // namespace

// =========================================================================================
extern "C" {

int lbm_b200_abi_version(void) { return LBM_B200_ABI_VERSION; }

const char* lbm_b200_last_error(void) { return g_error.c_str(); }

int lbm_b200_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int lbm_b200_selftest(int device, unsigned long long counts[3])
{
  if (!counts) return fail(LBM_B200_ERR_ARG, "NULL argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return fail(LBM_B200_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(LBM_B200_ERR_ARG, "device %d requested but %d device(s) are visible", device, ndev);
  CUDA_TRY(cudaSetDevice(device));
  unsigned long long* dev = nullptr;
  CUDA_TRY(cudaMalloc(&dev, 3 * sizeof(unsigned long long)));
  cudaError_t e = cudaMemset(dev, 0, 3 * sizeof(unsigned long long));
  if (e == cudaSuccess) {
    lbm::selftest_math<<<148 * 16, 256>>>(dev, 1ull << 28, -0.0f);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(counts, dev, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  cudaFree(dev);
  if (e != cudaSuccess) return fail(LBM_B200_ERR_CUDA, "self-test failed to run: %s", cudaGetErrorString(e));
  return LBM_B200_OK;
}

int lbm_b200_decompose(int ny, int n_slabs, int* rows, int* first_row)
{
  if (ny < 1 || n_slabs < 1 || !rows || !first_row) return fail(LBM_B200_ERR_ARG, "bad arguments to lbm_b200_decompose");
  const int base = ny / n_slabs;
  int extra = ny % n_slabs;
  int grow_last = 0, shrink_second_to_last = 0;
  if (base < 3) {                      // keep the last slab at >= 3 rows (see header)
    grow_last = 1;
    if (extra) extra--;
    else shrink_second_to_last = 1;
  }
  int next = 0;
  for (int i = 0; i < n_slabs; i++) {
    int r = base + (i < extra ? 1 : 0);
    if (i == n_slabs - 2) r -= shrink_second_to_last;
    if (i == n_slabs - 1) r += grow_last;
    rows[i] = r;
    first_row[i] = next;
    next += r;
  }
  for (int i = 0; i < n_slabs; i++)
    if (rows[i] < 3 && n_slabs > 1)
      return fail(LBM_B200_ERR_ARG, "slab %d of %d would have %d rows; at least 3 rows per slab are required", i, n_slabs, rows[i]);
  return LBM_B200_OK;
}

int lbm_b200_plan_bands(int rows, int nx, int band_rows, int sms, int* bands_out, int* rows_per_band)
{
  if (rows < 2 || nx < 4 || sms < 1 || band_rows < 0 || !bands_out || !rows_per_band)
    return fail(LBM_B200_ERR_ARG, "bad arguments to lbm_b200_plan_bands");
  // Automatic height: every work item recomputes two rows, and the items run in waves of one per resident warp
  // (kernel 5's default shape: 3 CTAs x 4 warps per SM) -- take the height with the cheapest waves x (rows + 2.5).
  // Balanced bands; the first and the last one hold at least two rows (a ring slab pushes two rows per direction
  // and publishes them from one work item).
  plan_bands_k(rows, nx, band_rows, sms, 2, 12, false, bands_out, rows_per_band);
  return LBM_B200_OK;
}

int lbm_b200_plan_bands_ex(int rows, int nx, int band_rows, int sms, int steps, int ring, int* bands_out, int* rows_per_band)
{
  if (rows < 2 || nx < 4 || sms < 1 || band_rows < 0 || steps < 2 || steps > lbm::kHalo || !bands_out || !rows_per_band)
    return fail(LBM_B200_ERR_ARG, "bad arguments to lbm_b200_plan_bands_ex");
  if (ring && steps >= 3 && rows < lbm::kHalo + 2)
    return fail(LBM_B200_ERR_ARG, "a ring slab needs at least %d rows for %d timesteps per pass", lbm::kHalo + 2, steps);
  int wpc = 0, ctas = 0;
  if (steps >= 3) stepsk_shape(steps, 1, &wpc, &ctas);
  plan_bands_k(rows, nx, band_rows, sms, steps, steps >= 3 ? wpc * ctas : 12, ring != 0, bands_out, rows_per_band);
  return LBM_B200_OK;
}

float lbm_b200_free_cells_inv(const int* obstacles, long n_cells)
{
  if (!obstacles || n_cells < 0) { fail(LBM_B200_ERR_ARG, "bad arguments to lbm_b200_free_cells_inv"); return 0.0f; }
  long free_cells = n_cells;
  for (long i = 0; i < n_cells; i++)
    if (obstacles[i]) free_cells--;
  // a fully blocked grid gives +inf, exactly as the reference's 1.0f/numOfFreeCells would (d2q9-bgk.c:950)
  return 1.0f / (float)free_cells;
}

static int create_whole(lbm_b200** handle, int nx, int ny, float density, float accel, float omega,
                        const void* obstacles, int format, int n_slabs, const int* devices, bool inplace)
{
  int rc = check_common(nx, ny, omega, obstacles, handle, inplace);
  if (rc) return rc;
  if (format < LBM_B200_OBST_INT32 || format > LBM_B200_OBST_BITS) return fail(LBM_B200_ERR_ARG, "unknown obstacle format %d", format);
  if (n_slabs < 1) return fail(LBM_B200_ERR_ARG, "n_slabs must be >= 1");
  std::vector<int> rows(n_slabs), first(n_slabs);
  rc = lbm_b200_decompose(ny, n_slabs, rows.data(), first.data());
  if (rc) return rc;
  int ndev = 0;
  cudaGetDeviceCount(&ndev);

  lbm_b200* h = new lbm_b200();
  init_common(h, nx, ny, density, accel, omega, 0.0f);   // free_cells_inv follows from the device-side count below
  h->inplace = inplace;
  if (inplace) { h->opt_kernel = 0; if (h->opt_cache_hint > 2) h->opt_cache_hint = 0; }
  h->n_ranks = n_slabs;
  h->slabs.resize(n_slabs);
  for (int i = 0; i < n_slabs; i++) {
    Slab& s = h->slabs[i];
    s.device = devices ? devices[i] : i;
    s.rows = rows[i];
    s.first_row = first[i];
    if (s.device < 0 || s.device >= ndev) {
      lbm_b200_destroy(h);
      return fail(LBM_B200_ERR_ARG, "slab %d asks for device %d but %d device(s) are visible", i, s.device, ndev);
    }
  }
  // one stream per device; slabs that share a device share its stream (lock step)
  for (int i = 0; i < n_slabs; i++) {
    Slab& s = h->slabs[i];
    for (int j = 0; j < i; j++)
      if (h->slabs[j].device == s.device) { s.stream = h->slabs[j].stream; break; }
    if (!s.stream) {
      if (cudaSetDevice(s.device) != cudaSuccess || cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) {
        lbm_b200_destroy(h);
        return fail(LBM_B200_ERR_CUDA, "cannot create a stream on device %d: %s", s.device, cudaGetErrorString(cudaGetLastError()));
      }
      s.own_stream = true;
    }
  }
  const size_t ob_row = obstacle_row_bytes(h, format);
  const char* const ob = static_cast<const char*>(obstacles);
  for (int i = 0; i < n_slabs; i++) {
    rc = alloc_slab(h, h->slabs[i], ob + (size_t)first[i] * ob_row, format);
    if (rc) { lbm_b200_destroy(h); return rc; }
  }
  // 1/free cells (d2q9-bgk.c:945-950) from the blocked-cell counts the packing kernels produced
  {
    long blocked = 0;
    for (Slab& s : h->slabs) {
      unsigned long long n = 0;
      cudaSetDevice(s.device);
      if (cudaMemcpy(&n, s.blocked_dev, sizeof n, cudaMemcpyDeviceToHost) != cudaSuccess) {
        lbm_b200_destroy(h);
        return fail(LBM_B200_ERR_CUDA, "reading the blocked-cell count failed: %s", cudaGetErrorString(cudaGetLastError()));
      }
      blocked += (long)n;
    }
    h->inv = 1.0f / ((long)nx * ny - blocked);
  }
  if (n_slabs > 1) {
    for (int i = 0; i < n_slabs; i++) {
      const int f = first[i], r = rows[i];
      const void* src[kPad];
      for (int j = 0; j < kPad; j++) {
        // word row rows+j holds the row of padded row rows+j, except the first two: rows -1 and `rows`
        const int yy = (j == 0) ? -1 : (j == 1 ? r : y_of_padded(r, r + j));
        src[j] = ob + (size_t)(((f + yy) % ny + ny) % ny) * ob_row;
      }
      rc = set_halo_mask(h, h->slabs[i], src, format);
      if (rc) { lbm_b200_destroy(h); return rc; }
    }
  }
  // ring wiring: direct peer pointers (d2q9-bgk.c:244-247 for the neighbour ranks)
  if (n_slabs > 1) {
    for (int i = 0; i < n_slabs; i++) {
      Slab& s = h->slabs[i];
      Slab& so = h->slabs[(i - 1 + n_slabs) % n_slabs];
      Slab& no = h->slabs[(i + 1) % n_slabs];
      for (Slab* o : {&so, &no}) {
        if (o->device == s.device) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, s.device, o->device);
        if (!can) { lbm_b200_destroy(h); return fail(LBM_B200_ERR_CUDA, "device %d cannot access peer device %d", s.device, o->device); }
        cudaSetDevice(s.device);
        cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
          lbm_b200_destroy(h);
          return fail(LBM_B200_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", s.device, o->device, cudaGetErrorString(e));
        }
        cudaGetLastError();
      }
      s.south.buf[0] = so.buf[0]; s.south.buf[1] = so.buf[1]; s.south.flags = so.flags; s.south.plane = so.plane; s.south.rows = so.rows;
      s.north.buf[0] = no.buf[0]; s.north.buf[1] = no.buf[1]; s.north.flags = no.flags; s.north.plane = no.plane; s.north.rows = no.rows;
    }
  }
  rc = finish_create(h);
  if (rc) { lbm_b200_destroy(h); return rc; }
  *handle = h;
  return LBM_B200_OK;
}

int lbm_b200_create(lbm_b200** handle, int nx, int ny, float density, float accel, float omega,
                    const int* obstacles, int n_slabs, const int* devices)
{
  return create_whole(handle, nx, ny, density, accel, omega, obstacles, LBM_B200_OBST_INT32, n_slabs, devices, false);
}

int lbm_b200_create_inplace(lbm_b200** handle, int nx, int ny, float density, float accel, float omega,
                            const int* obstacles, int n_slabs, const int* devices)
{
  return create_whole(handle, nx, ny, density, accel, omega, obstacles, LBM_B200_OBST_INT32, n_slabs, devices, true);
}

int lbm_b200_create_ex(lbm_b200** handle, int nx, int ny, float density, float accel, float omega,
                       const void* obstacles, int obstacles_format, int n_slabs, const int* devices, int inplace)
{
  return create_whole(handle, nx, ny, density, accel, omega, obstacles, obstacles_format, n_slabs, devices, inplace != 0);
}

static int create_slab(lbm_b200** handle, int nx, int ny_global, int first_row, int rows,
                       int rank, int n_ranks, float density, float accel, float omega,
                       float free_cells_inv, const void* obstacles_slab, int format, int device, bool inplace)
{
  int rc = check_common(nx, ny_global, omega, obstacles_slab, handle, inplace);
  if (rc) return rc;
  if (format < LBM_B200_OBST_INT32 || format > LBM_B200_OBST_BITS) return fail(LBM_B200_ERR_ARG, "unknown obstacle format %d", format);
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(LBM_B200_ERR_ARG, "bad rank %d of %d", rank, n_ranks);
  if (rows < 3 || first_row < 0 || first_row + rows > ny_global) return fail(LBM_B200_ERR_ARG, "bad slab rows [%d, %d) of %d (at least 3 rows)", first_row, first_row + rows, ny_global);
  if (n_ranks == 1 && rows != ny_global) return fail(LBM_B200_ERR_ARG, "a single rank must own the whole grid");
  int ndev = 0;
  cudaGetDeviceCount(&ndev);
  if (device < 0 || device >= ndev) return fail(LBM_B200_ERR_ARG, "device %d requested but %d device(s) are visible", device, ndev);

  lbm_b200* h = new lbm_b200();
  init_common(h, nx, ny_global, density, accel, omega, free_cells_inv);
  h->inplace = inplace;
  if (inplace) { h->opt_kernel = 0; if (h->opt_cache_hint > 2) h->opt_cache_hint = 0; }
  h->n_ranks = n_ranks;
  h->rank0 = rank;
  h->multi_process = n_ranks > 1;
  h->connected = n_ranks == 1;
  h->slabs.resize(1);
  Slab& s = h->slabs[0];
  s.device = device; s.rows = rows; s.first_row = first_row;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) {
    lbm_b200_destroy(h);
    return fail(LBM_B200_ERR_CUDA, "cannot create a stream on device %d", device);
  }
  s.own_stream = true;
  rc = alloc_slab(h, s, obstacles_slab, format);
  if (!rc) rc = finish_create(h);
  if (rc) { lbm_b200_destroy(h); return rc; }
  *handle = h;
  return LBM_B200_OK;
}

int lbm_b200_create_slab(lbm_b200** handle, int nx, int ny_global, int first_row, int rows,
                         int rank, int n_ranks, float density, float accel, float omega,
                         float free_cells_inv, const int* obstacles_slab, int device)
{
  return create_slab(handle, nx, ny_global, first_row, rows, rank, n_ranks, density, accel, omega, free_cells_inv,
                     obstacles_slab, LBM_B200_OBST_INT32, device, false);
}

int lbm_b200_create_slab_inplace(lbm_b200** handle, int nx, int ny_global, int first_row, int rows,
                                 int rank, int n_ranks, float density, float accel, float omega,
                                 float free_cells_inv, const int* obstacles_slab, int device)
{
  return create_slab(handle, nx, ny_global, first_row, rows, rank, n_ranks, density, accel, omega, free_cells_inv,
                     obstacles_slab, LBM_B200_OBST_INT32, device, true);
}

int lbm_b200_create_slab_ex(lbm_b200** handle, int nx, int ny_global, int first_row, int rows,
                            int rank, int n_ranks, float density, float accel, float omega,
                            float free_cells_inv, const void* obstacles_slab, int obstacles_format, int device, int inplace)
{
  return create_slab(handle, nx, ny_global, first_row, rows, rank, n_ranks, density, accel, omega, free_cells_inv,
                     obstacles_slab, obstacles_format, device, inplace != 0);
}

int lbm_b200_ipc_blob_bytes(void) { return (int)sizeof(IpcBlob); }

int lbm_b200_ipc_export(lbm_b200* h, void* blob)
{
  if (!h || !blob) return fail(LBM_B200_ERR_ARG, "NULL argument");
  if (h->slabs.size() != 1) return fail(LBM_B200_ERR_STATE, "ipc export is for slab handles");
  Slab& s = h->slabs[0];
  IpcBlob b{};
  CUDA_TRY(cudaSetDevice(s.device));
  CUDA_TRY(cudaIpcGetMemHandle(&b.buf[0], s.buf[0]));
  if (!h->inplace) CUDA_TRY(cudaIpcGetMemHandle(&b.buf[1], s.buf[1]));
  CUDA_TRY(cudaIpcGetMemHandle(&b.flags, s.flags));
  CUDA_TRY(cudaIpcGetMemHandle(&b.mask, s.mask));
  b.plane = s.plane; b.rows = s.rows; b.device = s.device; b.pid = (int)getpid(); b.inplace = h->inplace ? 1 : 0;
  memcpy(blob, &b, sizeof b);
  return LBM_B200_OK;
}

// unmaps whatever open_neighbour mapped (also after a partial failure)
static void close_neighbour(Neighbour& n)
{
  if (n.ipc) {
    for (int i = 0; i < 2; i++)
      if (n.buf[i]) cudaIpcCloseMemHandle(n.buf[i]);
    if (n.flags) cudaIpcCloseMemHandle(n.flags);
    if (n.mask) cudaIpcCloseMemHandle(n.mask);
  }
  n = Neighbour{};
}

static int open_neighbour(Neighbour& n, const IpcBlob& b)
{
  n = Neighbour{};
  n.ipc = true;                                     // from here on close_neighbour() releases what has been mapped
  for (int i = 0; i < (b.inplace ? 1 : 2); i++) CUDA_TRY(cudaIpcOpenMemHandle((void**)&n.buf[i], b.buf[i], cudaIpcMemLazyEnablePeerAccess));
  CUDA_TRY(cudaIpcOpenMemHandle((void**)&n.flags, b.flags, cudaIpcMemLazyEnablePeerAccess));
  CUDA_TRY(cudaIpcOpenMemHandle((void**)&n.mask, b.mask, cudaIpcMemLazyEnablePeerAccess));
  n.plane = (size_t)b.plane; n.rows = b.rows;
  return LBM_B200_OK;
}

// the part of lbm_b200_ipc_connect that can fail after mappings exist (the caller unmaps them on failure)
static int connect_mapped(lbm_b200* h, Slab& s, const IpcBlob& sb, const IpcBlob& nb)
{
  int rc = open_neighbour(s.south, sb);
  if (rc) return rc;
  if (memcmp(&sb, &nb, sizeof sb) == 0) {           // two ranks: both neighbours are the same slab
    s.north = s.south;
    s.north.ipc = false;
  } else {
    rc = open_neighbour(s.north, nb);
    if (rc) return rc;
  }
  // obstacle words of the neighbours' rows this slab recomputes in the two-steps-per-pass kernel (set_halo_mask)
  const size_t w = (size_t)h->mask_row_words, bytes = w * sizeof(uint32_t);
  for (int d = 1; d <= lbm::kHalo; d++) {
    if (s.south.rows >= d)
      CUDA_TRY(cudaMemcpy(s.mask + (size_t)(s.rows + 2 * (d - 1)) * w, s.south.mask + (size_t)(s.south.rows - d) * w, bytes, cudaMemcpyDefault));
    if (s.north.rows >= d)
      CUDA_TRY(cudaMemcpy(s.mask + (size_t)(s.rows + 2 * (d - 1) + 1) * w, s.north.mask + (size_t)(d - 1) * w, bytes, cudaMemcpyDefault));
  }
  return LBM_B200_OK;
}

int lbm_b200_ipc_connect(lbm_b200* h, const void* south_blob, const void* north_blob)
{
  if (!h || !south_blob || !north_blob) return fail(LBM_B200_ERR_ARG, "NULL argument");
  if (!h->multi_process) return fail(LBM_B200_ERR_STATE, "ipc connect is for slab handles with more than one rank");
  if (h->connected) return fail(LBM_B200_ERR_STATE, "already connected");
  Slab& s = h->slabs[0];
  IpcBlob sb, nb;
  memcpy(&sb, south_blob, sizeof sb);
  memcpy(&nb, north_blob, sizeof nb);
  if ((sb.inplace != 0) != h->inplace || (nb.inplace != 0) != h->inplace)
    return fail(LBM_B200_ERR_ARG, "ring neighbours must all be in-place handles or all ping-pong handles");
  if (sb.rows < 3 || nb.rows < 3 || sb.plane != (unsigned long long)(sb.rows + kPad) * h->nx || nb.plane != (unsigned long long)(nb.rows + kPad) * h->nx)
    return fail(LBM_B200_ERR_ARG, "a neighbour's blob does not describe a slab of this grid (nx %d)", h->nx);
  CUDA_TRY(cudaSetDevice(s.device));
  const int rc = connect_mapped(h, s, sb, nb);
  if (rc) {                                         // unmap what was mapped; g_error keeps the cause
    if (!s.north.ipc) s.north = Neighbour{};        // an alias of the southern mapping (two ranks)
    close_neighbour(s.north);
    close_neighbour(s.south);
    cudaGetLastError();
    return rc;
  }
  h->connected = true;
  return LBM_B200_OK;
}

int lbm_b200_enqueue(lbm_b200* h, int iters)
{
  if (!h) return fail(LBM_B200_ERR_ARG, "NULL handle");
  if (iters < 0) return fail(LBM_B200_ERR_ARG, "iters must be >= 0");
  if (!h->connected) return fail(LBM_B200_ERR_STATE, "slab handle is not connected to its neighbours (lbm_b200_ipc_connect)");
  if (h->failed) return fail(LBM_B200_ERR_STATE, "an earlier run gave up waiting for a ring neighbour; the state is invalid (destroy the handle)");
  if (h->fused2 && h->n_ranks > 1)
    for (const Slab& s : h->slabs)
      if (s.rows < 4 || s.south.rows < 4 || s.north.rows < 4)
        return fail(LBM_B200_ERR_STATE, "the two-steps-per-pass kernel needs at least 4 rows in every slab of a ring (this slab %d, "
                    "south %d, north %d): set the option fused2 = 0 on every rank", s.rows, s.south.rows, s.north.rows);
  if (h->fusedk && h->n_ranks > 1)
    for (const Slab& s : h->slabs)
      if (s.rows < lbm::kHalo + 2 || s.south.rows < lbm::kHalo + 2 || s.north.rows < lbm::kHalo + 2)
        return fail(LBM_B200_ERR_STATE, "the %d-steps-per-pass kernel needs at least %d rows in every slab of a ring (this slab %d, "
                    "south %d, north %d): set the option fused_steps = 2 on every rank", h->fusedk, lbm::kHalo + 2, s.rows,
                    s.south.rows, s.north.rows);
  NvtxRange nvtx_range("lbm_b200 mainloop");
  h->last_iters = iters;
  h->launches = 0;
  bool av_moved = false;
  for (Slab& s : h->slabs) {
    int rc = ensure_av_capacity(s, (size_t)iters, &av_moved);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(s.device));
    CUDA_TRY(cudaMemsetAsync(s.cursor, 0, sizeof(unsigned), s.stream));
  }
  if (av_moved) destroy_graphs(h);                   // they captured the old av_vels pointer
  int glen = (int)h->opt_graph_steps;
  if (h->resident || h->fused2 || h->cluster) glen = 0;
  if (glen < 0) {
    // auto: grids whose step kernel is launch-latency bound (a few microseconds) are replayed
    // from CUDA graphs -- measured 4.1 -> 2.7 us per step on the 128..256-wide decks
    // (per GPU: a ring slab of a small grid is even more launch bound, and one host thread may be feeding
    // several devices)
    const bool small = (long)h->nx * h->ny / h->n_ranks <= kGraphAutoCells;
    glen = (small && iters > 2 * kChunkSteps) ? kChunkSteps : 0;
  }
  if (glen > 0) {
    glen = std::min(glen, kChunkSteps) & ~1;         // even, so a replay preserves the buffer parity
    if (iters - 1 < glen) glen = 0;                  // too short a run to replay even once
    if (glen >= 2 && h->graph_len != glen) {
      const long before = h->launches;
      int rc = build_graphs(h, glen);
      if (rc) return rc;
      h->launches = before;                          // captured, not launched
    }
  }
  for (Slab& s : h->slabs) {
    CUDA_TRY(cudaSetDevice(s.device));
    CUDA_TRY(cudaEventRecord(s.ev_start, s.stream));
  }
  if (iters > 0) {
    // the first step's body force (d2q9-bgk.c:345-348); later ones are folded into the stores
    for (Slab& s : h->slabs) {
      if (s.accel_row < 0) continue;
      CUDA_TRY(cudaSetDevice(s.device));
      lbm::accelerate_row<<<(h->nx + 255) / 256, 256, 0, s.stream>>>(
          s.buf[h->inplace ? 0 : h->cur], layout_of(h, s), s.mask + (size_t)(s.accel_row - 1) * h->mask_row_words,
          s.accel_row, h->sc.aw1, h->sc.aw2);
      CUDA_TRY(cudaGetLastError());
      h->launches++;
    }
    if (h->fused2 && h->n_ranks > 1) {
      // The slab north of the driven row's owner recomputes the owner's last row(s) in the steps before the last of
      // a pass and pulls the driven row out of its own second halo row (kernel 5: planes 5 and 6 of it; kernel 7: the
      // whole row, which it recomputes as its row -2): its copy -- pushed un-forced at the end of the last run --
      // gets the same pre-pass; same inputs, same bits.
      for (Slab& s : h->slabs) {
        if (s.first_row != 0) continue;
        CUDA_TRY(cudaSetDevice(s.device));
        // (ordered against the owner's push of that copy by the strip flags, not by streams: the owner may be
        // another process whose last pass of the previous run is still in flight)
        StepArgs a = base_args(h, s, 0, false);
        peer_args(h, s, a);
        a.wait_from_south = s.flags + kFlagWords;
        lbm::accelerate_halo_row<<<(h->nx + 255) / 256, 256, 0, s.stream>>>(
            s.buf[h->cur], layout_of(h, s), s.mask + (size_t)(s.rows + 2) * h->mask_row_words, s.rows + 2, h->sc.aw1, h->sc.aw2,
            a, h->fused_strips);
        CUDA_TRY(cudaGetLastError());
        h->launches++;
      }
    }
    int t = 0;
    if (h->cluster) {
      // the grid lives in one cluster's shared memory for up to kChunkSteps steps per launch (kernel 6)
      Slab& s = h->slabs[0];
      CUDA_TRY(cudaSetDevice(s.device));
      const size_t bytes = cluster_smem_bytes(h);
      while (t < iters) {
        const int n = std::min(kChunkSteps, iters - t);
        lbm::ClusterArgs ca{};
        ca.in = s.buf[h->cur]; ca.out = s.buf[h->cur ^ 1]; ca.plane = s.plane;
        ca.mask = s.mask; ca.mask_row_words = h->mask_row_words;
        ca.nx = h->nx; ca.rows_per_cta = h->ny / lbm::kClusterCtas;
        ca.steps = n; ca.fold_last = (t + n != iters) ? 1 : 0;
        ca.accel_row = h->ny - 2;
        ca.c = h->sc;
        ca.partials = s.partials; ca.partial_stride = s.per_step;
        CUDA_TRY(cudaMemsetAsync(s.partials, 0, (size_t)n * s.per_step * sizeof(double), s.stream));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(lbm::kClusterCtas); cfg.blockDim = dim3(h->cluster_threads); cfg.dynamicSmemBytes = bytes;
        cfg.stream = s.stream;
        cudaLaunchAttribute attr{};
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = lbm::kClusterCtas; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        if (!h->cluster_rows) CUDA_TRY(cudaLaunchKernelEx(&cfg, lbm::steps_cluster, ca));
        else if (ca.rows_per_cta <= 8) CUDA_TRY(cudaLaunchKernelEx(&cfg, lbm::steps_cluster_rows<8>, ca));
        else CUDA_TRY(cudaLaunchKernelEx(&cfg, lbm::steps_cluster_rows<lbm::kClusterMaxRows>, ca));
        h->launches += 2;
        h->cur ^= 1;                                   // the state always lands in the other buffer
        int rc = enqueue_reduce(h, n);
        if (rc) return rc;
        t += n;
      }
    }
    if (h->resident) {
      // many steps per cooperative launch; the buffers swap roles inside the kernel
      Slab& s = h->slabs[0];
      CUDA_TRY(cudaSetDevice(s.device));
      while (t < iters) {
        const int n = std::min(kChunkSteps, iters - t);
        StepArgs a = base_args(h, s, 0, true);
        a.row_begin = 1; a.row_count = s.rows; a.row_stride = 1;
        a.south_of_first = s.rows;
        a.north_of_last = 1;
        lbm::ResidentArgs r{s.buf[h->cur], s.buf[h->cur ^ 1], n, t + n != iters, s.per_step};
        void* params[] = {&a, &r};
        CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(lbm::steps_resident<2>), dim3(s.grid_full),
                                             dim3(s.threads_full), params, 0, s.stream));
        h->launches++;
        h->cur ^= (n & 1);
        int rc = enqueue_reduce(h, n);
        if (rc) return rc;
        t += n;
      }
    }
    if (glen >= 2 && h->graph_len == glen) {
      while (iters - 1 - t >= glen) {                // the very last step never folds a force in
        for (Slab& s : h->slabs) {
          if (s.own_stream) {
            CUDA_TRY(cudaSetDevice(s.device));
            CUDA_TRY(cudaGraphLaunch(s.graphs[h->cur], s.stream));
          }
          h->launches += (long)glen * ((h->n_ranks == 1 || use_vec4(h)) ? 1 : 2) + 2;
        }
        t += glen;
      }
    }
    while (t < iters) {
      const int n = std::min(kChunkSteps, iters - t);
      if (h->fused2) {
        // fused passes and the single-step tail use different grids: slots are summed over per_step entries
        for (Slab& s : h->slabs) {
          CUDA_TRY(cudaSetDevice(s.device));
          CUDA_TRY(cudaMemsetAsync(s.partials, 0, (size_t)n * s.per_step * sizeof(double), s.stream));
        }
      }
      for (int i = 0; i < n; i++) {
        int rc;
        if (h->fusedk) {                               // up to K steps from t+i on in one pass over HBM
          const int k = std::min(h->fusedk, n - i);
          rc = enqueue_fusedk(h, i, k, t + i + k - 1 != iters - 1);
          i += k - 1;
        } else if (h->fused2 && i + 1 < n) {           // steps t+i and t+i+1 in one pass over HBM
          rc = enqueue_fused2(h, i, t + i + 1 != iters - 1, false);
          i++;
        } else if (h->fused2 && h->n_ranks > 1) {      // a ring's odd step: same strips, same handshake
          rc = enqueue_fused2(h, i, t + i != iters - 1, true);
        } else {
          rc = enqueue_step(h, i, t + i != iters - 1);
        }
        if (rc) return rc;
      }
      int rc = enqueue_reduce(h, n);
      if (rc) return rc;
      t += n;
    }
  }
  for (Slab& s : h->slabs) {
    CUDA_TRY(cudaSetDevice(s.device));
    CUDA_TRY(cudaEventRecord(s.ev_stop, s.stream));
  }
  return LBM_B200_OK;
}

int lbm_b200_sync(lbm_b200* h)
{
  if (!h) return fail(LBM_B200_ERR_ARG, "NULL handle");
  for (Slab& s : h->slabs) {
    CUDA_TRY(cudaSetDevice(s.device));
    CUDA_TRY(cudaStreamSynchronize(s.stream));
  }
  if (h->n_ranks > 1) {
    // did a kernel give up waiting for a neighbour (lbm::spin_until)?  The reference would sit in MPI_Waitall
    // (d2q9-bgk.c:364) for ever; here the run drains and the caller is told which exchange never arrived.
    for (size_t i = 0; i < h->slabs.size(); i++) {
      Slab& s = h->slabs[i];
      unsigned err = 0;
      CUDA_TRY(cudaSetDevice(s.device));
      CUDA_TRY(cudaMemcpy(&err, s.flags + kError, sizeof err, cudaMemcpyDeviceToHost));
      if (err) {
        h->failed = true;
        return fail(LBM_B200_ERR_STATE, "slab %d (rows %d..%d): gave up after %ld ms waiting for halo %s %u from the %s neighbour "
                    "(rank %d); the neighbour is not running or the ranks enqueued different work",
                    h->rank0 + (int)i, s.first_row, s.first_row + s.rows - 1, h->opt_spin_timeout_ms,
                    h->fused2 ? "strip" : "chunk", err >> 8, (err & 3u) == lbm::kWaitFromSouth ? "southern" : "northern",
                    ((err & 3u) == lbm::kWaitFromSouth ? h->rank0 + (int)i - 1 + h->n_ranks : h->rank0 + (int)i + 1) % h->n_ranks);
      }
    }
  }
  return LBM_B200_OK;
}

int lbm_b200_elapsed_ms(lbm_b200* h, float* ms)
{
  if (!h || !ms) return fail(LBM_B200_ERR_ARG, "NULL argument");
  float worst = 0.f;
  for (Slab& s : h->slabs) {
    float t = 0.f;
    CUDA_TRY(cudaSetDevice(s.device));
    CUDA_TRY(cudaEventElapsedTime(&t, s.ev_start, s.ev_stop));
    worst = std::max(worst, t);
  }
  *ms = worst;
  return LBM_B200_OK;
}

int lbm_b200_fetch_av_vels(lbm_b200* h, int iters, float* av_vels)
{
  if (!h || (!av_vels && iters > 0)) return fail(LBM_B200_ERR_ARG, "NULL argument");
  if (iters < 0 || iters > h->last_iters) return fail(LBM_B200_ERR_ARG, "iters exceeds the last enqueue");
  if (iters == 0) return LBM_B200_OK;
  if (int rc = lbm_b200_sync(h)) return rc;          // drains the streams and reports a timed-out halo wait
  std::vector<float> part;
  for (size_t i = 0; i < h->slabs.size(); i++) {
    Slab& s = h->slabs[i];
    CUDA_TRY(cudaSetDevice(s.device));
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    if (i == 0) {
      CUDA_TRY(cudaMemcpy(av_vels, s.av_dev, (size_t)iters * sizeof(float), cudaMemcpyDeviceToHost));
    } else {                                         // rank-order float sum, as MPI_Reduce (396)
      part.resize(iters);
      CUDA_TRY(cudaMemcpy(part.data(), s.av_dev, (size_t)iters * sizeof(float), cudaMemcpyDeviceToHost));
      for (int t = 0; t < iters; t++) av_vels[t] += part[t];
    }
  }
  return LBM_B200_OK;
}

int lbm_b200_run(lbm_b200* h, int iters, float* av_vels)
{
  int rc = lbm_b200_enqueue(h, iters);
  if (rc) return rc;
  rc = lbm_b200_sync(h);
  if (rc) return rc;
  if (av_vels) return lbm_b200_fetch_av_vels(h, iters, av_vels);
  return LBM_B200_OK;
}

int lbm_b200_shape(const lbm_b200* h, int* nx, int* rows, int* first_row)
{
  if (!h) return fail(LBM_B200_ERR_ARG, "NULL handle");
  int total = 0;
  for (const Slab& s : h->slabs) total += s.rows;
  if (nx) *nx = h->nx;
  if (rows) *rows = total;
  if (first_row) *first_row = h->slabs[0].first_row;
  return LBM_B200_OK;
}

int lbm_b200_get_cells(lbm_b200* h, float* cells)
{
  if (!h || !cells) return fail(LBM_B200_ERR_ARG, "NULL argument");
  size_t done = 0;
  for (Slab& s : h->slabs) {
    CUDA_TRY(cudaSetDevice(s.device));
    const size_t row_floats = (size_t)h->nx * 9;
    float* scratch = nullptr;
    int step = 0;
    int rc = staging(h, s, row_floats * sizeof(float), s.rows, &scratch, &step);
    if (rc) return rc;
    for (int r = 0; r < s.rows; r += step) {
      const size_t ncell = (size_t)std::min(step, s.rows - r) * h->nx;
      const unsigned blocks = (unsigned)((ncell * 9 + 255) / 256);
      lbm::soa_to_aos<<<blocks, 256, 0, s.stream>>>(s.buf[h->inplace ? 0 : h->cur], layout_of(h, s), 1 + r, ncell, scratch);
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaMemcpyAsync(cells + done * 9, scratch, ncell * 9 * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
      CUDA_TRY(cudaStreamSynchronize(s.stream));
      done += ncell;
    }
  }
  return LBM_B200_OK;
}

int lbm_b200_set_cells(lbm_b200* h, const float* cells)
{
  if (!h || !cells) return fail(LBM_B200_ERR_ARG, "NULL argument");
  if (h->multi_process) return fail(LBM_B200_ERR_STATE, "set_cells is not available on multi-process slab handles");
  const size_t row_floats = (size_t)h->nx * 9;
  if (h->inplace) { lbm_b200_sync(h); h->cur = 0; }   // the canonical layout is written
  for (Slab& s : h->slabs) {
    CUDA_TRY(cudaSetDevice(s.device));
    float* scratch = nullptr;
    int step = 0;
    int rc = staging(h, s, row_floats * sizeof(float), s.rows + kPad, &scratch, &step);
    if (rc) return rc;
    // owned rows plus all halo rows (see kPad), taken from the periodic global grid
    for (int r0 = 0; r0 < s.rows + kPad; r0 += step) {
      const int n = std::min(step, s.rows + kPad - r0);
      for (int r = r0; r < r0 + n; r++) {
        const int y = y_of_padded(s.rows, r);
        const int gy = ((s.first_row + y) % h->ny + h->ny) % h->ny;
        CUDA_TRY(cudaMemcpyAsync(scratch + (size_t)(r - r0) * row_floats, cells + (size_t)gy * row_floats,
                                 row_floats * sizeof(float), cudaMemcpyHostToDevice, s.stream));
      }
      const size_t ncell = (size_t)n * h->nx;
      const unsigned blocks = (unsigned)((ncell * 9 + 255) / 256);
      lbm::aos_to_soa<<<blocks, 256, 0, s.stream>>>(scratch, s.plane, (size_t)r0 * h->nx, ncell, s.buf[h->inplace ? 0 : h->cur]);
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaStreamSynchronize(s.stream));
    }
  }
  return LBM_B200_OK;
}

int lbm_b200_get_final_state(lbm_b200* h, float* u_x, float* u_y, float* u, float* pressure)
{
  if (!h) return fail(LBM_B200_ERR_ARG, "NULL handle");
  float* outs[4] = {u_x, u_y, u, pressure};
  size_t done = 0;
  for (Slab& s : h->slabs) {
    CUDA_TRY(cudaSetDevice(s.device));
    float* scratch = nullptr;
    int step = 0;
    int rc = staging(h, s, (size_t)h->nx * 4 * sizeof(float), s.rows, &scratch, &step);
    if (rc) return rc;
    if (step >= s.rows && s.rows >= 64) {
      // The scratch holds the whole slab (ping-pong handles: the idle buffer): cut it into row chunks and let the
      // epilogue kernel of chunk c+1 run while chunk c is on its way to the host -- kernels on the slab's stream,
      // copies on a second stream, one event per chunk.  (The reference computes these fields on the host while it
      // prints them, d2q9-bgk.c:1076-1111.)
      constexpr int kChunks = 8;
      if (!s.copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&s.copy_stream, cudaStreamNonBlocking));
      cudaEvent_t ev[kChunks] = {};
      const int rows_per = (s.rows + kChunks - 1) / kChunks;
      cudaError_t e = cudaSuccess;
      size_t off = 0;                                 // cells of the slab done so far
      for (int c = 0; c < kChunks && e == cudaSuccess; c++) {
        const int r0 = c * rows_per, n = std::min(rows_per, s.rows - r0);
        if (n <= 0) break;
        const size_t ncell = (size_t)n * h->nx;
        float* part = scratch + 4 * off;
        lbm::final_state<<<(unsigned)((ncell + 255) / 256), 256, 0, s.stream>>>(s.buf[h->inplace ? 0 : h->cur], layout_of(h, s), 1 + r0,
                                                                                s.mask, h->mask_row_words, ncell, h->density, part);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[c], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(ev[c], s.stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(s.copy_stream, ev[c], 0);
        for (int k = 0; k < 4 && e == cudaSuccess; k++)
          if (outs[k]) e = cudaMemcpyAsync(outs[k] + done + off, part + (size_t)k * ncell, ncell * sizeof(float), cudaMemcpyDeviceToHost, s.copy_stream);
        off += ncell;
      }
      cudaError_t e2 = cudaStreamSynchronize(s.copy_stream);
      cudaError_t e3 = cudaStreamSynchronize(s.stream);
      for (cudaEvent_t x : ev)
        if (x) cudaEventDestroy(x);
      if (e == cudaSuccess) e = (e2 != cudaSuccess) ? e2 : e3;
      if (e != cudaSuccess) return fail(LBM_B200_ERR_CUDA, "reading back the macroscopic fields failed: %s", cudaGetErrorString(e));
      done += (size_t)s.rows * h->nx;
      continue;
    }
    for (int r = 0; r < s.rows; r += step) {
      const size_t ncell = (size_t)std::min(step, s.rows - r) * h->nx;
      const unsigned blocks = (unsigned)((ncell + 255) / 256);
      lbm::final_state<<<blocks, 256, 0, s.stream>>>(s.buf[h->inplace ? 0 : h->cur], layout_of(h, s), 1 + r, s.mask,
                                                    h->mask_row_words, ncell, h->density, scratch);
      CUDA_TRY(cudaGetLastError());
      for (int k = 0; k < 4; k++)
        if (outs[k])
          CUDA_TRY(cudaMemcpyAsync(outs[k] + done, scratch + (size_t)k * ncell, ncell * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
      CUDA_TRY(cudaStreamSynchronize(s.stream));
      done += ncell;
    }
  }
  return LBM_B200_OK;
}

int lbm_b200_set_option(lbm_b200* h, const char* key, long value)
{
  if (!h || !key) return fail(LBM_B200_ERR_ARG, "NULL argument");
  if (h->inplace && (!strcmp(key, "kernel") || !strcmp(key, "min_ctas") || !strcmp(key, "resident")) )
    return fail(LBM_B200_ERR_STATE, "option '%s' does not apply to an in-place handle", key);
  if (!strcmp(key, "kernel")) {
    if (value < 0 || value > 2) return fail(LBM_B200_ERR_ARG, "kernel must be 0, 1 or 2");
    if (value == 2 && !(h->nx % 4 == 0 && h->nx >= 8)) return fail(LBM_B200_ERR_ARG, "kernel 2 needs nx %% 4 == 0 and nx >= 8");
    h->opt_kernel = value;
  } else if (!strcmp(key, "graph_steps")) {
    if (value < -1) return fail(LBM_B200_ERR_ARG, "graph_steps must be >= -1");
    h->opt_graph_steps = value;
  } else if (!strcmp(key, "ctas_per_sm")) {
    if (value < 0 || value > 65536) return fail(LBM_B200_ERR_ARG, "ctas_per_sm must be 0..65536");
    h->opt_ctas_per_sm = value;
  } else if (!strcmp(key, "min_ctas")) {
    if (value < 2 || value > 4) return fail(LBM_B200_ERR_ARG, "min_ctas must be 2, 3 or 4");
    h->opt_min_ctas = value;
  } else if (!strcmp(key, "resident")) {
    if (value < -1 || value > 1) return fail(LBM_B200_ERR_ARG, "resident must be -1, 0 or 1");
    h->opt_resident = value;
  } else if (!strcmp(key, "fused2")) {
    if (value < -1 || value > 1) return fail(LBM_B200_ERR_ARG, "fused2 must be -1, 0 or 1");
    h->opt_fused2 = value;
  } else if (!strcmp(key, "cluster")) {
    if (value < -1 || value > 1) return fail(LBM_B200_ERR_ARG, "cluster must be -1, 0 or 1");
    h->opt_cluster = value;
  } else if (!strcmp(key, "fused_k7")) {
    if (value < 0 || value > 1) return fail(LBM_B200_ERR_ARG, "fused_k7 must be 0 or 1");
    h->opt_fused_k7 = value;
  } else if (!strcmp(key, "fused_ctas")) {
    if (value < 0 || value > 12) return fail(LBM_B200_ERR_ARG, "fused_ctas must be 0 (automatic) .. 12");
    h->opt_fused_ctas = value;
  } else if (!strcmp(key, "cluster_rows")) {
    if (value < 0 || value > 1) return fail(LBM_B200_ERR_ARG, "cluster_rows must be 0 or 1");
    h->opt_cluster_rows = value;
  } else if (!strcmp(key, "fused_steps")) {
    if (value != 0 && (value < 2 || value > lbm::kHalo)) return fail(LBM_B200_ERR_ARG, "fused_steps must be 0 (automatic), 2, 3 or 4");
    h->opt_fused_steps = value;
  } else if (!strcmp(key, "band_rows")) {
    if (value < 0 || value > (1 << 20)) return fail(LBM_B200_ERR_ARG, "band_rows must be 0 (automatic) .. 2^20");
    h->opt_band_rows = value;
  } else if (!strcmp(key, "staging_bytes")) {
    if (value < 1) return fail(LBM_B200_ERR_ARG, "staging_bytes must be positive");
    h->opt_staging_bytes = value;
  } else if (!strcmp(key, "fused_deep")) {
    if (value < -1 || value > 1) return fail(LBM_B200_ERR_ARG, "fused_deep must be -1, 0 or 1");
    h->opt_fused_deep = value;
  } else if (!strcmp(key, "prefetch_rows")) {
    if (value < 0 || value > 16) return fail(LBM_B200_ERR_ARG, "prefetch_rows must be 0 .. 16");
    h->opt_prefetch_rows = value;
  } else if (!strcmp(key, "spin_timeout_ms")) {
    if (value < 1 || value > 3600000) return fail(LBM_B200_ERR_ARG, "spin_timeout_ms must be 1 .. 3600000");
    h->opt_spin_timeout_ms = value;
  } else if (!strcmp(key, "debug_skip_slab")) {
    if (value < -1 || value >= (long)h->slabs.size()) return fail(LBM_B200_ERR_ARG, "debug_skip_slab must be -1 or a slab index of this handle");
    h->opt_debug_skip_slab = value;
  } else if (!strcmp(key, "cache_hint")) {
    if (value < 0 || value > 4 || value == 3) return fail(LBM_B200_ERR_ARG, "cache_hint must be 0, 1, 2 or 4");
    if (h->inplace && value > 2) return fail(LBM_B200_ERR_ARG, "cache_hint of an in-place handle must be 0, 1 or 2");
    h->opt_cache_hint = value;
  } else {
    return fail(LBM_B200_ERR_ARG, "unknown option '%s'", key);
  }
  lbm_b200_sync(h);
  destroy_graphs(h);
  const bool was_fused = h->fused2;
  const int was_fusedk = h->fusedk;
  plan(h);
  if (!strcmp(key, "kernel") || !strcmp(key, "fused2")) {
    int rc = resync_flags(h);
    if (rc) return rc;
  }
  if ((h->fused2 && !was_fused) || (h->fusedk && !was_fusedk)) {   // whichever option brought a fused kernel (or its deeper halo) back on a ring
    int rc = pull_halos(h);
    if (rc) return rc;
  }
  return ensure_partials(h);
}

int lbm_b200_get_option(const lbm_b200* h, const char* key, long* value)
{
  if (!h || !key || !value) return fail(LBM_B200_ERR_ARG, "NULL argument");
  if (!strcmp(key, "kernel")) *value = h->inplace ? 4 : (h->cluster ? 6 : (h->fused2 ? (h->fusedk ? 7 : 5) : (h->resident ? 3 : (use_vec4(h) ? 2 : 1))));
  else if (!strcmp(key, "fused_steps")) *value = h->fused2 ? (h->fusedk ? h->fusedk : 2) : 1;
  else if (!strcmp(key, "cluster")) *value = h->cluster ? 1 : 0;
  else if (!strcmp(key, "fused_ctas")) *value = h->opt_fused_ctas;
  else if (!strcmp(key, "fused_k7")) *value = h->opt_fused_k7;
  else if (!strcmp(key, "cluster_rows")) *value = (h->cluster && h->cluster_rows) ? 1 : 0;
  else if (!strcmp(key, "fused2")) *value = h->fused2 ? 1 : 0;
  else if (!strcmp(key, "band_rows")) *value = h->fused2 ? h->slabs[0].fused_band_rows : h->opt_band_rows;
  else if (!strcmp(key, "inplace")) *value = h->inplace ? 1 : 0;
  else if (!strcmp(key, "staging_bytes")) *value = h->opt_staging_bytes;
  else if (!strcmp(key, "prefetch_rows")) *value = h->opt_prefetch_rows;
  else if (!strcmp(key, "fused_deep")) *value = stage_rows(h) - 1;
  else if (!strcmp(key, "spin_timeout_ms")) *value = h->opt_spin_timeout_ms;
  else if (!strcmp(key, "debug_skip_slab")) *value = h->opt_debug_skip_slab;
  else if (!strcmp(key, "graph_steps")) *value = h->opt_graph_steps;
  else if (!strcmp(key, "ctas_per_sm")) *value = h->opt_ctas_per_sm;
  else if (!strcmp(key, "min_ctas")) *value = h->opt_min_ctas;
  else if (!strcmp(key, "cache_hint")) *value = h->opt_cache_hint;
  else if (!strcmp(key, "resident")) *value = h->resident ? 1 : 0;
  else if (!strcmp(key, "grid")) *value = h->slabs[0].per_step;
  else if (!strcmp(key, "threads")) *value = h->slabs[0].grid_full ? h->slabs[0].threads_full : h->slabs[0].threads_int;
  else if (!strcmp(key, "launches_per_step")) *value = (h->n_ranks == 1 || use_vec4(h)) ? 1 : 2;
  else if (!strcmp(key, "launches")) *value = h->launches;
  else return fail(LBM_B200_ERR_ARG, "unknown option '%s'", key);
  return LBM_B200_OK;
}

void lbm_b200_destroy(lbm_b200* h)
{
  if (!h) return;
  destroy_graphs(h);
  // slabs may share a stream: drain everything first, free memory, destroy streams last
  for (Slab& s : h->slabs) {
    cudaSetDevice(s.device);
    if (s.stream) cudaStreamSynchronize(s.stream);
  }
  for (Slab& s : h->slabs) {
    cudaSetDevice(s.device);
    if (!s.north.ipc) s.north = Neighbour{};        // peer pointers, or an alias of the southern IPC mapping
    close_neighbour(s.north);
    close_neighbour(s.south);
    for (int b = 0; b < 2; b++)
      if (s.buf[b]) cudaFree(s.buf[b]);
    if (s.mask) cudaFree(s.mask);
    if (s.flags) cudaFree(s.flags);
    if (s.partials) cudaFree(s.partials);
    if (s.av_dev) cudaFree(s.av_dev);
    if (s.cursor) cudaFree(s.cursor);
    if (s.blocked_dev) cudaFree(s.blocked_dev);
    if (s.bounce) cudaFree(s.bounce);
    if (s.copy_stream) cudaStreamDestroy(s.copy_stream);
    if (s.ev_start) cudaEventDestroy(s.ev_start);
    if (s.ev_stop) cudaEventDestroy(s.ev_stop);
  }
  for (Slab& s : h->slabs) {
    if (!s.own_stream || !s.stream) continue;
    cudaSetDevice(s.device);
    cudaStreamDestroy(s.stream);
  }
  cudaGetLastError();
  delete h;
}

}  // extern "C"
