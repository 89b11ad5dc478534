// K2 / K3 / ridge field — np.gradient-exact stencils fused with the 2x2 eigen epilogues.
//
// K2 (FTLE): replaces R:src/dynlab/diagnostics/lagrangian.py:50-78 (and :130-167 with Xi_max)
// K3 (Eulerian): replaces R:src/dynlab/diagnostics/_base_classes.py:110-111 and the per-point
//     loops of R:src/dynlab/diagnostics/eulerian.py:47-59, :159-195, :317-343
// ridge: replaces R:src/dynlab/diagnostics/_base_classes.py:216-268 (up to the contour call)
// np.gradient semantics: numpy/lib/_function_base_impl.py:1246-1384 — per axis, uniform iff all
// coordinate differences equal the first one exactly (:1246-1251); uniform interior
// (f[i+1]-f[i-1])/(2h) (:1305), non-uniform interior a f[i-1] + b f[i] + c f[i+1] (:1307-1318),
// edge_order 1 two-point (:1321-1334), edge_order 2 three-point (:1337-1372).
//
// B200 mapping (24 B per point in, 8 out).  Three interior kernels, newest first:
//   * stencil_tma_kernel (default): TMA tile copies (cp.async.bulk.tensor.2d) into a staged
//     shared-memory ring, 128 threads = 128 tile columns, rows of a stage fully unrolled,
//     branch-free epilogue with predicated stores;
//   * stencil_bulk_kernel: 1-D bulk copies per row into an 8-row ring (fallback when no tensor
//     map can be encoded; DLB_STENCIL_TMA=0);
//   * stencil_interior_kernel: one warp walks 32 columns with a renamed register window and warp
//     shuffles (unaligned rows: SoA fields with odd xdim; DLB_STENCIL_BULK=0).
// The O(xdim + ydim) boundary nodes — numpy's one-sided formulas — run in a separate small edge
// kernel, which keeps the interior loops free of edge logic.  Coefficients are precomputed per
// row / per column on the host with numpy's own formulas (true divisions) and cached in the
// workspace, so the kernels have no FP64 divides on the interior path; sqrt and log are
// branch-free MUFU-seeded Newton / table versions.  Device helpers shared with eulerflow.cu live
// in stencil_device.cuh.
#include <math.h>
#include <math_constants.h>

#include <mutex>
#include <type_traits>
#include <vector>

#include <cuda.h>  // CUtensorMap types only; the encoder is fetched through the runtime

#include "stencil_device.cuh"

using dlbst::check_slab;
using dlbst::env_flag;
using dlbst::num_sms;
using dlbst::setup_axes;

namespace dlbst {

int env_flag(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// SM count of the CURRENT device (one process may drive several devices: per-device table)
int num_sms() {
  static int sms_of[DLB_MAX_DEVICES] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int& sms = sms_of[dev & (DLB_MAX_DEVICES - 1)];
  if (sms == 0 &&
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    sms = 148;
  return sms;
}

}  // namespace dlbst

namespace {

// Interior kernel: output nodes with all four neighbours (global rows 1..ydim-2, columns
// 1..xdim-2).  One warp walks 32 consecutive columns down ST_ROWS rows with the rows i-1, i,
// i+1 (and the prefetched i+2) of its column in registers: every element is requested from
// L2/DRAM once per warp, as one coalesced row segment.  The x neighbours come from the
// adjacent lanes by warp shuffle — lanes 0 and 31 only carry the halo columns, so a warp emits
// 30 columns (an ncu run of the load-based variant showed the L/R loads missing L1 43 % of the
// time and the kernel stalled on them: long_scoreboard 7.3 per issue).
constexpr int ST_COLS = 30;
#ifndef DLB_ST_W
#define DLB_ST_W 4
#endif
constexpr int ST_W = DLB_ST_W;  // rows held in registers: 3 in use + (ST_W - 3) prefetched

__device__ __forceinline__ double2 shfl_up2(double2 v) {
  return make_double2(__shfl_up_sync(FULL, v.x, 1), __shfl_up_sync(FULL, v.y, 1));
}
__device__ __forceinline__ double2 shfl_down2(double2 v) {
  return make_double2(__shfl_down_sync(FULL, v.x, 1), __shfl_down_sync(FULL, v.y, 1));
}

#ifndef DLB_ST_MINB
#define DLB_ST_MINB 4
#endif
template <int KIND, int MODE, bool XI, bool UX, bool UY, class Loader>
__global__ void __launch_bounds__(ST_BLOCK, DLB_ST_MINB)
    stencil_interior_kernel(const StencilParams P, const Loader ld) {
  const long long xdim = P.ax.n, ydim = P.ay.n;
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * ST_BLOCK + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * ST_BLOCK) >> 5;
  // interior rows of the requested range, interior columns
  const long long r_lo = (P.out_row0 > 1) ? P.out_row0 : 1;
  const long long r_hi = (P.out_row0 + P.out_rows < ydim - 1) ? P.out_row0 + P.out_rows : ydim - 1;
  const long long ncols = xdim - 2;
  if (r_hi <= r_lo || ncols <= 0) return;
  const long long nstrips = (ncols + ST_COLS - 1) / ST_COLS;
  const long long nchunks = (r_hi - r_lo + ST_ROWS - 1) / ST_ROWS;
  const long long ntasks = nstrips * nchunks;
  const double ix = P.ax.inv2h, iy = P.ay.inv2h;
  const int sx = (int)xdim;  // row stride in elements
  __shared__ double2 ysm[ST_BLOCK / 32][ST_ROWS][2];
  const int wib = threadIdx.x >> 5;

  for (long long task = warp; task < ntasks; task += nwarps) {  // warp-uniform trip count
    const long long chunk = task / nstrips;
    const long long strip = task - chunk * nstrips;
    const long long jj = strip * ST_COLS + lane;           // lane 0 / 31: halo columns
    const long long j = (jj < xdim) ? jj : xdim - 1;       // clamp the loads of tail lanes
    const bool emit_lane = (lane >= 1) && (lane <= ST_COLS) && (jj <= xdim - 2);
    const long long g0 = r_lo + chunk * ST_ROWS;
    const int nrows = (int)((g0 + ST_ROWS < r_hi) ? ST_ROWS : r_hi - g0);
    double xa = 0.0, xb = 0.0, xc = 0.0;
    if (!UX) {
      xa = __ldg(P.ax.ca + j);
      xb = __ldg(P.ax.cb + j);
      xc = __ldg(P.ax.cc + j);
    }
    const typename Loader::Cursor cur = ld.cursor(g0 - P.grow0, j);
    const long long obase = (g0 - P.out_row0) * xdim + j;
    // y coefficients of the task's rows -> shared memory, one 32-byte record per row (lane l
    // fetches row l).  Read straight from global they were consumed at once and, with the
    // streaming data evicting them from L1, cost ~35 % of all stall samples (long_scoreboard).
    if (!UY) {
      __syncwarp();  // previous task's readers are done
      if (lane < nrows) {
        const double2* src = P.ay4 + 2 * (g0 + lane);
        ysm[wib][lane][0] = __ldg(src);
        ysm[wib][lane][1] = __ldg(src + 1);
      }
      __syncwarp();
    }
    // Rolling window over rows: ST_W registers that rotate by *renaming* (the loop is unrolled
    // ST_W times, so every index below is static): r[s] = row k-1, r[s+1] = row k, r[s+2] =
    // row k+1, r[s+3..] = prefetched rows; the load issued in step k is row k+ST_W-1 and is
    // first used ST_W-3 steps later.  No register move ever touches a load that is still in
    // flight (a rotating `D = Dn` copy stalled on long_scoreboard for 45 % of all samples in
    // the first ncu capture).
    double2 r[ST_W];
#pragma unroll
    for (int i = 0; i < ST_W; ++i) {
      const int row = i - 1;  // relative to g0; rows up to nrows are inside the grid
      r[i] = cur.load(((row <= nrows) ? row : nrows) * sx);
    }
    int k = 0;      // row within the task
    int off = 0;    // k * sx
    auto step = [&](int s) {  // s is a compile-time constant after unrolling
      const double2 U = r[s % ST_W];
      const double2 C = r[(s + 1) % ST_W], D = r[(s + 2) % ST_W];
      if (k + ST_W - 1 <= nrows) r[s % ST_W] = cur.load(off + (ST_W - 1) * sx);
      const double2 L = shfl_up2(C), R = shfl_down2(C);
      double d0dx, d1dx, d0dy, d1dy;
      if (UX) {
        d0dx = (R.x - L.x) * ix;
        d1dx = (R.y - L.y) * ix;
      } else {
        d0dx = fma(xc, R.x, fma(xb, C.x, xa * L.x));
        d1dx = fma(xc, R.y, fma(xb, C.y, xa * L.y));
      }
      if (UY) {
        d0dy = (D.x - U.x) * iy;
        d1dy = (D.y - U.y) * iy;
      } else {
        const double2 yab = ysm[wib][k][0], ycz = ysm[wib][k][1];
        d0dy = fma(ycz.x, D.x, fma(yab.y, C.x, yab.x * U.x));
        d1dy = fma(ycz.x, D.y, fma(yab.y, C.y, yab.x * U.y));
      }
      if (emit_lane) emit<KIND, MODE, XI>(P, C, d0dx, d1dx, d0dy, d1dy, obase + off);
      off += sx;
      ++k;
    };
    while (k + ST_W <= nrows) {
#pragma unroll
      for (int s = 0; s < ST_W; ++s) step(s);
    }
#pragma unroll
    for (int s = 0; s < ST_W - 1; ++s)
      if (k < nrows) step(s);
  }
}

// ---------------------------------------------------------------------------------------------
// Bulk-copy (TMA engine) variant of the interior kernel.  The register-window kernel above is
// latency-bound (ncu: long_scoreboard dominates, DRAM ~55 %): a warp can only keep as many row
// segments in flight as it has registers for.  Here one elected thread per CTA streams row
// segments into a shared-memory ring with `cp.async.bulk` (SASS: UBLKCP) completing on
// mbarriers, so the bytes in flight (TB_SLOTS - 2 rows x 2 KB per CTA, 8 CTAs per SM) no longer
// depend on registers; the 128 threads of the CTA (one per loaded column, the two outermost
// are halo-only) read neighbours straight from shared memory.
// Requirements: 16-byte aligned row segments — always true for the AoS flow map, true for the
// SoA velocity fields when xdim is even (otherwise the register-window kernel is used).
constexpr int TB_THREADS = 128;          // loaded columns per row segment = threads per CTA
constexpr int TB_OUT = TB_THREADS - 2;   // output columns per CTA strip
#ifndef DLB_TB_SLOTS
#define DLB_TB_SLOTS 8
#endif
#ifndef DLB_TB_ROWS
#define DLB_TB_ROWS 64
#endif
constexpr int TB_SLOTS = DLB_TB_SLOTS;    // rows in the shared-memory ring
constexpr int TB_ROWS = DLB_TB_ROWS;      // output rows per task

__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes,
                                         unsigned long long* b) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(b))
      : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* b, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(b)), "r"(parity)
      : "memory");
  return ok != 0;
}

// ring row = 2048 bytes: AoS 128 x double2, SoA 128 x u then 128 x v
struct RingAoS {
  static __device__ __forceinline__ void issue(const LoaderAoS& ld, double* slot, long long lrow,
                                               long long c0, int ncol, unsigned long long* bar) {
    mbar_expect_tx(bar, (unsigned)ncol * 16u);
    bulk_g2s(slot, ld.p + lrow * ld.xdim + c0, (unsigned)ncol * 16u, bar);
  }
  static __device__ __forceinline__ double2 read(const double* slot, int t) {
    return reinterpret_cast<const double2*>(slot)[t];
  }
};
struct RingSoA {
  static __device__ __forceinline__ void issue(const LoaderSoA& ld, double* slot, long long lrow,
                                               long long c0, int ncol, unsigned long long* bar) {
    mbar_expect_tx(bar, (unsigned)ncol * 16u);
    const long long k = lrow * ld.xdim + c0;
    bulk_g2s(slot, ld.u + k, (unsigned)ncol * 8u, bar);
    bulk_g2s(slot + TB_THREADS, ld.v + k, (unsigned)ncol * 8u, bar);
  }
  static __device__ __forceinline__ double2 read(const double* slot, int t) {
    return make_double2(slot[t], slot[TB_THREADS + t]);
  }
};

#ifndef DLB_TB_MINB_XI
#define DLB_TB_MINB_XI 6  // the eigenvector epilogue (dlanv2 path) needs more registers
#endif
template <int KIND, int MODE, bool XI, bool UX, bool UY, class Loader, class Ring>
__global__ void __launch_bounds__(TB_THREADS, XI ? DLB_TB_MINB_XI : 8)
    stencil_bulk_kernel(const StencilParams P, const Loader ld) {
  __shared__ __align__(128) double ring[TB_SLOTS][2 * TB_THREADS];
  __shared__ __align__(8) unsigned long long full[TB_SLOTS];
  __shared__ double2 ysm[TB_ROWS][2];
  __shared__ double2 logtab_s[KIND == KIND_FTLE ? 128 : 1];
  const long long xdim = P.ax.n, ydim = P.ay.n;
  const int t = threadIdx.x;
  if (KIND == KIND_FTLE) logtab_s[t] = g_logtab[t];  // TB_THREADS == 128 entries
  const long long r_lo = (P.out_row0 > 1) ? P.out_row0 : 1;
  const long long r_hi = (P.out_row0 + P.out_rows < ydim - 1) ? P.out_row0 + P.out_rows : ydim - 1;
  const long long ncols_out = xdim - 2;
  if (r_hi <= r_lo || ncols_out <= 0) return;
  const long long nstrips = (ncols_out + TB_OUT - 1) / TB_OUT;
  const long long nchunks = (r_hi - r_lo + TB_ROWS - 1) / TB_ROWS;
  const long long ntasks = nstrips * nchunks;
  const double ix = P.ax.inv2h, iy = P.ay.inv2h;
  if (t == 0) {
#pragma unroll
    for (int i = 0; i < TB_SLOTS; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned nbase = 0;  // rows that went through the ring before this task (CTA-uniform)

  for (long long task = blockIdx.x; task < ntasks; task += gridDim.x) {
    const long long chunk = task / nstrips;
    const long long strip = task - chunk * nstrips;
    const long long c0 = strip * TB_OUT;  // first loaded column (even)
    const int ncol = (int)((c0 + TB_THREADS <= xdim) ? TB_THREADS : xdim - c0);
    const long long g0 = r_lo + chunk * TB_ROWS;
    const int nrows = (int)((g0 + TB_ROWS < r_hi) ? TB_ROWS : r_hi - g0);
    const int nload = nrows + 2;             // rows g0-1 .. g0+nrows
    const long long lrow0 = g0 - 1 - P.grow0;  // local row of the first loaded row
    const int tc = (t < ncol) ? t : ncol - 1;  // clamped column of this thread inside the segment
    const long long col = c0 + tc;
    const bool emit_thread = (t >= 1) && (t <= TB_OUT) && (c0 + t <= xdim - 2);

    __syncthreads();  // the previous task's ring and ysm are no longer read
    if (t == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const int first = (nload < TB_SLOTS) ? nload : TB_SLOTS;
      for (int i = 0; i < first; ++i)
        Ring::issue(ld, ring[(nbase + i) % TB_SLOTS], lrow0 + i, c0, ncol, &full[(nbase + i) % TB_SLOTS]);
    }
    if (!UY && t < nrows) {
      const double2* src = P.ay4 + 2 * (g0 + t);
      ysm[t][0] = __ldg(src);
      ysm[t][1] = __ldg(src + 1);
    }
    double xa = 0.0, xb = 0.0, xc = 0.0;
    if (!UX) {
      xa = __ldg(P.ax.ca + col);
      xb = __ldg(P.ax.cb + col);
      xc = __ldg(P.ax.cc + col);
    }
    __syncthreads();  // ysm visible

    auto wait_row = [&](int i) {  // loaded row i of this task has landed
      const unsigned n = nbase + (unsigned)i;
      unsigned long long* b = &full[n % TB_SLOTS];
      const unsigned parity = (n / TB_SLOTS) & 1u;
      while (!mbar_try_wait(b, parity)) {
      }
      return ring[n % TB_SLOTS];
    };
    double2 U = Ring::read(wait_row(0), tc);
    const double* slotC = wait_row(1);
    double2 C = Ring::read(slotC, tc);
    long long o = (g0 - P.out_row0) * xdim + col;
    for (int k = 0; k < nrows; ++k) {
      const double* slotD = wait_row(k + 2);
      const double2 D = Ring::read(slotD, tc);
      const double2 L = Ring::read(slotC, (tc > 0) ? tc - 1 : 0);
      const double2 R = Ring::read(slotC, (tc < ncol - 1) ? tc + 1 : ncol - 1);
      double d0dx, d1dx, d0dy, d1dy;
      if (UX) {
        d0dx = (R.x - L.x) * ix;
        d1dx = (R.y - L.y) * ix;
      } else {
        d0dx = fma(xc, R.x, fma(xb, C.x, xa * L.x));
        d1dx = fma(xc, R.y, fma(xb, C.y, xa * L.y));
      }
      if (UY) {
        d0dy = (D.x - U.x) * iy;
        d1dy = (D.y - U.y) * iy;
      } else {
        const double2 yab = ysm[k][0], ycz = ysm[k][1];
        d0dy = fma(ycz.x, D.x, fma(yab.y, C.x, yab.x * U.x));
        d1dy = fma(ycz.x, D.y, fma(yab.y, C.y, yab.x * U.y));
      }
      if (emit_thread)
        emit<KIND, MODE, XI>(P, C, d0dx, d1dx, d0dy, d1dy, o, KIND == KIND_FTLE ? logtab_s : nullptr);
      o += xdim;
      U = C;
      C = D;
      slotC = slotD;
      // Loaded rows <= k+1 are dead in shared memory once every thread has passed this point
      // (row k+2 is still read as the x-neighbour row of the next iteration): refill the slot
      // of row k with row k+TB_SLOTS, i.e. the ring runs TB_SLOTS-3 rows ahead of the compute.
      __syncthreads();
      if (t == 0 && k + TB_SLOTS < nload) {
        const unsigned n = nbase + (unsigned)(k + TB_SLOTS);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        Ring::issue(ld, ring[n % TB_SLOTS], lrow0 + k + TB_SLOTS, c0, ncol, &full[n % TB_SLOTS]);
      }
    }
    nbase += (unsigned)nload;
  }
}

// ---------------------------------------------------------------------------------------------
// TMA tile variant (default).  The ncu captures of the bulk kernel above showed that it is not
// the DRAM pipe that limits it but the issue slots: per 126-point row a CTA paid a barrier, an
// mbarrier wait with run-time slot / parity arithmetic and, in warp 0, one or two 1-D bulk-copy
// issues — 88 non-FP64 instructions per warp row against 31 FP64 (2 * FP64 + other = the measured
// time).  Here the TMA engine moves whole [TM_RT rows x 128 columns] tiles with ONE
// `cp.async.bulk.tensor.2d` per array (SASS: UTMALDG) described by a CUtensorMap, the pipeline
// is counted in stages instead of rows (one mbarrier wait and one __syncthreads per 8 rows, all
// shared-memory offsets inside a stage are compile-time constants because the row loop is fully
// unrolled) and it runs across task boundaries, so a new task does not start with an empty pipe.
// Out-of-range rows / columns of a tile are zero-filled by the TMA unit and never read back.
constexpr int TM_THREADS = 128;                 // tile width = threads per CTA
constexpr int TM_OUT = TM_THREADS - 2;          // output columns per tile
#ifndef DLB_TM_RT
#define DLB_TM_RT 4  // measured: 4 rows x 4 stages (32 KB, 6 CTAs/SM) best; 2..8 rows, 2..6 stages within 5 %
#endif
constexpr int TM_RT = DLB_TM_RT;                // rows per stage
constexpr int TM_SPT = 64 / TM_RT;              // stages per task
constexpr int TM_LOAD = TM_RT * TM_SPT;         // loaded rows per task (64)
constexpr int TM_ROWS = TM_LOAD - 2;            // output rows per task (62)
#ifndef DLB_TM_STAGES
#define DLB_TM_STAGES 4
#endif
constexpr int TM_STAGES = DLB_TM_STAGES;        // stages in flight
constexpr int TM_STAGE_BYTES = TM_RT * TM_THREADS * 16;
#ifndef DLB_TM_MINB
#define DLB_TM_MINB 5  // <= 102 registers (the Xi variants want ~90; 6 made them spill: +10 %)
#endif

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

template <int KIND, int MODE, bool XI, bool AOS, bool UX, bool UY>
__global__ void __launch_bounds__(TM_THREADS, DLB_TM_MINB)
    stencil_tma_kernel(const StencilParams P, const __grid_constant__ CUtensorMap map0,
                       const __grid_constant__ CUtensorMap map1) {
  extern __shared__ __align__(128) unsigned char tiles[];  // [TM_STAGES][TM_STAGE_BYTES]
  __shared__ __align__(8) unsigned long long full[TM_STAGES];
  __shared__ double2 ysm[TM_ROWS][2];
  __shared__ double2 logtab_s[KIND == KIND_FTLE ? 128 : 1];
  const long long xdim = P.ax.n, ydim = P.ay.n;
  const int t = threadIdx.x;
  if (KIND == KIND_FTLE) logtab_s[t] = g_logtab[t];  // TM_THREADS == 128 entries
  const long long r_lo = (P.out_row0 > 1) ? P.out_row0 : 1;
  const long long r_hi = (P.out_row0 + P.out_rows < ydim - 1) ? P.out_row0 + P.out_rows : ydim - 1;
  const long long ncols_out = xdim - 2;
  if (r_hi <= r_lo || ncols_out <= 0) return;
  const long long nstrips = (ncols_out + TM_OUT - 1) / TM_OUT;
  const long long nchunks = (r_hi - r_lo + TM_ROWS - 1) / TM_ROWS;
  const long long ntasks = nstrips * nchunks;
  if ((long long)blockIdx.x >= ntasks) return;
  const long long my_tasks = (ntasks - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const long long total_stages = my_tasks * TM_SPT;
  // np.gradient's uniform / non-uniform branch per axis is a template parameter: as a run-time
  // flag it cost a predicated copy of both x formulas and a real branch around the y formula
  // in every row
  constexpr bool ux = UX, uy = UY;
  const double ix = P.ax.inv2h, iy = P.ay.inv2h;

  // producer side (thread 0): stage q of this CTA's stream = stage (q % TM_SPT) of its task q / TM_SPT
  auto issue = [&](long long q) {
    const long long task = blockIdx.x + (q / TM_SPT) * gridDim.x;
    const int st = (int)(q % TM_SPT);
    const long long chunk = task / nstrips, strip = task - chunk * nstrips;
    const int c0 = (int)(strip * TM_OUT);
    const int lrow = (int)(r_lo + chunk * TM_ROWS - 1 - P.grow0) + st * TM_RT;
    const int slot = (int)(q % TM_STAGES);
    unsigned char* dst = tiles + slot * TM_STAGE_BYTES;
    mbar_expect_tx(&full[slot], TM_STAGE_BYTES);
    if (AOS) {
      tma_load_2d(dst, &map0, 2 * c0, lrow, &full[slot]);
    } else {
      tma_load_2d(dst, &map0, c0, lrow, &full[slot]);
      tma_load_2d(dst + TM_STAGE_BYTES / 2, &map1, c0, lrow, &full[slot]);
    }
  };
  if (t == 0) {
#pragma unroll
    for (int i = 0; i < TM_STAGES; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (long long q = 0; q < TM_STAGES && q < total_stages; ++q) issue(q);
  }
  __syncthreads();

  long long q = 0;  // stages consumed so far (CTA-uniform)
  for (long long it = 0; it < my_tasks; ++it) {
    const long long task = blockIdx.x + it * gridDim.x;
    const long long chunk = task / nstrips, strip = task - chunk * nstrips;
    const long long c0 = strip * TM_OUT;
    const long long g0 = r_lo + chunk * TM_ROWS;  // first output row of the task
    const int nrows = (int)((g0 + TM_ROWS < r_hi) ? TM_ROWS : r_hi - g0);
    const long long col = (c0 + t < xdim) ? c0 + t : xdim - 1;
    const bool emit_thread = (t >= 1) && (t <= TM_OUT) && (c0 + t <= xdim - 2);
    const int tl = (t > 0) ? t - 1 : 0, tr = (t < TM_THREADS - 1) ? t + 1 : TM_THREADS - 1;
    double xa = 0.0, xb = 0.0, xc = 0.0;
    if (!ux) {
      xa = __ldg(P.ax.ca + col);
      xb = __ldg(P.ax.cb + col);
      xc = __ldg(P.ax.cc + col);
    }
    if (!uy) {
      __syncthreads();  // the previous task's readers of ysm are done
      if (t < nrows) {
        const double2* src = P.ay4 + 2 * (g0 + t);
        ysm[t][0] = __ldg(src);
        ysm[t][1] = __ldg(src + 1);
      }
      __syncthreads();
    }
    // loaded row r of the task is global row g0 - 1 + r; while row r streams in as D, the centre
    // row r - 1 (registers C, L, R) with U = row r - 2 is emitted: output row index k = r - 2
    double2 U = make_double2(0.0, 0.0), C = U, L = U, R = U;
    long long o = (g0 - 2 - P.out_row0) * xdim + col;  // output offset of k = -2 (never stored)
#pragma unroll 1
    for (int st = 0; st < TM_SPT; ++st, ++q) {
      const int slot = (int)(q % TM_STAGES);
      const unsigned parity = (unsigned)((q / TM_STAGES) & 1);
      while (!mbar_try_wait(&full[slot], parity)) {
      }
      const unsigned char* tile = tiles + slot * TM_STAGE_BYTES;
#pragma unroll
      for (int j = 0; j < TM_RT; ++j) {
        double2 D, Ld, Rd;
        if (AOS) {
          const double2* row = reinterpret_cast<const double2*>(tile) + j * TM_THREADS;
          D = row[t];
          Ld = row[tl];
          Rd = row[tr];
        } else {
          const double* ru = reinterpret_cast<const double*>(tile) + j * TM_THREADS;
          const double* rv = ru + TM_RT * TM_THREADS;
          D = make_double2(ru[t], rv[t]);
          Ld = make_double2(ru[tl], rv[tl]);
          Rd = make_double2(ru[tr], rv[tr]);
        }
        const int k = st * TM_RT + j - 2;  // output row of the centre held in C
        double d0dx, d1dx, d0dy, d1dy;
        if (ux) {
          d0dx = (R.x - L.x) * ix;
          d1dx = (R.y - L.y) * ix;
        } else {
          d0dx = fma(xc, R.x, fma(xb, C.x, xa * L.x));
          d1dx = fma(xc, R.y, fma(xb, C.y, xa * L.y));
        }
        if (uy) {
          d0dy = (D.x - U.x) * iy;
          d1dy = (D.y - U.y) * iy;
        } else {
          const int kc = (k < 0) ? 0 : k;
          const double2 yab = ysm[kc][0], ycz = ysm[kc][1];
          d0dy = fma(ycz.x, D.x, fma(yab.y, C.x, yab.x * U.x));
          d1dy = fma(ycz.x, D.y, fma(yab.y, C.y, yab.x * U.y));
        }
        emit_pred<KIND, MODE, XI>(P, C, d0dx, d1dx, d0dy, d1dy, o,
                                  KIND == KIND_FTLE ? logtab_s : nullptr,
                                  emit_thread && k >= 0 && k < nrows);
        o += xdim;
        U = C;
        C = D;
        L = Ld;
        R = Rd;
      }
      // every thread has read this stage: hand its slot back to the TMA unit
      __syncthreads();
      if (t == 0 && q + TM_STAGES < total_stages) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(q + TM_STAGES);
      }
    }
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no libcuda link dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    (void)cudaGetLastError();
  }
  return fn;
}

// 2-D FP64 tensor [rows][inner] with row pitch `pitch_bytes`, box [TM_RT][box_inner]
bool make_tile_map(CUtensorMap* map, const void* base, long long inner, long long rows,
                   long long pitch_bytes, int box_inner) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn || ((uintptr_t)base & 15u) || (pitch_bytes & 15) || inner <= 0 || rows <= 0 ||
      inner >= (1ll << 32) || rows >= (1ll << 32) || pitch_bytes >= (1ll << 40))
    return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)pitch_bytes};
  const cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)TM_RT};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void*>(base), gdim, gstride, box,
            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int KIND, int MODE, bool XI, bool AOS, bool UX, bool UY>
int launch_tma_uv(const StencilParams& P, const CUtensorMap& m0, const CUtensorMap& m1,
                  long long ntasks, cudaStream_t stream) {
  auto kern = stencil_tma_kernel<KIND, MODE, XI, AOS, UX, UY>;
  constexpr int dyn = TM_STAGES * TM_STAGE_BYTES;
  static int per_sm_of[DLB_MAX_DEVICES] = {0};  // the attribute is a per-device setting
  int dev = 0;
  DLB_CUDA_CHECK(cudaGetDevice(&dev));
  int& per_sm = per_sm_of[dev & (DLB_MAX_DEVICES - 1)];
  if (per_sm == 0) {
    DLB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    DLB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TM_THREADS, dyn));
    if (per_sm < 1) per_sm = 1;
  }
  long long blocks = (long long)num_sms() * per_sm;
  if (blocks > ntasks) blocks = ntasks;
  kern<<<(unsigned)blocks, TM_THREADS, dyn, stream>>>(P, m0, m1);
  DLB_CUDA_CHECK(cudaGetLastError());
  return DLB_OK;
}

// (fl(1/c), -ln(fl(1/c))) for c = 1 + i/128, uploaded once per device (a __device__ symbol has
// one instance per device).
int ensure_logtab() {
  static std::mutex mu;
  static bool done[DLB_MAX_DEVICES] = {false};
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  DLB_CUDA_CHECK(cudaGetDevice(&dev));
  if (done[dev & (DLB_MAX_DEVICES - 1)]) return DLB_OK;
  double2 h[128];
  for (int i = 0; i < 128; ++i) {
    const double c = 1.0 + i / 128.0;
    const double inv = (i == 0) ? 1.0 : 1.0 / c;
    h[i].x = inv;
    h[i].y = (i == 0) ? 0.0 : (double)(-logl((long double)inv));
  }
  DLB_CUDA_CHECK(cudaMemcpyToSymbol(g_logtab, h, sizeof(h)));
  done[dev & (DLB_MAX_DEVICES - 1)] = true;
  return DLB_OK;
}

template <int KIND, int MODE, bool XI, bool AOS>
int launch_tma(const StencilParams& P, const CUtensorMap& m0, const CUtensorMap& m1,
               long long ntasks, cudaStream_t stream) {
  if (P.ax.uniform && P.ay.uniform)
    return launch_tma_uv<KIND, MODE, XI, AOS, true, true>(P, m0, m1, ntasks, stream);
  if (P.ax.uniform)
    return launch_tma_uv<KIND, MODE, XI, AOS, true, false>(P, m0, m1, ntasks, stream);
  if (P.ay.uniform)
    return launch_tma_uv<KIND, MODE, XI, AOS, false, true>(P, m0, m1, ntasks, stream);
  return launch_tma_uv<KIND, MODE, XI, AOS, false, false>(P, m0, m1, ntasks, stream);
}

template <int KIND, int MODE, bool XI, class Loader>
int launch_stencil(const StencilParams& P, const Loader& ld, cudaStream_t stream) {
  if (KIND == KIND_FTLE) {
    const int rc = ensure_logtab();
    if (rc) return rc;
  }
  const long long xdim = P.ax.n, ydim = P.ay.n;
  // edges
  const long long r0 = P.out_row0, r1 = P.out_row0 + P.out_rows;
  const long long n_full = ((r0 == 0) ? 1 : 0) + ((r1 == ydim && ydim > 1) ? 1 : 0);
  const long long total_edge = n_full * xdim + (P.out_rows - n_full) * 2;
  long long eb = (total_edge + ST_BLOCK - 1) / ST_BLOCK;
  const long long cap = (long long)num_sms() * 8;
  if (eb > cap) eb = cap;
  if (eb > 0) {
    stencil_edge_kernel<KIND, MODE, XI, Loader><<<(unsigned)eb, ST_BLOCK, 0, stream>>>(P, ld);
    DLB_CUDA_CHECK(cudaGetLastError());
  }
  // interior
  const long long r_lo = (r0 > 1) ? r0 : 1, r_hi = (r1 < ydim - 1) ? r1 : ydim - 1;
  if (r_hi > r_lo && xdim > 2) {
    constexpr bool aos = std::is_same<Loader, LoaderAoS>::value;
    using Ring = typename std::conditional<aos, RingAoS, RingSoA>::type;
    const bool bulk_ok = env_flag("DLB_STENCIL_BULK", 1) && (aos || (xdim % 2 == 0)) &&
                         ld.aligned16();
    if (bulk_ok && env_flag("DLB_STENCIL_TMA", 1)) {
      CUtensorMap m0, m1;
      bool ok;
      if constexpr (aos) {
        ok = make_tile_map(&m0, ld.p, 2 * xdim, P.nloc, 16 * xdim, 2 * TM_THREADS);
        m1 = m0;
      } else {
        ok = make_tile_map(&m0, ld.u, xdim, P.nloc, 8 * xdim, TM_THREADS) &&
             make_tile_map(&m1, ld.v, xdim, P.nloc, 8 * xdim, TM_THREADS);
      }
      if (ok) {
        const long long ntasks = ((xdim - 2 + TM_OUT - 1) / TM_OUT) * ((r_hi - r_lo + TM_ROWS - 1) / TM_ROWS);
        return launch_tma<KIND, MODE, XI, aos>(P, m0, m1, ntasks, stream);
      }
    }
    if (bulk_ok) {
      const long long nstrips = (xdim - 2 + TB_OUT - 1) / TB_OUT;
      const long long nchunks = (r_hi - r_lo + TB_ROWS - 1) / TB_ROWS;
      long long blocks = nstrips * nchunks;
      if (blocks > cap) blocks = cap;
      const unsigned g = (unsigned)blocks;
      if (P.ax.uniform && P.ay.uniform)
        stencil_bulk_kernel<KIND, MODE, XI, true, true, Loader, Ring><<<g, TB_THREADS, 0, stream>>>(P, ld);
      else if (P.ax.uniform)
        stencil_bulk_kernel<KIND, MODE, XI, true, false, Loader, Ring><<<g, TB_THREADS, 0, stream>>>(P, ld);
      else if (P.ay.uniform)
        stencil_bulk_kernel<KIND, MODE, XI, false, true, Loader, Ring><<<g, TB_THREADS, 0, stream>>>(P, ld);
      else
        stencil_bulk_kernel<KIND, MODE, XI, false, false, Loader, Ring><<<g, TB_THREADS, 0, stream>>>(P, ld);
      DLB_CUDA_CHECK(cudaGetLastError());
      return DLB_OK;
    }
    const long long nstrips = (xdim - 2 + ST_COLS - 1) / ST_COLS;
    const long long nchunks = (r_hi - r_lo + ST_ROWS - 1) / ST_ROWS;
    long long blocks = (nstrips * nchunks + ST_BLOCK / 32 - 1) / (ST_BLOCK / 32);
    if (blocks > cap) blocks = cap;
    const unsigned g = (unsigned)blocks;
    if (P.ax.uniform && P.ay.uniform)
      stencil_interior_kernel<KIND, MODE, XI, true, true, Loader><<<g, ST_BLOCK, 0, stream>>>(P, ld);
    else if (P.ax.uniform)
      stencil_interior_kernel<KIND, MODE, XI, true, false, Loader><<<g, ST_BLOCK, 0, stream>>>(P, ld);
    else if (P.ay.uniform)
      stencil_interior_kernel<KIND, MODE, XI, false, true, Loader><<<g, ST_BLOCK, 0, stream>>>(P, ld);
    else
      stencil_interior_kernel<KIND, MODE, XI, false, false, Loader><<<g, ST_BLOCK, 0, stream>>>(P, ld);
    DLB_CUDA_CHECK(cudaGetLastError());
  }
  return DLB_OK;
}

// ---- ridge field: two passes over a scalar field ----
struct LoaderScalar {
  const double* f;
  long long xdim;
  __device__ __forceinline__ double2 at(long long lrow, long long j) const {
    return make_double2(__ldg(f + lrow * xdim + j), 0.0);
  }
};

// pass 1: (dfdx, dfdy) of the field -> scratch as AoS double2
__global__ void __launch_bounds__(ST_BLOCK) ridge_pass1(const StencilParams P, const LoaderScalar ld,
                                                        double2* grad) {
  const long long xdim = P.ax.n, ydim = P.ay.n;
  const long long total = xdim * ydim;
  for (long long k = (long long)blockIdx.x * ST_BLOCK + threadIdx.x; k < total;
       k += (long long)gridDim.x * ST_BLOCK) {
    const long long i = k / xdim, j = k - i * xdim;
    const double2 C = ld.at(i, j);
    const double2 L = ld.at(i, j > 0 ? j - 1 : 0), R = ld.at(i, j < xdim - 1 ? j + 1 : xdim - 1);
    const double2 U = ld.at(i > 0 ? i - 1 : 0, j), D = ld.at(i < ydim - 1 ? i + 1 : ydim - 1, j);
    double gx, gy, t0, t1;
    edge_axis_x(P, ld, i, j, C, L, R, &gx, &t0);
    edge_axis_y(P, ld, i, j, C, U, D, &gy, &t1);
    grad[k] = make_double2(gx, gy);
  }
}

// pass 2: Hessian from the first derivatives, directional derivative, concavity, masks
// (_base_classes.py:218-268).
__global__ void __launch_bounds__(ST_BLOCK) ridge_pass2(const StencilParams P, const LoaderAoS ld,
                                                        const double* field) {
  const long long xdim = P.ax.n, ydim = P.ay.n;
  const long long total = xdim * ydim;
  for (long long k = (long long)blockIdx.x * ST_BLOCK + threadIdx.x; k < total;
       k += (long long)gridDim.x * ST_BLOCK) {
    const long long i = k / xdim, j = k - i * xdim;
    const double2 C = ld.at(i, j);  // (dfdx, dfdy)
    const double2 L = ld.at(i, j > 0 ? j - 1 : 0), R = ld.at(i, j < xdim - 1 ? j + 1 : xdim - 1);
    const double2 U = ld.at(i > 0 ? i - 1 : 0, j), D = ld.at(i < ydim - 1 ? i + 1 : ydim - 1, j);
    double dfdxdx, dfdydx, dfdxdy, dfdydy;
    edge_axis_x(P, ld, i, j, C, L, R, &dfdxdx, &dfdydx);  // d/dx of (dfdx, dfdy)
    edge_axis_y(P, ld, i, j, C, U, D, &dfdxdy, &dfdydy);  // d/dy of (dfdx, dfdy)
    const double x0 = P.xi_in[2 * k], x1 = P.xi_in[2 * k + 1];
    // np.dot([dfdx, dfdy], Xi)   (:225-227)
    const double dd = fma(C.y, x1, C.x * x0);
    // np.dot(np.dot([[dfdxdx, dfdxdy], [dfdydx, dfdydy]], Xi), Xi)   (:228-235)
    const double h0 = fma(dfdxdy, x1, dfdxdx * x0);
    const double h1 = fma(dfdydy, x1, dfdydx * x0);
    const double conc = fma(h1, x1, h0 * x0);
    const double fv = field[k];
    bool m = (fv <= 0.0) || (conc >= 0.0);           // :251-258
    if (P.use_threshold) m = m || (fv <= P.threshold);  // :266-268
    P.field[k] = dd;
    P.conc[k] = conc;
    P.mask[k] = m ? 1 : 0;
  }
}

// ---- host: numpy's coefficient formulas ----
struct HostAxis {
  std::vector<double> ca, cb, cc;
  double inv2h = 0, w_first = 0, w_last = 0;
  int uniform = 1;
};

// numpy/lib/_function_base_impl.py:1246-1251, 1305-1372
void build_axis(const double* c, long long n, int edge_order, HostAxis* H) {
  H->ca.assign(n, 0.0);
  H->cb.assign(n, 0.0);
  H->cc.assign(n, 0.0);
  std::vector<double> dx(n > 1 ? n - 1 : 0);
  for (long long i = 0; i + 1 < n; ++i) dx[i] = c[i + 1] - c[i];
  H->uniform = 1;
  for (long long i = 1; i + 1 < n; ++i)
    if (dx[i] != dx[0]) H->uniform = 0;
  const double h = dx[0];
  if (H->uniform) {
    H->inv2h = 1.0 / (2. * h);
    H->w_first = H->w_last = 1.0 / h;
  } else {
    for (long long i = 1; i + 1 < n; ++i) {
      const double dx1 = dx[i - 1], dx2 = dx[i];
      H->ca[i] = -(dx2) / (dx1 * (dx1 + dx2));
      H->cb[i] = (dx2 - dx1) / (dx1 * dx2);
      H->cc[i] = dx1 / (dx2 * (dx1 + dx2));
    }
    H->w_first = 1.0 / dx[0];
    H->w_last = 1.0 / dx[n - 2];
  }
  if (edge_order == 2) {
    if (H->uniform) {
      H->ca[0] = -1.5 / h; H->cb[0] = 2. / h; H->cc[0] = -0.5 / h;
      H->ca[n - 1] = 0.5 / h; H->cb[n - 1] = -2. / h; H->cc[n - 1] = 1.5 / h;
    } else {
      double dx1 = dx[0], dx2 = dx[1];
      H->ca[0] = -(2. * dx1 + dx2) / (dx1 * (dx1 + dx2));
      H->cb[0] = (dx1 + dx2) / (dx1 * dx2);
      H->cc[0] = -dx1 / (dx2 * (dx1 + dx2));
      dx1 = dx[n - 3];
      dx2 = dx[n - 2];
      H->ca[n - 1] = (dx2) / (dx1 * (dx1 + dx2));
      H->cb[n - 1] = -(dx2 + dx1) / (dx1 * dx2);
      H->cc[n - 1] = (2. * dx2 + dx1) / (dx2 * (dx1 + dx2));
    }
  }
}

// (x, y, edge_order) whose coefficients are resident in a workspace, one entry per workspace
// (= per device / engine; a process that drives several devices keeps all of them warm).
struct AxesEntry {
  bool valid = false;
  const void* ws = nullptr;
  int device = -1;
  int edge_order = 0;
  std::vector<double> x, y;
  HostAxis hx, hy;
  cudaStream_t stream = nullptr;   // stream of the upload / the last user
  cudaEvent_t uploaded = nullptr;  // recorded after the upload (owned by the entry)
  unsigned long long tick = 0;
};
struct AxesCache {
  std::mutex mu;
  AxesEntry e[DLB_AXES_ENTRIES];
  unsigned long long clock = 0;
  AxesEntry* find(const void* ws) {
    for (auto& a : e)
      if (a.ws == ws) return &a;
    return nullptr;
  }
  AxesEntry* victim() {  // an unused slot, else the least recently used one
    AxesEntry* v = &e[0];
    for (auto& a : e) {
      if (!a.ws) return &a;
      if (a.tick < v->tick) v = &a;
    }
    return v;
  }
};
AxesCache& axes_cache() {
  static AxesCache c;
  return c;
}

}  // namespace

namespace dlbst {

// Validate geometry, build + upload both axes into the workspace (cached), fill P.ax / P.ay.
int setup_axes(StencilParams& P, const double* h_x, int64_t xdim, const double* h_y, int64_t ydim,
               int edge_order, void* d_workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (!h_x || !h_y || !d_workspace) {
    dlb_set_error("stencil: null pointer");
    return DLB_ERR_ARG;
  }
  if (edge_order > 2) {  // numpy :1255-1256
    dlb_set_error("'edge_order' greater than 2 not supported");
    return DLB_ERR_ARG;
  }
  if (edge_order < 1) {
    dlb_set_error("'edge_order' must be 1 or 2");
    return DLB_ERR_ARG;
  }
  if (xdim < edge_order + 1 || ydim < edge_order + 1) {  // numpy :1288-1291
    dlb_set_error(
        "Shape of array too small to calculate a numerical gradient, at least (edge_order + 1) "
        "elements are required.");
    return DLB_ERR_TOO_SMALL;
  }
  if (workspace_bytes < dlb_workspace_bytes(xdim, ydim)) {
    dlb_set_error("workspace too small");
    return DLB_ERR_WORKSPACE;
  }
  // workspace: 256 B header, x[xdim], y[ydim] (rk45.cu), then coefficients
  double* base = (double*)((char*)d_workspace + 256) + (xdim + ydim);
  double* dxa = base;
  double* dya = base + 3 * xdim;
  AxesCache& all = axes_cache();
  std::lock_guard<std::mutex> lock(all.mu);
  int dev = 0;
  DLB_CUDA_CHECK(cudaGetDevice(&dev));
  AxesEntry* found = all.find(d_workspace);
  const bool hit = found && found->valid && found->device == dev &&
                   found->edge_order == edge_order && (int64_t)found->x.size() == xdim &&
                   (int64_t)found->y.size() == ydim &&
                   memcmp(found->x.data(), h_x, sizeof(double) * xdim) == 0 &&
                   memcmp(found->y.data(), h_y, sizeof(double) * ydim) == 0;
  AxesEntry& cache = found ? *found : *all.victim();
  cache.tick = ++all.clock;
  if (!hit) {
    if (found && found->stream != stream && found->device == dev) {
      // kernels enqueued on another stream may still read the coefficients this upload is
      // about to overwrite; their completion cannot be observed from here (the stream handle
      // may even be gone), so drain the device.  Rare: same workspace, new grid, new stream.
      DLB_CUDA_CHECK(cudaDeviceSynchronize());
    }
    if (cache.uploaded && cache.device != dev) {  // slot reused for another device
      cudaEventDestroy(cache.uploaded);
      (void)cudaGetLastError();
      cache.uploaded = nullptr;
    }
    cache.valid = false;
    cache.ws = d_workspace;
    cache.device = dev;
    build_axis(h_x, xdim, edge_order, &cache.hx);
    build_axis(h_y, ydim, edge_order, &cache.hy);
    const HostAxis &hx = cache.hx, &hy = cache.hy;
    std::vector<double> stage(3 * (xdim + ydim) + 4 * ydim + xdim + ydim);
    memcpy(&stage[0], hx.ca.data(), sizeof(double) * xdim);
    memcpy(&stage[xdim], hx.cb.data(), sizeof(double) * xdim);
    memcpy(&stage[2 * xdim], hx.cc.data(), sizeof(double) * xdim);
    memcpy(&stage[3 * xdim], hy.ca.data(), sizeof(double) * ydim);
    memcpy(&stage[3 * xdim + ydim], hy.cb.data(), sizeof(double) * ydim);
    memcpy(&stage[3 * xdim + 2 * ydim], hy.cc.data(), sizeof(double) * ydim);
    double* y4 = &stage[3 * (xdim + ydim)];
    for (int64_t i = 0; i < ydim; ++i) {
      y4[4 * i] = hy.ca[i];
      y4[4 * i + 1] = hy.cb[i];
      y4[4 * i + 2] = hy.cc[i];
      y4[4 * i + 3] = 0.0;
    }
    memcpy(y4 + 4 * ydim, h_x, sizeof(double) * xdim);  // private copy of the coordinates
    memcpy(y4 + 4 * ydim + xdim, h_y, sizeof(double) * ydim);
    // pageable source: cudaMemcpyAsync stages it before returning, so `stage` may die here
    DLB_CUDA_CHECK(cudaMemcpyAsync(base, stage.data(), sizeof(double) * stage.size(),
                                   cudaMemcpyHostToDevice, stream));
    if (!cache.uploaded)
      DLB_CUDA_CHECK(cudaEventCreateWithFlags(&cache.uploaded, cudaEventDisableTiming));
    DLB_CUDA_CHECK(cudaEventRecord(cache.uploaded, stream));
    cache.edge_order = edge_order;
    cache.x.assign(h_x, h_x + xdim);
    cache.y.assign(h_y, h_y + ydim);
    cache.stream = stream;
    cache.valid = true;
  } else if (cache.stream != stream) {
    // uploaded on another stream: order this stream after the upload (device-side wait on the
    // entry's own event; the other stream's handle is never touched again)
    DLB_CUDA_CHECK(cudaStreamWaitEvent(stream, cache.uploaded, 0));
    cache.stream = stream;
  }
  const HostAxis &hx = cache.hx, &hy = cache.hy;
  P.ax = Axis{dxa, dxa + xdim, dxa + 2 * xdim, xdim, hx.inv2h, hx.w_first, hx.w_last, hx.uniform};
  P.ay = Axis{dya, dya + ydim, dya + 2 * ydim, ydim, hy.inv2h, hy.w_first, hy.w_last, hy.uniform};
  P.ay4 = (const double2*)(base + 3 * (xdim + ydim));  // 32-byte aligned: 256 + 32 (xdim+ydim)
  P.xs = base + 3 * (xdim + ydim) + 4 * ydim;
  P.ys = P.xs + xdim;
  P.edge_order = edge_order;
  return DLB_OK;
}

int check_slab(const StencilParams& P, int64_t ydim) {
  if (P.out_rows < 0 || P.out_row0 < 0 || P.out_row0 + P.out_rows > ydim || P.grow0 < 0 ||
      P.grow0 + P.nloc > ydim) {
    dlb_set_error("stencil: slab outside the grid");
    return DLB_ERR_ARG;
  }
  if (P.out_rows == 0) return DLB_OK;
  // every row an output row's stencil touches must be local
  long long lo = P.out_row0 > 0 ? P.out_row0 - 1 : 0;
  long long hi = P.out_row0 + P.out_rows < ydim ? P.out_row0 + P.out_rows : ydim - 1;
  if (P.edge_order == 2) {
    if (P.out_row0 == 0 && hi < 2) hi = 2;
    if (P.out_row0 + P.out_rows == ydim && lo > ydim - 3) lo = ydim - 3;
  }
  if (lo < P.grow0 || hi >= P.grow0 + P.nloc) {
    dlb_set_error("stencil: halo rows missing (need global rows %lld..%lld, have %lld..%lld)", lo,
                  hi, P.grow0, P.grow0 + P.nloc - 1);
    return DLB_ERR_ARG;
  }
  return DLB_OK;
}

}  // namespace dlbst


void dlb_axes_cache_note_layout(const void* d_workspace, int64_t xdim, int64_t ydim) {
  AxesCache& all = axes_cache();
  std::lock_guard<std::mutex> lock(all.mu);
  AxesEntry* c = all.find(d_workspace);
  if (c && c->valid && ((int64_t)c->x.size() != xdim || (int64_t)c->y.size() < ydim))
    c->valid = false;
}

extern "C" int dlb_ftle_stencil(const double* d_flow_map, int64_t nloc, int64_t grow0,
                                const double* h_x, int64_t xdim, const double* h_y, int64_t ydim,
                                int edge_order, double t_span, int64_t out_row0, int64_t out_rows,
                                double* d_ftle, double* d_xi, int xi_mode, void* d_workspace,
                                size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!d_flow_map || !d_ftle || (xi_mode != DLB_XI_NONE && !d_xi)) {
    dlb_set_error("dlb_ftle_stencil: null pointer");
    return DLB_ERR_ARG;
  }
  StencilParams P;
  memset(&P, 0, sizeof(P));
  int rc = setup_axes(P, h_x, xdim, h_y, ydim, edge_order, d_workspace, workspace_bytes, stream);
  if (rc) return rc;
  P.nloc = nloc;
  P.grow0 = grow0;
  P.out_row0 = out_row0;
  P.out_rows = out_rows;
  rc = check_slab(P, ydim);
  if (rc || out_rows == 0) return rc;
  P.fm = (const double2*)d_flow_map;
  P.field = d_ftle;
  P.xi = (double2*)d_xi;
  P.xi_mode = xi_mode;
  P.scale = 1.0 / (2.0 * fabs(t_span));  // lagrangian.py:71
  LoaderAoS ld{P.fm, xdim};
  if (xi_mode == DLB_XI_NONE) return launch_stencil<KIND_FTLE, 0, false>(P, ld, stream);
  return launch_stencil<KIND_FTLE, 0, true>(P, ld, stream);
}

// K2 fused with the row gather (csrc/peer.cu has the DMA flavour): same kernels, the epilogue
// stores every value locally and into the peers' buffers, then one warp releases the flags.
int dlb_push_flags_after(uint64_t* const* d_flags, int n, uint64_t value, cudaStream_t stream);

extern "C" int dlb_ftle_stencil_gather(const double* d_flow_map, int64_t nloc, int64_t grow0,
                                       const double* h_x, int64_t xdim, const double* h_y,
                                       int64_t ydim, int edge_order, double t_span,
                                       int64_t out_row0, int64_t out_rows, double* d_ftle,
                                       double* const* d_peer_ftle, int n_peer,
                                       uint64_t* const* d_flags, int n_flags, uint64_t flag_value,
                                       void* d_workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!d_flow_map || !d_ftle || n_peer < 0 || n_peer > DLB_MAX_PEERS || (n_peer && !d_peer_ftle) ||
      n_flags < 0 || n_flags > DLB_MAX_PEERS || (n_flags && !d_flags)) {
    dlb_set_error("dlb_ftle_stencil_gather: bad argument");
    return DLB_ERR_ARG;
  }
  StencilParams P;
  memset(&P, 0, sizeof(P));
  int rc = setup_axes(P, h_x, xdim, h_y, ydim, edge_order, d_workspace, workspace_bytes, stream);
  if (rc) return rc;
  P.nloc = nloc;
  P.grow0 = grow0;
  P.out_row0 = out_row0;
  P.out_rows = out_rows;
  rc = check_slab(P, ydim);
  if (rc) return rc;
  if (out_rows > 0) {
    P.fm = (const double2*)d_flow_map;
    P.field = d_ftle;
    P.xi_mode = DLB_XI_NONE;
    P.scale = 1.0 / (2.0 * fabs(t_span));  // lagrangian.py:71
    for (int k = 0; k < n_peer; ++k)
      if (d_peer_ftle[k] && d_peer_ftle[k] != d_ftle) P.peer[P.npeer++] = d_peer_ftle[k];
    LoaderAoS ld{P.fm, xdim};
    rc = P.npeer ? launch_stencil<KIND_FTLE, FTLE_PUSH, false>(P, ld, stream)
                 : launch_stencil<KIND_FTLE, FTLE_PLAIN, false>(P, ld, stream);
    if (rc) return rc;
  }
  return dlb_push_flags_after(d_flags, n_flags, flag_value, stream);
}

extern "C" int dlb_euler_stencil(int mode, const double* d_u, const double* d_v, int64_t nloc,
                                 int64_t grow0, const double* h_x, int64_t xdim, const double* h_y,
                                 int64_t ydim, int edge_order, int64_t out_row0, int64_t out_rows,
                                 int negate, double* d_field, uint8_t* d_mask, double* d_xi,
                                 int xi_mode, void* d_workspace, size_t workspace_bytes,
                                 void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!d_u || !d_v || !d_field || (xi_mode != DLB_XI_NONE && !d_xi)) {
    dlb_set_error("dlb_euler_stencil: null pointer");
    return DLB_ERR_ARG;
  }
  if (mode < DLB_EULER_S_MIN || mode > DLB_EULER_RATIO) {
    dlb_set_error("dlb_euler_stencil: unknown mode %d", mode);
    return DLB_ERR_ARG;
  }
  if ((mode == DLB_EULER_RATE || mode == DLB_EULER_RATIO) && (!d_mask || xi_mode != DLB_XI_NONE)) {
    dlb_set_error("dlb_euler_stencil: rate/ratio need d_mask and produce no eigenvector");
    return DLB_ERR_ARG;
  }
  StencilParams P;
  memset(&P, 0, sizeof(P));
  int rc = setup_axes(P, h_x, xdim, h_y, ydim, edge_order, d_workspace, workspace_bytes, stream);
  if (rc) return rc;
  P.nloc = nloc;
  P.grow0 = grow0;
  P.out_row0 = out_row0;
  P.out_rows = out_rows;
  rc = check_slab(P, ydim);
  if (rc || out_rows == 0) return rc;
  P.u = d_u;
  P.v = d_v;
  P.field = d_field;
  P.mask = d_mask;
  P.xi = (double2*)d_xi;
  P.xi_mode = xi_mode;
  P.negate = negate;
  LoaderSoA ld{d_u, d_v, xdim};
  const bool xi = xi_mode != DLB_XI_NONE;
  switch (mode) {
    case DLB_EULER_S_MIN:
      return xi ? launch_stencil<KIND_EULER, DLB_EULER_S_MIN, true>(P, ld, stream)
                : launch_stencil<KIND_EULER, DLB_EULER_S_MIN, false>(P, ld, stream);
    case DLB_EULER_S_MAX:
      return xi ? launch_stencil<KIND_EULER, DLB_EULER_S_MAX, true>(P, ld, stream)
                : launch_stencil<KIND_EULER, DLB_EULER_S_MAX, false>(P, ld, stream);
    case DLB_EULER_RATE:
      return launch_stencil<KIND_EULER, DLB_EULER_RATE, false>(P, ld, stream);
    default:
      return launch_stencil<KIND_EULER, DLB_EULER_RATIO, false>(P, ld, stream);
  }
}

extern "C" int dlb_ridge_field(const double* d_field, const double* d_xi, const double* h_x,
                               int64_t xdim, const double* h_y, int64_t ydim, int edge_order,
                               int use_threshold, double threshold, double* d_dd, double* d_conc,
                               uint8_t* d_mask, double* d_scratch, void* d_workspace,
                               size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!d_field || !d_xi || !d_dd || !d_conc || !d_mask || !d_scratch) {
    dlb_set_error("dlb_ridge_field: null pointer");
    return DLB_ERR_ARG;
  }
  StencilParams P;
  memset(&P, 0, sizeof(P));
  int rc = setup_axes(P, h_x, xdim, h_y, ydim, edge_order, d_workspace, workspace_bytes, stream);
  if (rc) return rc;
  P.nloc = ydim;
  P.grow0 = 0;
  P.out_row0 = 0;
  P.out_rows = ydim;
  P.xi_in = d_xi;
  P.field = d_dd;
  P.conc = d_conc;
  P.mask = d_mask;
  P.use_threshold = use_threshold;
  P.threshold = threshold;
  const long long total = (long long)xdim * ydim;
  long long blocks = (total + ST_BLOCK - 1) / ST_BLOCK;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  LoaderScalar l1{d_field, xdim};
  ridge_pass1<<<(unsigned)blocks, ST_BLOCK, 0, stream>>>(P, l1, (double2*)d_scratch);
  DLB_CUDA_CHECK(cudaGetLastError());
  LoaderAoS l2{(const double2*)d_scratch, xdim};
  ridge_pass2<<<(unsigned)blocks, ST_BLOCK, 0, stream>>>(P, l2, d_field);
  DLB_CUDA_CHECK(cudaGetLastError());
  return DLB_OK;
}
