// K1 — FP64 Dormand-Prince 5(4) flow-map integrator with scipy solve_ivp(RK45) semantics.
//
// Replaces the reference's seed loop + integrator
//   R:src/dynlab/diagnostics/_base_classes.py:175-189   (product(y, x) seed order, (x, y) seeds)
//   R:src/dynlab/utils.py:6-25                           (integrator protocol; endpoint only)
// and restates scipy's RK45 (SURVEY §8a "K1 spec"):
//   scipy/integrate/_ivp/rk.py:14-71 (rk_step), :111-176 (_step_impl), :538-552 (tableau)
//   scipy/integrate/_ivp/common.py:44-65 (validate_tol, norm), :68-134 (select_initial_step)
//   scipy/integrate/_ivp/base.py:164 (direction), :179-210 (step / finished test)
//
// B200 mapping.  One thread owns one trajectory; state (t, h, y, f and the seven stage
// vectors) lives in registers.  The kernel is persistent: gridDim = resident CTAs, each warp
// reserves chunks of consecutive seeds from a global counter (one atomic per chunk) and
// refills lanes whose trajectory ended.  Adaptive step counts diverge between seeds, so a
// lane that finishes early would otherwise idle until the slowest lane of its warp is done;
// refilling is batched (threshold) because initialising a seed is divergent code.
// Every iteration of the main loop is one RK attempt (6 RHS evaluations); accept/reject only
// changes which values are kept, so the loop body is uniform across lanes.
#include <math_constants.h>

#include "flows.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
#ifndef DLB_RK_BLOCK
#define DLB_RK_BLOCK 128
#endif
constexpr int RK_BLOCK = DLB_RK_BLOCK;
#ifndef DLB_RK_MINBLOCKS
#define DLB_RK_MINBLOCKS 4
#endif

// Dormand-Prince tableau exactly as scipy states it (rk.py:541-552); B[1] = E[1] = 0.
// Kept in __constant__ memory (see dmath.cuh: literals would cost two UMOVs per use).
__constant__ double RKT[30] = {
    1.0 / 5,  // A10
    3.0 / 40,  // A20
    9.0 / 40,  // A21
    44.0 / 45,  // A30
    -56.0 / 15,  // A31
    32.0 / 9,  // A32
    19372.0 / 6561,  // A40
    -25360.0 / 2187,  // A41
    64448.0 / 6561,  // A42
    -212.0 / 729,  // A43
    9017.0 / 3168,  // A50
    -355.0 / 33,  // A51
    46732.0 / 5247,  // A52
    49.0 / 176,  // A53
    -5103.0 / 18656,  // A54
    35.0 / 384,  // B0
    500.0 / 1113,  // B2
    125.0 / 192,  // B3
    -2187.0 / 6784,  // B4
    11.0 / 84,  // B5
    -71.0 / 57600,  // E0
    71.0 / 16695,  // E2
    -71.0 / 1920,  // E3
    17253.0 / 339200,  // E4
    -22.0 / 525,  // E5
    1.0 / 40,  // E6
    1.0 / 5,  // C1
    3.0 / 10,  // C2
    4.0 / 5,  // C3
    8.0 / 9,  // C4
};
#define A10 RKT[0]
#define A20 RKT[1]
#define A21 RKT[2]
#define A30 RKT[3]
#define A31 RKT[4]
#define A32 RKT[5]
#define A40 RKT[6]
#define A41 RKT[7]
#define A42 RKT[8]
#define A43 RKT[9]
#define A50 RKT[10]
#define A51 RKT[11]
#define A52 RKT[12]
#define A53 RKT[13]
#define A54 RKT[14]
#define B0 RKT[15]
#define B2 RKT[16]
#define B3 RKT[17]
#define B4 RKT[18]
#define B5 RKT[19]
#define E0 RKT[20]
#define E2 RKT[21]
#define E3 RKT[22]
#define E4 RKT[23]
#define E5 RKT[24]
#define E6 RKT[25]
#define C1 RKT[26]
#define C2 RKT[27]
#define C3 RKT[28]
#define C4 RKT[29]

struct Rk45Params {
  FlowConsts fc;
  const double* x;      // grid mode: [xdim]
  const double* y;      // grid mode: [rows] (slab rows)
  const double* seeds;  // generic mode: [total][N]
  long long total;
  long long xdim;
  double t0, tf, rtol, max_step, first_step, direction;
  double atol[3];
  double* out;
  uint32_t* nfev;
  int8_t* status;
  unsigned long long* counters;
  unsigned long long* next;
  int chunk;
  int refill_threshold;
  // seeds [stat_lo, stat_hi) of this launch enter the counters (a slab's redundantly integrated
  // halo rows do not: the counters of a sharded run then equal the single-device ones)
  long long stat_lo, stat_hi;
};

template <int N>
__device__ __forceinline__ double rms_norm(const double* v) {
  // common.py:63-65: ||v||_2 / sqrt(n)
  double s = v[0] * v[0];
#pragma unroll
  for (int i = 1; i < N; ++i) s = fma(v[i], v[i], s);
  return sqrt(s) * (N == 2 ? 0.70710678118654752440 : 0.57735026918962576451);
}

// The six stage evaluations of one attempt (rk_step, rk.py:60-69) for state (t, y, f = K0)
// and step h.  Stage times are evaluated first in one block; stages 6 and 7 share t + h
// (C[5] = 1).  Returns K1..K6 (K6 = f_new) and y_new.
template <int N>
struct StageOut {
  double K[6][N];
  double yn[N];
};

template <int ID, int N, class Trig>
__device__ __forceinline__ void rk_stages(const FlowConsts& fc, Trig& trig, double t, double h,
                                          const double* y, const double* f, StageOut<N>& o) {
  constexpr int NT = FlowTime<ID>::N > 0 ? FlowTime<ID>::N : 1;
  double ts[5], tc[5][NT], ys[N];
  ts[0] = fma(C1, h, t);
  ts[1] = fma(C2, h, t);
  ts[2] = fma(C3, h, t);
  ts[3] = fma(C4, h, t);
  ts[4] = t + h;
  flow_time<ID, 5>(fc, trig, ts, tc);
  double(*K)[N] = o.K;
#pragma unroll
  for (int i = 0; i < N; ++i) ys[i] = fma(f[i] * A10, h, y[i]);
  flow_space<ID>(fc, trig, tc[0], ys, K[0]);
#pragma unroll
  for (int i = 0; i < N; ++i) ys[i] = fma(fma(K[0][i], A21, f[i] * A20), h, y[i]);
  flow_space<ID>(fc, trig, tc[1], ys, K[1]);
#pragma unroll
  for (int i = 0; i < N; ++i)
    ys[i] = fma(fma(K[1][i], A32, fma(K[0][i], A31, f[i] * A30)), h, y[i]);
  flow_space<ID>(fc, trig, tc[2], ys, K[2]);
#pragma unroll
  for (int i = 0; i < N; ++i)
    ys[i] = fma(fma(K[2][i], A43, fma(K[1][i], A42, fma(K[0][i], A41, f[i] * A40))), h, y[i]);
  flow_space<ID>(fc, trig, tc[3], ys, K[3]);
#pragma unroll
  for (int i = 0; i < N; ++i)
    ys[i] = fma(
        fma(K[3][i], A54, fma(K[2][i], A53, fma(K[1][i], A52, fma(K[0][i], A51, f[i] * A50)))),
        h, y[i]);
  flow_space<ID>(fc, trig, tc[4], ys, K[4]);
#pragma unroll
  for (int i = 0; i < N; ++i)
    o.yn[i] = fma(
        h, fma(K[4][i], B5, fma(K[3][i], B4, fma(K[2][i], B3, fma(K[1][i], B2, f[i] * B0)))),
        y[i]);
  flow_space<ID>(fc, trig, tc[4], o.yn, K[5]);  // f_new
}

// Out-of-line redo of an attempt whose fast trigonometry saw a huge / non-finite argument.
template <int ID, int N>
__device__ __noinline__ StageOut<N> rk_stages_safe(const FlowConsts& fc, double t, double h,
                                                   StageOut<1> yv, StageOut<1> fv) {
  // (y and f travel by value inside small structs: yv.K[i][0] = y[i], fv.K[i][0] = f[i])
  double y[N], f[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    y[i] = yv.K[i][0];
    f[i] = fv.K[i][0];
  }
  dm::TrigSafe trig;
  StageOut<N> o;
  rk_stages<ID, N>(fc, trig, t, h, y, f, o);
  return o;
}

// |nextafter(t, dir * inf) - t| (rk.py:119) from the bit pattern: the spacing of doubles
// at t in the direction of integration.
__device__ __forceinline__ double ulp_towards(double t, double dir) {
  // t is finite here: t0 and tf are validated on the host and t stays between them.
  const long long b = __double_as_longlong(t);
  const bool away = ((b < 0) == (dir < 0));  // moving away from zero?
  const double nb = __longlong_as_double(away ? b + 1 : b - 1);
  const double d = fabs(nb - t);
  // t == +-0: nextafter(0, +-inf) = +-denorm_min
  return ((b & 0x7fffffffffffffffLL) == 0) ? 4.9406564584124654e-324 : d;
}

template <int ID, int N, bool GRID>
__global__ void __launch_bounds__(RK_BLOCK, DLB_RK_MINBLOCKS) rk45_kernel(const Rk45Params P) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  const double dir = P.direction;
  const double tf = P.tf;
  const double rtol = P.rtol;

  long long chunk_next = 0, chunk_end = 0;  // warp-uniform
  bool exhausted = false;                   // warp-uniform
  bool active = false;

  long long idx = 0;
  double t = 0.0, h_abs = 0.0;
  double y[N], f[N];
  bool fresh = true, rejected = false;
  unsigned nfev_l = 0, rej_l = 0;  // per trajectory: RHS evaluations, rejected attempts
  unsigned long long c_nfev = 0, c_acc = 0, c_rej = 0, c_fail = 0, c_seeds = 0;

  auto finish = [&](int st) {
    if (N == 2) {
      reinterpret_cast<double2*>(P.out)[idx] = make_double2(y[0], y[1]);
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) P.out[idx * N + i] = y[i];
    }
    if (P.nfev) P.nfev[idx] = nfev_l;
    if (P.status) P.status[idx] = (int8_t)st;
    if (idx >= P.stat_lo && idx < P.stat_hi) {  // (a slab's redundant halo rows are not counted)
      // attempts = (nfev - 1) / 6 whether the start-up took one RHS evaluation or two
      const unsigned attempts = (nfev_l - 1u) / 6u;
      c_nfev += nfev_l;
      c_acc += attempts - rej_l;
      c_rej += rej_l;
      c_seeds += 1;
      if (st) c_fail += 1;
    }
    active = false;
  };

  for (;;) {
    // ---------------- lane refill (warp-uniform control flow) ----------------
    const unsigned idle = __ballot_sync(FULL, !active);
    if (idle != 0u && !exhausted && (__popc(idle) >= P.refill_threshold || idle == FULL)) {
#pragma unroll 1
      for (;;) {
        const unsigned need = __ballot_sync(FULL, !active);
        if (need == 0u) break;
        if (chunk_next >= chunk_end) {
          long long b = 0;
          if (lane == 0) b = (long long)atomicAdd(P.next, (unsigned long long)P.chunk);
          b = __shfl_sync(FULL, b, 0);
          if (b >= P.total) {
            exhausted = true;
            break;
          }
          chunk_next = b;
          chunk_end = (b + P.chunk < P.total) ? b + P.chunk : P.total;
        }
        const long long avail = chunk_end - chunk_next;
        const int rank = __popc(need & lt_mask);
        if (!active && rank < avail) {
          // ---- start a trajectory: rk.py:86-104 (RungeKutta.__init__) ----
          idx = chunk_next + rank;
          if (GRID) {
            const long long row = idx / P.xdim;
            const long long col = idx - row * P.xdim;
            y[0] = P.x[col];  // seed is (x, y): R:_base_classes.py:181
            y[1] = P.y[row];
          } else {
#pragma unroll
            for (int i = 0; i < N; ++i) y[i] = P.seeds[idx * N + i];
          }
          t = P.t0;
          flow_rhs<ID>(P.fc, t, y, f);  // rk.py:94
          nfev_l = 1;
          rej_l = 0;
          fresh = true;
          rejected = false;
          active = true;
          const double L = fabs(tf - t);
          if (L == 0.0) {
            finish(0);  // base.py:193-198: t == t_bound, no step
          } else if (P.first_step > 0.0) {
            h_abs = P.first_step;
          } else {
            // select_initial_step, common.py:109-134 (order = 4)
            double sc[N], w[N], y1[N], f1[N];
#pragma unroll
            for (int i = 0; i < N; ++i) sc[i] = P.atol[i] + fabs(y[i]) * rtol;
#pragma unroll
            for (int i = 0; i < N; ++i) w[i] = y[i] / sc[i];
            const double d0 = rms_norm<N>(w);
#pragma unroll
            for (int i = 0; i < N; ++i) w[i] = f[i] / sc[i];
            const double d1 = rms_norm<N>(w);
            double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
            h0 = fmin(h0, L);
#pragma unroll
            for (int i = 0; i < N; ++i) y1[i] = y[i] + h0 * dir * f[i];
            flow_rhs<ID>(P.fc, t + h0 * dir, y1, f1);
            nfev_l = 2;
#pragma unroll
            for (int i = 0; i < N; ++i) w[i] = (f1[i] - f[i]) / sc[i];
            const double d2 = rms_norm<N>(w) / h0;
            double h1;
            if (d1 <= 1e-15 && d2 <= 1e-15)
              h1 = fmax(1e-6, h0 * 1e-3);
            else
              h1 = pow(0.01 / fmax(d1, d2), 0.2);
            h_abs = fmin(fmin(100 * h0, h1), fmin(L, P.max_step));
          }
        }
        const long long want = (long long)__popc(need);
        chunk_next += (want < avail) ? want : avail;
      }
    }
    if (!__any_sync(FULL, active)) {
      if (exhausted) break;
      continue;  // (only reachable if every fetched seed finished at once: t0 == tf)
    }

    // ---------------- one RK45 attempt: rk.py:111-176 ----------------
    if (active) {
      const double min_step = 10.0 * ulp_towards(t, dir);  // rk.py:119
      if (fresh) {  // start of _step_impl: clamp once per step (rk.py:121-126)
        if (h_abs > P.max_step)
          h_abs = P.max_step;
        else if (h_abs < min_step)
          h_abs = min_step;
        fresh = false;
        rejected = false;
      }
      if (h_abs < min_step) {  // rk.py:132-133 TOO_SMALL_STEP -> status -1, keep last state
        finish(-1);
      } else {
        double h = h_abs * dir;
        double t_new = t + h;
        if (dir * (t_new - tf) > 0) t_new = tf;  // rk.py:138-139
        h = t_new - t;
        h_abs = fabs(h);

        // rk_step with branch-free trigonometry; one validity test for the whole attempt
        StageOut<N> so;
        {
          dm::TrigFast trig;
          rk_stages<ID, N>(P.fc, trig, t, h, y, f, so);
          if (!trig.ok()) {
            StageOut<1> yv, fv;
#pragma unroll
            for (int i = 0; i < N; ++i) {
              yv.K[i][0] = y[i];
              fv.K[i][0] = f[i];
            }
            so = rk_stages_safe<ID, N>(P.fc, t, h, yv, fv);
          }
        }
        const double(*K)[N] = so.K;
        const double* yn = so.yn;
        nfev_l += 6;

        // error estimate, rk.py:105-109,146-147; s = error_norm^2 (no square root needed:
        // error_norm < 1 <=> s < 1, error_norm ** -0.2 == s ** -0.1)
        double ssum = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
          const double sc = P.atol[i] + dm::max_sel(fabs(y[i]), fabs(yn[i])) * rtol;
          const double ee =
              fma(K[5][i], E6,
                  fma(K[4][i], E5,
                      fma(K[3][i], E4, fma(K[2][i], E3, fma(K[1][i], E2, f[i] * E0)))));
          const double e = ee * h * dm::rcp(sc);
          ssum = fma(e, e, ssum);
        }
        const double s = ssum * (N == 2 ? 0.5 : (1.0 / 3.0));

        // step-size controller, rk.py:149-165.  The power is taken on s clamped to
        // [1e-12, 1e8] (error_norm in [1e-6, 1e4]): outside it MAX_FACTOR / MIN_FACTOR win, so
        // the result is unchanged; a NaN norm maps to MIN_FACTOR like Python's max().
        double sc2 = (s < 1e8) ? s : 1e8;
        sc2 = dm::max_sel(sc2, 1e-12);
        const double pw = 0.9 * dm::pow_m01(sc2);
        if (s < 1.0) {
          double factor = (s == 0.0) ? 10.0 : dm::min_sel(pw, 10.0);
          if (rejected) factor = dm::min_sel(factor, 1.0);
          h_abs *= factor;
          t = t_new;
#pragma unroll
          for (int i = 0; i < N; ++i) {
            y[i] = yn[i];
            f[i] = K[5][i];
          }
          fresh = true;
          // (accepted steps are derived from nfev when the trajectory ends: nothing to count here)
          if (dir * (t - tf) >= 0) finish(0);  // base.py:207-208
        } else {
          h_abs *= dm::max_sel(pw, 0.2);
          rejected = true;
          rej_l += 1;
        }
      }
    }
  }

  // ---------------- counters ----------------
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c_nfev += __shfl_down_sync(FULL, c_nfev, o);
    c_acc += __shfl_down_sync(FULL, c_acc, o);
    c_rej += __shfl_down_sync(FULL, c_rej, o);
    c_fail += __shfl_down_sync(FULL, c_fail, o);
    c_seeds += __shfl_down_sync(FULL, c_seeds, o);
  }
  if (lane == 0) {
    atomicAdd(&P.counters[DLB_CNT_NFEV], c_nfev);
    atomicAdd(&P.counters[DLB_CNT_ACCEPTED], c_acc);
    atomicAdd(&P.counters[DLB_CNT_REJECTED], c_rej);
    atomicAdd(&P.counters[DLB_CNT_FAILED], c_fail);
    atomicAdd(&P.counters[DLB_CNT_SEEDS], c_seeds);
  }
}

template <int ID, int N, bool GRID>
int launch_rk45(const Rk45Params& P, cudaStream_t stream) {
  // per device and kernel instance: SM count x resident CTAs per SM (one process may drive
  // several devices)
  static int resident_of[DLB_MAX_DEVICES] = {0};
  int dev = 0;
  DLB_CUDA_CHECK(cudaGetDevice(&dev));
  int& resident = resident_of[dev & (DLB_MAX_DEVICES - 1)];
  if (resident == 0) {
    int sms = 0, per_sm = 0;
    DLB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    DLB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &per_sm, rk45_kernel<ID, N, GRID>, RK_BLOCK, 0));
    if (per_sm < 1) per_sm = 1;
    resident = sms * per_sm;
  }
  long long blocks = resident;
  const long long needed = (P.total + RK_BLOCK - 1) / RK_BLOCK;
  if (blocks > needed) blocks = needed;
  if (blocks < 1) blocks = 1;
  rk45_kernel<ID, N, GRID><<<(unsigned)blocks, RK_BLOCK, 0, stream>>>(P);
  DLB_CUDA_CHECK(cudaGetLastError());
  return DLB_OK;
}

int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

int fill_common(Rk45Params& P, int flow_id, const double* h_params, int n_params, double t0,
                double tf, double rtol, double max_step, double first_step) {
  if (!flow_prepare(flow_id, h_params, n_params, &P.fc)) {
    dlb_set_error("unknown flow id %d or wrong parameter count %d", flow_id, n_params);
    return DLB_ERR_ARG;
  }
  if (!isfinite(t0) || !isfinite(tf)) {
    dlb_set_error("t0 and tf must be finite");
    return DLB_ERR_ARG;
  }
  if (!(max_step > 0)) {  // common.py:19-23 validate_max_step
    dlb_set_error("`max_step` must be positive.");
    return DLB_ERR_ARG;
  }
  if (first_step < 0 || first_step > fabs(tf - t0)) {  // common.py:10-16
    dlb_set_error("`first_step` must be positive and within bounds.");
    return DLB_ERR_ARG;
  }
  const double eps = 2.220446049250313e-16;
  P.rtol = (rtol < 100 * eps) ? 100 * eps : rtol;  // common.py:47-51
  P.t0 = t0;
  P.tf = tf;
  P.max_step = max_step;
  P.first_step = first_step;
  P.direction = (tf != t0) ? ((tf > t0) ? 1.0 : -1.0) : 1.0;  // base.py:164
  P.chunk = env_int("DLB_RK45_CHUNK", 32);
  P.refill_threshold = env_int("DLB_RK45_REFILL", 32);
  if (P.chunk < 32) P.chunk = 32;
  if (P.refill_threshold < 1) P.refill_threshold = 1;
  if (P.refill_threshold > 32) P.refill_threshold = 32;
  return DLB_OK;
}

}  // namespace

// Workspace layout (bytes): [0,64) work counter; then x[xdim], y[ydim]; then the stencil
// coefficient arrays (stencil.cu).
extern "C" size_t dlb_workspace_bytes(int64_t xdim, int64_t ydim) {
  if (xdim < 0) xdim = 0;
  if (ydim < 0) ydim = 0;
  // x, y, 3 + 3 coefficient arrays, the interleaved y record (4 per row), the stencils' own
  // cached copy of x and y, and 32-byte slack
  return 256 + sizeof(double) * (size_t)(5 * (xdim + ydim) + 4 * ydim + 64);
}

namespace {
// K1 over a row slab with the coordinates already on the device: d_x[xdim], d_yrows[rows].
int flowmap_launch(int flow_id, const double* h_params, int n_params, const double* d_x,
                   int64_t xdim, const double* d_yrows, int64_t rows, double t0, double tf,
                   double rtol, double atol, double max_step, double first_step,
                   double* d_flow_map, uint32_t* d_nfev, int8_t* d_status, uint64_t* d_counters,
                   void* d_workspace, cudaStream_t stream, int64_t stat_row0 = 0,
                   int64_t stat_rows = -1) {
  Rk45Params P;
  memset(&P, 0, sizeof(P));
  if (stat_rows < 0) stat_rows = rows;
  P.stat_lo = stat_row0 * xdim;
  P.stat_hi = (stat_row0 + stat_rows) * xdim;
  int rc = fill_common(P, flow_id, h_params, n_params, t0, tf, rtol, max_step, first_step);
  if (rc) return rc;
  P.atol[0] = P.atol[1] = P.atol[2] = atol;
  char* ws = (char*)d_workspace;
  DLB_CUDA_CHECK(cudaMemsetAsync(ws, 0, 256, stream));
  DLB_CUDA_CHECK(cudaMemsetAsync(d_counters, 0, sizeof(uint64_t) * DLB_CNT_WORDS, stream));
  if (rows == 0) return DLB_OK;
  P.x = d_x;
  P.y = d_yrows;
  P.xdim = xdim;
  P.total = (long long)xdim * rows;
  P.out = d_flow_map;
  P.nfev = d_nfev;
  P.status = d_status;
  P.counters = (unsigned long long*)d_counters;
  P.next = (unsigned long long*)ws;
  rc = DLB_ERR_ARG;
  DLB_DISPATCH_FLOW(flow_id, {
    if constexpr (ID < DLB_FLOW_ABC) rc = launch_rk45<ID, 2, true>(P, stream);
  });
  return rc;
}

int flowmap_check(const char* who, int flow_id, const void* x, const void* y, const void* out,
                  const void* counters, const void* ws, int64_t xdim, int64_t rows, int64_t row0,
                  double atol) {
  if (!x || !y || !out || !counters || !ws || xdim < 1 || rows < 0 || row0 < 0) {
    dlb_set_error("%s: bad argument", who);
    return DLB_ERR_ARG;
  }
  if (flow_id < 0 || flow_id >= DLB_FLOW_COUNT || k_flow_info[flow_id].dim != 2) {
    dlb_set_error("%s: flow %d is not a 2-D flow", who, flow_id);
    return DLB_ERR_ARG;
  }
  if (!(atol >= 0)) {  // common.py:57-58
    dlb_set_error("`atol` must be positive.");
    return DLB_ERR_ARG;
  }
  return DLB_OK;
}
}  // namespace

extern "C" int dlb_flowmap_rk45(int flow_id, const double* h_params, int n_params,
                                const double* h_x, int64_t xdim, const double* h_y, int64_t row0,
                                int64_t rows, double t0, double tf, double rtol, double atol,
                                double max_step, double first_step, double* d_flow_map,
                                uint32_t* d_nfev, int8_t* d_status, uint64_t* d_counters,
                                void* d_workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = flowmap_check("dlb_flowmap_rk45", flow_id, h_x, h_y, d_flow_map, d_counters,
                         d_workspace, xdim, rows, row0, atol);
  if (rc) return rc;
  if (workspace_bytes < dlb_workspace_bytes(xdim, rows)) {
    dlb_set_error("workspace too small");
    return DLB_ERR_WORKSPACE;
  }
  {  // argument validation happens before any CUDA call
    Rk45Params probe;
    memset(&probe, 0, sizeof(probe));
    rc = fill_common(probe, flow_id, h_params, n_params, t0, tf, rtol, max_step, first_step);
    if (rc) return rc;
  }
  char* ws = (char*)d_workspace;
  double* d_x = (double*)(ws + 256);
  double* d_y = d_x + xdim;
  dlb_axes_cache_note_layout(d_workspace, xdim, rows);
  if (rows > 0) {
    DLB_CUDA_CHECK(
        cudaMemcpyAsync(d_x, h_x, sizeof(double) * xdim, cudaMemcpyHostToDevice, stream));
    DLB_CUDA_CHECK(
        cudaMemcpyAsync(d_y, h_y + row0, sizeof(double) * rows, cudaMemcpyHostToDevice, stream));
  }
  return flowmap_launch(flow_id, h_params, n_params, d_x, xdim, d_y, rows, t0, tf, rtol, atol,
                        max_step, first_step, d_flow_map, d_nfev, d_status, d_counters,
                        d_workspace, stream);
}

extern "C" int dlb_flowmap_rk45_resident(int flow_id, const double* h_params, int n_params,
                                         const double* d_x, int64_t xdim, const double* d_y,
                                         int64_t row0, int64_t rows, double t0, double tf,
                                         double rtol, double atol, double max_step,
                                         double first_step, double* d_flow_map, uint32_t* d_nfev,
                                         int8_t* d_status, uint64_t* d_counters,
                                         int64_t stat_row0, int64_t stat_rows,
                                         void* d_workspace, size_t workspace_bytes, void* stream_) {
  if (stat_row0 < row0 || stat_rows < 0 || stat_row0 + stat_rows > row0 + rows) {
    dlb_set_error("dlb_flowmap_rk45_resident: counted rows outside [row0, row0 + rows)");
    return DLB_ERR_ARG;
  }
  int rc = flowmap_check("dlb_flowmap_rk45_resident", flow_id, d_x, d_y, d_flow_map, d_counters,
                         d_workspace, xdim, rows, row0, atol);
  if (rc) return rc;
  if (workspace_bytes < 256) {
    dlb_set_error("workspace too small");
    return DLB_ERR_WORKSPACE;
  }
  return flowmap_launch(flow_id, h_params, n_params, d_x, xdim, d_y + row0, rows, t0, tf, rtol,
                        atol, max_step, first_step, d_flow_map, d_nfev, d_status, d_counters,
                        d_workspace, (cudaStream_t)stream_, stat_row0 - row0, stat_rows);
}

extern "C" int dlb_rk45_endpoints(int flow_id, const double* h_params, int n_params, int n_dim,
                                  const double* d_seeds, int64_t m, double t0, double tf,
                                  double rtol, const double* h_atol, double max_step,
                                  double first_step, double* d_out, uint32_t* d_nfev,
                                  int8_t* d_status, uint64_t* d_counters, void* d_workspace,
                                  size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!d_seeds || !d_out || !d_counters || !d_workspace || !h_atol || m < 0) {
    dlb_set_error("dlb_rk45_endpoints: bad argument");
    return DLB_ERR_ARG;
  }
  if (flow_id < 0 || flow_id >= DLB_FLOW_COUNT || k_flow_info[flow_id].dim != n_dim) {
    dlb_set_error("dlb_rk45_endpoints: flow %d does not have dimension %d", flow_id, n_dim);
    return DLB_ERR_ARG;
  }
  if (workspace_bytes < dlb_workspace_bytes(0, 0)) {
    dlb_set_error("workspace too small");
    return DLB_ERR_WORKSPACE;
  }
  Rk45Params P;
  memset(&P, 0, sizeof(P));
  int rc = fill_common(P, flow_id, h_params, n_params, t0, tf, rtol, max_step, first_step);
  if (rc) return rc;
  for (int i = 0; i < n_dim; ++i) {
    if (!(h_atol[i] >= 0)) {
      dlb_set_error("`atol` must be positive.");
      return DLB_ERR_ARG;
    }
    P.atol[i] = h_atol[i];
  }
  char* ws = (char*)d_workspace;
  DLB_CUDA_CHECK(cudaMemsetAsync(ws, 0, 256, stream));
  DLB_CUDA_CHECK(cudaMemsetAsync(d_counters, 0, sizeof(uint64_t) * DLB_CNT_WORDS, stream));
  if (m == 0) return DLB_OK;
  P.seeds = d_seeds;
  P.total = m;
  P.stat_lo = 0;
  P.stat_hi = m;
  P.xdim = 1;
  P.out = d_out;
  P.nfev = d_nfev;
  P.status = d_status;
  P.counters = (unsigned long long*)d_counters;
  P.next = (unsigned long long*)ws;
  rc = DLB_ERR_ARG;
  DLB_DISPATCH_FLOW(flow_id, {
    if constexpr (ID < DLB_FLOW_ABC) {
      rc = launch_rk45<ID, 2, false>(P, stream);
    } else {
      rc = launch_rk45<ID, 3, false>(P, stream);
    }
  });
  return rc;
}
