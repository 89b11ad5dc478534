// FP64 device math used by the flow twins and the integrator (sm_100a).
//
// The integrator kernel is issue-bound around its FP64 instructions (~1 IPC per SM
// sub-partition, a DFMA issues every other cycle), so these routines are written to spend as
// few non-FP64 instructions as possible:
//   * every constant lives in one __constant__ table — sm_100a has no 64-bit immediates, a
//     literal costs two UMOVs per use, consecutive table entries load pairwise (LDCU.128);
//   * round-to-nearest by the 1.5*2^52 magic constant (quadrant = low word of the sum) instead
//     of FRND + F2I + I2F;
//   * quadrant fix-up with one select pair per output and an integer XOR on the sign bit;
//   * the large-argument test (|x| >= 105615, where the 3-term Cody-Waite reduction runs out
//     of bits, and inf/nan) is an integer compare on the high word; callers branch once per
//     RHS evaluation to CUDA's own sin/cos/sincos (Payne-Hanek) for those;
//   * flows.cuh interleaves the polynomial chains of independent arguments (the five stage
//     times of an RK attempt; the two spatial arguments of the double gyre) so that each
//     coefficient is loaded once and the DFMA latency is covered by ILP.
// Accuracy: the reduction keeps ~1 ulp relative accuracy of r even next to multiples of pi/2
// (with FMA the step x - q*P1 is exact); kernels are the fdlibm minimax sets on
// [-pi/4, pi/4] (< 1 ulp).  Checked on the device against numpy in tests/test_gpu_parity.py
// (test_device_trig_accuracy: <= 2 ulp over small, large and near-multiple-of-pi arguments).
//
// dm::pow_m01(s) = s^(-1/10) = err^(-1/5) for s = err^2: the step-size controller's
// `error_norm ** -0.2` (scipy/integrate/_ivp/rk.py:153,164) without the square root: MUFU
// lg2/ex2 seed in FP32 refined by two Newton steps in FP64 (relative error < 1e-15).
// dm::rcp(x): MUFU.RCP64H seed + two Newton steps, for the error-norm scaling.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace dm {

//   [0] 2/pi   [1..3] pi/2 = P1 + P2 + P3 (three full binary64 terms, 159 bits)
//   [4] 1.5 * 2^52 (rint magic)
//   [5..10]  k_sin minimax coefficients S6..S1 on [-pi/4, pi/4] (fdlibm set)
//   [11..16] k_cos minimax coefficients C6..C1
__constant__ double DM_K[18] = {
    0x1.45f306dc9c883p-1,          // 2/pi
    0x1.921fb54442d18p+0,          // P1
    0x1.1a62633145c07p-54,         // P2
    -0x1.f1976b7ed8fbcp-110,       // P3
    6755399441055744.0,            // 1.5 * 2^52
    1.58969099521155010221e-10,    // S6
    -2.50507602534068634195e-08,   // S5
    2.75573137070700676789e-06,    // S4
    -1.98412698298579493134e-04,   // S3
    8.33333333332248946124e-03,    // S2
    -1.66666666666666324348e-01,   // S1
    -1.13596475577881948265e-11,   // C6
    2.08757232129817482790e-09,    // C5
    -2.75573143513906633035e-07,   // C4
    2.48015872894767294178e-05,    // C3
    -1.38888888888741095749e-03,   // C2
    4.16666666666666019037e-02,    // C1
    0.0,
};
#define DM_2_OVER_PI dm::DM_K[0]
#define DM_PIO2_1 dm::DM_K[1]
#define DM_PIO2_2 dm::DM_K[2]
#define DM_PIO2_3 dm::DM_K[3]
#define DM_MAGIC dm::DM_K[4]
#define DM_S(i) dm::DM_K[11 - (i)]  // S1..S6
#define DM_C(i) dm::DM_K[17 - (i)]  // C1..C6

// high word of 105615.0 (0x40F9C8F000000000): |x| < 105615 <=> (hi & 0x7fffffff) < this
#define DM_TRIG_BIG_HI 0x40F9C8F0

__device__ __forceinline__ int hi_abs(double x) { return __double2hiint(x) & 0x7fffffff; }
__device__ __forceinline__ bool trig_is_big(double x) { return hi_abs(x) >= DM_TRIG_BIG_HI; }

__device__ __forceinline__ double xor_sign(double v, int signbit) {
  return __hiloint2double(__double2hiint(v) ^ signbit, __double2loint(v));
}

// x -> r in [-pi/4, pi/4] and quadrant q (only q mod 4 matters); valid for |x| < 105615.
__device__ __forceinline__ double reduce_pio2(double x, int* q) {
  const double t = fma(x, DM_2_OVER_PI, DM_MAGIC);
  *q = __double2loint(t);
  const double fq = t - DM_MAGIC;
  double r = fma(-fq, DM_PIO2_1, x);
  r = fma(-fq, DM_PIO2_2, r);
  r = fma(-fq, DM_PIO2_3, r);
  return r;
}

__device__ __forceinline__ double k_sin(double r, double r2) {
  double p = fma(DM_S(6), r2, DM_S(5));
  p = fma(p, r2, DM_S(4));
  p = fma(p, r2, DM_S(3));
  p = fma(p, r2, DM_S(2));
  p = fma(p, r2, DM_S(1));
  return fma(r * r2, p, r);
}

__device__ __forceinline__ double k_cos(double r2) {
  double p = fma(DM_C(6), r2, DM_C(5));
  p = fma(p, r2, DM_C(4));
  p = fma(p, r2, DM_C(3));
  p = fma(p, r2, DM_C(2));
  p = fma(p, r2, DM_C(1));
  return fma(r2 * r2, p, fma(-0.5, r2, 1.0));  // 1 - r2/2 + r2^2 p
}

// quadrant fix-up: (sin r, cos r, q) -> sin x, cos x
__device__ __forceinline__ double fix_sin(double sn, double cs, int q) {
  return xor_sign((q & 1) ? cs : sn, (q << 30) & 0x80000000);
}
__device__ __forceinline__ double fix_cos(double sn, double cs, int q) {
  return xor_sign((q & 1) ? sn : cs, ((q + 1) << 30) & 0x80000000);
}

// Fast paths (|x| < 105615 assumed — caller checks trig_is_big).
__device__ __forceinline__ void sincos_fast(double x, double* s, double* c) {
  int q;
  const double r = reduce_pio2(x, &q);
  const double r2 = r * r;
  const double sn = k_sin(r, r2), cs = k_cos(r2);
  *s = fix_sin(sn, cs, q);
  *c = fix_cos(sn, cs, q);
}
__device__ __forceinline__ double sin_fast(double x) {
  int q;
  const double r = reduce_pio2(x, &q);
  const double r2 = r * r;
  return fix_sin(k_sin(r, r2), k_cos(r2), q);
}
__device__ __forceinline__ double cos_fast(double x) {
  int q;
  const double r = reduce_pio2(x, &q);
  const double r2 = r * r;
  return fix_cos(k_sin(r, r2), k_cos(r2), q);
}

// General entry points: one well-predicted branch to the library for huge / inf / nan.
__device__ __forceinline__ void sincos(double x, double* s, double* c) {
  if (trig_is_big(x)) {
    ::sincos(x, s, c);
    return;
  }
  sincos_fast(x, s, c);
}
__device__ __forceinline__ double sin(double x) { return trig_is_big(x) ? ::sin(x) : sin_fast(x); }
__device__ __forceinline__ double cos(double x) { return trig_is_big(x) ? ::cos(x) : cos_fast(x); }

// 1/x for finite normal x (no special-case handling: used on positive error scales).
__device__ __forceinline__ double rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));  // MUFU.RCP64H, ~20 bits
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// s^(-1/10) for s in [1e-12, 1e8]  (= err^(-1/5) for err = sqrt(s) in [1e-6, 1e4]).
__device__ __forceinline__ double pow_m01(double s) {
  double r = (double)exp2f(-0.1f * __log2f((float)s));  // ~1e-6 relative
  // Newton on g(r) = r^-10 - s:  r <- r * (11 - s r^10) / 10   (error -> 5.5 e^2 per step)
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double r2 = r * r;
    const double r4 = r2 * r2;
    const double r10 = r4 * r4 * r2;
    r = r * (fma(-s, r10, 11.0) * 0.1);
  }
  return r;
}

}  // namespace dm
