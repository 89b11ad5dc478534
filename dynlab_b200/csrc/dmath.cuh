// FP64 device math used by the flow twins and the integrator.
//
// dm::sincos / dm::sin / dm::cos: Cody-Waite range reduction by pi/2 (three FMA steps, used
// for |x| < 105615; larger arguments, inf and nan take CUDA's Payne-Hanek path through
// ::sincos) followed by the classical minimax kernels on
// [-pi/4, pi/4] (fdlibm k_sin/k_cos coefficient sets, < 1 ulp).  Against CUDA's own
// sincos() this saves the slow-path plumbing and lets sin and cos of one argument share the
// reduction; accuracy is checked on the device against the CPU libm in tests/test_gpu_math.py.
//
// dm::pow_m02(x) = x^(-1/5) for x > 0, used by the step-size controller
// (scipy/integrate/_ivp/rk.py:153,164: SAFETY * error_norm ** -0.2): MUFU lg2/ex2 seed in
// FP32 refined by two Newton steps in FP64 (relative error < 4e-16).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#ifndef DLB_CUSTOM_MATH
#define DLB_CUSTOM_MATH 1
#endif

namespace dm {

#if DLB_CUSTOM_MATH

// All FP64 constants live in one __constant__ table: sm_100a has no 64-bit immediates, so a
// literal costs two UMOVs per use, while consecutive table entries load pairwise with one
// LDCU.128 into uniform registers.  Order = order of use.
//   [0] 2/pi   [1..3] pi/2 = P1 + P2 + P3 (three full binary64 terms, 159 bits)
//   [4..9]  k_sin minimax coefficients S6..S1 on [-pi/4, pi/4] (fdlibm set)
//   [10..15] k_cos minimax coefficients C6..C1
// With FMA the first reduction step x - q*P1 is exact (the difference is a multiple of 2^-52
// below 1 in magnitude) and the two correction steps round only at the size of the reduced
// argument, so r keeps ~1 ulp relative accuracy even next to multiples of pi/2.
__constant__ double DM_K[16] = {
    0x1.45f306dc9c883p-1,          // 2/pi
    0x1.921fb54442d18p+0,          // P1
    0x1.1a62633145c07p-54,         // P2
    -0x1.f1976b7ed8fbcp-110,       // P3
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
};
#define DM_2_OVER_PI DM_K[0]
#define DM_PIO2_1 DM_K[1]
#define DM_PIO2_2 DM_K[2]
#define DM_PIO2_3 DM_K[3]
#define DM_S6 DM_K[4]
#define DM_S5 DM_K[5]
#define DM_S4 DM_K[6]
#define DM_S3 DM_K[7]
#define DM_S2 DM_K[8]
#define DM_S1 DM_K[9]
#define DM_C6 DM_K[10]
#define DM_C5 DM_K[11]
#define DM_C4 DM_K[12]
#define DM_C3 DM_K[13]
#define DM_C2 DM_K[14]
#define DM_C1 DM_K[15]
#define DM_TRIG_FAST_MAX 105615.0

__device__ __forceinline__ double k_sin(double r, double r2) {
  double p = fma(DM_S6, r2, DM_S5);
  p = fma(p, r2, DM_S4);
  p = fma(p, r2, DM_S3);
  p = fma(p, r2, DM_S2);
  p = fma(p, r2, DM_S1);
  return fma(r * r2, p, r);
}

__device__ __forceinline__ double k_cos(double r2) {
  double p = fma(DM_C6, r2, DM_C5);
  p = fma(p, r2, DM_C4);
  p = fma(p, r2, DM_C3);
  p = fma(p, r2, DM_C2);
  p = fma(p, r2, DM_C1);
  // 1 - r2/2 + r2^2 * p
  return fma(r2 * r2, p, fma(-0.5, r2, 1.0));
}

// Reduce x to r in [-pi/4, pi/4], returns quadrant q (mod 4 matters).
__device__ __forceinline__ double reduce_pio2(double x, int* q) {
  const double fq = rint(x * DM_2_OVER_PI);
  *q = (int)fq;
  double r = fma(-fq, DM_PIO2_1, x);
  r = fma(-fq, DM_PIO2_2, r);
  r = fma(-fq, DM_PIO2_3, r);
  return r;
}

__device__ __forceinline__ void sincos(double x, double* s, double* c) {
  if (!(fabs(x) < DM_TRIG_FAST_MAX)) {  // rare: huge / inf / nan -> library slow path
    ::sincos(x, s, c);
    return;
  }
  int q;
  const double r = reduce_pio2(x, &q);
  const double r2 = r * r;
  const double sn = k_sin(r, r2), cs = k_cos(r2);
  const double a = (q & 1) ? cs : sn;
  const double b = (q & 1) ? sn : cs;
  *s = (q & 2) ? -a : a;
  *c = ((q + 1) & 2) ? -b : b;
}

__device__ __forceinline__ double sin(double x) {
  if (!(fabs(x) < DM_TRIG_FAST_MAX)) return ::sin(x);
  int q;
  const double r = reduce_pio2(x, &q);
  const double r2 = r * r;
  const double v = (q & 1) ? k_cos(r2) : k_sin(r, r2);
  return (q & 2) ? -v : v;
}

__device__ __forceinline__ double cos(double x) {
  if (!(fabs(x) < DM_TRIG_FAST_MAX)) return ::cos(x);
  int q;
  const double r = reduce_pio2(x, &q);
  const double r2 = r * r;
  const double v = (q & 1) ? k_sin(r, r2) : k_cos(r2);
  return ((q + 1) & 2) ? -v : v;
}

// x^(-1/5) for x in [1e-6, 1e4] (the controller clamps: outside that range the
// MIN_FACTOR / MAX_FACTOR limits of rk.py:10-11 decide, see rk45.cu).
__device__ __forceinline__ double pow_m02(double x) {
  double r = (double)exp2f(-0.2f * __log2f((float)x));  // ~1e-6 relative
  // Newton on g(r) = r^-5 - x:  r <- r * (6 - x r^5) / 5   (error -> 3 e^2 per step)
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double r2 = r * r;
    const double r5 = r2 * r2 * r;
    r = r * (fma(-x, r5, 6.0) * 0.2);
  }
  return r;
}

#else  // plain CUDA libm (kept for A/B measurements)

__device__ __forceinline__ void sincos(double x, double* s, double* c) { ::sincos(x, s, c); }
__device__ __forceinline__ double sin(double x) { return ::sin(x); }
__device__ __forceinline__ double cos(double x) { return ::cos(x); }
__device__ __forceinline__ double pow_m02(double x) { return ::pow(x, -0.2); }

#endif

}  // namespace dm
