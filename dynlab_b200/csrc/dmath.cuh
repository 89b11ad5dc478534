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
// lg2/ex2 seed in FP32 refined by one third-order series step in FP64 (relative error < 1e-15).
// dm::rcp(x): MUFU.RCP64H seed + two Newton steps, for the error-norm scaling.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace dm {

//   [0] 2/pi   [1..3] pi/2 = P1 + P2 + P3 (three full binary64 terms, 159 bits)
//   [4] 1.5 * 2^52 (rint magic)
//   [5..10]  k_sin minimax coefficients S6..S1 on [-pi/4, pi/4] (fdlibm set)
//   [11..16] k_cos minimax coefficients C6..C1
__constant__ double DM_K[34] = {
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
    3.141592653589793,             // [18] np.pi = 2 P1 (sincospi2; high part of pi in sinN)
    0x1.1a62633145c07p-53,         // [19] 2 P2: pi - fl(pi) to 53 bits
    -0.16666666666666666,  // [20] sin Taylor coefficient of r^3
    0.008333333333333333,  // [21] sin Taylor coefficient of r^5
    -0.0001984126984126984,  // [22] sin Taylor coefficient of r^7
    2.7557319223985893e-06,  // [23] sin Taylor coefficient of r^9
    -2.505210838544172e-08,  // [24] sin Taylor coefficient of r^11
    1.6059043836821613e-10,  // [25] sin Taylor coefficient of r^13
    -7.647163731819816e-13,  // [26] sin Taylor coefficient of r^15
    2.8114572543455206e-15,  // [27] sin Taylor coefficient of r^17
    -8.22063524662433e-18,  // [28] sin Taylor coefficient of r^19
    1.9572941063391263e-20,  // [29] sin Taylor coefficient of r^21
    0x1.45f306dc9c883p-2,          // [30] 1/pi
    0.0, 0.0, 0.0,
};
// Constant provider: reads the table through the constant cache (LDCU into uniform registers).
// (A register-resident copy was measured: the extra ~34 registers cost more in occupancy than
// the ~100 LDCU per RK attempt it saves — 83 ms vs 76 ms on BASELINE config 3.)
struct KC {
  __device__ __forceinline__ double operator[](int i) const { return DM_K[i]; }
};
#define DM_2_OVER_PI K[0]
#define DM_PIO2_1 K[1]
#define DM_PIO2_2 K[2]
#define DM_PIO2_3 K[3]
#define DM_MAGIC K[4]
#define DM_S(i) K[11 - (i)]  // S1..S6
#define DM_C(i) K[17 - (i)]  // C1..C6

// high word of 105615.0 (0x40F9C8F000000000): |x| < 105615 <=> (hi & 0x7fffffff) < this
#define DM_TRIG_BIG_HI 0x40F9C8F0

__device__ __forceinline__ int hi_abs(double x) { return __double2hiint(x) & 0x7fffffff; }
__device__ __forceinline__ bool trig_is_big(double x) { return hi_abs(x) >= DM_TRIG_BIG_HI; }

__device__ __forceinline__ double xor_sign(double v, int signbit) {
  return __hiloint2double(__double2hiint(v) ^ signbit, __double2loint(v));
}

// x -> r in [-pi/4, pi/4] and quadrant q (only q mod 4 matters); valid for |x| < 105615.
template <class KT>
__device__ __forceinline__ double reduce_pio2(const KT& K, double x, int* q) {
  const double t = fma(x, DM_2_OVER_PI, DM_MAGIC);
  *q = __double2loint(t);
  const double fq = t - DM_MAGIC;
  double r = fma(-fq, DM_PIO2_1, x);
  r = fma(-fq, DM_PIO2_2, r);
  r = fma(-fq, DM_PIO2_3, r);
  return r;
}

// The same with pi/2 = P1 + P2 only (106 bits): the dropped term is |q| * 1.2e-33, below half an
// ulp of r unless |r| < |q| * 1e-17 - for the bounded arguments of an integration (|q| < 100)
// that is the exact lattice x = q pi/2 itself, where it changes r by 2e-17 relative.  One FP64
// instruction less per argument; used by TrigFast (the integrator's attempt loop) only.
template <class KT>
__device__ __forceinline__ double reduce_pio2_fast(const KT& K, double x, int* q) {
#ifdef DLB_TRIG_3TERM
  return reduce_pio2(K, x, q);
#else
  const double t = fma(x, DM_2_OVER_PI, DM_MAGIC);
  *q = __double2loint(t);
  const double fq = t - DM_MAGIC;
  return fma(-fq, DM_PIO2_2, fma(-fq, DM_PIO2_1, x));
#endif
}

template <class KT>
__device__ __forceinline__ double k_sin(const KT& K, double r, double r2) {
  double p = fma(DM_S(6), r2, DM_S(5));
  p = fma(p, r2, DM_S(4));
  p = fma(p, r2, DM_S(3));
  p = fma(p, r2, DM_S(2));
  p = fma(p, r2, DM_S(1));
  return fma(r * r2, p, r);
}

template <class KT>
__device__ __forceinline__ double k_cos(const KT& K, double r2) {
  double p = fma(DM_C(6), r2, DM_C(5));
  p = fma(p, r2, DM_C(4));
  p = fma(p, r2, DM_C(3));
  p = fma(p, r2, DM_C(2));
  p = fma(p, r2, DM_C(1));
  return fma(r2, fma(r2, p, -0.5), 1.0);  // 1 + r2 (-1/2 + r2 p)
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
  const KC K;
  int q;
  const double r = reduce_pio2(K, x, &q);
  const double r2 = r * r;
  const double sn = k_sin(K, r, r2), cs = k_cos(K, r2);
  *s = fix_sin(sn, cs, q);
  *c = fix_cos(sn, cs, q);
}
__device__ __forceinline__ double sin_fast(double x) {
  const KC K;
  int q;
  const double r = reduce_pio2(K, x, &q);
  const double r2 = r * r;
  return fix_sin(k_sin(K, r, r2), k_cos(K, r2), q);
}
__device__ __forceinline__ double cos_fast(double x) {
  const KC K;
  int q;
  const double r = reduce_pio2(K, x, &q);
  const double r2 = r * r;
  return fix_cos(k_sin(K, r, r2), k_cos(K, r2), q);
}

// Out-of-line library paths for huge / inf / nan arguments (CUDA's Payne-Hanek reduction);
// kept out of the hot loops so that they cost neither registers nor instruction cache.
static __device__ __noinline__ double2 sincos_slow(double x) {
  double s, c;
  ::sincos(x, &s, &c);
  return make_double2(s, c);
}
static __device__ __noinline__ double sin_slow(double x) { return ::sin(x); }
static __device__ __noinline__ double cos_slow(double x) { return ::cos(x); }

// General entry points: one well-predicted branch for huge / inf / nan.
__device__ __forceinline__ void sincos(double x, double* s, double* c) {
  if (trig_is_big(x)) {
    const double2 sc = sincos_slow(x);
    *s = sc.x;
    *c = sc.y;
    return;
  }
  sincos_fast(x, s, c);
}
__device__ __forceinline__ double sin(double x) { return trig_is_big(x) ? sin_slow(x) : sin_fast(x); }
__device__ __forceinline__ double cos(double x) { return trig_is_big(x) ? cos_slow(x) : cos_fast(x); }

// ---- trig policies ---------------------------------------------------------------------
// TrigFast: unconditional fast paths, branch-free; remembers the largest |argument| (high
// word) it saw.  The integrator runs a whole RK attempt with it as ONE basic block — so the
// scheduler can overlap the independent polynomial chains of different stages — and tests
// ok() once at the end; in the (practically never taken) failing case it redoes the attempt
// with TrigSafe.  TrigSafe: per-call branch to the out-of-line library path.
struct TrigFast {
  int big = 0;
  __device__ __forceinline__ bool ok() const { return big < DM_TRIG_BIG_HI; }
  __device__ __forceinline__ void sincos(double x, double* s, double* c) {
    big = max(big, hi_abs(x));
    sincos_fast(x, s, c);
  }
  __device__ __forceinline__ double sin(double x) {
    big = max(big, hi_abs(x));
    return sin_fast(x);
  }
  __device__ __forceinline__ double cos(double x) {
    big = max(big, hi_abs(x));
    return cos_fast(x);
  }
  // two independent sincos, polynomial chains interleaved
  __device__ __forceinline__ void sincos2(double a0, double a1, double* s0, double* c0,
                                          double* s1, double* c1) {
    const KC K;
    big = max(big, max(hi_abs(a0), hi_abs(a1)));
    int q0, q1;
    const double r0 = reduce_pio2_fast(K, a0, &q0), r1 = reduce_pio2_fast(K, a1, &q1);
    const double z0 = r0 * r0, z1 = r1 * r1;
    double ps0 = fma(DM_S(6), z0, DM_S(5)), ps1 = fma(DM_S(6), z1, DM_S(5));
    double pc0 = fma(DM_C(6), z0, DM_C(5)), pc1 = fma(DM_C(6), z1, DM_C(5));
#pragma unroll
    for (int k = 4; k >= 1; --k) {
      ps0 = fma(ps0, z0, DM_S(k));
      ps1 = fma(ps1, z1, DM_S(k));
      pc0 = fma(pc0, z0, DM_C(k));
      pc1 = fma(pc1, z1, DM_C(k));
    }
    const double sn0 = fma(r0 * z0, ps0, r0), sn1 = fma(r1 * z1, ps1, r1);
    const double cs0 = fma(z0, fma(z0, pc0, -0.5), 1.0), cs1 = fma(z1, fma(z1, pc1, -0.5), 1.0);
    *s0 = fix_sin(sn0, cs0, q0);
    *c0 = fix_cos(sn0, cs0, q0);
    *s1 = fix_sin(sn1, cs1, q1);
    *c1 = fix_cos(sn1, cs1, q1);
  }
  // sin / cos of pi*a0 and of a1*pi (the double gyre's sin(pi f), cos(pi y)), formed like the
  // reference: fl(pi * a) first.  (Reducing a itself in units of pi - q = rint(2a), r = pi (a -
  // q/2), 4 FP64 instructions instead of 5 - was measured 3.7 % faster on K1 and is NOT used: it
  // makes sin(pi * 1.0) exactly 0 where the reference has sin(fl(pi)) = 1.2e-16, and the
  // reference's eigenvector signs on the invariant wall y = 1 are decided by that residue:
  // test_lcs_fields_against_reference failed.)
  __device__ __forceinline__ void sincospi2(double a0, double a1, double* s0, double* c0,
                                            double* s1, double* c1) {
    sincos2(DM_K[18] * a0, a1 * DM_K[18], s0, c0, s1, c1);
  }
  // NS independent sines (COS: cosines), chains interleaved; out has stride `stride` doubles
  // Only the sine is wanted here (the stage-time factor sin(w t) of the forced flows), so the
  // argument is reduced modulo pi instead of pi/2: sin x = (-1)^q sin r, |r| <= pi/2, and sin r
  // is ONE odd polynomial to r^21 (Taylor; the first dropped term is 1.2e-18) instead of the sin
  // and cos kernels plus a select between them: 16 FP64 instructions per sine instead of 19, and
  // a shift + XOR instead of two selects and a quadrant test.  <= 1.8 ulp (99 % within 1 ulp),
  // checked against long-double sin on 2.2e5 points; pi = K[18] + K[19] (106 bits).
  template <int NS>
  __device__ __forceinline__ void sinN_halfturn(const double* arg, double* out, int stride) {
    const KC K;
    double r[NS], z[NS], p[NS];
    int q[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      big = max(big, hi_abs(arg[s]));
      const double t = fma(arg[s], K[30], DM_MAGIC);
      q[s] = __double2loint(t);
      const double fq = t - DM_MAGIC;
      r[s] = fma(-fq, K[19], fma(-fq, K[18], arg[s]));
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      z[s] = r[s] * r[s];
      p[s] = fma(K[29], z[s], K[28]);
    }
#pragma unroll
    for (int k = 27; k >= 20; --k) {
#pragma unroll
      for (int s = 0; s < NS; ++s) p[s] = fma(p[s], z[s], K[k]);
    }
#pragma unroll
    for (int s = 0; s < NS; ++s)
      out[s * stride] = xor_sign(fma(r[s] * z[s], p[s], r[s]), q[s] << 31);
  }
  template <int NS, bool COS>
  __device__ __forceinline__ void sinN(const double* arg, double* out, int stride) {
#ifndef DLB_SIN_QUADRANT
    if constexpr (!COS) {
      sinN_halfturn<NS>(arg, out, stride);
      return;
    }
#endif
    const KC K;
    double r[NS], z[NS], ps[NS], pc[NS];
    int q[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      big = max(big, hi_abs(arg[s]));
      r[s] = reduce_pio2_fast(K, arg[s], &q[s]);
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      z[s] = r[s] * r[s];
      ps[s] = fma(DM_S(6), z[s], DM_S(5));
      pc[s] = fma(DM_C(6), z[s], DM_C(5));
    }
#pragma unroll
    for (int k = 4; k >= 1; --k) {
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        ps[s] = fma(ps[s], z[s], DM_S(k));
        pc[s] = fma(pc[s], z[s], DM_C(k));
      }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const double sn = fma(r[s] * z[s], ps[s], r[s]);
      const double cs = fma(z[s], fma(z[s], pc[s], -0.5), 1.0);
      out[s * stride] = COS ? fix_cos(sn, cs, q[s]) : fix_sin(sn, cs, q[s]);
    }
  }
};

struct TrigSafe {
  __device__ __forceinline__ bool ok() const { return true; }
  __device__ __forceinline__ void sincos(double x, double* s, double* c) { dm::sincos(x, s, c); }
  __device__ __forceinline__ double sin(double x) { return dm::sin(x); }
  __device__ __forceinline__ double cos(double x) { return dm::cos(x); }
  __device__ __forceinline__ void sincos2(double a0, double a1, double* s0, double* c0,
                                          double* s1, double* c1) {
    dm::sincos(a0, s0, c0);
    dm::sincos(a1, s1, c1);
  }
  // sin / cos of pi*a0, pi*a1 exactly as the reference forms them: fl(pi * a) first
  __device__ __forceinline__ void sincospi2(double a0, double a1, double* s0, double* c0,
                                            double* s1, double* c1) {
    dm::sincos(3.141592653589793 * a0, s0, c0);
    dm::sincos(a1 * 3.141592653589793, s1, c1);
  }
  template <int NS, bool COS>
  __device__ __forceinline__ void sinN(const double* arg, double* out, int stride) {
#pragma unroll
    for (int s = 0; s < NS; ++s) out[s * stride] = COS ? dm::cos(arg[s]) : dm::sin(arg[s]);
  }
};

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
  // FP32 seed through the MUFU lg2 / ex2 units (.ftz: s is far from the subnormal range, so the
  // scaling fix-ups of the IEEE-compliant forms are dead weight): ~1e-6 relative
  float lg, sd;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"((float)s));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(sd) : "f"(-0.1f * lg));
  const double r = (double)sd;
  // with e = 1 - s r^10 (|e| ~ 1e-5):  s^(-1/10) = r (1 - e)^(-1/10)
  //   = r (1 + e/10 + 11 e^2/200 + 77 e^3/2000 + O(e^4)),  remainder < 1e-20
  const double r2 = r * r;
  const double r4 = r2 * r2;
  const double r10 = r4 * r4 * r2;
  const double e = fma(-s, r10, 1.0);
  const double p = fma(fma(0.0385, e, 0.055), e, 0.1);
  return fma(r, e * p, r);
}

// max / min by one compare and a select.  fmax / fmin compile to DSETP.MAX plus five integer
// instructions of NaN bookkeeping; on this kernel every issue slot next to the FP64 pipe counts.
// NaN behaviour: a NaN in `b` is returned, a NaN in `a` is dropped.
__device__ __forceinline__ double max_sel(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double min_sel(double a, double b) { return a < b ? a : b; }

}  // namespace dm
