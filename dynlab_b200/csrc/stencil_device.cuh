// Device-side building blocks shared by the stencil kernels (stencil.cu) and the fused
// flow-evaluation + stencil kernels (eulerflow.cu): axis / parameter structs, the np.gradient
// edge formulas, the 2x2 eigen epilogues (`emit`), loaders and the boundary kernel.  Everything
// here is header-only and lives in an anonymous namespace (one copy per translation unit);
// the host helpers with external linkage are declared at the bottom and defined in stencil.cu.
#pragma once
#include <math.h>
#include <math_constants.h>

#include "common.cuh"

// ---- types shared across translation units (external linkage) ----
namespace dlbst {

// Per-axis gradient description (device pointers into the workspace).
struct Axis {
  const double* ca;  // [n] 3-point coefficient on f[i-1] (or shifted at the edges)
  const double* cb;  // [n]
  const double* cc;  // [n]
  long long n;
  double inv2h;    // uniform: 1/(2h)
  double w_first;  // edge_order 1: 1/dx_0
  double w_last;   // edge_order 1: 1/dx_{n-1}
  int uniform;
};

struct StencilParams {
  Axis ax, ay;
  const double2* ay4;  // y coefficients interleaved (a, b | c, 0) per row, 32-byte records
  const double* xs;    // the coordinates themselves (cached upload, fused flow kernels)
  const double* ys;
  long long nloc, grow0;        // local slab: rows [grow0, grow0+nloc)
  long long out_row0, out_rows; // global rows to produce
  int edge_order;
  // inputs
  const double2* fm;  // FTLE: [nloc][xdim] (Phi_x, Phi_y)
  const double* u;    // Euler: [nloc][xdim]
  const double* v;
  const double* xi_in;  // ridge: [ydim][xdim][2]
  // outputs
  double* field;
  double2* xi;
  uint8_t* mask;
  double* conc;
  double scale;      // FTLE: 1/(2|T|)
  double threshold;  // ridge
  int use_threshold;
  int xi_mode;
  int negate;
  // FTLE with MODE == FTLE_PUSH (K2 fused with the row gather): every value is also stored at
  // the same offset of these buffers (peer GPUs' copies of the field, addressed over NVLink)
  int npeer;
  double* peer[DLB_MAX_PEERS];
};

}  // namespace dlbst

namespace {

using dlbst::Axis;
using dlbst::StencilParams;


constexpr unsigned FULL = 0xffffffffu;
constexpr int ST_BLOCK = 256;
constexpr int ST_ROWS = 32;  // rows per warp task


enum { KIND_FTLE = 0, KIND_EULER = 1 };
// build-time switch of the tile kernel's epilogue stores (see st_f64_if below)
#ifndef DLB_ASM_STORE
#define DLB_ASM_STORE 1
#endif
// MODE of a KIND_FTLE kernel: plain, or with the peer stores of the fused gather
enum { FTLE_PLAIN = 0, FTLE_PUSH = 1 };

// Store under a predicate WITHOUT a branch.  `if (store) p[o] = v` lets the compiler wrap the
// whole epilogue that produced v into a BSSY / BRA region per row (it did: the rows of a stage
// were separate basic blocks, their dependent FP64 chains could not be interleaved; ncu's top
// stall was the fixed-latency "wait").  An asm volatile store is executed unconditionally as far
// as the optimiser knows, so the value is always computed and the rows of a stage stay ONE
// basic block; the hardware predicate only suppresses the write.
__device__ __forceinline__ void st_f64_if(double* p, double v, bool pred) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p st.global.f64 [%0], %1;\n\t}" ::"l"(p),
      "d"(v), "r"((int)pred));  // no "memory" clobber: nothing in these kernels reads what they
                                // write, and the clobber would pin the next row's shared-memory
                                // loads behind this store (the rows' chains could not overlap)
}
__device__ __forceinline__ void st_u8_if(uint8_t* p, int v, bool pred) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p st.global.u8 [%0], %1;\n\t}" ::"l"(p),
      "r"(v), "r"((int)pred));
}

template <int KIND, int MODE>
__device__ __forceinline__ void store_ftle(const StencilParams& P, long long o, double out) {
  P.field[o] = out;
  if constexpr (KIND == KIND_FTLE && MODE == FTLE_PUSH) {
#pragma unroll 1
    for (int k = 0; k < P.npeer; ++k) P.peer[k][o] = out;  // posted NVLink writes
  }
}

// ---- symmetric 2x2 eigen helpers ----

// Reciprocal / square root / reciprocal square root from the MUFU seeds plus Newton steps in
// FP64 (1-2 ulp), no special-case branches.  Arguments are positive and normal here.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));  // MUFU.RCP64H, ~20 bits
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}
// g ~ sqrt(x), r ~ 1/sqrt(x): coupled (Goldschmidt) iteration, final residual step on each
__device__ __forceinline__ void fast_sqrt_rsqrt(double x, double* g_out, double* r_out) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double g = x * r, h = 0.5 * r;
  double e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  *g_out = fma(fma(-g, g, x), h, g);
  *r_out = h + h;
}

// Eigenpair with the ordering and eigenvector sign np.linalg.eig (LAPACK dgeev: dlahqr's
// deflation test + dlanv2's standardisation) produces for a symmetric 2x2 matrix, followed by
// the reference's argsort selection (R:lagrangian.py:150-152, R:eulerian.py:327-330).
__device__ __forceinline__ void eig_sym_lapack(double a, double b, double d, bool want_max,
                                               double* w, double* vx, double* vy) {
  const double ulp = 2.220446049250313e-16;      // dlamch('P')
  const double smlnum = 2.0041683600089728e-292;  // dlamch('S') * (n / ulp), n = 2
  bool deflate = (b == 0.0) || (fabs(b) <= smlnum);
  if (!deflate) {  // dlahqr's conservative small-subdiagonal test
    const double tst = fabs(a) + fabs(d);
    if (fabs(b) <= ulp * tst) {
      const double ab = fabs(b), ba = fabs(b);
      const double aa = fmax(fabs(d), fabs(a - d)), bb = fmin(fabs(d), fabs(a - d));
      const double s = aa + ab;
      if (ba * (ab / s) <= fmax(smlnum, ulp * (bb * (aa / s)))) deflate = true;
    }
  }
  double w0, w1, cs, sn;
  if (deflate) {
    w0 = a;
    w1 = d;
    cs = 1.0;
    sn = 0.0;
  } else {  // dlanv2 with c == b: always the real-eigenvalue branch
    // LAPACK:  p = (a-d)/2, scale = max(|p|, |b|), z = p/scale*p + |b|/scale*|b|,
    //          z = p + sign(sqrt(scale)*sqrt(z), p), w0 = d + z, w1 = d - |b|/z*|b|,
    //          tau = dlapy2(b, z), cs = z/tau, sn = b/tau.
    // Same quantities with reciprocals and one sqrt / one rsqrt from the MUFU seeds (the five
    // IEEE divisions, two sqrt and the hypot of the literal transcription made this epilogue,
    // not HBM, the bound of the Xi kernels): sqrt(scale)*sqrt(z) = scale*sqrt((p/scale)^2 +
    // (b/scale)^2), |z| >= |b| > 0, tau = |z| sqrt(1 + (b/z)^2).  Within 2 ulp of LAPACK; the
    // signs (of p, b, z) that fix the eigenvector orientation are untouched.
    const double p = 0.5 * (a - d);
    const double ab = fabs(b), ap = fabs(p);
    const double sc = (ap > ab) ? ap : ab;
    const double rs = fast_rcp(sc);
    const double ps = p * rs, bs = ab * rs;
    double g, r;
    fast_sqrt_rsqrt(fma(ps, ps, bs * bs), &g, &r);
    const double z = p + copysign(sc * g, p);
    const double rz = fast_rcp(z);
    w0 = d + z;
    w1 = d - (ab * rz) * ab;
    const double t = b * rz;
    fast_sqrt_rsqrt(fma(t, t, 1.0), &g, &r);
    cs = copysign(r, z);
    sn = t * cs;
  }
  // eigenvectors: column 0 = (cs, sn) for w0, column 1 = (-sn, cs) for w1.
  // argsort ascending: idx[-1] is 0 only if w0 > w1; idx[0] is 1 only if w0 > w1.
  const bool first_is_larger = (w0 > w1);
  const bool take0 = want_max ? first_is_larger : !first_is_larger;
  *w = take0 ? w0 : w1;
  *vx = take0 ? cs : -sn;
  *vy = take0 ? sn : cs;
}

// R:src/dynlab/utils.py:41-45
__device__ __forceinline__ void force_eigenvector(double* vx, double* vy) {
  const double a0 = fabs(*vx), a1 = fabs(*vy);
  if ((a0 >= a1 && *vx < 0) || (a0 < a1 && *vy < 0)) {
    *vx = -*vx;
    *vy = -*vy;
  }
}

// ---- gradient of a 2-component field at one node ----
// s0, s1, s2 are the three samples the formula uses along the axis (already shifted at the
// edges); `kind`: 0 interior, 1 first edge, 2 last edge.
__device__ __forceinline__ double grad3(const Axis& A, long long i, int edge_order, double m,
                                        double c, double p) {
  // interior: (m, c, p) = f[i-1], f[i], f[i+1]
  if (A.uniform) return (p - m) * A.inv2h;
  return fma(A.cc[i], p, fma(A.cb[i], c, A.ca[i] * m));
}

struct Grad2 {
  double d0dx, d0dy, d1dx, d1dy;
};

// Loader: at(local_row, col) -> double2.
template <class Loader>
__device__ __forceinline__ void edge_axis_x(const StencilParams& P, const Loader& ld,
                                            long long lrow, long long j, double2 C, double2 L,
                                            double2 R, double* g0, double* g1) {
  const Axis& A = P.ax;
  const long long n = A.n;
  if (j > 0 && j < n - 1) {
    *g0 = grad3(A, j, P.edge_order, L.x, C.x, R.x);
    *g1 = grad3(A, j, P.edge_order, L.y, C.y, R.y);
  } else if (P.edge_order == 1) {
    if (j == 0) {  // (f[1] - f[0]) / dx_0
      *g0 = (R.x - C.x) * A.w_first;
      *g1 = (R.y - C.y) * A.w_first;
    } else {  // (f[n-1] - f[n-2]) / dx_n
      *g0 = (C.x - L.x) * A.w_last;
      *g1 = (C.y - L.y) * A.w_last;
    }
  } else {
    const double a = A.ca[j], b = A.cb[j], c = A.cc[j];
    if (j == 0) {  // a f[0] + b f[1] + c f[2]
      const double2 X = ld.at(lrow, 2);
      *g0 = fma(c, X.x, fma(b, R.x, a * C.x));
      *g1 = fma(c, X.y, fma(b, R.y, a * C.y));
    } else {  // a f[n-3] + b f[n-2] + c f[n-1]
      const double2 X = ld.at(lrow, n - 3);
      *g0 = fma(c, C.x, fma(b, L.x, a * X.x));
      *g1 = fma(c, C.y, fma(b, L.y, a * X.y));
    }
  }
}

template <class Loader>
__device__ __forceinline__ void edge_axis_y(const StencilParams& P, const Loader& ld,
                                            long long gi, long long j, double2 C, double2 U,
                                            double2 D, double* g0, double* g1) {
  const Axis& A = P.ay;
  const long long n = A.n;
  if (gi > 0 && gi < n - 1) {
    *g0 = grad3(A, gi, P.edge_order, U.x, C.x, D.x);
    *g1 = grad3(A, gi, P.edge_order, U.y, C.y, D.y);
  } else if (P.edge_order == 1) {
    if (gi == 0) {
      *g0 = (D.x - C.x) * A.w_first;
      *g1 = (D.y - C.y) * A.w_first;
    } else {
      *g0 = (C.x - U.x) * A.w_last;
      *g1 = (C.y - U.y) * A.w_last;
    }
  } else {
    const double a = A.ca[gi], b = A.cb[gi], c = A.cc[gi];
    if (gi == 0) {
      const double2 X = ld.at(2 - P.grow0, j);
      *g0 = fma(c, X.x, fma(b, D.x, a * C.x));
      *g1 = fma(c, X.y, fma(b, D.y, a * C.y));
    } else {
      const double2 X = ld.at(n - 3 - P.grow0, j);
      *g0 = fma(c, C.x, fma(b, U.x, a * X.x));
      *g1 = fma(c, C.y, fma(b, U.y, a * X.y));
    }
  }
}

// Loaders: at(local_row, col) for the edge formulas; Cursor for the rolling interior walk: a
// task-base pointer plus a 32-bit element offset (one IMAD.WIDE per access instead of 64-bit
// pointer bumps; a task spans at most ST_ROWS + 2 rows, far below 2^31 elements).
struct LoaderAoS {
  const double2* p;
  long long xdim;
  __device__ __forceinline__ double2 at(long long lrow, long long j) const {
    return __ldg(p + lrow * xdim + j);
  }
  bool aligned16() const { return ((uintptr_t)p & 15u) == 0; }
  struct Cursor {
    const double2* q;
    __device__ __forceinline__ double2 load(int off) const { return __ldg(q + off); }
  };
  __device__ __forceinline__ Cursor cursor(long long lrow, long long j) const {
    return Cursor{p + lrow * xdim + j};
  }
};

struct LoaderSoA {
  const double* u;
  const double* v;
  long long xdim;
  __device__ __forceinline__ double2 at(long long lrow, long long j) const {
    const long long k = lrow * xdim + j;
    return make_double2(__ldg(u + k), __ldg(v + k));
  }
  bool aligned16() const { return (((uintptr_t)u | (uintptr_t)v) & 15u) == 0; }
  struct Cursor {
    const double* qu;
    const double* qv;
    __device__ __forceinline__ double2 load(int off) const {
      return make_double2(__ldg(qu + off), __ldg(qv + off));
    }
  };
  __device__ __forceinline__ Cursor cursor(long long lrow, long long j) const {
    const long long k = lrow * xdim + j;
    return Cursor{u + k, v + k};
  }
};

// sqrt(x) for finite x >= 0: MUFU.RSQ64H seed (~2^-20), ONE coupled Newton (Goldschmidt) step
// (g ~ sqrt x and h ~ 1 / (2 sqrt x) to ~2^-39) and a final residual step g + (x - g^2) h, which
// is itself a Newton step and squares the error again (2^-78: the result is rounding only,
// <= 1 ulp).  Round 1 ran two Goldschmidt steps before the residual one; the second bought
// nothing (3 FP64 instructions of ~55 per FTLE point on an FP64-issue-bound kernel).  0 maps
// to 0.  No special-case branches.
__device__ __forceinline__ double fast_sqrt(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double g = x * r, h = 0.5 * r;
  const double e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  g = fma(fma(-g, g, x), h, g);
  // x == 0 (rsqrt = inf, g = NaN) -> 0.  Tested on the high word: an integer compare costs one
  // issue slot, an FP64 compare two (x >= 0 or NaN here; a NaN's high word passes and g is NaN).
  return (__double2hiint(x) >= 0x00100000) ? g : 0.0;
}

// Constants of fast_log_ge1 in constant memory (a 64-bit literal costs two UMOVs per use on
// sm_100a; the first ncu capture of this kernel showed 14 % of its instructions were UMOV).
__constant__ double ST_LOGK[19] = {
    2.0 / 19.0, 2.0 / 17.0, 2.0 / 15.0, 2.0 / 13.0, 2.0 / 11.0, 2.0 / 9.0, 2.0 / 7.0, 2.0 / 5.0,
    2.0 / 3.0,
    6.93147180369123816490e-01,  // ln2 high part (exact when multiplied by |e| < 2^11)
    1.90821492927058770002e-10,  // ln2 low part
    1.4142135623730951,          // sqrt(2)
    // ln1p(r) = r + r^2 (-1/2 + r/3 - r^2/4 + ... - r^6/8), Horner from the top  [12..18]
    -1.0 / 8.0, 1.0 / 7.0, -1.0 / 6.0, 1.0 / 5.0, -1.0 / 4.0, 1.0 / 3.0, -1.0 / 2.0,
};

// ln(x) for finite x >= 1 (the FTLE clamp guarantees it): x = 2^e m, m in [sqrt(1/2), sqrt 2),
// ln m = 2 atanh(s), s = (m - 1) / (m + 1), odd series in s to s^19 (|s| <= 0.1716).
__device__ __forceinline__ double fast_log_ge1(double x) {
  int hi = __double2hiint(x);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;
  double m = __hiloint2double(hi, __double2loint(x));  // [1, 2)
  const bool up = m > ST_LOGK[11];
  m = up ? 0.5 * m : m;
  e = up ? e + 1 : e;
  const double num = m - 1.0, den = m + 1.0;
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
  double t = fma(-den, r, 1.0);
  r = fma(r, t, r);
  t = fma(-den, r, 1.0);
  r = fma(r, t, r);
  double sq = num * r;
  sq = fma(fma(-sq, den, num), r, sq);  // s = num / den to ~0.5 ulp
  const double z = sq * sq;
  double p = ST_LOGK[0];
#pragma unroll
  for (int k = 1; k <= 8; ++k) p = fma(p, z, ST_LOGK[k]);
  const double ed = (double)e;
  // e ln2 + 2 s + s z p, ln2 split so that e * LN2_HI is exact
  const double hi_part = fma(ed, ST_LOGK[9], 2.0 * sq);
  const double lo_part = fma(ed, ST_LOGK[10], sq * z * p);
  return hi_part + lo_part;
}

// Table-driven ln(x), x >= 1 finite: x = 2^e m, m in [1, 2), c = 1 + i/128 with i the top seven
// mantissa bits; ln x = e ln2 + ln c + ln1p(r), r = m / c - 1 in [0, 2^-7).  The table holds
// (fl(1/c), -ln(fl(1/c))) so that the rounding of 1/c cancels; entry 0 is (1, 0), which keeps
// full relative accuracy for x -> 1.  13 FP64 instructions instead of 27 for the series
// version: the FTLE kernel is co-limited by the FP64 pipe (sqrt + log per point).
__device__ double2 g_logtab[128];

__device__ __forceinline__ double fast_log_ge1_tab(double x, const double2* tab) {
  const int hi = __double2hiint(x);
  const int e = (hi >> 20) - 1023;
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
  const double2 t = tab[(hi >> 13) & 0x7f];
  const double r = fma(m, t.x, -1.0);
  double q = ST_LOGK[12];
#pragma unroll
  for (int k = 13; k <= 18; ++k) q = fma(q, r, ST_LOGK[k]);  // -1/8 ... -1/2
  const double ed = (double)e;
  const double hi_part = fma(ed, ST_LOGK[9], t.y);
  const double lo_part = fma(ed, ST_LOGK[10], fma(r * r, q, r));
  return hi_part + lo_part;
}

// Per-point epilogue shared by the interior and the edge kernels.
// MODE: for KIND_FTLE unused; for KIND_EULER one of dlb_euler_mode.
template <int KIND, int MODE, bool XI>
__device__ __forceinline__ void emit(const StencilParams& P, double2 C, double d0dx, double d1dx,
                                     double d0dy, double d1dy, long long o,
                                     const double2* logtab = nullptr) {
  if (KIND == KIND_FTLE) {
    // JF = [[dfxdx, dfxdy], [dfydx, dfydy]], C = JF^T JF   (lagrangian.py:65-66)
    const double c11 = fma(d0dx, d0dx, d1dx * d1dx);
    const double c12 = fma(d0dx, d0dy, d1dx * d1dy);
    const double c22 = fma(d0dy, d0dy, d1dy * d1dy);
    double lam, vx = 0.0, vy = 0.0;
    if (XI) {
      eig_sym_lapack(c11, c12, c22, true, &lam, &vx, &vy);
      if (P.xi_mode == DLB_XI_FORCED) force_eigenvector(&vx, &vy);
    } else {
      const double hm = 0.5 * (c11 - c22);
      lam = 0.5 * (c11 + c22) + fast_sqrt(fma(hm, hm, c12 * c12));
    }
    // lagrangian.py:70-73; huge / non-finite lambda (blown-up trajectories): library log
    double out = 0.0;
    if (lam >= 1.0)
      out = P.scale * ((lam < 1.0e300)
                           ? (logtab ? fast_log_ge1_tab(lam, logtab) : fast_log_ge1(lam))
                           : log(lam));
    else if (lam != lam)
      out = lam;  // NaN propagates like np.log
    store_ftle<KIND, MODE>(P, o, out);
    if (XI) P.xi[o] = make_double2(vx, vy);
  } else {
    // S = 0.5 (G + G^T), G = [[dudx, dudy], [dvdx, dvdy]]   (eulerian.py:52-53)
    const double a = d0dx, d = d1dy, b = 0.5 * (d0dy + d1dx);
    double out, vx = 0.0, vy = 0.0;
    bool masked = false;
    if (MODE == DLB_EULER_S_MIN || MODE == DLB_EULER_S_MAX) {
      const bool want_max = (MODE == DLB_EULER_S_MAX);
      if (XI) {
        eig_sym_lapack(a, b, d, want_max, &out, &vx, &vy);
        if (P.xi_mode == DLB_XI_FORCED) force_eigenvector(&vx, &vy);
      } else {
        const double hm = 0.5 * (a - d);
        const double rad = fast_sqrt(fma(hm, hm, b * b));
        out = want_max ? 0.5 * (a + d) + rad : 0.5 * (a + d) - rad;
      }
      if (P.negate) out = -out;
    } else {
      const double uu = C.x, vv = C.y;
      const double v2 = fma(uu, uu, vv * vv);  // eulerian.py:187
      double num;
      if (MODE == DLB_EULER_RATE)  // V . (J^T S J) V, eulerian.py:160-165
        num = fma(a * vv, vv, fma(-2.0 * b * uu, vv, d * uu * uu));
      else  // V . (tr(S) I - 2 S) V, eulerian.py:167-171
        num = fma((a - d) * vv, vv, fma(-4.0 * b * uu, vv, (d - a) * uu * uu));
      masked = (v2 == 0.0);  // eulerian.py:188-192
      out = masked ? 0.0 : num / v2;
    }
    P.field[o] = out;
    if (XI) P.xi[o] = make_double2(vx, vy);
    if (P.mask) P.mask[o] = masked ? 1 : 0;
  }
}

// Same epilogue for the streaming interior kernels: everything is computed unconditionally and
// only the stores are predicated, so a fully unrolled group of rows stays ONE basic block and
// the scheduler can interleave the rows' dependent FP64 chains (with `if (lane emits) emit(...)`
// every row was its own BSSY/BSYNC region; ncu: fixed-latency "wait" was the top stall).  The
// eigenvector and rate / ratio variants keep their data-dependent branches (deflation test,
// IEEE division) and go through emit().
template <int KIND, int MODE, bool XI>
__device__ __forceinline__ void emit_pred(const StencilParams& P, double2 C, double d0dx,
                                          double d1dx, double d0dy, double d1dy, long long o,
                                          const double2* logtab, bool store) {
  if constexpr (KIND == KIND_FTLE && !XI) {
    const double c11 = fma(d0dx, d0dx, d1dx * d1dx);
    const double c12 = fma(d0dx, d0dy, d1dx * d1dy);
    const double c22 = fma(d0dy, d0dy, d1dy * d1dy);
    const double hm = 0.5 * (c11 - c22);
    const double lam = 0.5 * (c11 + c22) + fast_sqrt(fma(hm, hm, c12 * c12));
    const double fl = fast_log_ge1_tab(lam, logtab);  // meaningless below 1, selected away
    // lam >= 0, +inf or NaN.  Classified on the high word (integer compares: one issue slot
    // each instead of two per FP64 compare): [0x3ff00000, 0x7ff00000) finite >= 1 -> scale *
    // ln(lam); >= 0x7ff00000 (unsigned: +inf, NaN of either sign) -> scale * lam = inf / NaN like
    // np.log; below -> 0   (lagrangian.py:70-73)
    const unsigned lh = (unsigned)__double2hiint(lam);
    const double picked = (lh >= 0x7ff00000u) ? lam : fl;
    const double out = (lh >= 0x3ff00000u) ? P.scale * picked : 0.0;
    if constexpr (MODE == FTLE_PUSH || !DLB_ASM_STORE) {
      if (store) store_ftle<KIND, MODE>(P, o, out);  // + peer stores (the last rows of a slab only)
    } else {
      st_f64_if(P.field + o, out, store);
    }
  } else if constexpr (KIND == KIND_EULER && !XI &&
                       (MODE == DLB_EULER_S_MIN || MODE == DLB_EULER_S_MAX)) {
    const double a = d0dx, d = d1dy, b = 0.5 * (d0dy + d1dx);
    const double hm = 0.5 * (a - d);
    const double rad = fast_sqrt(fma(hm, hm, b * b));
    double out = (MODE == DLB_EULER_S_MAX) ? 0.5 * (a + d) + rad : 0.5 * (a + d) - rad;
    if (P.negate) out = -out;
#if DLB_ASM_STORE
    st_f64_if(P.field + o, out, store);
    st_u8_if(P.mask + o, 0, store && P.mask != nullptr);
#else
    if (store) {
      P.field[o] = out;
      if (P.mask) P.mask[o] = 0;
    }
#endif
  } else {
    if (store) emit<KIND, MODE, XI>(P, C, d0dx, d1dx, d0dy, d1dy, o, logtab);
  }
}

// Edge kernel: the O(xdim + ydim) output nodes on the grid boundary (global rows 0 and
// ydim-1 in full, columns 0 and xdim-1 of the other rows) with numpy's one-sided formulas.
template <int KIND, int MODE, bool XI, class Loader>
__global__ void __launch_bounds__(ST_BLOCK) stencil_edge_kernel(const StencilParams P,
                                                                const Loader ld) {
  const long long xdim = P.ax.n, ydim = P.ay.n;
  const long long r0 = P.out_row0, r1 = P.out_row0 + P.out_rows;
  const bool has_top = (r0 == 0), has_bot = (r1 == ydim) && (ydim > 1 || !has_top);
  const long long n_full = (has_top ? 1 : 0) + (has_bot ? 1 : 0);
  const long long n_side_rows = P.out_rows - n_full;
  const long long side_cols = (xdim > 1) ? 2 : 1;
  const long long total = n_full * xdim + n_side_rows * side_cols;
  for (long long k = (long long)blockIdx.x * ST_BLOCK + threadIdx.x; k < total;
       k += (long long)gridDim.x * ST_BLOCK) {
    long long gi, j;
    if (k < n_full * xdim) {
      const long long which = k / xdim;
      j = k - which * xdim;
      gi = (which == 0 && has_top) ? 0 : ydim - 1;
    } else {
      const long long kk = k - n_full * xdim;
      const long long rr = kk / side_cols;
      j = (kk - rr * side_cols == 0) ? 0 : xdim - 1;
      gi = r0 + (has_top ? 1 : 0) + rr;
    }
    const long long lrow = gi - P.grow0;
    const double2 C = ld.at(lrow, j);
    const double2 L = ld.at(lrow, (j > 0) ? j - 1 : 0);
    const double2 R = ld.at(lrow, (j < xdim - 1) ? j + 1 : xdim - 1);
    const double2 U = ld.at(((gi > 0) ? gi - 1 : 0) - P.grow0, j);
    const double2 D = ld.at(((gi < ydim - 1) ? gi + 1 : ydim - 1) - P.grow0, j);
    double d0dx, d1dx, d0dy, d1dy;
    edge_axis_x(P, ld, lrow, j, C, L, R, &d0dx, &d1dx);
    edge_axis_y(P, ld, gi, j, C, U, D, &d0dy, &d1dy);
    emit<KIND, MODE, XI>(P, C, d0dx, d1dx, d0dy, d1dy, (gi - P.out_row0) * xdim + j);
  }
}

}  // namespace

// ---- host helpers (stencil.cu) ----
namespace dlbst {
int env_flag(const char* name, int dflt);
int num_sms();
// Validate geometry, build + upload both axes into the workspace (cached), fill P.ax / P.ay.
int setup_axes(StencilParams& P, const double* h_x, int64_t xdim, const double* h_y, int64_t ydim,
               int edge_order, void* d_workspace, size_t workspace_bytes, cudaStream_t stream);
int check_slab(const StencilParams& P, int64_t ydim);
}  // namespace dlbst
