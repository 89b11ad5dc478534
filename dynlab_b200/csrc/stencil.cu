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
// B200 mapping (HBM-bound, 24 B per point).  Interior kernel: one warp walks 32 consecutive
// columns down a chunk of rows keeping the rows i-1, i, i+1 of its column in registers (next
// row prefetched), so every element is requested once as a coalesced 256/512 B row segment;
// the x neighbours come from the adjacent lanes by warp shuffle (30 output columns per warp).
// The O(xdim + ydim) boundary nodes — numpy's one-sided formulas — run in a separate small
// edge kernel, which keeps the interior loop free of edge logic (64 registers, no spills).
// Coefficients are precomputed per row / per column on the host with numpy's own formulas
// (true divisions), so the kernels have no FP64 divides on the interior path; sqrt and log
// are branch-free MUFU-seeded Newton / series versions.  Grids are multiples of the SM count
// and warps grid-stride over (strip, row-chunk) tasks.
#include <math.h>
#include <math_constants.h>

#include <mutex>
#include <type_traits>
#include <vector>

#include "common.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int ST_BLOCK = 256;
constexpr int ST_ROWS = 32;  // rows per warp task

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
};

enum { KIND_FTLE = 0, KIND_EULER = 1 };

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

// sqrt(x) for finite x >= 0: MUFU.RSQ64H seed + two coupled Newton (Goldschmidt) steps and a
// final residual correction; 0 maps to 0.  No special-case branches (the library routine's
// slow-path test is what the HBM-bound kernels cannot afford).
__device__ __forceinline__ double fast_sqrt(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double g = x * r, h = 0.5 * r;
  double e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  g = fma(fma(-g, g, x), h, g);
  return (x > 0.0) ? g : 0.0;
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

int env_flag(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

int g_sms = 0;
int num_sms() {
  if (g_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      g_sms = 148;
  }
  return g_sms;
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
    P.field[o] = out;
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

// (fl(1/c), -ln(fl(1/c))) for c = 1 + i/128, uploaded once per process and device.
int ensure_logtab() {
  static std::mutex mu;
  static int done_dev = -1;
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  DLB_CUDA_CHECK(cudaGetDevice(&dev));
  if (done_dev == dev) return DLB_OK;
  double2 h[128];
  for (int i = 0; i < 128; ++i) {
    const double c = 1.0 + i / 128.0;
    const double inv = (i == 0) ? 1.0 : 1.0 / c;
    h[i].x = inv;
    h[i].y = (i == 0) ? 0.0 : (double)(-logl((long double)inv));
  }
  DLB_CUDA_CHECK(cudaMemcpyToSymbol(g_logtab, h, sizeof(h)));
  done_dev = dev;
  return DLB_OK;
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

// Last (workspace, x, y, edge_order) whose coefficients are resident in the workspace.
struct AxesCache {
  std::mutex mu;
  bool valid = false;
  const void* ws = nullptr;
  int edge_order = 0;
  std::vector<double> x, y;
  HostAxis hx, hy;
  cudaStream_t stream = nullptr;
};
AxesCache& axes_cache() {
  static AxesCache c;
  return c;
}

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
  AxesCache& cache = axes_cache();
  std::lock_guard<std::mutex> lock(cache.mu);
  const bool hit = cache.valid && cache.ws == d_workspace && cache.edge_order == edge_order &&
                   (int64_t)cache.x.size() == xdim && (int64_t)cache.y.size() == ydim &&
                   memcmp(cache.x.data(), h_x, sizeof(double) * xdim) == 0 &&
                   memcmp(cache.y.data(), h_y, sizeof(double) * ydim) == 0;
  if (!hit) {
    cache.valid = false;
    build_axis(h_x, xdim, edge_order, &cache.hx);
    build_axis(h_y, ydim, edge_order, &cache.hy);
    const HostAxis &hx = cache.hx, &hy = cache.hy;
    std::vector<double> stage(3 * (xdim + ydim) + 4 * ydim);
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
    // pageable source: cudaMemcpyAsync stages it before returning, so `stage` may die here
    DLB_CUDA_CHECK(cudaMemcpyAsync(base, stage.data(), sizeof(double) * stage.size(),
                                   cudaMemcpyHostToDevice, stream));
    cache.ws = d_workspace;
    cache.edge_order = edge_order;
    cache.x.assign(h_x, h_x + xdim);
    cache.y.assign(h_y, h_y + ydim);
    cache.stream = stream;
    cache.valid = true;
  } else if (cache.stream != stream) {
    // uploaded on another stream: order this stream after it
    DLB_CUDA_CHECK(cudaStreamSynchronize(cache.stream));
    cache.stream = stream;
  }
  const HostAxis &hx = cache.hx, &hy = cache.hy;
  P.ax = Axis{dxa, dxa + xdim, dxa + 2 * xdim, xdim, hx.inv2h, hx.w_first, hx.w_last, hx.uniform};
  P.ay = Axis{dya, dya + ydim, dya + 2 * ydim, ydim, hy.inv2h, hy.w_first, hy.w_last, hy.uniform};
  P.ay4 = (const double2*)(base + 3 * (xdim + ydim));  // 32-byte aligned: 256 + 32 (xdim+ydim)
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

}  // namespace

void dlb_axes_cache_note_layout(const void* d_workspace, int64_t xdim, int64_t ydim) {
  AxesCache& c = axes_cache();
  std::lock_guard<std::mutex> lock(c.mu);
  if (c.valid && c.ws == d_workspace &&
      ((int64_t)c.x.size() != xdim || (int64_t)c.y.size() < ydim))
    c.valid = false;
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
