// Device twins of dynlab.flows (R:src/dynlab/flows.py:13-331).
//
// Each flow is split into
//   flow_prepare(id, params, consts)  — host: parameter-only sub-expressions, evaluated in the
//                                       reference's own operation order (separately rounded);
//   flow_rhs<ID>(consts, t, Y, out)   — device: f(t, Y), inlined into the calling kernel.
// The operation order of the reference's Python expressions is kept so that results differ
// from numpy only through libm-vs-CUDA transcendental rounding and FMA contraction.
#pragma once
#include "common.cuh"
#include "dmath.cuh"

#define DLB_PI 3.141592653589793  // np.pi

struct FlowInfo {
  const char* name;
  int dim;
  int nparams;
};

static const FlowInfo k_flow_info[DLB_FLOW_COUNT] = {
    {"double_gyre", 2, 3},
    {"autonomous_double_gyre", 2, 0},
    {"bickley_jet", 2, 10},
    {"glider", 2, 0},
    {"hills_vortex", 2, 0},
    {"sigmoidal", 2, 0},
    {"autonomous_duffing_oscillator", 2, 0},
    {"duffing_oscillator", 2, 3},
    {"pendulum", 2, 3},
    {"hurricane", 2, 3},
    {"van_der_pol_oscillator", 2, 4},
    {"bead_on_a_rotating_hoop", 2, 2},
    {"lotka_volterra", 2, 4},
    {"abc", 3, 4},
    {"lorenz", 3, 3},
};

// Host-side parameter preparation.  Returns false on a bad id / parameter count.
static inline bool flow_prepare(int id, const double* p, int n_params, FlowConsts* fc) {
  if (id < 0 || id >= DLB_FLOW_COUNT) return false;
  if (n_params != k_flow_info[id].nparams) return false;
  if (n_params > 0 && p == nullptr) return false;
  memset(fc, 0, sizeof(*fc));
  double* c = fc->c;
  switch (id) {
    case DLB_FLOW_DOUBLE_GYRE: {  // A, w, e  (flows.py:27-33)
      const double A = p[0], w = p[1], e = p[2];
      c[0] = w;
      c[1] = e;
      c[2] = 2 * e;            // `2*e*np.sin(w*t)` associates as (2*e)*sin
      c[3] = -DLB_PI * A;      // `-np.pi*A*...`
      c[4] = DLB_PI * A;       // `np.pi*A*...`
    } break;
    case DLB_FLOW_BICKLEY_JET: {  // flows.py:84-96
      const double scale = p[9];
      const double U0 = p[0] / scale, L = p[1] / scale, re = p[2] / scale;
      const double A2 = p[3], A3 = p[4];
      const double c2 = p[5] * U0, c3 = p[6] * U0, k2 = p[7] / re, k3 = p[8] / re;
      c[0] = U0;
      c[1] = L;
      c[2] = c2;
      c[3] = c3;
      c[4] = k2;
      c[5] = k3;
      c[6] = 2 * A3 * U0;        // u: 2*A3*U0*tanh*sech2*cos(...)
      c[7] = 2 * A2 * U0;
      c[8] = -k3 * A3 * L * U0;  // v: -k3*A3*L*U0*sech2*sin(...)
      c[9] = k2 * A2 * L * U0;
    } break;
    case DLB_FLOW_HURRICANE: {  // eye_alpha, eye_omega, eye_amplitude (flows.py:216-221)
      const double sq = sqrt(1.0 + 4.0 * p[0]);
      c[0] = p[0];
      c[1] = p[1];
      c[2] = p[2];
      c[3] = 1.0 / sq;           // beta
      c[4] = 0.5 * (1.0 + sq);   // y offset
    } break;
    case DLB_FLOW_BEAD:  // epsilon, gamma (flows.py:260-262)
      c[0] = 1 / p[0];
      c[1] = p[1];
      break;
    default:
      for (int i = 0; i < n_params; ++i) c[i] = p[i];
  }
  return true;
}

// ---- time / space split --------------------------------------------------------------------
// f(t, Y) = flow_space(flow_time(t), Y).  Within one Runge-Kutta attempt the six stage times
// are known up front (t + c_s h), so the integrator evaluates the time parts of all stages in
// one interleaved block (independent polynomial chains, constants loaded once) and reuses the
// value at t + h for the two stages that share it (c_6 = c_7 = 1).  tc[] holds FlowTime<ID>::N
// doubles: the separable forcing term where there is one, otherwise t itself.
template <int ID>
struct FlowTime {
  static constexpr int N = 1;  // default: tc[0] = t
};
#define DLB_AUTONOMOUS(id)          \
  template <>                       \
  struct FlowTime<id> {             \
    static constexpr int N = 0;     \
  };
DLB_AUTONOMOUS(DLB_FLOW_AUTONOMOUS_DOUBLE_GYRE)
DLB_AUTONOMOUS(DLB_FLOW_GLIDER)
DLB_AUTONOMOUS(DLB_FLOW_HILLS_VORTEX)
DLB_AUTONOMOUS(DLB_FLOW_AUTONOMOUS_DUFFING)
DLB_AUTONOMOUS(DLB_FLOW_BEAD)
DLB_AUTONOMOUS(DLB_FLOW_LOTKA_VOLTERRA)
DLB_AUTONOMOUS(DLB_FLOW_LORENZ)
#undef DLB_AUTONOMOUS

// Time part for NS stage times at once (NS = 1 or 5).  tc is [NS][max(N, 1)].
template <int ID, int NS, class Trig>
__device__ __forceinline__ void flow_time(const FlowConsts& fc, Trig& trig, const double* t,
                                          double (*tc)[FlowTime<ID>::N > 0 ? FlowTime<ID>::N : 1]) {
  const double* c = fc.c;
  if constexpr (ID == DLB_FLOW_DOUBLE_GYRE || ID == DLB_FLOW_DUFFING || ID == DLB_FLOW_PENDULUM ||
                ID == DLB_FLOW_VAN_DER_POL || ID == DLB_FLOW_HURRICANE || ID == DLB_FLOW_ABC) {
    // one sin (cos for hurricane) of (frequency * t) per stage
    constexpr bool is_cos = (ID == DLB_FLOW_HURRICANE);
    const double w = (ID == DLB_FLOW_DOUBLE_GYRE) ? c[0]
                     : (ID == DLB_FLOW_DUFFING || ID == DLB_FLOW_PENDULUM) ? c[2]
                     : (ID == DLB_FLOW_VAN_DER_POL) ? c[3]
                     : (ID == DLB_FLOW_HURRICANE) ? c[1]
                                                  : DLB_PI;
    double arg[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) arg[s] = w * t[s];
    trig.template sinN<NS, is_cos>(arg, &tc[0][0], 1);
  } else if constexpr (FlowTime<ID>::N == 1) {
#pragma unroll
    for (int s = 0; s < NS; ++s) tc[s][0] = t[s];
  }
}

// ---- x / y separation (grid evaluation) --------------------------------------------------------
// On meshgrid(x, y) at a fixed time some flows factor into functions of x only and of y only
// (double gyre: sin(pi f(x)) * cos(pi y); Bickley jet: sech^2(y/L) * cos(k (x - c t))).  The
// fused flow + stencil kernels (eulerflow.cu) evaluate the x part once per column and the y part
// once per row of a tile and only *combine* them per node - the transcendentals leave the
// per-node cost.  FlowXY<ID>::NX / NY doubles per column / row; NX == 0: not separable, the
// node is evaluated by flow_space.  combine() performs exactly the operations flow_space does
// on the same intermediate values, so both give bit-identical velocities.
template <int ID>
struct FlowXY {
  static constexpr int NX = 0, NY = 0;
};
template <>
struct FlowXY<DLB_FLOW_DOUBLE_GYRE> {
  static constexpr int NX = 3, NY = 2;  // (c3 sin(pi f), c4 cos(pi f), dfdx) | (sin, cos)(pi y)
};
template <>
struct FlowXY<DLB_FLOW_AUTONOMOUS_DOUBLE_GYRE> {
  static constexpr int NX = 2, NY = 2;  // (-pi sin(pi x), pi cos(pi x)) | (sin, cos)(pi y)
};
template <>
struct FlowXY<DLB_FLOW_BICKLEY_JET> {
  static constexpr int NX = 4, NY = 5;  // (s3, c3, s2, c2) | sech^2, tanh products
};

template <int ID, class Trig>
__device__ __forceinline__ void flow_xpart(const FlowConsts& fc, Trig& trig, const double* tc,
                                           double x, double* xp) {
  const double* c = fc.c;
  if constexpr (ID == DLB_FLOW_DOUBLE_GYRE) {
    const double s = tc[0];  // sin(w t)
    const double a = c[1] * s;
    const double b = fma(-c[2], s, 1.0);
    const double f = x * fma(a, x, b);
    double sf, cf;
    trig.sincos(DLB_PI * f, &sf, &cf);
    xp[0] = c[3] * sf;
    xp[1] = c[4] * cf;
    xp[2] = fma(2 * a, x, b);
  } else if constexpr (ID == DLB_FLOW_AUTONOMOUS_DOUBLE_GYRE) {
    double sx, cx;
    trig.sincos(DLB_PI * x, &sx, &cx);
    xp[0] = -DLB_PI * sx;
    xp[1] = DLB_PI * cx;
  } else if constexpr (ID == DLB_FLOW_BICKLEY_JET) {
    const double t = tc[0];
    trig.sincos(c[5] * fma(-c[3], t, x), &xp[0], &xp[1]);
    trig.sincos(c[4] * fma(-c[2], t, x), &xp[2], &xp[3]);
  }
}

template <int ID, class Trig>
__device__ __forceinline__ void flow_ypart(const FlowConsts& fc, Trig& trig, const double* tc,
                                           double y, double* yp) {
  const double* c = fc.c;
  if constexpr (ID == DLB_FLOW_DOUBLE_GYRE || ID == DLB_FLOW_AUTONOMOUS_DOUBLE_GYRE) {
    trig.sincos(y * DLB_PI, &yp[0], &yp[1]);
  } else if constexpr (ID == DLB_FLOW_BICKLEY_JET) {
    const double yl = y / c[1];
    const double ch = 1 / cosh(yl);
    const double sech2 = ch * ch, th = tanh(yl);
    yp[0] = c[0] * sech2;
    yp[1] = c[6] * th * sech2;
    yp[2] = c[7] * th * sech2;
    yp[3] = c[8] * sech2;
    yp[4] = c[9] * sech2;
  }
}

template <int ID>
__device__ __forceinline__ void flow_combine(const double* xp, const double* yp, double* out) {
  if constexpr (ID == DLB_FLOW_DOUBLE_GYRE) {
    out[0] = xp[0] * yp[1];            // c3 sf cy
    out[1] = xp[1] * yp[0] * xp[2];    // c4 cf sy dfdx
  } else if constexpr (ID == DLB_FLOW_AUTONOMOUS_DOUBLE_GYRE) {
    out[0] = xp[0] * yp[1];
    out[1] = xp[1] * yp[0];
  } else if constexpr (ID == DLB_FLOW_BICKLEY_JET) {
    // U0 sech2 + 2 A3 U0 tanh sech2 cos(k3 ..) + 2 A2 U0 tanh sech2 cos(k2 ..)   (flows.py:94)
    out[0] = fma(yp[2], xp[3], fma(yp[1], xp[1], yp[0]));
    // -k3 A3 L U0 sech2 sin(k3 ..) - k2 A2 L U0 sech2 sin(k2 ..)                 (flows.py:95)
    out[1] = fma(-yp[4], xp[2], yp[3] * xp[0]);
  }
}

// Space part.  Y/out have k_flow_info[ID].dim entries.  Every multiply-add is written as an
// explicit fma(): nothing is left to the compiler's contraction heuristics, so each inlined copy
// (K0, K1, the fused Eulerian kernels, loop prologue vs. loop body) rounds identically.
template <int ID, class Trig>
__device__ __forceinline__ void flow_space(const FlowConsts& fc, Trig& trig, const double* tc,
                                           const double* Y, double* out) {
  const double* c = fc.c;
  const double x = Y[0], y = Y[1];
  if constexpr (ID == DLB_FLOW_DOUBLE_GYRE) {
    const double s = tc[0];  // sin(w t)
    const double a = c[1] * s;
    const double b = fma(-c[2], s, 1.0);
    const double f = x * fma(a, x, b);       // a x^2 + b x
    const double dfdx = fma(2 * a, x, b);    // 2 a x + b
    double sf, cf, sy, cy;
    trig.sincospi2(f, y, &sf, &cf, &sy, &cy);  // sin / cos of pi f and of y pi
    out[0] = c[3] * sf * cy;
    out[1] = c[4] * cf * sy * dfdx;
  } else if constexpr (ID == DLB_FLOW_AUTONOMOUS_DOUBLE_GYRE) {
    double xp[2], yp[2];
    flow_xpart<ID>(fc, trig, tc, x, xp);
    flow_ypart<ID>(fc, trig, tc, y, yp);
    flow_combine<ID>(xp, yp, out);
  } else if constexpr (ID == DLB_FLOW_BICKLEY_JET) {
    double xp[4], yp[5];
    flow_xpart<ID>(fc, trig, tc, x, xp);
    flow_ypart<ID>(fc, trig, tc, y, yp);
    flow_combine<ID>(xp, yp, out);
  } else if constexpr (ID == DLB_FLOW_GLIDER) {
    const double r = sqrt(fma(x, x, y * y));
    const double th2 = -2 * atan2(y, x);
    double s, cs;
    trig.sincos(th2, &s, &cs);
    s = 1.2 * s;
    cs = 1.4 - cs;
    out[0] = r * -fma(y, s, cs * x);
    out[1] = fma(r, fma(x, s, -(cs * y)), -1.0);
  } else if constexpr (ID == DLB_FLOW_HILLS_VORTEX) {
    out[0] = 2.0 * x * y;
    out[1] = fma(-2.0 * y, y, -2.0 * fma(2.0 * x, x, -1.0));
  } else if constexpr (ID == DLB_FLOW_SIGMOIDAL) {
    out[0] = 0.4 * tanh(10.0 * fma(-0.5, tc[0], x));
    out[1] = y * 0;
  } else if constexpr (ID == DLB_FLOW_AUTONOMOUS_DUFFING) {
    out[0] = y;
    out[1] = fma(-(x * x), x, x);
  } else if constexpr (ID == DLB_FLOW_DUFFING) {  // A, gamma, omega; tc = sin(omega t)
    out[0] = y;
    out[1] = fma(c[0], tc[0], fma(-c[1], y, fma(-(x * x), x, x)));
  } else if constexpr (ID == DLB_FLOW_PENDULUM) {  // A, gamma, omega; tc = sin(omega t)
    out[0] = y;
    out[1] = fma(c[0], tc[0], fma(-c[1], y, -trig.sin(x)));
  } else if constexpr (ID == DLB_FLOW_HURRICANE) {  // tc = cos(eye_omega t)
    const double yt = y - c[4];
    const double den = fma(yt, yt, x * x) + c[0];
    out[0] = -yt / den - c[3];
    out[1] = fma(c[2] * y, tc[0], x / den);
  } else if constexpr (ID == DLB_FLOW_VAN_DER_POL) {  // A, mu, gamma, omega; tc = sin(omega t)
    out[0] = y;
    out[1] = fma(c[0], tc[0], fma(-c[2], y, fma(c[1] * fma(-x, x, 1.0), y, -x)));
  } else if constexpr (ID == DLB_FLOW_BEAD) {  // 1/epsilon, gamma
    double s, cs;
    trig.sincos(x, &s, &cs);
    out[0] = y;
    out[1] = c[0] * fma(s, fma(c[1], cs, -1.0), -y);
  } else if constexpr (ID == DLB_FLOW_LOTKA_VOLTERRA) {  // alpha, beta, gamma, delta
    out[0] = fma(-(c[1] * x), y, c[0] * x);
    out[1] = fma(c[3] * x, y, -(c[2] * y));
  } else if constexpr (ID == DLB_FLOW_ABC) {  // ABC_Amplitude, A, B, C; tc = sin(pi t)
    const double z = Y[2];
    const double At = fma(c[0], tc[0], c[1]);
    double sx, cx, sy, cy, sz, cz;
    trig.sincos2(x, y, &sx, &cx, &sy, &cy);
    trig.sincos(z, &sz, &cz);
    out[0] = fma(At, sz, c[3] * cy);
    out[1] = fma(c[2], sx, At * cz);
    out[2] = fma(c[3], sy, c[2] * cx);
  } else if constexpr (ID == DLB_FLOW_LORENZ) {  // sigma, rho, beta
    const double z = Y[2];
    out[0] = c[0] * (y - x);
    out[1] = fma(x, c[1] - z, -y);
    out[2] = fma(x, y, -(c[2] * z));
  }
}

// f(t, Y) in one call (K0 and the integrator's start-up evaluations).
template <int ID>
__device__ __forceinline__ void flow_rhs(const FlowConsts& fc, double t, const double* Y,
                                         double* out) {
  dm::TrigSafe trig;
  double tc[1][FlowTime<ID>::N > 0 ? FlowTime<ID>::N : 1];
  flow_time<ID, 1>(fc, trig, &t, tc);
  flow_space<ID>(fc, trig, tc[0], Y, out);
}

// Run the statement block with `constexpr int ID` bound to the runtime flow id.
#define DLB_DISPATCH_FLOW(id, ...)                                           \
  switch (id) {                                                              \
    case 0: { constexpr int ID = 0; __VA_ARGS__; } break;                           \
    case 1: { constexpr int ID = 1; __VA_ARGS__; } break;                           \
    case 2: { constexpr int ID = 2; __VA_ARGS__; } break;                           \
    case 3: { constexpr int ID = 3; __VA_ARGS__; } break;                           \
    case 4: { constexpr int ID = 4; __VA_ARGS__; } break;                           \
    case 5: { constexpr int ID = 5; __VA_ARGS__; } break;                           \
    case 6: { constexpr int ID = 6; __VA_ARGS__; } break;                           \
    case 7: { constexpr int ID = 7; __VA_ARGS__; } break;                           \
    case 8: { constexpr int ID = 8; __VA_ARGS__; } break;                           \
    case 9: { constexpr int ID = 9; __VA_ARGS__; } break;                           \
    case 10: { constexpr int ID = 10; __VA_ARGS__; } break;                         \
    case 11: { constexpr int ID = 11; __VA_ARGS__; } break;                         \
    case 12: { constexpr int ID = 12; __VA_ARGS__; } break;                         \
    case 13: { constexpr int ID = 13; __VA_ARGS__; } break;                         \
    case 14: { constexpr int ID = 14; __VA_ARGS__; } break;                         \
    default: break;                                                          \
  }
