/* TEST INFRASTRUCTURE — plain-C restatement of dynlab's Lagrangian hot path.
 *
 * This is the CPU oracle / CPU baseline.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (dynlab_b200) never links or calls anything in here.
 *
 * What is restated (R: = /root/reference/, SP: = site-packages/):
 *   flows            R:src/dynlab/flows.py:13-331       (scalar evaluation)
 *   seed loop        R:src/dynlab/diagnostics/_base_classes.py:175-189
 *   RK45             SP:scipy/integrate/_ivp/rk.py:14-71,111-176,538-552
 *                    SP:scipy/integrate/_ivp/common.py:44-65,68-134
 *                    SP:scipy/integrate/_ivp/base.py:164,179-210
 *   np.gradient      SP:numpy/lib/_function_base_impl.py:1246-1384
 *   FTLE eigen       R:src/dynlab/diagnostics/lagrangian.py:60-76 (closed form for
 *                    the symmetric 2x2 eigenvalue; agrees with dgeev to ~5e-15)
 *
 * The integrator arithmetic lives in scipy (unpinned in R:pyproject.toml:23,
 * 1.18.1 in this image): this file restates its published algorithm.  It is
 * pinned by tests/test_oracle_golden.py against fixtures produced by the
 * unmodified reference driving scipy's RK45 (tests/golden/), including equal
 * per-seed nfev counts.
 *
 * Build: see oracle/c/Makefile (gcc -O2 -ffp-contract=off -pthread).
 * -ffp-contract=off keeps every multiply and add separately rounded, like the
 * Python/numpy scalar arithmetic of the reference.
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#define ORC_MAXDIM 3
#define ORC_MAXPAR 16

enum {
  FLOW_DOUBLE_GYRE = 0, FLOW_AUTONOMOUS_DOUBLE_GYRE, FLOW_BICKLEY_JET, FLOW_GLIDER,
  FLOW_HILLS_VORTEX, FLOW_SIGMOIDAL, FLOW_AUTONOMOUS_DUFFING, FLOW_DUFFING, FLOW_PENDULUM,
  FLOW_HURRICANE, FLOW_VAN_DER_POL, FLOW_BEAD, FLOW_LOTKA_VOLTERRA, FLOW_ABC, FLOW_LORENZ,
  FLOW_COUNT
};

static const int k_flow_dim[FLOW_COUNT] = {2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3};
static const int k_flow_npar[FLOW_COUNT] = {3, 0, 10, 0, 0, 0, 0, 3, 3, 3, 4, 2, 4, 4, 3};

int orc_flow_dim(int id) { return (id >= 0 && id < FLOW_COUNT) ? k_flow_dim[id] : -1; }
int orc_flow_npar(int id) { return (id >= 0 && id < FLOW_COUNT) ? k_flow_npar[id] : -1; }

static const double PI = 3.141592653589793; /* np.pi */

/* f(t, Y) -> out.  p = the Python signature's keyword parameters in order. */
static void flow_rhs(int id, const double* p, double t, const double* Y, double* out) {
  const double x = Y[0], y = Y[1];
  switch (id) {
    case FLOW_DOUBLE_GYRE: { /* R:flows.py:27-33  p = A, w, e */
      const double A = p[0], w = p[1], e = p[2];
      const double s = sin(w * t);
      const double a = e * s;
      const double b = 1 - 2 * e * s;
      const double f = a * (x * x) + b * x;
      const double dfdx = 2 * a * x + b;
      out[0] = -PI * A * sin(PI * f) * cos(y * PI);
      out[1] = PI * A * cos(PI * f) * sin(y * PI) * dfdx;
    } break;
    case FLOW_AUTONOMOUS_DOUBLE_GYRE: /* R:flows.py:46-48 */
      out[0] = -PI * sin(PI * x) * cos(y * PI);
      out[1] = PI * cos(PI * x) * sin(y * PI);
      break;
    case FLOW_BICKLEY_JET: { /* R:flows.py:84-96  p = U0,L,re,A2,A3,c2,c3,k2,k3,scale */
      const double scale = p[9];
      const double U0 = p[0] / scale, L = p[1] / scale, re = p[2] / scale;
      const double A2 = p[3], A3 = p[4];
      const double c2 = p[5] * U0, c3 = p[6] * U0, k2 = p[7] / re, k3 = p[8] / re;
      const double ch = 1 / cosh(y / L);
      const double sech2 = ch * ch, th = tanh(y / L);
      out[0] = U0 * sech2 + 2 * A3 * U0 * th * sech2 * cos(k3 * (x - c3 * t)) +
               2 * A2 * U0 * th * sech2 * cos(k2 * (x - c2 * t));
      out[1] = -k3 * A3 * L * U0 * sech2 * sin(k3 * (x - c3 * t)) -
               k2 * A2 * L * U0 * sech2 * sin(k2 * (x - c2 * t));
    } break;
    case FLOW_GLIDER: { /* R:flows.py:109-113 */
      const double r = sqrt(x * x + y * y);
      const double th2 = -2 * atan2(y, x);
      const double s = 1.2 * sin(th2), c = 1.4 - cos(th2);
      out[0] = r * (-y * s - c * x);
      out[1] = r * (x * s - c * y) - 1;
    } break;
    case FLOW_HILLS_VORTEX: /* R:flows.py:126-128 */
      out[0] = 2.0 * x * y;
      out[1] = -2.0 * (2.0 * x * x - 1.0) - 2.0 * y * y;
      break;
    case FLOW_SIGMOIDAL: /* R:flows.py:140-142 */
      out[0] = 0.4 * tanh(10.0 * (x - 0.5 * t));
      out[1] = y * 0;
      break;
    case FLOW_AUTONOMOUS_DUFFING: /* R:flows.py:155-157 */
      out[0] = y;
      out[1] = x - x * x * x;
      break;
    case FLOW_DUFFING: /* R:flows.py:174-176  p = A, gamma, omega */
      out[0] = y;
      out[1] = x - x * x * x - p[1] * y + p[0] * sin(p[2] * t);
      break;
    case FLOW_PENDULUM: /* R:flows.py:193-195  p = A, gamma, omega */
      out[0] = y;
      out[1] = -sin(x) - p[1] * y + p[0] * sin(p[2] * t);
      break;
    case FLOW_HURRICANE: { /* R:flows.py:216-221  p = eye_alpha, eye_omega, eye_amplitude */
      const double sq = sqrt(1.0 + 4.0 * p[0]);
      const double beta = 1.0 / sq;
      const double yt = y - 0.5 * (1.0 + sq);
      const double den = x * x + yt * yt + p[0];
      out[0] = -yt / den - beta;
      out[1] = x / den + p[2] * y * cos(p[1] * t);
    } break;
    case FLOW_VAN_DER_POL: /* R:flows.py:239-241  p = A, mu, gamma, omega */
      out[0] = y;
      out[1] = p[1] * (1 - x * x) * y - x - p[2] * y + p[0] * sin(p[3] * t);
      break;
    case FLOW_BEAD: /* R:flows.py:260-262  p = epsilon, gamma */
      out[0] = y;
      out[1] = (1 / p[0]) * (sin(x) * (p[1] * cos(x) - 1) - y);
      break;
    case FLOW_LOTKA_VOLTERRA: /* R:flows.py:282-284  p = alpha, beta, gamma, delta */
      out[0] = p[0] * x - p[1] * x * y;
      out[1] = p[3] * x * y - p[2] * y;
      break;
    case FLOW_ABC: { /* R:flows.py:304-307  p = ABC_Amplitude, A, B, C */
      const double z = Y[2];
      const double At = p[1] + p[0] * sin(PI * t);
      out[0] = At * sin(z) + p[3] * cos(y);
      out[1] = p[2] * sin(x) + At * cos(z);
      out[2] = p[3] * sin(y) + p[2] * cos(x);
    } break;
    case FLOW_LORENZ: { /* R:flows.py:328-331  p = sigma, rho, beta */
      const double z = Y[2];
      out[0] = p[0] * (y - x);
      out[1] = x * (p[1] - z) - y;
      out[2] = x * y - p[2] * z;
    } break;
    default:
      out[0] = out[1] = NAN;
  }
}

int orc_flow_eval(int id, const double* p, double t, const double* Y, double* out) {
  if (id < 0 || id >= FLOW_COUNT) return -1;
  flow_rhs(id, p, t, Y, out);
  return 0;
}

/* Dormand-Prince tableau, SP:scipy/integrate/_ivp/rk.py:541-552 */
static const double DP_C[6] = {0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1};
static const double DP_A[6][5] = {
    {0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
static const double DP_B[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
static const double DP_E[7] = {-71.0 / 57600, 0,           71.0 / 16695, -71.0 / 1920,
                               17253.0 / 339200, -22.0 / 525, 1.0 / 40};

/* SP:scipy/integrate/_ivp/common.py:63-65  norm(x) = ||x||_2 / sqrt(n) */
static double rms(const double* v, int n) {
  double s = 0;
  for (int i = 0; i < n; ++i) s += v[i] * v[i];
  return sqrt(s) / sqrt((double)n);
}

typedef struct {
  int64_t nfev, naccept, nreject;
  int status; /* 0 finished, -1 step size too small (SP:.../ivp.py:658-666) */
} orc_stats;

/* One solve_ivp(method='RK45') call; only the endpoint is kept, as the
 * reference does with [-1, :] (R:_base_classes.py:180-182). */
static void rk45_solve(int id, const double* p, int n, const double* y0, double t0, double tf,
                       double rtol, const double* atol, double max_step, double first_step,
                       double* y_out, orc_stats* st) {
  double y[ORC_MAXDIM], f[ORC_MAXDIM], K[7][ORC_MAXDIM], ynew[ORC_MAXDIM] = {0}, tmp[ORC_MAXDIM],
      scale[ORC_MAXDIM];
  st->nfev = st->naccept = st->nreject = 0;
  st->status = 0;
  /* validate_tol, common.py:47-51 */
  if (rtol < 100 * DBL_EPSILON) rtol = 100 * DBL_EPSILON;
  const double direction = (tf != t0) ? ((tf > t0) ? 1.0 : -1.0) : 1.0; /* base.py:164 */
  double t = t0;
  memcpy(y, y0, sizeof(double) * n);
  flow_rhs(id, p, t, y, f); /* rk.py:94 */
  st->nfev++;
  double h_abs;
  if (first_step > 0) {
    h_abs = first_step; /* validate_first_step, common.py:10-16 (caller validates) */
  } else {             /* select_initial_step, common.py:109-134 */
    const double L = fabs(tf - t0);
    if (L == 0.0) {
      h_abs = 0.0;
    } else {
      for (int i = 0; i < n; ++i) scale[i] = atol[i] + fabs(y[i]) * rtol;
      for (int i = 0; i < n; ++i) tmp[i] = y[i] / scale[i];
      const double d0 = rms(tmp, n);
      for (int i = 0; i < n; ++i) tmp[i] = f[i] / scale[i];
      const double d1 = rms(tmp, n);
      double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
      h0 = fmin(h0, L);
      for (int i = 0; i < n; ++i) ynew[i] = y[i] + h0 * direction * f[i];
      double f1[ORC_MAXDIM];
      flow_rhs(id, p, t + h0 * direction, ynew, f1);
      st->nfev++;
      for (int i = 0; i < n; ++i) tmp[i] = (f1[i] - f[i]) / scale[i];
      const double d2 = rms(tmp, n) / h0;
      double h1;
      if (d1 <= 1e-15 && d2 <= 1e-15)
        h1 = fmax(1e-6, h0 * 1e-3);
      else
        h1 = pow(0.01 / fmax(d1, d2), 1.0 / 5.0);
      h_abs = fmin(fmin(100 * h0, h1), fmin(L, max_step));
    }
  }
  /* solve_ivp loop: solver.step() until finished, base.py:179-210 */
  for (;;) {
    if (t == tf) break; /* base.py:193-198 */
    /* _step_impl, rk.py:111-176 */
    const double min_step = 10 * fabs(nextafter(t, direction * INFINITY) - t);
    if (h_abs > max_step)
      h_abs = max_step;
    else if (h_abs < min_step)
      h_abs = min_step;
    int rejected = 0;
    double t_new, h;
    for (;;) {
      if (h_abs < min_step) {
        st->status = -1;
        memcpy(y_out, y, sizeof(double) * n);
        return;
      }
      h = h_abs * direction;
      t_new = t + h;
      if (direction * (t_new - tf) > 0) t_new = tf;
      h = t_new - t;
      h_abs = fabs(h);
      /* rk_step, rk.py:60-69: dy = dot(K[:s].T, a[:s]) * h */
      memcpy(K[0], f, sizeof(double) * n);
      for (int s = 1; s < 6; ++s) {
        for (int i = 0; i < n; ++i) {
          double acc = 0;
          for (int j = 0; j < s; ++j) acc += K[j][i] * DP_A[s][j];
          tmp[i] = y[i] + acc * h;
        }
        flow_rhs(id, p, t + DP_C[s] * h, tmp, K[s]);
      }
      for (int i = 0; i < n; ++i) {
        double acc = 0;
        for (int j = 0; j < 6; ++j) acc += K[j][i] * DP_B[j];
        ynew[i] = y[i] + h * acc;
      }
      flow_rhs(id, p, t + h, ynew, K[6]);
      st->nfev += 6;
      for (int i = 0; i < n; ++i) {
        scale[i] = atol[i] + fmax(fabs(y[i]), fabs(ynew[i])) * rtol; /* rk.py:146 */
        double acc = 0;
        for (int j = 0; j < 7; ++j) acc += K[j][i] * DP_E[j];
        tmp[i] = acc * h / scale[i]; /* rk.py:105-109 */
      }
      const double err = rms(tmp, n);
      if (err < 1) {
        double factor;
        if (err == 0)
          factor = 10;
        else
          factor = fmin(10, 0.9 * pow(err, -0.2));
        if (rejected) factor = fmin(1, factor);
        h_abs *= factor;
        break;
      } else {
        h_abs *= fmax(0.2, 0.9 * pow(err, -0.2));
        rejected = 1;
        st->nreject++;
      }
    }
    st->naccept++;
    t = t_new;
    memcpy(y, ynew, sizeof(double) * n);
    memcpy(f, K[6], sizeof(double) * n);
    if (direction * (t - tf) >= 0) break; /* base.py:207-208 */
  }
  memcpy(y_out, y, sizeof(double) * n);
}

/* ---- tiny pthread parallel-for (libgomp is not in this image) ---- */
typedef void (*orc_body)(int64_t k, void* ctx);
typedef struct {
  atomic_llong next;
  int64_t n, chunk;
  orc_body body;
  void* ctx;
} orc_pf;

static void* orc_pf_worker(void* arg) {
  orc_pf* pf = (orc_pf*)arg;
  for (;;) {
    int64_t b = atomic_fetch_add(&pf->next, pf->chunk);
    if (b >= pf->n) break;
    int64_t e = b + pf->chunk < pf->n ? b + pf->chunk : pf->n;
    for (int64_t k = b; k < e; ++k) pf->body(k, pf->ctx);
  }
  return NULL;
}

static void parallel_for(int64_t n, int64_t chunk, orc_body body, void* ctx, int nthreads) {
  if (nthreads <= 0) nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
  if (nthreads > 256) nthreads = 256;
  orc_pf pf;
  atomic_init(&pf.next, 0);
  pf.n = n; pf.chunk = chunk; pf.body = body; pf.ctx = ctx;
  if (nthreads <= 1 || n <= chunk) { orc_pf_worker(&pf); return; }
  pthread_t th[256];
  int started = 0;
  for (int i = 0; i < nthreads - 1; ++i)
    if (pthread_create(&th[started], NULL, orc_pf_worker, &pf) == 0) ++started;
  orc_pf_worker(&pf);
  for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
}

typedef struct {
  int id, n;
  const double *p, *seeds, *x, *y, *atol;
  int64_t xdim;
  double t0, tf, rtol, max_step, first_step;
  double* out;
  int64_t *nfev, *naccept;
  int32_t* status;
} orc_job;

static void endpoint_body(int64_t k, void* c) {
  orc_job* j = (orc_job*)c;
  orc_stats st;
  rk45_solve(j->id, j->p, j->n, j->seeds + k * j->n, j->t0, j->tf, j->rtol, j->atol, j->max_step,
             j->first_step, j->out + k * j->n, &st);
  if (j->nfev) j->nfev[k] = st.nfev;
  if (j->naccept) j->naccept[k] = st.naccept;
  if (j->status) j->status[k] = st.status;
}

static void flowmap_body(int64_t k, void* c) {
  orc_job* j = (orc_job*)c;
  const double seed[2] = {j->x[k % j->xdim], j->y[k / j->xdim]};
  orc_stats st;
  rk45_solve(j->id, j->p, 2, seed, j->t0, j->tf, j->rtol, j->atol, j->max_step, 0.0,
             j->out + 2 * k, &st);
  if (j->nfev) j->nfev[k] = st.nfev;
  if (j->status) j->status[k] = st.status;
}

/* Generic endpoint integration of m seeds of dimension n (seeds[m][n]). */
int orc_rk45_endpoints(int id, const double* p, int n, const double* seeds, int64_t m, double t0,
                       double tf, double rtol, const double* atol, double max_step,
                       double first_step, double* out, int64_t* nfev, int64_t* naccept,
                       int32_t* status, int nthreads) {
  if (id < 0 || id >= FLOW_COUNT || n != k_flow_dim[id]) return -1;
  orc_job j = {.id = id, .n = n, .p = p, .seeds = seeds, .atol = atol, .t0 = t0, .tf = tf,
               .rtol = rtol, .max_step = max_step, .first_step = first_step, .out = out,
               .nfev = nfev, .naccept = naccept, .status = status};
  parallel_for(m, 64, endpoint_body, &j, nthreads);
  return 0;
}

/* Flow map over product(y, x): y outer, x inner, seed passed as (x, y)
 * (R:_base_classes.py:181-187); out[ydim][xdim][2]. */
int orc_flow_map(int id, const double* p, const double* x, int64_t xdim, const double* y,
                 int64_t ydim, double t0, double tf, double rtol, double atol_s, double max_step,
                 double* out, int64_t* nfev, int32_t* status, int nthreads) {
  if (id < 0 || id >= FLOW_COUNT || k_flow_dim[id] != 2) return -1;
  const double atol[2] = {atol_s, atol_s};
  orc_job j = {.id = id, .n = 2, .p = p, .x = x, .y = y, .xdim = xdim, .atol = atol, .t0 = t0,
               .tf = tf, .rtol = rtol, .max_step = max_step, .out = out, .nfev = nfev,
               .status = status};
  parallel_for(xdim * ydim, 64, flowmap_body, &j, nthreads);
  return 0;
}

/* np.gradient along one axis of f[n0][n1] (C order) with coordinate array c
 * (length = that axis), SP:numpy/lib/_function_base_impl.py:1246-1384.
 * axis 0: stride n1, axis 1: stride 1. */
static void gradient_axis(const double* f, int64_t n0, int64_t n1, int axis, const double* c,
                          int edge_order, double* out) {
  const int64_t n = axis == 0 ? n0 : n1;
  const int64_t other = axis == 0 ? n1 : n0;
  const int64_t sa = axis == 0 ? n1 : 1, so = axis == 0 ? 1 : n1;
  /* uniform iff all diffs equal diff[0] (:1246-1251) */
  int uniform = 1;
  const double d0 = c[1] - c[0];
  for (int64_t i = 1; i + 1 < n; ++i)
    if (c[i + 1] - c[i] != d0) uniform = 0;
  for (int64_t o = 0; o < other; ++o) {
    const double* g = f + o * so;
    double* r = out + o * so;
    for (int64_t i = 1; i + 1 < n; ++i) {
      if (uniform) {
        r[i * sa] = (g[(i + 1) * sa] - g[(i - 1) * sa]) / (2. * d0); /* :1305 */
      } else {
        const double dx1 = c[i] - c[i - 1], dx2 = c[i + 1] - c[i]; /* :1307-1318 */
        const double a = -(dx2) / (dx1 * (dx1 + dx2));
        const double b = (dx2 - dx1) / (dx1 * dx2);
        const double cc = dx1 / (dx2 * (dx1 + dx2));
        r[i * sa] = a * g[(i - 1) * sa] + b * g[i * sa] + cc * g[(i + 1) * sa];
      }
    }
    if (edge_order == 1) { /* :1321-1334 */
      r[0] = (g[sa] - g[0]) / (uniform ? d0 : (c[1] - c[0]));
      r[(n - 1) * sa] = (g[(n - 1) * sa] - g[(n - 2) * sa]) / (uniform ? d0 : (c[n - 1] - c[n - 2]));
    } else { /* :1337-1372 */
      double a, b, cc;
      if (uniform) {
        a = -1.5 / d0; b = 2. / d0; cc = -0.5 / d0;
      } else {
        const double dx1 = c[1] - c[0], dx2 = c[2] - c[1];
        a = -(2. * dx1 + dx2) / (dx1 * (dx1 + dx2));
        b = (dx1 + dx2) / (dx1 * dx2);
        cc = -dx1 / (dx2 * (dx1 + dx2));
      }
      r[0] = a * g[0] + b * g[sa] + cc * g[2 * sa];
      if (uniform) {
        a = 0.5 / d0; b = -2. / d0; cc = 1.5 / d0;
      } else {
        const double dx1 = c[n - 2] - c[n - 3], dx2 = c[n - 1] - c[n - 2];
        a = (dx2) / (dx1 * (dx1 + dx2));
        b = -(dx2 + dx1) / (dx1 * dx2);
        cc = (2. * dx2 + dx1) / (dx2 * (dx1 + dx2));
      }
      r[(n - 1) * sa] = a * g[(n - 3) * sa] + b * g[(n - 2) * sa] + cc * g[(n - 1) * sa];
    }
  }
}

int orc_gradient(const double* f, int64_t ydim, int64_t xdim, const double* y, const double* x,
                 int edge_order, double* dfdy, double* dfdx) {
  if (edge_order < 1 || edge_order > 2) return -1;        /* :1255-1256 */
  if (ydim < edge_order + 1 || xdim < edge_order + 1) return -2; /* :1288-1291 */
  gradient_axis(f, ydim, xdim, 0, y, edge_order, dfdy);
  gradient_axis(f, ydim, xdim, 1, x, edge_order, dfdx);
  return 0;
}

/* FTLE field from a flow map, R:lagrangian.py:50-76; closed-form lambda_max. */
int orc_ftle(const double* fm, int64_t ydim, int64_t xdim, const double* y, const double* x,
             double t0, double tf, int edge_order, double* field) {
  const int64_t N = ydim * xdim;
  double* buf = (double*)malloc(sizeof(double) * N * 6);
  if (!buf) return -3;
  double *fx = buf, *fy = buf + N, *dfxdy = buf + 2 * N, *dfxdx = buf + 3 * N,
         *dfydy = buf + 4 * N, *dfydx = buf + 5 * N;
  for (int64_t k = 0; k < N; ++k) { fx[k] = fm[2 * k]; fy[k] = fm[2 * k + 1]; }
  int rc = orc_gradient(fx, ydim, xdim, y, x, edge_order, dfxdy, dfxdx);
  if (!rc) rc = orc_gradient(fy, ydim, xdim, y, x, edge_order, dfydy, dfydx);
  if (!rc) {
    const double c = 1.0 / (2.0 * fabs(tf - t0));
    for (int64_t k = 0; k < N; ++k) {
      /* JF = [[dfxdx, dfxdy],[dfydx, dfydy]],  C = JF^T JF  (:65-66) */
      const double a = dfxdx[k], b = dfxdy[k], cc = dfydx[k], d = dfydy[k];
      const double c11 = a * a + cc * cc, c12 = a * b + cc * d, c22 = b * b + d * d;
      const double hm = 0.5 * (c11 - c22);
      const double lam = 0.5 * (c11 + c22) + sqrt(hm * hm + c12 * c12);
      field[k] = (lam >= 1) ? c * log(lam) : 0.0; /* :70-73 */
    }
  }
  free(buf);
  return rc;
}
