// K0 fused into K3 — Eulerian diagnostics of an analytic flow without materialising (u, v).
//
// The reference evaluates `u, v = f(t, np.meshgrid(x, y))`
// (R:src/dynlab/diagnostics/_base_classes.py:101-102), takes four np.gradient fields (:110-111)
// and loops over the nodes (R:src/dynlab/diagnostics/eulerian.py:47-59, :179-195, :317-333).
// With K0 and K3 as separate kernels that is 16 B/node written, 16 B/node read back and 8 B/node
// written; here the velocity at a node is produced in registers right where the
// stencil consumes it, so only the 8 B/node result (+ mask / eigenvector) touches HBM.
//
// Work decomposition: a CTA of four warps owns four adjacent strips of 32 columns (30 outputs +
// 2 halo each) and walks EF_ROWS = 126 rows.  Per task every thread evaluates the x part of its
// column once (flows.cuh: flow_xpart), the 128 threads evaluate the 128 y parts of the task's
// rows (halo included) once into shared memory (flow_ypart), and each node only costs
// flow_combine — for the double gyre and the Bickley jet the sin/cos/tanh/cosh calls are
// amortised over 126 rows / 120 columns.  Flows
// that do not separate evaluate flow_space per node (still no HBM round trip).  Rows i-1, i,
// i+1 of a column live in registers, x neighbours come from the adjacent lanes by shuffle,
// boundary nodes go through the shared edge kernel with a loader that evaluates the flow.
// The gradients / epilogues are the ones of dlb_euler_stencil and the velocities are bit-identical
// to dlb_flow_eval_grid's (flows.cuh writes every multiply-add as an explicit fma), so the result
// equals the two-kernel path exactly.
#include "flows.cuh"
#include "stencil_device.cuh"

namespace {

constexpr int EF_BLOCK = 128;          // 4 warps side by side
constexpr int EF_ROWS = EF_BLOCK - 2;  // rows per CTA task: EF_ROWS + 2 y parts = one per thread
constexpr int EF_COLS = 30;            // output columns per warp
constexpr int EF_WARPS = EF_BLOCK / 32;

template <int ID>
struct FlowGrid {  // what the kernels need to evaluate the flow at grid node (row, col)
  FlowConsts fc;
  double t;
  const double* xs;
  const double* ys;
};

// time part (sin(w t) for the double gyre, t itself for the Bickley jet, nothing for autonomous
// flows), computed on the device exactly like flow_rhs does
template <int ID>
__device__ __forceinline__ void time_part(const FlowGrid<ID>& G, double* tc) {
  dm::TrigSafe trig;
  double tcs[1][FlowTime<ID>::N > 0 ? FlowTime<ID>::N : 1] = {};
  flow_time<ID, 1>(G.fc, trig, &G.t, tcs);
  tc[0] = tcs[0][0];
}

template <int ID>
__device__ __forceinline__ double2 node_value(const FlowGrid<ID>& G, const double* tc, double x,
                                              double y) {
  dm::TrigSafe trig;
  const double Y[2] = {x, y};
  double out[2];
  flow_space<ID>(G.fc, trig, tc, Y, out);
  return make_double2(out[0], out[1]);
}

// Loader for the edge kernel: at(local row, column) evaluates the flow (grow0 == 0 here).
template <int ID>
struct LoaderFlow {
  FlowGrid<ID> G;
  __device__ __forceinline__ double2 at(long long lrow, long long j) const {
    double tc[1];
    time_part<ID>(G, tc);
    return node_value<ID>(G, tc, __ldg(G.xs + j), __ldg(G.ys + lrow));
  }
};

template <int ID, int MODE, bool XI>
__global__ void __launch_bounds__(EF_BLOCK)
    euler_flow_interior_kernel(const StencilParams P, const FlowGrid<ID> G) {
  constexpr int NX = FlowXY<ID>::NX, NY = FlowXY<ID>::NY;
  constexpr bool SEP = NX > 0;
  constexpr int NYS = SEP ? NY : 1;  // non-separable: the row's y itself
  constexpr int NXS = SEP ? NX : 1;
  __shared__ double yps[EF_ROWS + 2][NYS];
  __shared__ double2 ysm[EF_ROWS][2];
  const long long xdim = P.ax.n, ydim = P.ay.n;
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  const long long r_lo = (P.out_row0 > 1) ? P.out_row0 : 1;
  const long long r_hi = (P.out_row0 + P.out_rows < ydim - 1) ? P.out_row0 + P.out_rows : ydim - 1;
  const long long ncols = xdim - 2;
  if (r_hi <= r_lo || ncols <= 0) return;
  const long long nstrips = (ncols + EF_COLS * EF_WARPS - 1) / (EF_COLS * EF_WARPS);  // CTA strips
  const long long nchunks = (r_hi - r_lo + EF_ROWS - 1) / EF_ROWS;
  const long long ntasks = nstrips * nchunks;
  const bool ux = P.ax.uniform != 0, uy = P.ay.uniform != 0;
  const double ix = P.ax.inv2h, iy = P.ay.inv2h;
  dm::TrigSafe trig;
  double tc[1];
  time_part<ID>(G, tc);

  for (long long task = blockIdx.x; task < ntasks; task += gridDim.x) {  // CTA-uniform
    const long long chunk = task / nstrips;
    const long long strip = task - chunk * nstrips;
    const long long jj = (strip * EF_WARPS + wib) * EF_COLS + lane;  // lane 0 / 31: halo columns
    const long long j = (jj < xdim) ? jj : xdim - 1;                 // clamp the tail lanes
    const bool emit_lane = (lane >= 1) && (lane <= EF_COLS) && (jj <= xdim - 2);
    const long long g0 = r_lo + chunk * EF_ROWS;
    const int nrows = (int)((g0 + EF_ROWS < r_hi) ? EF_ROWS : r_hi - g0);
    double xa = 0.0, xb = 0.0, xc = 0.0;
    if (!ux) {
      xa = __ldg(P.ax.ca + j);
      xb = __ldg(P.ax.cb + j);
      xc = __ldg(P.ax.cc + j);
    }
    // x part of this thread's column; y parts of rows g0-1 .. g0+nrows, one row per thread
    double xp[NXS];
    const double xj = __ldg(G.xs + j);
    if constexpr (SEP)
      flow_xpart<ID>(G.fc, trig, tc, xj, xp);
    else
      xp[0] = xj;
    __syncthreads();  // the previous task's readers are done
    if (tid < nrows + 2) {
      const double yr = __ldg(G.ys + (g0 - 1 + tid));
      if constexpr (SEP) {
        double yp[NYS];
        flow_ypart<ID>(G.fc, trig, tc, yr, yp);
#pragma unroll
        for (int q = 0; q < NYS; ++q) yps[tid][q] = yp[q];
      } else {
        yps[tid][0] = yr;
      }
    }
    if (!uy && tid < nrows) {
      const double2* src = P.ay4 + 2 * (g0 + tid);
      ysm[tid][0] = __ldg(src);
      ysm[tid][1] = __ldg(src + 1);
    }
    __syncthreads();
    auto value = [&](int r) -> double2 {  // velocity at (row g0-1+r, column j)
      if constexpr (SEP) {
        double yp[NYS], out[2];
#pragma unroll
        for (int q = 0; q < NYS; ++q) yp[q] = yps[r][q];
        flow_combine<ID>(xp, yp, out);
        return make_double2(out[0], out[1]);
      } else {
        return node_value<ID>(G, tc, xp[0], yps[r][0]);
      }
    };
    double2 U = value(0), C = value(1);
    long long o = (g0 - P.out_row0) * xdim + j;
    for (int k = 0; k < nrows; ++k) {
      const double2 D = value(k + 2);
      const double2 L = make_double2(__shfl_up_sync(FULL, C.x, 1), __shfl_up_sync(FULL, C.y, 1));
      const double2 R = make_double2(__shfl_down_sync(FULL, C.x, 1), __shfl_down_sync(FULL, C.y, 1));
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
        const double2 yab = ysm[k][0], ycz = ysm[k][1];
        d0dy = fma(ycz.x, D.x, fma(yab.y, C.x, yab.x * U.x));
        d1dy = fma(ycz.x, D.y, fma(yab.y, C.y, yab.x * U.y));
      }
      emit_pred<KIND_EULER, MODE, XI>(P, C, d0dx, d1dx, d0dy, d1dy, o, nullptr, emit_lane);
      o += xdim;
      U = C;
      C = D;
    }
  }
}

template <int ID, int MODE, bool XI>
int launch_euler_flow(const StencilParams& P, const FlowGrid<ID>& G, cudaStream_t stream) {
  const long long xdim = P.ax.n, ydim = P.ay.n;
  const long long r0 = P.out_row0, r1 = P.out_row0 + P.out_rows;
  const long long cap = (long long)dlbst::num_sms() * 16;
  const long long n_full = ((r0 == 0) ? 1 : 0) + ((r1 == ydim && ydim > 1) ? 1 : 0);
  const long long total_edge = n_full * xdim + (P.out_rows - n_full) * 2;
  long long eb = (total_edge + ST_BLOCK - 1) / ST_BLOCK;
  if (eb > cap) eb = cap;
  if (eb > 0) {
    const LoaderFlow<ID> ld{G};
    stencil_edge_kernel<KIND_EULER, MODE, XI, LoaderFlow<ID>><<<(unsigned)eb, ST_BLOCK, 0, stream>>>(P, ld);
    DLB_CUDA_CHECK(cudaGetLastError());
  }
  const long long r_lo = (r0 > 1) ? r0 : 1, r_hi = (r1 < ydim - 1) ? r1 : ydim - 1;
  if (r_hi > r_lo && xdim > 2) {
    const long long nstrips = (xdim - 2 + EF_COLS * EF_WARPS - 1) / (EF_COLS * EF_WARPS);
    const long long nchunks = (r_hi - r_lo + EF_ROWS - 1) / EF_ROWS;
    long long blocks = nstrips * nchunks;
    if (blocks > cap) blocks = cap;
    euler_flow_interior_kernel<ID, MODE, XI><<<(unsigned)blocks, EF_BLOCK, 0, stream>>>(P, G);
    DLB_CUDA_CHECK(cudaGetLastError());
  }
  return DLB_OK;
}

template <int ID>
int dispatch_mode(int mode, bool xi, const StencilParams& P, const FlowGrid<ID>& G,
                  cudaStream_t stream) {
  switch (mode) {
    case DLB_EULER_S_MIN:
      return xi ? launch_euler_flow<ID, DLB_EULER_S_MIN, true>(P, G, stream)
                : launch_euler_flow<ID, DLB_EULER_S_MIN, false>(P, G, stream);
    case DLB_EULER_S_MAX:
      return xi ? launch_euler_flow<ID, DLB_EULER_S_MAX, true>(P, G, stream)
                : launch_euler_flow<ID, DLB_EULER_S_MAX, false>(P, G, stream);
    case DLB_EULER_RATE:
      return launch_euler_flow<ID, DLB_EULER_RATE, false>(P, G, stream);
    default:
      return launch_euler_flow<ID, DLB_EULER_RATIO, false>(P, G, stream);
  }
}

}  // namespace

extern "C" int dlb_euler_stencil_flow(int mode, int flow_id, const double* h_params, int n_params,
                                      double t, const double* h_x, int64_t xdim, const double* h_y,
                                      int64_t ydim, int edge_order, int64_t out_row0,
                                      int64_t out_rows, int negate, double* d_field,
                                      uint8_t* d_mask, double* d_xi, int xi_mode,
                                      void* d_workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FlowConsts fc;
  if (!flow_prepare(flow_id, h_params, n_params, &fc)) {
    dlb_set_error("unknown flow id %d or wrong parameter count %d", flow_id, n_params);
    return DLB_ERR_ARG;
  }
  if (k_flow_info[flow_id].dim != 2) {
    dlb_set_error("dlb_euler_stencil_flow: flow %d is not a 2-D flow", flow_id);
    return DLB_ERR_ARG;
  }
  if (!d_field || (xi_mode != DLB_XI_NONE && !d_xi)) {
    dlb_set_error("dlb_euler_stencil_flow: null pointer");
    return DLB_ERR_ARG;
  }
  if (mode < DLB_EULER_S_MIN || mode > DLB_EULER_RATIO) {
    dlb_set_error("dlb_euler_stencil_flow: unknown mode %d", mode);
    return DLB_ERR_ARG;
  }
  if ((mode == DLB_EULER_RATE || mode == DLB_EULER_RATIO) && (!d_mask || xi_mode != DLB_XI_NONE)) {
    dlb_set_error("dlb_euler_stencil_flow: rate/ratio need d_mask and produce no eigenvector");
    return DLB_ERR_ARG;
  }
  StencilParams P;
  memset(&P, 0, sizeof(P));
  int rc = dlbst::setup_axes(P, h_x, xdim, h_y, ydim, edge_order, d_workspace, workspace_bytes,
                             stream);
  if (rc) return rc;
  P.nloc = ydim;  // every row can be evaluated: no slab, no halo to provide
  P.grow0 = 0;
  P.out_row0 = out_row0;
  P.out_rows = out_rows;
  rc = dlbst::check_slab(P, ydim);
  if (rc || out_rows == 0) return rc;
  P.field = d_field;
  P.mask = d_mask;
  P.xi = (double2*)d_xi;
  P.xi_mode = xi_mode;
  P.negate = negate;
  // coordinates: setup_axes keeps a cached device copy next to the coefficients
  const bool xi = xi_mode != DLB_XI_NONE;
  rc = DLB_ERR_ARG;
  DLB_DISPATCH_FLOW(flow_id, {
    if constexpr (ID < DLB_FLOW_ABC) {
      FlowGrid<ID> G;
      G.fc = fc;
      G.t = t;
      G.xs = P.xs;
      G.ys = P.ys;
      rc = dispatch_mode<ID>(mode, xi, P, G, stream);
    }
  });
  return rc;
}
