// K0 — evaluate a flow on a point list or on meshgrid(x, y).
// Replaces `self.u, self.v = f(t, np.meshgrid(x, y))` in EulerianDiagnostic2D.compute
// (R:src/dynlab/diagnostics/_base_classes.py:101-102) and lets the reference's flow golden
// vectors (R:tests/test_flows.py) exercise the device twins directly.
// Elementwise and HBM-bound: 16 B written per grid node (grid mode reads only the two
// coordinate vectors), grid-stride with coalesced stores.
#include "flows.cuh"

namespace {

constexpr int FE_BLOCK = 256;

struct EvalParams {
  FlowConsts fc;
  double t;
  const double* px;
  const double* py;
  const double* pz;
  long long n;
  long long xdim;  // grid mode: n = xdim * ydim, point k = (x[k % xdim], y[k / xdim])
  double* u;
  double* v;
  double* w;
};

template <int ID, bool GRID>
__global__ void __launch_bounds__(FE_BLOCK) flow_eval_kernel(const EvalParams P) {
  constexpr int N = (ID >= DLB_FLOW_ABC) ? 3 : 2;
  for (long long k = (long long)blockIdx.x * FE_BLOCK + threadIdx.x; k < P.n;
       k += (long long)gridDim.x * FE_BLOCK) {
    double Y[3] = {0.0, 0.0, 0.0}, out[3] = {0.0, 0.0, 0.0};
    if (GRID) {
      const long long i = k / P.xdim;
      Y[0] = P.px[k - i * P.xdim];
      Y[1] = P.py[i];
    } else {
      Y[0] = P.px[k];
      Y[1] = P.py[k];
      if (N == 3) Y[2] = P.pz[k];
    }
    flow_rhs<ID>(P.fc, P.t, Y, out);
    P.u[k] = out[0];
    P.v[k] = out[1];
    if (N == 3) P.w[k] = out[2];
  }
}

// Grid mode for x/y-separable flows (flows.cuh: FlowXY): a CTA owns 128 columns x FE_TROWS
// rows; every thread evaluates the x part of its column once, the CTA evaluates the y parts of
// the tile's rows once into shared memory, and a node costs flow_combine plus two coalesced
// stores - the kernel becomes write-bound (16 B/node) instead of transcendental-bound.  Values
// are bit-identical to the per-node kernel above (same operations on the same intermediates).
constexpr int FE_TCOLS = 128;
constexpr int FE_TROWS = 128;

template <int ID>
__global__ void __launch_bounds__(FE_TCOLS) flow_eval_tile_kernel(const EvalParams P, long long ydim) {
  constexpr int NX = FlowXY<ID>::NX > 0 ? FlowXY<ID>::NX : 1;
  constexpr int NY = FlowXY<ID>::NY > 0 ? FlowXY<ID>::NY : 1;
  __shared__ double yps[FE_TROWS][NY];
  const long long xdim = P.xdim;
  const long long nstrips = (xdim + FE_TCOLS - 1) / FE_TCOLS;
  const long long nchunks = (ydim + FE_TROWS - 1) / FE_TROWS;
  const int t = threadIdx.x;
  dm::TrigSafe trig;
  double tcs[1][FlowTime<ID>::N > 0 ? FlowTime<ID>::N : 1] = {};
  flow_time<ID, 1>(P.fc, trig, &P.t, tcs);
  for (long long task = blockIdx.x; task < nstrips * nchunks; task += gridDim.x) {
    const long long chunk = task / nstrips, strip = task - chunk * nstrips;
    const long long j = strip * FE_TCOLS + t, r0 = chunk * FE_TROWS;
    const int nrows = (int)((r0 + FE_TROWS <= ydim) ? FE_TROWS : ydim - r0);
    double xp[NX];
    flow_xpart<ID>(P.fc, trig, tcs[0], P.px[(j < xdim) ? j : xdim - 1], xp);
    __syncthreads();  // the previous task's readers are done
    if (t < nrows) {
      double yp[NY];
      flow_ypart<ID>(P.fc, trig, tcs[0], P.py[r0 + t], yp);
#pragma unroll
      for (int q = 0; q < NY; ++q) yps[t][q] = yp[q];
    }
    __syncthreads();
    if (j < xdim) {
      long long o = r0 * xdim + j;
      for (int r = 0; r < nrows; ++r, o += xdim) {
        double yp[NY], out[2];
#pragma unroll
        for (int q = 0; q < NY; ++q) yp[q] = yps[r][q];
        flow_combine<ID>(xp, yp, out);
        P.u[o] = out[0];
        P.v[o] = out[1];
      }
    }
  }
}

template <int ID, bool GRID>
int launch_eval(const EvalParams& P, cudaStream_t stream) {
  int dev = 0, sms = 0;
  DLB_CUDA_CHECK(cudaGetDevice(&dev));
  DLB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if constexpr (GRID && FlowXY<ID>::NX > 0) {
    const long long ydim = P.n / P.xdim;
    long long tiles = ((P.xdim + FE_TCOLS - 1) / FE_TCOLS) * ((ydim + FE_TROWS - 1) / FE_TROWS);
    if (tiles > (long long)sms * 16) tiles = (long long)sms * 16;
    flow_eval_tile_kernel<ID><<<(unsigned)tiles, FE_TCOLS, 0, stream>>>(P, ydim);
    DLB_CUDA_CHECK(cudaGetLastError());
    return DLB_OK;
  }
  long long blocks = (P.n + FE_BLOCK - 1) / FE_BLOCK;
  if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
  if (blocks < 1) blocks = 1;
  flow_eval_kernel<ID, GRID><<<(unsigned)blocks, FE_BLOCK, 0, stream>>>(P);
  DLB_CUDA_CHECK(cudaGetLastError());
  return DLB_OK;
}

}  // namespace

extern "C" int dlb_flow_eval_points(int flow_id, const double* h_params, int n_params, double t,
                                    const double* d_px, const double* d_py, const double* d_pz,
                                    int64_t n, double* d_u, double* d_v, double* d_w,
                                    void* stream_) {
  EvalParams P;
  memset(&P, 0, sizeof(P));
  if (!flow_prepare(flow_id, h_params, n_params, &P.fc)) {
    dlb_set_error("unknown flow id %d or wrong parameter count %d", flow_id, n_params);
    return DLB_ERR_ARG;
  }
  const bool three = k_flow_info[flow_id].dim == 3;
  if (n < 0 || !d_px || !d_py || !d_u || !d_v || (three && (!d_pz || !d_w))) {
    dlb_set_error("dlb_flow_eval_points: bad argument");
    return DLB_ERR_ARG;
  }
  if (n == 0) return DLB_OK;
  P.t = t;
  P.px = d_px;
  P.py = d_py;
  P.pz = d_pz;
  P.n = n;
  P.u = d_u;
  P.v = d_v;
  P.w = d_w;
  int rc = DLB_ERR_ARG;
  DLB_DISPATCH_FLOW(flow_id, { rc = launch_eval<ID, false>(P, (cudaStream_t)stream_); });
  return rc;
}

extern "C" int dlb_flow_eval_grid(int flow_id, const double* h_params, int n_params, double t,
                                  const double* h_x, int64_t xdim, const double* h_y, int64_t ydim,
                                  double* d_u, double* d_v, void* d_workspace,
                                  size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EvalParams P;
  memset(&P, 0, sizeof(P));
  if (!flow_prepare(flow_id, h_params, n_params, &P.fc)) {
    dlb_set_error("unknown flow id %d or wrong parameter count %d", flow_id, n_params);
    return DLB_ERR_ARG;
  }
  if (k_flow_info[flow_id].dim != 2) {
    dlb_set_error("dlb_flow_eval_grid: flow %d is not a 2-D flow", flow_id);
    return DLB_ERR_ARG;
  }
  if (!h_x || !h_y || !d_u || !d_v || !d_workspace || xdim < 0 || ydim < 0) {
    dlb_set_error("dlb_flow_eval_grid: bad argument");
    return DLB_ERR_ARG;
  }
  if (workspace_bytes < dlb_workspace_bytes(xdim, ydim)) {
    dlb_set_error("workspace too small");
    return DLB_ERR_WORKSPACE;
  }
  if (xdim == 0 || ydim == 0) return DLB_OK;
  double* d_x = (double*)((char*)d_workspace + 256);
  double* d_y = d_x + xdim;
  dlb_axes_cache_note_layout(d_workspace, xdim, ydim);
  DLB_CUDA_CHECK(cudaMemcpyAsync(d_x, h_x, sizeof(double) * xdim, cudaMemcpyHostToDevice, stream));
  DLB_CUDA_CHECK(cudaMemcpyAsync(d_y, h_y, sizeof(double) * ydim, cudaMemcpyHostToDevice, stream));
  P.t = t;
  P.px = d_x;
  P.py = d_y;
  P.n = (long long)xdim * ydim;
  P.xdim = xdim;
  P.u = d_u;
  P.v = d_v;
  int rc = DLB_ERR_ARG;
  DLB_DISPATCH_FLOW(flow_id, {
    if constexpr (ID < DLB_FLOW_ABC) rc = launch_eval<ID, true>(P, stream);
  });
  return rc;
}
