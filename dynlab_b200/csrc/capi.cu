// C-ABI plumbing: error string, flow registry queries, FP64 peak probe.
#include <stdarg.h>

#include "flows.cuh"

static thread_local char g_err[512] = "";

void dlb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int dlb_abi_version(void) { return DLB_ABI_VERSION; }
extern "C" const char* dlb_last_error(void) { return g_err; }

extern "C" int dlb_flow_dim(int id) {
  return (id >= 0 && id < DLB_FLOW_COUNT) ? k_flow_info[id].dim : DLB_ERR_ARG;
}
extern "C" int dlb_flow_nparams(int id) {
  return (id >= 0 && id < DLB_FLOW_COUNT) ? k_flow_info[id].nparams : DLB_ERR_ARG;
}
extern "C" const char* dlb_flow_name(int id) {
  return (id >= 0 && id < DLB_FLOW_COUNT) ? k_flow_info[id].name : nullptr;
}

// ---------------------------------------------------------------------------------------
// FP64 roofline denominator.  MEASURED_PEAKS.json holds no FP64 figure, so the bench measures
// one in the same run: 8 independent DFMA chains per thread (enough ILP to cover the DFMA
// latency at 16 warps/SM), 1024 threads/SM-wave, no memory traffic.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_chain_kernel(double* sink, int iters, double a,
                                                         double b) {
  double v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5,
         v6 = v0 + 6, v7 = v0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      v0 = fma(v0, a, b);
      v1 = fma(v1, a, b);
      v2 = fma(v2, a, b);
      v3 = fma(v3, a, b);
      v4 = fma(v4, a, b);
      v5 = fma(v5, a, b);
      v6 = fma(v6, a, b);
      v7 = fma(v7, a, b);
    }
  }
  const double s = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;
  if (s == 12345.678) sink[0] = s;  // never true; keeps the chains alive
}

extern "C" int dlb_dfma_peak(int iters, double* flops_per_s, double* elapsed_ms) {
  if (iters < 1 || !flops_per_s) {
    dlb_set_error("dlb_dfma_peak: bad argument");
    return DLB_ERR_ARG;
  }
  int dev = 0, sms = 0;
  DLB_CUDA_CHECK(cudaGetDevice(&dev));
  DLB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* sink = nullptr;
  DLB_CUDA_CHECK(cudaMalloc(&sink, sizeof(double)));
  cudaEvent_t e0, e1;
  DLB_CUDA_CHECK(cudaEventCreate(&e0));
  DLB_CUDA_CHECK(cudaEventCreate(&e1));
  const int blocks = sms * 8, threads = 256;
  dfma_chain_kernel<<<blocks, threads>>>(sink, 16, 0.999999, 1e-9);  // warm-up
  DLB_CUDA_CHECK(cudaDeviceSynchronize());
  DLB_CUDA_CHECK(cudaEventRecord(e0));
  dfma_chain_kernel<<<blocks, threads>>>(sink, iters, 0.999999, 1e-9);
  DLB_CUDA_CHECK(cudaEventRecord(e1));
  DLB_CUDA_CHECK(cudaEventSynchronize(e1));
  float ms = 0.f;
  DLB_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
  const double fl = 2.0 * 8 * 16 * (double)iters * (double)blocks * threads;
  *flops_per_s = fl / (ms * 1e-3);
  if (elapsed_ms) *elapsed_ms = ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  return DLB_OK;
}
