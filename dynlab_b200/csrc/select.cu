// Order statistics of a field on the device (the `percentile` argument of
// LCS / iLES).
//
// The reference thresholds its ridge mask with `np.percentile(field, percentile)`
// (R:src/dynlab/diagnostics/_base_classes.py:266-268).  numpy's default 'linear' method needs two
// neighbouring order statistics a[lo], a[lo+1] of the flattened field and interpolates between
// them on the host; finding them with np.partition means a full-field D2H copy plus an O(n)
// introselect on one core — 0.1-0.5 s per slice against 40 ms for the FTLE field itself.  Here the
// two values are found where the field lives by an 8-pass most-significant-digit radix select:
// doubles are mapped to order-preserving 64-bit keys, each pass histograms one byte of the keys
// that still match the prefix found so far (256 bins, privatised per CTA in shared memory), and a
// one-warp kernel picks the bin that holds the wanted rank.  The selected values are exact
// elements of the input, so the host-side interpolation reproduces numpy bit for bit.  HBM bound:
// 8 reads of the field (0.1 ms for 8.4 M values).
#include "common.cuh"

namespace {

constexpr int SEL_BLOCK = 256;
constexpr int SEL_MAXQ = 2;

// order-preserving map double -> uint64 (negative values reversed, positives above them)
__device__ __forceinline__ unsigned long long key_of(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double value_of(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

struct SelState {                        // lives in d_scratch
  unsigned long long prefix[SEL_MAXQ];   // key bits decided so far
  unsigned long long rank[SEL_MAXQ];     // rank still to find among the keys matching the prefix
  unsigned long long nan_count;
  unsigned long long pad[3];
  unsigned int hist[SEL_MAXQ][256];
};

__global__ void sel_init_kernel(SelState* S, long long k0, long long k1) {
  const int t = threadIdx.x;
  if (t == 0) {
    S->prefix[0] = S->prefix[1] = 0;
    S->rank[0] = (unsigned long long)k0;
    S->rank[1] = (unsigned long long)k1;
    S->nan_count = 0;
  }
  for (int i = t; i < SEL_MAXQ * 256; i += blockDim.x) (&S->hist[0][0])[i] = 0;
}

template <int NQ>
__global__ void __launch_bounds__(SEL_BLOCK) sel_hist_kernel(const double* __restrict__ v,
                                                             long long n, int pass, SelState* S) {
  __shared__ unsigned int h[NQ][256];
  __shared__ unsigned int nans;
  for (int i = threadIdx.x; i < NQ * 256; i += SEL_BLOCK) (&h[0][0])[i] = 0;
  if (threadIdx.x == 0) nans = 0;
  __syncthreads();
  const int shift = 56 - 8 * pass;
  unsigned long long pre[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) pre[q] = S->prefix[q];
  // bits above the current digit must equal the prefix (pass 0: everything matches)
  const unsigned long long himask = pass == 0 ? 0ull : (~0ull << (shift + 8));
  unsigned int my_nans = 0;
  for (long long i = (long long)blockIdx.x * SEL_BLOCK + threadIdx.x; i < n;
       i += (long long)gridDim.x * SEL_BLOCK) {
    const double x = __ldg(v + i);
    if (pass == 0 && x != x) ++my_nans;
    const unsigned long long k = key_of(x);
    const unsigned d = (unsigned)(k >> shift) & 255u;
#pragma unroll
    for (int q = 0; q < NQ; ++q)
      if (((k ^ pre[q]) & himask) == 0) atomicAdd(&h[q][d], 1u);
  }
  if (my_nans) atomicAdd(&nans, my_nans);
  __syncthreads();
  for (int i = threadIdx.x; i < NQ * 256; i += SEL_BLOCK) {
    const unsigned c = (&h[0][0])[i];
    if (c) atomicAdd(&(&S->hist[0][0])[i], c);
  }
  if (threadIdx.x == 0 && nans) atomicAdd(&S->nan_count, (unsigned long long)nans);
}

// one warp per query: find the bin holding the wanted rank, extend the prefix, clear the bins
__global__ void sel_pick_kernel(SelState* S, int nq, int pass, double* out,
                                unsigned long long* nan_out) {
  const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (q < nq) {
    // lane owns bins [8 lane, 8 lane + 8)
    unsigned c[8];
    unsigned sum = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      c[b] = S->hist[q][8 * lane + b];
      sum += c[b];
    }
    unsigned long long incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    unsigned long long below = incl - sum;
    const unsigned long long r = S->rank[q];
    const int shift = 56 - 8 * pass;
    __syncwarp();
    if (r >= below && r < incl) {
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        if (r < below + c[b]) {
          const unsigned long long p = S->prefix[q] | ((unsigned long long)(8 * lane + b) << shift);
          S->prefix[q] = p;
          S->rank[q] = r - below;
          if (pass == 7) out[q] = value_of(p);
          break;
        }
        below += c[b];
      }
    }
#pragma unroll
    for (int b = 0; b < 8; ++b) S->hist[q][8 * lane + b] = 0;
  }
  if (threadIdx.x == 0 && pass == 7) *nan_out = S->nan_count;
}

}  // namespace

extern "C" size_t dlb_select_scratch_bytes(void) { return sizeof(SelState); }

extern "C" int dlb_order_statistics(const double* d_values, int64_t n, const int64_t* h_ranks,
                                    int nranks, double* d_out, uint64_t* d_nan_count,
                                    void* d_scratch, size_t scratch_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!d_values || !h_ranks || !d_out || !d_nan_count || !d_scratch || n <= 0 || nranks < 1 ||
      nranks > SEL_MAXQ) {
    dlb_set_error("dlb_order_statistics: bad argument (n >= 1, 1 <= nranks <= 2)");
    return DLB_ERR_ARG;
  }
  if (n >= ((int64_t)1 << 32)) {  // the per-digit histograms are 32-bit counters
    dlb_set_error("dlb_order_statistics: more than 2^32 - 1 values");
    return DLB_ERR_ARG;
  }
  if (scratch_bytes < sizeof(SelState)) {
    dlb_set_error("dlb_order_statistics: scratch too small (dlb_select_scratch_bytes)");
    return DLB_ERR_WORKSPACE;
  }
  for (int q = 0; q < nranks; ++q)
    if (h_ranks[q] < 0 || h_ranks[q] >= n) {
      dlb_set_error("dlb_order_statistics: rank outside [0, n)");
      return DLB_ERR_ARG;
    }
  SelState* S = (SelState*)d_scratch;
  int dev = 0, sms = 0;
  DLB_CUDA_CHECK(cudaGetDevice(&dev));
  DLB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long long blocks = (n + SEL_BLOCK - 1) / SEL_BLOCK;
  if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
  sel_init_kernel<<<1, 256, 0, stream>>>(S, h_ranks[0], h_ranks[nranks - 1]);
  for (int pass = 0; pass < 8; ++pass) {
    if (nranks == 1)
      sel_hist_kernel<1><<<(unsigned)blocks, SEL_BLOCK, 0, stream>>>(d_values, n, pass, S);
    else
      sel_hist_kernel<2><<<(unsigned)blocks, SEL_BLOCK, 0, stream>>>(d_values, n, pass, S);
    sel_pick_kernel<<<1, 32 * SEL_MAXQ, 0, stream>>>(S, nranks, pass, d_out,
                                                     (unsigned long long*)d_nan_count);
  }
  DLB_CUDA_CHECK(cudaGetLastError());
  return DLB_OK;
}
