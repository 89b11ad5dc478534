// Contour segments of a masked field on the device.
//
// The reference ends LCS / iLES with contourpy's `contour_generator(x, y, z).lines(0)`
// (R:src/dynlab/diagnostics/_base_classes.py:270).  Tracing the polylines is O(ridge length) and
// stays on the host (dynlab_b200/_contour.py), but finding the cells a contour crosses is
// O(cells) — 4 s of numpy for one 4096x2048 slice of BASELINE config 5, against 36 ms for the
// FTLE field itself.  This kernel does that part: one thread per cell, contourpy's conventions
//   * a node is "upper" iff z > level; masked or non-finite nodes are excluded,
//   * quads with no masked corner: marching squares, saddles resolved by the mean of the four
//     corners; quads with exactly one masked corner: the triangle of the other three (corner
//     masking); two or more masked corners: nothing,
//   * every segment runs with the upper region on its left, from the cell edge where the
//     counter-clockwise boundary goes upper -> lower to the edge where it goes lower -> upper,
// and appends compact records (cell, start-edge key, end-edge key, both vertices) through a
// warp-aggregated atomic counter.  Vertices are the same linear interpolations the host tracer
// computes (IEEE division), so both paths give bit-identical polylines.
#include "common.cuh"

namespace {

constexpr int CT_BLOCK = 256;
constexpr unsigned FULL = 0xffffffffu;

struct ContourParams {
  const double* z;
  const uint8_t* mask;  // may be null
  const double* x;
  const double* y;
  long long xdim, ydim;
  double level;
  long long* cell;      // [max]
  long long* key_a;     // [max] start-edge key
  long long* key_b;     // [max] end-edge key
  double* pts;          // [max][4]  (sx, sy, ex, ey)
  long long max_segments;
  unsigned long long* count;
};

struct Seg {
  int a0, a1, b0, b1;  // corner indices 0..3 of the cell
};

__global__ void __launch_bounds__(CT_BLOCK) contour_kernel(const ContourParams P) {
  const long long nx = P.xdim, ny = P.ydim;
  const long long ncell = (nx - 1) * (ny - 1);
  const long long nn = nx * ny;
  const unsigned lane = threadIdx.x & 31u;
  for (long long base = (long long)blockIdx.x * CT_BLOCK; base < ncell;
       base += (long long)gridDim.x * CT_BLOCK) {
    const long long c = base + threadIdx.x;
    Seg seg[2];
    int nseg = 0;
    long long node[4] = {0, 0, 0, 0};
    double zc[4] = {0, 0, 0, 0};
    if (c < ncell) {
      const long long j = c / (nx - 1), i = c - j * (nx - 1);
      // corners counter-clockwise: (j,i) (j,i+1) (j+1,i+1) (j+1,i)
      node[0] = j * nx + i;
      node[1] = node[0] + 1;
      node[2] = node[0] + nx + 1;
      node[3] = node[0] + nx;
      bool m[4], up[4];
      int nmask = 0, nup = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        zc[k] = P.z[node[k]];
        m[k] = (P.mask && P.mask[node[k]]) || !isfinite(zc[k]);
        up[k] = !m[k] && (zc[k] > P.level);
        nmask += m[k];
        nup += up[k];
      }
      if (nmask == 0 && nup > 0 && nup < 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int k1 = (k + 1) & 3, k2 = (k + 2) & 3, k3 = (k + 3) & 3;
          if (!(up[k] && !up[k1])) continue;
          Seg s;
          s.a0 = k;
          s.a1 = k1;
          if (nup == 2 && up[k2] && !up[k3]) {  // saddle
            const double zmid = 0.25 * (zc[0] + zc[1] + zc[2] + zc[3]);
            if (zmid > P.level) {  // centre upper: the lower corner k+1 is cut off
              s.b0 = k1;
              s.b1 = k2;
            } else {  // centre lower: the upper corner k is cut off
              s.b0 = k3;
              s.b1 = k;
            }
          } else {  // the only lower -> upper edge
            s.b0 = 0;
            s.b1 = 1;
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (!up[e] && up[(e + 1) & 3]) {
                s.b0 = e;
                s.b1 = (e + 1) & 3;
              }
          }
          if (nseg < 2) seg[nseg++] = s;
        }
      } else if (nmask == 1) {
        int cm = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (m[k]) cm = k;
        const int v[3] = {(cm + 1) & 3, (cm + 2) & 3, (cm + 3) & 3};  // still counter-clockwise
#pragma unroll
        for (int s3 = 0; s3 < 3; ++s3) {
          const int a = v[s3], a1 = v[(s3 + 1) % 3];
          if (!(up[a] && !up[a1])) continue;
#pragma unroll
          for (int e = 0; e < 3; ++e) {
            const int b = v[e], b1 = v[(e + 1) % 3];
            if (!up[b] && up[b1] && nseg < 2) seg[nseg++] = Seg{a, a1, b, b1};
          }
        }
      }
    }
    // warp-aggregated append
    const unsigned has1 = __ballot_sync(FULL, nseg >= 1), has2 = __ballot_sync(FULL, nseg >= 2);
    const int total = __popc(has1) + __popc(has2);
    if (total == 0) continue;
    unsigned long long wbase = 0;
    if (lane == 0) wbase = atomicAdd(P.count, (unsigned long long)total);
    wbase = __shfl_sync(FULL, wbase, 0);
    const unsigned lt = (1u << lane) - 1u;
    long long slot = (long long)wbase + __popc(has1 & lt) + __popc(has2 & lt);
    for (int s = 0; s < nseg; ++s, ++slot) {
      if (slot >= P.max_segments) break;
      const Seg g = seg[s];
      const long long p = node[g.a0], q = node[g.a1], r = node[g.b0], t = node[g.b1];
      const long long loa = p < q ? p : q, hia = p < q ? q : p;
      const long long lob = r < t ? r : t, hib = r < t ? t : r;
      P.cell[slot] = c;
      P.key_a[slot] = loa * nn + hia;
      P.key_b[slot] = lob * nn + hib;
      // vertex on an edge, contourpy's expression (src/base_impl.h, interp(), linear z):
      //   frac = (z1 - level) / (z1 - z0);  v = v0 * frac + v1 * (1 - frac)
      // with point0 the edge's upper node (z > level; on the line's left) and point1 its lower
      // node: (a0, a1) starts upper -> lower, (b0, b1) ends lower -> upper.  Separately rounded
      // (no FMA contraction), like the C++ it restates and like oracle/contour_oracle.py.
      const double fa = __ddiv_rn(__dsub_rn(zc[g.a1], P.level), __dsub_rn(zc[g.a1], zc[g.a0]));
      const double fb = __ddiv_rn(__dsub_rn(zc[g.b0], P.level), __dsub_rn(zc[g.b0], zc[g.b1]));
      const double ga = __dsub_rn(1.0, fa), gb = __dsub_rn(1.0, fb);
      const double xp = P.x[p % nx], xq = P.x[q % nx], yp = P.y[p / nx], yq = P.y[q / nx];
      const double xr = P.x[r % nx], xt = P.x[t % nx], yr = P.y[r / nx], yt = P.y[t / nx];
      double* o = P.pts + 4 * slot;
      o[0] = __dadd_rn(__dmul_rn(xp, fa), __dmul_rn(xq, ga));
      o[1] = __dadd_rn(__dmul_rn(yp, fa), __dmul_rn(yq, ga));
      o[2] = __dadd_rn(__dmul_rn(xt, fb), __dmul_rn(xr, gb));
      o[3] = __dadd_rn(__dmul_rn(yt, fb), __dmul_rn(yr, gb));
    }
  }
}

}  // namespace

extern "C" int dlb_contour_segments(const double* d_z, const uint8_t* d_mask, const double* h_x,
                                    int64_t xdim, const double* h_y, int64_t ydim, double level,
                                    int64_t* d_cell, int64_t* d_key_a, int64_t* d_key_b,
                                    double* d_pts, int64_t max_segments, uint64_t* d_count,
                                    void* d_workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!d_z || !h_x || !h_y || !d_cell || !d_key_a || !d_key_b || !d_pts || !d_count ||
      !d_workspace || max_segments < 0) {
    dlb_set_error("dlb_contour_segments: bad argument");
    return DLB_ERR_ARG;
  }
  if (workspace_bytes < dlb_workspace_bytes(xdim, ydim)) {
    dlb_set_error("workspace too small");
    return DLB_ERR_WORKSPACE;
  }
  DLB_CUDA_CHECK(cudaMemsetAsync(d_count, 0, sizeof(uint64_t), stream));
  if (xdim < 2 || ydim < 2) return DLB_OK;
  double* d_x = (double*)((char*)d_workspace + 256);
  double* d_y = d_x + xdim;
  dlb_axes_cache_note_layout(d_workspace, xdim, ydim);
  DLB_CUDA_CHECK(cudaMemcpyAsync(d_x, h_x, sizeof(double) * xdim, cudaMemcpyHostToDevice, stream));
  DLB_CUDA_CHECK(cudaMemcpyAsync(d_y, h_y, sizeof(double) * ydim, cudaMemcpyHostToDevice, stream));
  ContourParams P;
  P.z = d_z;
  P.mask = d_mask;
  P.x = d_x;
  P.y = d_y;
  P.xdim = xdim;
  P.ydim = ydim;
  P.level = level;
  P.cell = (long long*)d_cell;
  P.key_a = (long long*)d_key_a;
  P.key_b = (long long*)d_key_b;
  P.pts = d_pts;
  P.max_segments = max_segments;
  P.count = (unsigned long long*)d_count;
  int dev = 0, sms = 0;
  DLB_CUDA_CHECK(cudaGetDevice(&dev));
  DLB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long ncell = (long long)(xdim - 1) * (ydim - 1);
  long long blocks = (ncell + CT_BLOCK - 1) / CT_BLOCK;
  if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
  contour_kernel<<<(unsigned)blocks, CT_BLOCK, 0, stream>>>(P);
  DLB_CUDA_CHECK(cudaGetLastError());
  return DLB_OK;
}
