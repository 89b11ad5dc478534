// Polyline linking of contour segments (host side of SURVEY §8 f3).
//
// dlb_contour_segments leaves one record per crossed cell; chaining them into the polylines
// contourpy's `lines(level)` returns (R:src/dynlab/diagnostics/_base_classes.py:270) is
// O(ridge length) pointer chasing, done here on the host in C++ — the Python loop it replaces
// cost 67 ms per 4096x2048 slice of BASELINE config 5 (2 700 lines, 146 k vertices), twice the
// device time of the whole slice.  Same rules as dynlab_b200/_contour.py:link_segments (kept as
// the test cross-check):
//   * segments are ordered by (cell, start key) = row-major scan order of the owning cell,
//   * a segment's successor is the segment whose start-edge key equals its end-edge key (it lies
//     in the cell on the other side of that grid edge),
//   * open lines start at segments nothing leads into; whatever is left are closed loops, entered
//     at their first segment in that order and closed by repeating the first vertex,
//   * lines are emitted in the order of their first segment.
// Host-only code: no device memory is touched.
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace {
// floor(k / d), k >= 0, without a 64-bit integer division (the divisors are loop
// invariants; idiv costs more than everything else in the loops below)
struct FastDiv {
  int64_t d;
  double inv;
  explicit FastDiv(int64_t d_) : d(d_), inv(1.0 / (double)d_) {}
  inline int64_t operator()(int64_t k) const {
    if (k >> 52) return k / d;  // beyond the exactly representable range: plain division
    int64_t q = (int64_t)((double)k * inv);
    const int64_t r = k - q * d;
    if (r < 0) --q;
    else if (r >= d) ++q;
    return q;
  }
};
}  // namespace

extern "C" int dlb_link_segments(const int64_t* h_cell, const int64_t* h_key_a,
                                 const int64_t* h_key_b, const double* h_pts, int64_t n,
                                 int64_t xdim, int64_t ydim, int64_t* h_line_start,
                                 double* h_vertices, int64_t* h_n_lines) {
  if (n < 0 || !h_n_lines || xdim < 2 || ydim < 2 ||
      (n > 0 && (!h_cell || !h_key_a || !h_key_b || !h_pts || !h_line_start || !h_vertices))) {
    dlb_set_error("dlb_link_segments: bad argument");
    return DLB_ERR_ARG;
  }
  *h_n_lines = 0;
  if (n == 0) return DLB_OK;
  const int64_t w = xdim - 1, rows = ydim - 1, nn = xdim * ydim;
  const FastDiv div_nn(nn), div_x(xdim);
  for (int64_t i = 0; i < n; ++i)
    if (h_cell[i] < 0 || h_cell[i] >= w * rows) {
      dlb_set_error("dlb_link_segments: cell index outside the grid");
      return DLB_ERR_ARG;
    }
  try {
    // ---- scan order of the owning cell: LSD radix sort by cell (stable), then order the (at
    // most two) segments of a cell by start key
    struct Rec {
      int64_t cell, key_a, key_b, idx;
    };
    std::vector<Rec> rec(n), tmp(n);
    for (int64_t i = 0; i < n; ++i) rec[i] = Rec{h_cell[i], h_key_a[i], h_key_b[i], i};
    {
      constexpr int BITS = 11;
      int need = 1;
      while (need < 63 && ((int64_t)1 << need) < w * rows) ++need;
      std::vector<int64_t> cnt((size_t)1 << BITS);
      for (int shift = 0; shift < need; shift += BITS) {
        std::fill(cnt.begin(), cnt.end(), 0);
        for (int64_t i = 0; i < n; ++i) ++cnt[(rec[i].cell >> shift) & ((1 << BITS) - 1)];
        int64_t run = 0;
        for (auto& c : cnt) {
          const int64_t t = c;
          c = run;
          run += t;
        }
        for (int64_t i = 0; i < n; ++i) tmp[cnt[(rec[i].cell >> shift) & ((1 << BITS) - 1)]++] = rec[i];
        rec.swap(tmp);
      }
    }
    auto before = [](const Rec& a, const Rec& b) {
      return a.key_a != b.key_a ? a.key_a < b.key_a : a.idx < b.idx;
    };
    for (int64_t p = 0; p < n;) {
      int64_t e = p + 1;
      while (e < n && rec[e].cell == rec[p].cell) ++e;
      if (e - p == 2) {
        if (before(rec[p + 1], rec[p])) std::swap(rec[p], rec[p + 1]);
      } else if (e - p > 2) {
        std::sort(rec.begin() + p, rec.begin() + e, before);
      }
      p = e;
    }
    std::vector<int64_t> row_start(rows + 1, 0);  // first position of each grid row in `rec`
    {
      int64_t p = 0;
      for (int64_t r = 0; r < rows; ++r) {
        row_start[r] = p;
        const int64_t end_cell = (r + 1) * w;
        while (p < n && rec[p].cell < end_cell) ++p;
      }
      row_start[rows] = n;
    }
    // ---- successor: the segment starting on the grid edge this one ends on lies in the cell on
    // the other side of that edge.  Walking a row in order, the wanted cells of the rows above
    // and below only move forwards: one cursor each, no searching.
    std::vector<int64_t> nxt(n, -1);
    std::vector<char> has_prev(n, 0), used(n, 0);
    for (int64_t r = 0; r < rows; ++r) {
      int64_t cur_up = r > 0 ? row_start[r - 1] : 0, cur_dn = r + 1 < rows ? row_start[r + 1] : n;
      const int64_t end_up = row_start[r], end_dn = r + 1 < rows ? row_start[r + 2] : n;
      for (int64_t p = row_start[r]; p < row_start[r + 1]; ++p) {
        const int64_t c = rec[p].cell, kb = rec[p].key_b;
        const int64_t lo = div_nn(kb), hi = kb - lo * nn;
        const int64_t j = div_x(lo), i = lo - j * xdim;
        int64_t c1 = -1, c2 = -1;
        // The kind of grid edge decides where the neighbour lies - not cn - c, which is
        // ambiguous on a one-cell-wide grid (xdim == 2: the cell above is c - 1).  hi == lo + 1
        // is an x-edge only if lo is not in the last column: for xdim == 2 the diagonal of a
        // corner-masked triangle, (j, 1)-(j+1, 0), has the same key difference.
        bool along_x = false;
        if (hi == lo + 1 && i < xdim - 1) {  // between nodes (j,i) (j,i+1): cells (j-1,i), (j,i)
          along_x = true;
          if (j >= 1) c1 = (j - 1) * w + i;
          if (j < rows) c2 = j * w + i;
        } else if (hi == lo + xdim) {  // edge along y between (j,i) (j+1,i): cells (j,i-1), (j,i)
          if (i >= 1) c1 = j * w + i - 1;
          if (i < w) c2 = j * w + i;
        } else {
          continue;  // diagonal of a corner-masked triangle: the contoured region ends here
        }
        const int64_t cn = (c == c1) ? c2 : c1;
        if (cn < 0) continue;
        int64_t q = -1, q_end = 0;
        if (!along_x && cn > c) {  // right neighbour, same row
          q = p + 1;
          q_end = row_start[r + 1];
          while (q < q_end && rec[q].cell == c) ++q;
        } else if (!along_x) {  // left neighbour, same row
          q = p;
          while (q > row_start[r] && rec[q - 1].cell >= cn) --q;
          q_end = row_start[r + 1];
        } else if (cn < c) {  // row above
          while (cur_up < end_up && rec[cur_up].cell < cn) ++cur_up;
          q = cur_up;
          q_end = end_up;
        } else {  // row below
          while (cur_dn < end_dn && rec[cur_dn].cell < cn) ++cur_dn;
          q = cur_dn;
          q_end = end_dn;
        }
        for (; q < q_end && rec[q].cell == cn; ++q)
          if (rec[q].key_a == kb) {
            nxt[p] = q;
            has_prev[q] = 1;
            break;
          }
      }
    }
    std::vector<int64_t> order(n);
    for (int64_t p = 0; p < n; ++p) order[p] = rec[p].idx;
    std::vector<int64_t> heads;
    for (int64_t p = 0; p < n; ++p)
      if (!has_prev[p]) {
        heads.push_back(p);
        for (int64_t k = p; k != -1 && !used[k]; k = nxt[k]) used[k] = 1;
      }
    for (int64_t p = 0; p < n; ++p)
      if (!used[p]) {  // closed loop, entered at its first segment in scan order
        heads.push_back(p);
        for (int64_t k = p; !used[k]; k = nxt[k]) used[k] = 1;
      }
    std::sort(heads.begin(), heads.end());
    int64_t v = 0, nl = 0;
    for (int64_t s : heads) {
      h_line_start[nl++] = v;
      const double* p0 = h_pts + 4 * order[s];
      h_vertices[2 * v] = p0[0];
      h_vertices[2 * v + 1] = p0[1];
      ++v;
      int64_t k = s, steps = 0;
      do {
        if (++steps > n || v >= 2 * n) {  // malformed input (a chain running into a foreign loop)
          dlb_set_error("dlb_link_segments: segment chain does not terminate");
          return DLB_ERR_ARG;
        }
        const double* pk = h_pts + 4 * order[k];
        h_vertices[2 * v] = pk[2];
        h_vertices[2 * v + 1] = pk[3];
        ++v;
        k = nxt[k];
      } while (k != -1 && k != s);
    }
    h_line_start[nl] = v;
    *h_n_lines = nl;
  } catch (const std::exception& e) {
    dlb_set_error(e.what());
    return DLB_ERR_ARG;
  }
  return DLB_OK;
}
