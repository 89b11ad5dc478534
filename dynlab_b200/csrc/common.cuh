// dynlab_b200 — shared declarations for the sm_100a kernels and the C-ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/dynlab_b200.h"

#define DLB_MAX_CONSTS 16
#define DLB_MAX_DEVICES 64  // per-device tables (power of two)
#define DLB_AXES_ENTRIES 16 // stencil coefficient cache: workspaces kept warm

// Flow parameters after host-side preparation (flows.cuh: flow_prepare).  Passed to
// kernels by value, so every thread reads them from the constant bank.
struct FlowConsts {
  double c[DLB_MAX_CONSTS];
};

void dlb_set_error(const char* fmt, ...);

#define DLB_CUDA_CHECK(expr)                                                        \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      dlb_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,            \
                    cudaGetErrorString(_e));                                        \
      return DLB_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

// Workspace layout helpers (doubles).  The caller owns the workspace (a device buffer of
// dlb_workspace_bytes(xdim, ydim) bytes); kernels use it for coordinate/coefficient arrays
// and the work-stealing counter.
static inline size_t dlb_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// The stencil entry points keep their per-axis coefficient arrays in the caller's workspace
// and skip the rebuild + upload when the same (x, y, edge_order) come again (stencil.cu:
// setup_axes).  Entry points that write the x / y region of a workspace with another grid
// shape call this so that a stale cache entry is never trusted.
void dlb_axes_cache_note_layout(const void* d_workspace, int64_t xdim, int64_t ydim);
