/* dynlab_b200 — C-ABI of the B200-native FTLE / Eulerian hot path.
 *
 * The reference (hokiepete/dynlab) is pure Python and has no FFI of its own; its
 * boundary for this path is the Python class API.  These entry points
 * are what a binding for that path would call — one per reference hot loop — and are
 * what dynlab_b200's Python host (ctypes, dynlab_b200/_native.py) calls.
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success or a negative DLB_ERR_*;
 *     dlb_last_error() gives the message (thread-local).  Nothing throws, nothing owns
 *     caller memory.
 *   - pointers named d_* are DEVICE pointers on the current CUDA device; pointers named
 *     h_* are HOST pointers (coordinate vectors and parameters: O(xdim+ydim) data that the
 *     library inspects on the host, e.g. numpy's uniform-spacing test).
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream).  All work is
 *     enqueued on it; the functions do not synchronise unless stated.
 *   - all floating point data is IEEE binary64, C order (x fastest).
 *   - a flow is named by (flow_id, params[]) where params are the Python signature's keyword
 *     parameters of dynlab.flows.<name> in declaration order (R:src/dynlab/flows.py).
 */
#ifndef DYNLAB_B200_H
#define DYNLAB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DLB_ABI_VERSION 2

enum dlb_error {
  DLB_OK = 0,
  DLB_ERR_ARG = -1,       /* bad argument (flow id, dimensions, edge_order, null pointer) */
  DLB_ERR_CUDA = -2,      /* a CUDA runtime call failed */
  DLB_ERR_TOO_SMALL = -3, /* axis shorter than edge_order+1 (numpy raises ValueError) */
  DLB_ERR_WORKSPACE = -4  /* workspace too small */
};

/* R:src/dynlab/flows.py — the 15 analytic flows; ids are stable ABI. */
enum dlb_flow {
  DLB_FLOW_DOUBLE_GYRE = 0,             /* flows.py:13-33   params A, w, e */
  DLB_FLOW_AUTONOMOUS_DOUBLE_GYRE = 1,  /* flows.py:36-48   */
  DLB_FLOW_BICKLEY_JET = 2,             /* flows.py:51-96   params U0,L,re,A2,A3,c2,c3,k2,k3,scale */
  DLB_FLOW_GLIDER = 3,                  /* flows.py:99-113  */
  DLB_FLOW_HILLS_VORTEX = 4,            /* flows.py:116-128 */
  DLB_FLOW_SIGMOIDAL = 5,               /* flows.py:131-142 */
  DLB_FLOW_AUTONOMOUS_DUFFING = 6,      /* flows.py:145-157 */
  DLB_FLOW_DUFFING = 7,                 /* flows.py:160-176 params A, gamma, omega */
  DLB_FLOW_PENDULUM = 8,                /* flows.py:179-195 params A, gamma, omega */
  DLB_FLOW_HURRICANE = 9,               /* flows.py:198-221 params eye_alpha, eye_omega, eye_amplitude */
  DLB_FLOW_VAN_DER_POL = 10,            /* flows.py:224-241 params A, mu, gamma, omega */
  DLB_FLOW_BEAD = 11,                   /* flows.py:244-262 params epsilon, gamma */
  DLB_FLOW_LOTKA_VOLTERRA = 12,         /* flows.py:265-284 params alpha, beta, gamma, delta */
  DLB_FLOW_ABC = 13,                    /* flows.py:287-307 params ABC_Amplitude, A, B, C (3-D) */
  DLB_FLOW_LORENZ = 14,                 /* flows.py:310-331 params sigma, rho, beta (3-D) */
  DLB_FLOW_COUNT = 15
};

/* Eulerian stencil modes (kernel K3). */
enum dlb_euler_mode {
  DLB_EULER_S_MIN = 0,  /* AttractionRate: smallest eigenvalue of S   R:eulerian.py:47-59, :90 */
  DLB_EULER_S_MAX = 1,  /* RepulsionRate: largest eigenvalue of S     R:eulerian.py:120 */
  DLB_EULER_RATE = 2,   /* TrajectoryRepulsionRate  rho-dot           R:eulerian.py:159-165 */
  DLB_EULER_RATIO = 3   /* TrajectoryRepulsionRatio nu-dot            R:eulerian.py:167-171 */
};

/* Eigenvector output conventions (SURVEY §8a "Eigen conventions"). */
enum dlb_xi_mode {
  DLB_XI_NONE = 0,
  DLB_XI_LAPACK = 1, /* sign/order as LAPACK dgeev (np.linalg.eig) returns it */
  DLB_XI_FORCED = 2  /* then force_eigenvectors2D, R:src/dynlab/utils.py:41-45 */
};

/* Counters written by the integrator kernels (device array of 8 uint64). */
enum dlb_counter {
  DLB_CNT_NFEV = 0,     /* total RHS evaluations = 2*seeds + 6*attempts (scipy's nfev summed) */
  DLB_CNT_ACCEPTED = 1, /* accepted steps */
  DLB_CNT_REJECTED = 2, /* rejected attempts */
  DLB_CNT_FAILED = 3,   /* seeds that ended with scipy status -1 (step size too small) */
  DLB_CNT_SEEDS = 4,    /* seeds integrated */
  DLB_CNT_WORDS = 8
};

int dlb_abi_version(void);
const char* dlb_last_error(void);

/* Flow registry. */
int dlb_flow_dim(int flow_id);        /* 2 or 3; <0 if unknown */
int dlb_flow_nparams(int flow_id);    /* number of keyword parameters */
const char* dlb_flow_name(int flow_id);

/* Bytes of device workspace the grid entry points need for a (xdim, ydim) grid.
 * The workspace is scratch owned by the caller but written only by this library: the stencil
 * entry points keep their per-axis coefficient arrays in it and skip the rebuild + upload when
 * the same (workspace, x, y, edge_order) come again, so the caller must not modify it between
 * calls (freeing it is fine). */
size_t dlb_workspace_bytes(int64_t xdim, int64_t ydim);

/* K0 — evaluate a flow on points or on meshgrid(x, y).
 * Replaces `f(t, np.meshgrid(x, y))` in EulerianDiagnostic2D.compute
 * (R:src/dynlab/diagnostics/_base_classes.py:101-102).
 * points: d_px/d_py/d_pz are n-element device arrays (d_pz NULL for 2-D flows).
 * grid  : point (i, j) = (x[j], y[i]); outputs are [ydim][xdim]. */
int dlb_flow_eval_points(int flow_id, const double* h_params, int n_params, double t,
                         const double* d_px, const double* d_py, const double* d_pz, int64_t n,
                         double* d_u, double* d_v, double* d_w, void* stream);
int dlb_flow_eval_grid(int flow_id, const double* h_params, int n_params, double t,
                       const double* h_x, int64_t xdim, const double* h_y, int64_t ydim,
                       double* d_u, double* d_v, void* d_workspace, size_t workspace_bytes,
                       void* stream);

/* K1 — flow map of a row slab: for every node (x[j], y[row0+i]), i<rows, the endpoint of the
 * trajectory integrated from t0 to tf with scipy-RK45 semantics.
 * Replaces LagrangianDiagnostic2D.compute's seed loop + integrator
 * (R:src/dynlab/diagnostics/_base_classes.py:175-189, R:src/dynlab/utils.py:6-25 with the
 * RK45 semantics of scipy/integrate/_ivp/rk.py:14-176, common.py:44-134, base.py:164-210).
 * d_flow_map: [rows][xdim][2] (Phi_x, Phi_y).  max_step = INFINITY for none; first_step = 0
 * to select the initial step as scipy does.  Optional per-seed outputs (NULL to skip):
 * d_nfev [rows][xdim] uint32, d_status [rows][xdim] int8 (0 ok, -1 step size too small).
 * d_counters: DLB_CNT_WORDS uint64, zeroed by this call. */
int dlb_flowmap_rk45(int flow_id, const double* h_params, int n_params,
                     const double* h_x, int64_t xdim, const double* h_y, int64_t row0, int64_t rows,
                     double t0, double tf, double rtol, double atol, double max_step,
                     double first_step, double* d_flow_map, uint32_t* d_nfev, int8_t* d_status,
                     uint64_t* d_counters, void* d_workspace, size_t workspace_bytes,
                     void* stream);

/* Same kernel with the coordinate vectors already resident on the device (d_x[xdim], d_y[at
 * least row0+rows]): nothing is uploaded, so a launch is two 64-byte memsets and the kernel.
 * Chunked / multi-GPU pipelines call this once per row chunk; an upload per launch would queue
 * behind the peer copies on the copy engines.  The workspace only holds the 256-byte work
 * counter here.  Only the seeds of rows [stat_row0, stat_row0 + stat_rows) (a sub-range of the
 * integrated rows; pass row0, rows for all) enter d_counters: a slab that integrates its halo
 * rows redundantly reports the same totals as a single-device run. */
int dlb_flowmap_rk45_resident(int flow_id, const double* h_params, int n_params,
                              const double* d_x, int64_t xdim, const double* d_y, int64_t row0,
                              int64_t rows, double t0, double tf, double rtol, double atol,
                              double max_step, double first_step, double* d_flow_map,
                              uint32_t* d_nfev, int8_t* d_status, uint64_t* d_counters,
                              int64_t stat_row0, int64_t stat_rows,
                              void* d_workspace, size_t workspace_bytes, void* stream);

/* K1 (generic) — endpoints of m seeds of dimension n_dim (2 or 3; must match the flow).
 * d_seeds/d_out: [m][n_dim].  h_atol: n_dim values.  Same semantics as above; this is the
 * entry the integrator-protocol wrapper `integrator(f, t, Y0)[-1, :]` maps to
 * (R:src/dynlab/diagnostics/_base_classes.py:180-182). */
int dlb_rk45_endpoints(int flow_id, const double* h_params, int n_params, int n_dim,
                       const double* d_seeds, int64_t m, double t0, double tf, double rtol,
                       const double* h_atol, double max_step, double first_step, double* d_out,
                       uint32_t* d_nfev, int8_t* d_status, uint64_t* d_counters,
                       void* d_workspace, size_t workspace_bytes, void* stream);

/* K2 — FTLE field from a flow map: np.gradient (both branches, edge_order 1|2), Cauchy-Green
 * tensor, largest eigenvalue, log.  Replaces FTLE.compute / LCS.compute after the flow map
 * (R:src/dynlab/diagnostics/lagrangian.py:50-78, :130-167; numpy/lib/_function_base_impl.py
 * :1246-1384).
 * The caller holds `nloc` consecutive rows of the global [ydim][xdim] grid starting at global
 * row `grow0` (own rows plus halo rows) in d_flow_map [nloc][xdim][2]; the call produces
 * global rows [out_row0, out_row0+out_rows) into d_ftle [out_rows][xdim] (and d_xi
 * [out_rows][xdim][2] when xi_mode != DLB_XI_NONE).  Every row the stencil of an output row
 * touches must be inside [grow0, grow0+nloc).  t_span = |tf - t0|. */
int dlb_ftle_stencil(const double* d_flow_map, int64_t nloc, int64_t grow0,
                     const double* h_x, int64_t xdim, const double* h_y, int64_t ydim,
                     int edge_order, double t_span, int64_t out_row0, int64_t out_rows,
                     double* d_ftle, double* d_xi, int xi_mode,
                     void* d_workspace, size_t workspace_bytes, void* stream);

/* K3 — Eulerian stencil + eigen: velocity gradients, rate-of-strain tensor S and one of the
 * dlb_euler_mode fields.  Replaces EulerianDiagnostic2D.compute's gradients and the per-point
 * loops of _AttractionRepulsionRate / _TrajectoryRepulsionRateRatio / iLES
 * (R:src/dynlab/diagnostics/_base_classes.py:110-111, R:src/dynlab/diagnostics/eulerian.py
 * :47-59, :159-195, :317-343).  Same slab geometry as dlb_ftle_stencil; d_u/d_v are
 * [nloc][xdim].  d_mask ([out_rows][xdim] uint8, 1 = masked) is written for RATE/RATIO where
 * u*u+v*v == 0 (R:eulerian.py:188-192) and may be NULL for the S modes.  `negate` != 0 flips
 * the sign of the field (iLES kind='attracting', R:eulerian.py:339-340). */
int dlb_euler_stencil(int mode, const double* d_u, const double* d_v, int64_t nloc, int64_t grow0,
                      const double* h_x, int64_t xdim, const double* h_y, int64_t ydim,
                      int edge_order, int64_t out_row0, int64_t out_rows, int negate,
                      double* d_field, uint8_t* d_mask, double* d_xi, int xi_mode,
                      void* d_workspace, size_t workspace_bytes, void* stream);

/* K0 fused into K3 ("8 B/pt when f, t is supplied and K0 is fused") — the same
 * fields as dlb_euler_stencil for an analytic flow at time t on meshgrid(x, y)
 * (R:src/dynlab/diagnostics/_base_classes.py:101-102 followed by R:eulerian.py:47-59 / :179-195 /
 * :317-333) without writing or reading (u, v): the velocity of a node is produced in registers
 * where the stencil consumes it (x / y separable flows: transcendentals once per column / row of
 * a tile: double gyre, autonomous double gyre, Bickley jet).  Velocities are bit-identical to
 * dlb_flow_eval_grid's, so the output equals dlb_flow_eval_grid + dlb_euler_stencil exactly.  Rows [out_row0, out_row0 + out_rows) of any
 * rank's slab can be produced without a halo.  2-D flows only. */
int dlb_euler_stencil_flow(int mode, int flow_id, const double* h_params, int n_params, double t,
                           const double* h_x, int64_t xdim, const double* h_y, int64_t ydim,
                           int edge_order, int64_t out_row0, int64_t out_rows, int negate,
                           double* d_field, uint8_t* d_mask, double* d_xi, int xi_mode,
                           void* d_workspace, size_t workspace_bytes, void* stream);

/* Ridge field — gradient + Hessian of a scalar field, directional derivative
 * grad(f).xi and concavity xi^T H xi, with the reference's masks (field <= 0, concavity >= 0,
 * field <= threshold when use_threshold).  Replaces RidgeExtractor2D.extract_ridges up to the
 * contour call (R:src/dynlab/diagnostics/_base_classes.py:216-268).  Whole grid on one device:
 * d_field [ydim][xdim], d_xi [ydim][xdim][2] -> d_dd, d_conc [ydim][xdim], d_mask uint8.
 * d_scratch: 2*ydim*xdim doubles (first derivatives). */
int dlb_ridge_field(const double* d_field, const double* d_xi, const double* h_x, int64_t xdim,
                    const double* h_y, int64_t ydim, int edge_order, int use_threshold,
                    double threshold, double* d_dd, double* d_conc, uint8_t* d_mask,
                    double* d_scratch, void* d_workspace, size_t workspace_bytes, void* stream);

/* Contour segments — the O(cells) part of contourpy's
 * `contour_generator(x, y, z).lines(level)` (R:src/dynlab/diagnostics/_base_classes.py:270) on the
 * device: for every grid cell the `level` contour crosses, records (cell index, start-edge key,
 * end-edge key, start vertex, end vertex) with contourpy's conventions (upper iff z > level,
 * corner masking, saddles by the cell-centre mean, upper region on the left).  d_mask ([ydim]
 * [xdim] uint8, 1 = masked) may be NULL.  Records are appended in arbitrary order to d_cell /
 * d_key_a / d_key_b ([max_segments] int64; an edge key is lo_node * xdim*ydim + hi_node) and d_pts
 * ([max_segments][4]: sx, sy, ex, ey).  *d_count receives the number of segments found; if it
 * exceeds max_segments only the first max_segments slots were written (call again with more
 * room).  Linking segments into polylines is left to the host (O(ridge length)). */
int dlb_contour_segments(const double* d_z, const uint8_t* d_mask, const double* h_x, int64_t xdim,
                         const double* h_y, int64_t ydim, double level, int64_t* d_cell,
                         int64_t* d_key_a, int64_t* d_key_b, double* d_pts, int64_t max_segments,
                         uint64_t* d_count, void* d_workspace, size_t workspace_bytes,
                         void* stream);

/* Polyline linking (host side) — chains the records of dlb_contour_segments
 * (copied to the host) into the polylines `contour_generator(...).lines(level)` returns
 * (R:src/dynlab/diagnostics/_base_classes.py:270): lines in the scan order of the cell they start
 * in, open lines from where nothing leads in, closed loops repeating their first vertex.  Pure
 * host code, O(n) apart from small per-row sorts; xdim, ydim: the grid the keys refer to.
 * h_line_start [n + 1] and h_vertices [2 n][2] are caller-allocated;
 * line l is h_vertices[h_line_start[l] .. h_line_start[l + 1]) as (x, y) pairs. */
int dlb_link_segments(const int64_t* h_cell, const int64_t* h_key_a, const int64_t* h_key_b,
                      const double* h_pts, int64_t n, int64_t xdim, int64_t ydim,
                      int64_t* h_line_start, double* h_vertices, int64_t* h_n_lines);

/* Order statistics — the two neighbouring sorted values `np.percentile(field,
 * percentile)` interpolates between (R:src/dynlab/diagnostics/_base_classes.py:266-268), found on
 * the device by an 8-pass radix select instead of a full-field D2H copy + host partition.
 * d_values [n] (any layout, read flat); h_ranks [nranks] (1 or 2 zero-based ranks in the sorted
 * order, ascending); d_out [nranks] receives the selected elements (exact input values; -0.0 sorts
 * before +0.0); *d_nan_count the number of NaNs in d_values (numpy returns NaN if it is non-zero;
 * the ranks are then meaningless).  d_scratch: dlb_select_scratch_bytes() bytes. */
size_t dlb_select_scratch_bytes(void);
int dlb_order_statistics(const double* d_values, int64_t n, const int64_t* h_ranks, int nranks,
                         double* d_out, uint64_t* d_nan_count, void* d_scratch,
                         size_t scratch_bytes, void* stream);

/* ---- multi-GPU helpers (SURVEY §8b "multi-GPU helpers", §8e) -------------------------------
 * The reference's only parallel construct is a process pool whose per-seed results `p.map`
 * collects into one array (R:src/dynlab/diagnostics/_base_classes.py:139-148, :178-187).  Here
 * the seed grid is split into row slabs over the GPUs of one box; these entry points are the
 * collection step and the stencil's 1-row halo, done by the GPUs themselves over NVLink peer
 * memory: a copy engine writes the rows into the receivers' buffers while the SMs go on
 * integrating, then one warp releases a generation flag in each receiver's memory
 * (st.release.sys); the receiver's stream waits on its own flags with dlb_wait_flags.
 * "Symmetric" buffers are plain device allocations every participating GPU can address:
 * inside one process after dlb_peer_enable, between the one-process-per-GPU ranks through the
 * 64-byte handle of dlb_ipc_export opened with dlb_ipc_open (CUDA IPC; the handle bytes travel
 * by whatever channel the host has - dynlab_b200 uses torch.distributed object collectives). */
#define DLB_MAX_PEERS 16
#define DLB_IPC_HANDLE_BYTES 64
int dlb_device_count(int* n);
/* Let `device` address `peer`'s memory (idempotent).  Single-process multi-device use. */
int dlb_peer_enable(int device, int peer);
/* Zero-filled allocation on the current device that can be exported with dlb_ipc_export. */
int dlb_peer_alloc(size_t bytes, void** d_ptr);
int dlb_peer_free(void* d_ptr);
int dlb_ipc_export(const void* d_ptr, unsigned char handle[DLB_IPC_HANDLE_BYTES]);
/* Map another process's dlb_peer_alloc block into this process (peer access enabled lazily). */
int dlb_ipc_open(const unsigned char handle[DLB_IPC_HANDLE_BYTES], void** d_ptr);
int dlb_ipc_close(void* d_ptr);
/* Stream-ordered copy between any two addresses this process can address (own device, peer
 * device, IPC-mapped peer, pinned host): the one-sided read used to collect `flow_map` rows. */
int dlb_peer_copy(void* d_dst, const void* d_src, size_t bytes, void* stream);
/* Row gather, push side: copy [d_src, d_src + bytes) to each d_dst[i] (i < n_dst <=
 * DLB_MAX_PEERS; NULL or == d_src entries are skipped), then publish `flag_value` at every
 * non-NULL d_flags[i] with release semantics at system scope.  Flags are uint64 generation
 * counters in the receivers' memory and must only grow. */
int dlb_gather_rows(const void* d_src, size_t bytes, void* const* d_dst, int n_dst,
                    uint64_t* const* d_flags, uint64_t flag_value, void* stream);
/* 1-row halo exchange, push side: row `first_own` of d_local ([nloc][row_bytes]) goes to the
 * upper neighbour's bottom halo slot d_up_slot, row `last_own` to the lower neighbour's top halo
 * slot d_down_slot (NULL = no such neighbour), then `flag_value` is released at d_up_flag /
 * d_down_flag (in the neighbours' memory). */
int dlb_halo_exchange(const void* d_local, size_t row_bytes, int64_t first_own, int64_t last_own,
                      void* d_up_slot, void* d_down_slot, uint64_t* d_up_flag,
                      uint64_t* d_down_flag, uint64_t flag_value, void* stream);
/* K2 fused with the row gather: dlb_ftle_stencil (xi_mode NONE) whose epilogue stores every FTLE
 * value into d_ftle AND at the same offset of the n_peer buffers d_peer_ftle[k] (the other GPUs'
 * copies of these rows, i.e. their field base + out_row0 * xdim; NULL or == d_ftle entries are
 * skipped) — NVLink stores issued by the SMs as the values are produced, no copy afterwards —
 * and then releases `flag_value` at the n_flags flags like dlb_gather_rows.  Used for the last
 * rows of a slab, where a copy engine pass after the kernel would be exposed latency. */
int dlb_ftle_stencil_gather(const double* d_flow_map, int64_t nloc, int64_t grow0,
                            const double* h_x, int64_t xdim, const double* h_y, int64_t ydim,
                            int edge_order, double t_span, int64_t out_row0, int64_t out_rows,
                            double* d_ftle, double* const* d_peer_ftle, int n_peer,
                            uint64_t* const* d_flags, int n_flags, uint64_t flag_value,
                            void* d_workspace, size_t workspace_bytes, void* stream);
/* Receive side of both: enqueue a one-warp kernel that polls d_flags[0..n) (n <= 32, this
 * device's memory) until every one is >= min_value (acquire at system scope), so later work
 * on `stream` sees the rows.  Bounded: after timeout_s seconds it gives up and stores
 * 1 + (index of a late flag) to *d_timed_out (int, may be NULL) instead of hanging the GPU. */
int dlb_wait_flags(const uint64_t* d_flags, int n, uint64_t min_value, double timeout_s,
                   int* d_timed_out, void* stream);

/* FP64 roofline denominator: runs a register-resident DFMA chain kernel and returns the
 * sustained FP64 throughput in FLOP/s (FMA = 2) through *flops_per_s.  Synchronises. */
int dlb_dfma_peak(int iters, double* flops_per_s, double* elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* DYNLAB_B200_H */
