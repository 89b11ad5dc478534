#!/usr/bin/env python
"""bench.py — FTLE seeds/s on BASELINE config 3 (double gyre 8192x4096, t=(10,0), T=10,
edge_order=2, rtol=atol=1e-8), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A step = one pass of the hot path over the whole seed grid with the library's defaults:
K1 (RK45 flow map of this rank's row slab, cut by integrator work, halo rows integrated locally)
-> K2 (fused stencil / eigen / log) -> the FTLE rows to every rank (peer-memory push behind K1,
the last rows by K2's own epilogue; csrc/peer.cu).
  value : seeds/s with everything resident in HBM (CUDA events, max over ranks).
  e2e   : seeds/s through the public API FTLE().compute(x, y, f, t, ...) with host numpy
          coordinates in and the host FTLE field out on every rank (D2H inside the timed region).
  result_check  : sha256 of the value / e2e fields (all ranks compared) and the C-oracle error
                  of the flow map on the every-64th sample - the same at every N.
  roofline      : K1, FP64-bound: algorithmic FLOP (SURVEY §8d model, from the kernel's exact
                  attempt counter) / K1's launch time per step, against the DFMA peak measured
                  in this run (MEASURED_PEAKS.json has no FP64 figure).
  roofline_hbm  : K2: 24 B/point / launch duration against MEASURED_PEAKS hbm_gbs.
  cpu_baseline  : N = 1: the UNMODIFIED reference (oracle/_ref) FTLE(num_threads=cores) with
                  its default LSODA integrator on a strided sample, in a fresh interpreter;
                  beside it the reference with a scipy RK45 integrator and the plain-C port.
  configs       : driver-timed lines for BASELINE configs[1], [3], [4] (N = 1; the sweep at all N).
  single_process: N > 1: rank 0 alone drives all N GPUs in one process (no torch.distributed).
  nccl_transport_ms_per_step: N > 1: the same step with NCCL carrying halo + gather (A/B).
--impl reference times the reference's own CPU implementation of the path: the unmodified
package from oracle/_ref (kind "reference"; oracle/py_oracle.py, kind "port", only if the copy
is absent) on a bounded strided sample, all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "FTLE seeds/sec at 8192x4096 double gyre T=10"
XDIM, YDIM = 8192, 4096
T_SPAN = (10, 0)
TOL = 1e-8
EDGE_ORDER = 2

# SURVEY §8d FLOP model (FMA = 2, W = 20 flop per transcendental call), double gyre, n = 2
F_RHS = 18 + 5 * 20
F_ATTEMPT = 6 * F_RHS + 64 * 2 + 16 + 20   # 872
F_SEED = 2 * F_RHS + 10 * 2                # 256
BYTES_PER_POINT_K2 = 24
TRAFFIC_FILE = "r02_traffic.json"


def grid(stride: int = 1):
    x = np.linspace(0, 2, XDIM)[::stride]
    y = np.linspace(0, 1, YDIM)[::stride]
    return np.ascontiguousarray(x), np.ascontiguousarray(y)


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6),
                              ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


# ------------------------------------------------------------------------- CPU legs
def cpu_baseline_c_port(target_s: float = 12.0):
    """oracle/c (plain-C restatement of scipy RK45 + np.gradient + closed-form eigen), all host
    cores, on a strided sub-grid of config 3 sized for ~target_s seconds."""
    from oracle import c_oracle as co
    from oracle import py_oracle as po

    cores = host_cores()
    p = po.flow_params("double_gyre")
    x, y = grid(64)
    t0 = time.perf_counter()
    co.flow_map("double_gyre", p, x, y, 10.0, 0.0, TOL, TOL, nthreads=cores)
    rate = x.size * y.size / (time.perf_counter() - t0)
    stride = 64
    while stride > 2 and (XDIM // (stride // 2)) * (YDIM // (stride // 2)) / rate < target_s:
        stride //= 2
    x, y = grid(stride)
    t0 = time.perf_counter()
    fm, nfev, st = co.flow_map("double_gyre", p, x, y, 10.0, 0.0, TOL, TOL, nthreads=cores)
    co.ftle(fm, y, x, 10.0, 0.0, EDGE_ORDER)
    dt = time.perf_counter() - t0
    n = x.size * y.size
    return {"value": n / dt, "unit": "seeds/s", "cores": cores, "kind": "port",
            "sample": f"every {stride}th row and column of the 8192x4096 grid "
                      f"({x.size}x{y.size} = {n} seeds, {dt:.1f} s), oracle/c RK45 port, "
                      f"{cores} threads", "mean_nfev": float(nfev.mean())}


def workload_config(world: int) -> dict:
    """`config` of the JSON line: identical for our arm and the reference arm."""
    return {"workload": "FTLE double_gyre 8192x4096, t=(10,0), edge_order=2, rtol=atol=1e-8 "
                        "(BASELINE configs[2])",
            "grid": [XDIM, YDIM], "t": list(T_SPAN), "edge_order": EDGE_ORDER, "rtol": TOL,
            "atol": TOL, "n_gpus": world,
            "l2": "inputs/outputs (537 MB flow map + 268 MB field per step) exceed the 126 MB "
                  "L2; no flush needed"}


def sample_grid(seeds_target: int):
    """Strided sub-grid of config 3 (same domain, every stride-th row and column) with at least
    ``seeds_target`` seeds."""
    stride = 64
    while stride > 1 and (XDIM // stride) * (YDIM // stride) < seeds_target:
        stride //= 2
    x, y = grid(stride)
    return x, y, (f"every {stride}th row and column of the 8192x4096 grid "
                  f"({x.size}x{y.size} = {x.size * y.size} seeds)")


def reference_ftle(seeds_target: int, integrator: str = "lsoda"):
    """One `FTLE(num_threads=cores).compute(...)` of the reference on a strided sample of config 3.

    With oracle/_ref present (oracle/build_ref.sh) this is the UNMODIFIED reference package:
    `dynlab.diagnostics.FTLE` (R:lagrangian.py:23-78) with its default integrator (scipy LSODA,
    R:utils.py:23-25) or the scipy RK45 wrapper passed through its `integrator=` argument, its own
    process pool over all cores (R:_base_classes.py:139-146), np.gradient and np.linalg.eig loop.
    Without it (kind "port") the same steps restated in oracle/py_oracle.py.
    Returns (seeds/s, seconds, sample description, cores, kind, what)."""
    from oracle import ref_loader

    cores = host_cores()
    x, y, sample = sample_grid(seeds_target)
    if ref_loader.reference_available():
        ref_loader.load_reference()
        from dynlab.diagnostics import FTLE as RefFTLE
        from dynlab.flows import double_gyre as ref_double_gyre

        kw = {} if integrator == "lsoda" else {"integrator": ref_loader.rk45_wrapper}
        t0 = time.perf_counter()
        d = RefFTLE(num_threads=cores, **kw)
        field = d.compute(x, y, ref_double_gyre, T_SPAN, edge_order=EDGE_ORDER, rtol=TOL, atol=TOL)
        dt = time.perf_counter() - t0
        assert field.shape == (y.size, x.size)
        kind = "reference"
        what = ("unmodified hokiepete/dynlab from oracle/_ref: dynlab.diagnostics.FTLE("
                f"num_threads={cores}).compute, "
                + ("default integrator (scipy odeint / LSODA)" if integrator == "lsoda"
                   else "integrator = scipy solve_ivp RK45 wrapper")
                + ", its own process pool + np.gradient + np.linalg.eig loop")
    else:
        from oracle import py_oracle as po

        integ = po.odeint_wrapper if integrator == "lsoda" else po.rk45_wrapper
        t0 = time.perf_counter()
        fm = po.flow_map(x, y, po.FlowSpec("double_gyre"), T_SPAN, integrator=integ,
                         num_threads=cores, rtol=TOL, atol=TOL)
        po.ftle_from_flow_map(fm, x, y, T_SPAN, EDGE_ORDER)
        dt = time.perf_counter() - t0
        kind = "port"
        what = ("oracle/_ref absent (oracle/build_ref.sh was not run where /root/reference "
                "exists): oracle/py_oracle.py restatement, scipy "
                + ("odeint (LSODA)" if integrator == "lsoda" else "solve_ivp RK45")
                + " per seed through a process pool + np.gradient + np.linalg.eig")
    n = x.size * y.size
    return n / dt, dt, sample, cores, kind, what


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = host_cores()
    times, n_seeds = [], 0
    target = args.ref_seeds or max(2048, 600 * cores)
    sample = kind = what = ""
    for i in range(args.warmup + args.steps):
        v, dt, sample, cores, kind, what = reference_ftle(target, args.ref_integrator)
        if i >= args.warmup:
            times.append(dt)
            n_seeds = v * dt
    ms = 1e3 * float(np.mean(times))
    value = n_seeds / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "seeds/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(world),
        "cpu_baseline": {"value": value, "unit": "seeds/s", "cores": cores, "kind": kind,
                         "sample": "each step = " + sample + "; " + what},
        "e2e": {"value": value, "unit": "seeds/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def reference_subprocess(seeds: int, integrator: str, steps: int = 1):
    """The reference arm in a fresh interpreter (its process pool forks; this process holds a
    CUDA context).  Returns the parsed JSON line."""
    env = {k: v for k, v in os.environ.items()
           if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference",
                          "--steps", str(steps), "--warmup", "0", "--ref-seeds", str(seeds),
                          "--ref-integrator", integrator], env=env, check=True,
                         capture_output=True, text=True, timeout=900).stdout
    return json.loads(out.strip().splitlines()[-1])


# ------------------------------------------------------------------------- GPU arm
def sha256_of(a) -> str:
    import hashlib

    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def oracle_sample_error(flow_map_sample, stride=64):
    """max |flow_map - C oracle| on the every-`stride`-th row/column sample of config 3 (the
    oracle leg bench.py is allowed: checker of the result, never the thing measured)."""
    from oracle import c_oracle as co
    from oracle import py_oracle as po

    x, y = grid(stride)
    fm, _, _ = co.flow_map("double_gyre", po.flow_params("double_gyre"), x, y, float(T_SPAN[0]),
                           float(T_SPAN[1]), TOL, TOL, nthreads=host_cores())
    return float(np.abs(np.asarray(flow_map_sample) - fm).max()), int(x.size * y.size)


def extra_configs(eng, hbm_peak, world, rank, quick):
    """Driver-timed lines for BASELINE configs[1], [3], [4] (the headline is configs[2]).  Each
    carries an oracle check on a sample.  CUDA events for kernels, wall clock for API calls."""
    import torch

    from dynlab_b200 import _native as nat
    from dynlab_b200 import flows, sharding
    from dynlab_b200._resolve import resolve_flow
    from dynlab_b200.diagnostics import FTLE, AttractionRate
    from dynlab_b200.sweep import ftle_time_sweep
    from oracle import c_oracle as co
    from oracle import py_oracle as po

    def ev_ms(fn, n=5, warm=2, inner=1):
        """Median of n CUDA-event intervals around `inner` back-to-back calls, per call (sub-ms
        kernels: inner > 1 keeps the queue full so the interval is device time)."""
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(inner):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / inner)
        return float(np.median(ts))

    out = {}
    sharding.set_distributed(False)  # every section below runs on this rank's GPU alone
    try:
        if rank == 0 and not quick:
            # ---- configs[1]: Eulerian stencil + eigen on the Bickley jet, 8192 x 8192
            N = 8192
            _, fid, _, params = resolve_flow(flows.bickley_jet)
            x, y = np.linspace(0, 20000, N), np.linspace(-4000, 4000, N)
            u, v = eng.flow_eval_grid(fid, params, 0.0, x, y)
            c2 = {"grid": [N, N], "edge_order": 2, "hbm_peak_gbs": hbm_peak, "kernels": {}}
            for name, mode, xi, nbytes in (("attraction_rate_s_min", nat.EULER_S_MIN, 0, 24),
                                           ("iles_eigen_s_min_xi", nat.EULER_S_MIN, 2, 40)):
                ms = ev_ms(lambda: eng.euler_stencil(mode, u, v, 0, x, y, 2, 0, N, False, xi), inner=10)
                gbs = nbytes * N * N / ms / 1e6
                c2["kernels"][name] = {"ms": ms, "bytes_per_point": nbytes, "GB_per_s": gbs,
                                       "frac_of_hbm_peak": gbs / hbm_peak}
            ms = ev_ms(lambda: eng.euler_stencil_flow(nat.EULER_S_MIN, fid, params, 0.0, x, y, 2,
                                                      0, N, False, 0), inner=10)
            c2["kernels"]["attraction_rate_fused_flow"] = {"ms": ms, "bytes_per_point": 8,
                                                           "GB_per_s": 8 * N * N / ms / 1e6}
            uh, vh = u.cpu().numpy(), v.cpu().numpy()
            d = AttractionRate()
            walls = []
            for k in range(5):  # two warm-up calls (pinned staging buffers, allocator), median of 3
                t0 = time.perf_counter()
                att = d.compute(x, y, u=uh, v=vh, edge_order=2)
                walls.append((time.perf_counter() - t0) * 1e3)
            c2["e2e_host_uv_ms"] = float(np.median(walls[2:]))
            c2["e2e_host_uv_ms_all_calls"] = walls
            c2["e2e_h2d_bytes"], c2["e2e_d2h_bytes"] = 2 * 8 * N * N, 8 * N * N
            sl = (slice(4000, 4064), slice(3000, 3064))
            ref = po.attraction_repulsion_rate(x[sl[1]], y[sl[0]], u=uh[sl], v=vh[sl], edge_order=2)
            gs = float(np.abs(uh).max() / (x[1] - x[0]))
            c2["oracle_check"] = {"sample": "64x64 interior patch vs oracle/py_oracle (np.gradient + "
                                            "np.linalg.eig), 3844 interior points",
                                  "max_abs_err_over_gradient_scale": float(
                                      np.abs(np.ma.getdata(att)[sl][1:-1, 1:-1]
                                             - np.ma.getdata(ref)[1:-1, 1:-1]).max() / gs)}
            out["config2_eulerian_bickley_8192"] = c2
            del u, v, uh, vh, att, d
            torch.cuda.empty_cache()

            # ---- configs[3]: FTLE on flows whose step counts diverge, 4096 x 4096
            c4 = []
            for name, x, y, t in (
                    ("bead_on_a_rotating_hoop", np.linspace(-1, 1, 4096) * 3, np.linspace(-1, 1, 4096) * 2.5, (0.0, 2.0)),
                    ("duffing_oscillator", np.linspace(-2, 2, 4096), np.linspace(-2, 2, 4096), (0.0, 10.0)),
                    ("duffing_oscillator", np.linspace(-2, 2, 4096), np.linspace(-2, 2, 4096), (10.0, 0.0))):
                f = getattr(flows, name)
                d = FTLE()
                kw = dict(edge_order=2, rtol=TOL, atol=TOL)
                ms = ev_ms(lambda: d.compute_device(x, y, f, t, **kw), n=3, warm=1)
                st = eng.read_counters()
                attempts = (st["nfev"] - 2 * st["seeds"]) / 6.0
                fm_s = d.flow_map[::64, ::64]
                fm_o, _, _ = co.flow_map(name, po.flow_params(name), x[::64], y[::64], t[0], t[1],
                                         TOL, TOL, nthreads=host_cores())
                scale = np.maximum(np.abs(fm_o), 1.0)
                c4.append({"flow": name, "grid": [4096, 4096], "t": list(t), "ms": ms,
                           "seeds_per_s": 4096 * 4096 / ms * 1e3,
                           "attempts_per_s": attempts / ms * 1e3,
                           "mean_nfev_per_seed": st["nfev"] / st["seeds"], "failed_seeds": st["failed"],
                           "oracle_check": {"sample": "every 64th row and column, 4096 seeds, C oracle",
                                            "max_rel_err_flow_map": float((np.abs(fm_s - fm_o) / scale).max())}})
                del d
            out["config4_divergent_step_counts_4096"] = c4
            torch.cuda.empty_cache()
        # ---- configs[4]: 64 t0 slices of the double gyre 4096 x 2048, T = 20, LCS ridge lines
        sharding.set_distributed(True)  # slices are dealt round-robin over the ranks, no comms
        x, y = np.linspace(0, 2, 4096), np.linspace(0, 1, 2048)
        spans = [(k * 10.0 / 64 + 20.0, k * 10.0 / 64) for k in range(64)]
        kw = dict(edge_order=2, percentile=80, ridge=True, to_host=False, rtol=TOL, atol=TOL)
        ftle_time_sweep(x, y, flows.double_gyre, spans[:world], **kw)  # warm-up: one slice per rank
        torch.cuda.synchronize()
        sharding.host_barrier() if world > 1 else None
        t0 = time.perf_counter()
        res = ftle_time_sweep(x, y, flows.double_gyre, spans, **kw)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        c5 = {"grid": [4096, 2048], "slices": 64, "slices_this_rank": len(res), "T": 20.0,
              "percentile": 80, "wall_s_rank0": wall,
              "ridge_lines_rank0": int(sum(len(v[3]) for v in res.values()))}
        if rank == 0:
            k = min(res)
            sharding.set_distributed(False)
            d = FTLE()
            d.compute_device(x, y, flows.double_gyre, spans[k], edge_order=2, rtol=TOL, atol=TOL)
            fm_s = d.flow_map[::32, ::64]
            fm_o, _, _ = co.flow_map("double_gyre", po.flow_params("double_gyre"), x[::64], y[::32],
                                     spans[k][0], spans[k][1], TOL, TOL, nthreads=host_cores())
            dev = float((d.field_device - res[k][0]).abs().max().item())
            c5["oracle_check"] = {"sample": f"slice {k}: every 64th column / 32nd row, 4096 seeds, C oracle",
                                  "max_abs_err_flow_map": float(np.abs(fm_s - fm_o).max()),
                                  "max_abs_diff_sweep_field_vs_FTLE_compute": dev,
                                  "note": "the sweep's K2 also emits the eigenvector (LAPACK-"
                                          "convention eigenvalue expression), FTLE.compute's does "
                                          "not: fields agree to rounding, not bitwise"}
        out["config5_time_sweep_64x4096x2048"] = c5
    finally:
        sharding.set_distributed(True)
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: dynlab_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from dynlab_b200 import peer, sharding
    from dynlab_b200._engine import Engine
    from dynlab_b200.diagnostics import FTLE
    from dynlab_b200.flows import double_gyre

    if world == 1:
        peer.set_devices(None)  # N = 1 means one GPU even if the box shows more
    if args.halo:
        sharding.set_halo_mode(args.halo)
    if args.transport:
        sharding.set_transport(args.transport)

    eng = Engine.get()
    x, y = grid(1)
    n_seeds = XDIM * YDIM
    diag = FTLE()
    kw = dict(edge_order=EDGE_ORDER, rtol=TOL, atol=TOL)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=eng.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def total_launches():
        return sum(e.launches for e in Engine._by_device.values())

    # FP64 peak of this GPU, measured in this run (roofline denominator for K1)
    peak_flops, _ = eng.dfma_peak(2048)
    peak_flops, _ = eng.dfma_peak(8192)

    # ---------------- device-resident loop: `value` ----------------
    def time_value(steps, warmup, collect=False):
        for _ in range(warmup):
            diag.compute_device(x, y, double_gyre, T_SPAN, **kw)
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0 and collect:
            sampler.start()
        eng.profile = {} if collect else None
        l0 = total_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            out = diag.compute_device(x, y, double_gyre, T_SPAN, **kw)
        e1.record()
        barrier()
        clocks = sampler.stop() if (rank == 0 and collect) else None
        ms = max_over_ranks(e0.elapsed_time(e1)) / steps
        prof, eng.profile = eng.profile, None
        return ms, out, prof, total_launches() - l0, clocks

    ms_step, field_dev, prof, launches, clocks = time_value(args.steps, args.warmup, collect=True)
    k1_ms = float(sum(a.elapsed_time(b) for a, b in prof["rk45_flowmap"])) / args.steps
    k2_ms = float(sum(a.elapsed_time(b) for a, b in prof["ftle_stencil"])) / args.steps
    k1_launches = len(prof["rk45_flowmap"]) / args.steps
    diag._collect_stats() if world > 1 else None
    stats = diag.stats if world > 1 else eng.read_counters()  # this rank's seeds, last step
    value = n_seeds / (ms_step * 1e-3)
    group = peer.current_group(n_seeds) if world > 1 else None
    bounds = diag._slab_bounds(group, diag._plan(x, y, double_gyre, T_SPAN, dict(rtol=TOL, atol=TOL))) \
        if group is not None and sharding.transport() == "peer" else None
    my_rows = (bounds[rank + 1] - bounds[rank]) if bounds else sharding.partition(YDIM, world, rank).rows

    # result check of the device path: hash of the whole field + oracle error on a sample
    value_sha = sha256_of(field_dev.cpu().numpy())
    check = {"value_field_sha256": value_sha}
    if rank == 0:
        fm = diag.flow_map  # N > 1: one-sided read of the other ranks' rows (they wait below)
        err, n_s = oracle_sample_error(fm[::64, ::64])
        check.update({"flow_map_max_abs_err_vs_c_oracle": err,
                      "oracle_sample": f"every 64th row and column ({n_s} seeds), oracle/c RK45 port"})
        del fm
    del diag.flow_map
    barrier()

    # the same step with NCCL carrying halo + gather after the kernels (round-1 data path), A/B
    nccl_ms = None
    if world > 1 and sharding.transport() == "peer":
        sharding.set_transport("nccl")
        nccl_ms, f_nccl, _, _, _ = time_value(max(3, args.steps // 2), 2)
        check["nccl_transport_field_sha256"] = sha256_of(f_nccl.cpu().numpy())
        del f_nccl
        sharding.set_transport("peer")
        barrier()

    # ---------------- end-to-end through the public API: `e2e` ----------------
    for _ in range(max(1, min(args.warmup, 2))):
        diag.compute(x, y, double_gyre, T_SPAN, **kw)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        field = diag.compute(x, y, double_gyre, T_SPAN, **kw)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / args.steps
    check["e2e_field_sha256"] = sha256_of(np.ma.getdata(field))
    h2d = (XDIM + my_rows) * 8 + 3 * 8 * (XDIM + YDIM)      # coordinates + stencil coefficients
    d2h = field.size * 8 + 64 * world                          # FTLE field (once) + counters
    if world > 1:  # every rank must hold the same bytes
        mine = torch.tensor([int(check["value_field_sha256"][:15], 16),
                             int(check["e2e_field_sha256"][:15], 16)], dtype=torch.int64, device=eng.device)
        lo, hi = mine.clone(), mine.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        check["all_ranks_hold_identical_fields"] = bool(torch.equal(lo, hi))
    check["e2e_equals_value_field"] = check["e2e_field_sha256"] == value_sha
    del field
    barrier()

    # ---------------- one plain process driving all N GPUs (rank 0; the others wait) ----------------
    single = None
    if world > 1 and not args.no_single_process:
        if rank == 0:
            sharding.set_distributed(False)
            peer.set_devices(list(range(world)))
            try:
                d1 = FTLE()
                for _ in range(2):
                    d1.compute_device(x, y, double_gyre, T_SPAN, **kw)
                for dv in range(world):
                    torch.cuda.synchronize(dv)
                n_sp = max(3, args.steps // 2)
                t0 = time.perf_counter()
                for _ in range(n_sp):
                    fdev = d1.compute_device(x, y, double_gyre, T_SPAN, **kw)
                for dv in range(world):
                    torch.cuda.synchronize(dv)
                sp_dev = (time.perf_counter() - t0) / n_sp
                for _ in range(3):
                    d1.compute(x, y, double_gyre, T_SPAN, **kw)
                t0 = time.perf_counter()
                for _ in range(n_sp):
                    fh = d1.compute(x, y, double_gyre, T_SPAN, **kw)
                sp_e2e = (time.perf_counter() - t0) / n_sp
                single = {"devices": world, "steps": n_sp,
                          "compute_device_ms_per_step": sp_dev * 1e3,
                          "compute_host_field_ms_per_step": sp_e2e * 1e3,
                          "seeds_per_s_e2e": n_seeds / sp_e2e,
                          "field_sha256_device": sha256_of(fdev.cpu().numpy()),
                          "field_sha256_host": sha256_of(np.ma.getdata(fh)),
                          "what": "FTLE().compute in ONE process with no torch.distributed: one "
                                  "host thread drives every GPU (dlb_peer_enable + the same "
                                  "chunk pipeline); wall clock"}
                single["equals_value_field"] = (single["field_sha256_device"] == value_sha
                                                and single["field_sha256_host"] == value_sha)
                del d1, fdev, fh
            except Exception as exc:  # noqa: BLE001 - a side leg must not take the bench line down
                single = {"error": repr(exc)}
            finally:
                peer.set_devices("auto")
                sharding.set_distributed(True)
        sharding.host_barrier(timeout_s=600.0)  # host-side: the idle ranks' GPUs stay idle

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    configs = None
    if not args.no_configs:
        try:
            configs = extra_configs(eng, hbm_peak, world, rank, quick=world > 1)
        except Exception as exc:  # noqa: BLE001 - side legs must not take the bench line down
            configs = {"error": repr(exc)}
        barrier()

    if rank == 0:
        attempts = (stats["nfev"] - 2 * stats["seeds"]) / 6.0
        k1_flop = attempts * F_ATTEMPT + stats["seeds"] * F_SEED
        achieved = k1_flop / (k1_ms * 1e-3)
        traffic = {}
        try:  # DRAM bytes per launch from the committed ncu --set full captures (1-GPU, config 3)
            traffic = json.load(open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)))
        except OSError:
            pass

        def dram(kernel):
            t = traffic.get(kernel)
            return (t["dram_bytes_read"] + t["dram_bytes_write"]) if (t and world == 1) else None

        k2_gbs = BYTES_PER_POINT_K2 * my_rows * XDIM / (k2_ms * 1e-3) / 1e9
        src = f"profiles/{TRAFFIC_FILE} (ncu --set full capture of this kernel at config 3, 1 GPU; not measured in this run)"
        line = {
            "metric": METRIC, "value": value, "unit": "seeds/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(world),
            "parallelism": (f"{world} row slabs cut by integrator work ({bounds}), halo rows: "
                            f"{sharding.halo_mode()}, transport: {sharding.transport()} "
                            "(dlb_gather_rows pushes finished FTLE rows into every rank's field over "
                            "NVLink peer memory while the next row chunk integrates)"
                            if world > 1 else "single GPU"),
            "integrator": {"mean_nfev_per_seed": stats["nfev"] / max(stats["seeds"], 1),
                           "rejected_attempts": stats["rejected"], "failed_seeds": stats["failed"],
                           "seeds_this_rank": stats["seeds"]},
            "result_check": check,
            "e2e": {"value": n_seeds / e2e_s, "unit": "seeds/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3,
                    "note": "public API FTLE().compute with library defaults: host coordinates in, "
                            "full host field out on EVERY rank (pinned D2H of finished rows "
                            "overlapped with K1 chunk by chunk; N>1: each rank streams its own "
                            "rows into one pinned host segment mapped by all ranks)"},
            "gpu_launches": launches,
            "kernels_ms": {"rk45_flowmap": k1_ms, "ftle_stencil": k2_ms,
                           "rk45_launches_per_step": k1_launches},
            "clocks": clocks,
            "roofline": {"bound": "fp64", "achieved": achieved / 1e12, "peak": peak_flops / 1e12,
                         "unit": "TFLOP/s", "frac": achieved / peak_flops,
                         "traffic": dram("rk45_kernel"), "traffic_source": src,
                         "fp64_pipe_active_pct_ncu": (traffic.get("rk45_kernel") or {}).get(
                             "fp64_pipe_active_pct"),
                         "fp64_pipe_active_pct_ncu_source": src,
                         "kernel": "rk45_kernel<double_gyre> (K1)",
                         "note": "algorithmic FLOP = attempts*872 + seeds*256 (SURVEY §8d model, "
                                 "attempts counted by the kernel) / summed K1 launch time per step "
                                 "(CUDA events); peak = DFMA chain measured in this run "
                                 "(dlb_dfma_peak), theoretical 37.2 TFLOP/s at 1965 MHz"},
            "roofline_hbm": {"bound": "hbm", "achieved": k2_gbs, "peak": hbm_peak, "unit": "GB/s",
                             "frac": k2_gbs / hbm_peak,
                             "traffic": dram("stencil_tma_kernel_ftle"), "traffic_source": src,
                             "kernel": "stencil_tma_kernel<FTLE> + stencil_edge_kernel (K2)",
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
        }
        if nccl_ms is not None:
            line["nccl_transport_ms_per_step"] = nccl_ms
        if single is not None:
            line["single_process"] = single
        if configs is not None:
            line["configs"] = configs
        if world == 1 and not args.no_cpu:
            ref = reference_subprocess(max(2048, 600 * host_cores()), "lsoda")
            cb = dict(ref["cpu_baseline"])
            try:
                rk = reference_subprocess(max(2048, 300 * host_cores()), "rk45")
                cb["reference_with_scipy_rk45_integrator"] = {
                    "value": rk["value"], "unit": "seeds/s", "sample": rk["cpu_baseline"]["sample"]}
            except Exception as exc:  # noqa: BLE001 - a baseline leg must not kill the bench line
                cb["reference_with_scipy_rk45_integrator"] = {"error": repr(exc)}
            cb["c_port_rk45"] = cpu_baseline_c_port()
            line["cpu_baseline"] = cb
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--halo", choices=["exchange", "redundant"], default=None,
                    help="multi-GPU: how ranks obtain the halo rows (default: the library's)")
    ap.add_argument("--transport", choices=["peer", "nccl"], default=None,
                    help="multi-GPU: what carries halo + gather (default: the library's)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the lines for BASELINE configs[1], [3], [4]")
    ap.add_argument("--no-single-process", action="store_true",
                    help="N>1: skip the one-process-drives-all-GPUs measurement")
    ap.add_argument("--ref-seeds", type=int, default=0,
                    help="--impl reference: seeds per step (default 600 per host core)")
    ap.add_argument("--ref-integrator", choices=["lsoda", "rk45"], default="lsoda",
                    help="--impl reference: the reference's default integrator or scipy RK45")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
