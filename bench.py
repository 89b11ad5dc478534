#!/usr/bin/env python
"""bench.py — FTLE seeds/s on BASELINE config 3 (double gyre 8192x4096, t=(10,0), T=10,
edge_order=2, rtol=atol=1e-8), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A step = one pass of the hot path over the whole seed grid:
K1 (RK45 flow map of this rank's row slab) -> 1-row halo exchange -> K2 (fused stencil/eigen)
-> gather of the FTLE rows.
  value : seeds/s with everything resident in HBM (CUDA events, max over ranks).
  e2e   : seeds/s through the public API FTLE().compute(x, y, f, t, ...) with host numpy
          coordinates in and the host FTLE field out (pinned D2H inside the timed region).
  roofline      : K1, FP64-bound: algorithmic FLOP (SURVEY §8d model, from the kernel's exact
                  attempt counter) / K1's mean launch duration, against the DFMA peak measured
                  in this run (MEASURED_PEAKS.json has no FP64 figure).
  roofline_hbm  : K2, HBM-bound: 24 B/point / launch duration against MEASURED_PEAKS hbm_gbs.
  cpu_baseline  : the plain-C oracle port of the same algorithm on all host cores, bounded
                  strided sample of the same grid (rank 0, N=1 only).
--impl reference times the reference's own CPU implementation of the path (faithful Python
port: scipy LSODA through the seed loop with a process pool over all cores + np.gradient +
np.linalg.eig loop, oracle/py_oracle.py) on a bounded strided sample.
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

    from dynlab_b200 import sharding
    from dynlab_b200._engine import Engine
    from dynlab_b200.diagnostics import FTLE
    from dynlab_b200.flows import double_gyre

    if args.halo:
        sharding.set_halo_mode(args.halo)

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

    # FP64 peak of this GPU, measured in this run (roofline denominator for K1)
    peak_flops, _ = eng.dfma_peak(2048)
    peak_flops, _ = eng.dfma_peak(8192)

    # ---------------- device-resident loop: `value` ----------------
    for _ in range(args.warmup):
        diag.compute_device(x, y, double_gyre, T_SPAN, **kw)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    eng.profile = {}
    launches0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        diag.compute_device(x, y, double_gyre, T_SPAN, **kw)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = eng.launches - launches0
    prof, eng.profile = eng.profile, None
    k1_ms = float(np.mean([a.elapsed_time(b) for a, b in prof["rk45_flowmap"]]))
    k2_ms = float(np.mean([a.elapsed_time(b) for a, b in prof["ftle_stencil"]]))
    stats = eng.read_counters()  # last K1 launch on this rank
    slab = sharding.partition(YDIM, world, rank)
    ms_step = ms_total / args.steps
    value = n_seeds / (ms_step * 1e-3)

    # ---------------- end-to-end through the public API: `e2e` ----------------
    def time_e2e(mode):
        sharding.set_host_result(mode)
        for _ in range(max(1, min(args.warmup, 2))):
            diag.compute(x, y, double_gyre, T_SPAN, **kw)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out = diag.compute(x, y, double_gyre, T_SPAN, **kw)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0) / args.steps
        sharding.set_host_result("all")
        return dt, out

    # N = 1: private pinned copy, overlapped with K1.  N > 1: the ranks of the box share one
    # pinned host segment — every rank returns the full field, each copies only its own rows.
    e2e_s, field = time_e2e("all" if world == 1 else "shared")
    e2e_all_s = e2e_rank0_s = None
    if world > 1:  # same call with a private copy of the gathered field on every rank / rank 0
        e2e_all_s, _ = time_e2e("all")
        e2e_rank0_s, _ = time_e2e("rank0")
    h2d = (XDIM + slab.rows) * 8 + 3 * 8 * (XDIM + YDIM)      # coordinates + stencil coefficients
    d2h = field.size * 8 + 64 * world                          # FTLE field (once) + counters

    if rank == 0:
        attempts = (stats["nfev"] - 2 * stats["seeds"]) / 6.0
        k1_flop = attempts * F_ATTEMPT + stats["seeds"] * F_SEED
        achieved = k1_flop / (k1_ms * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic = {}
        try:  # DRAM bytes per launch from the committed ncu --set full captures (1-GPU, config 3)
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        except OSError:
            pass

        def dram(kernel):
            t = traffic.get(kernel)
            return (t["dram_bytes_read"] + t["dram_bytes_write"]) if (t and world == 1) else None

        k2_gbs = BYTES_PER_POINT_K2 * slab.rows * XDIM / (k2_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "seeds/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "FTLE double_gyre 8192x4096, t=(10,0), edge_order=2, "
                                   "rtol=atol=1e-8 (BASELINE configs[2])",
                       "parallelism": f"row slabs over {world} GPU(s); halo rows: "
                                      f"{sharding.halo_mode()}; all-gather" if world > 1 else "single GPU",
                       "l2": "inputs/outputs (537 MB flow map + 268 MB field per step) exceed "
                             "the 126 MB L2; no flush needed",
                       "integrator": {"mean_nfev_per_seed": stats["nfev"] / max(stats["seeds"], 1),
                                      "rejected_attempts": stats["rejected"],
                                      "failed_seeds": stats["failed"]}},
            "e2e": {"value": n_seeds / e2e_s, "unit": "seeds/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3,
                    "note": "public API FTLE().compute: host coordinates in, full host field out "
                            "on EVERY rank (N=1: pinned D2H overlapped with K1 in row chunks; N>1: "
                            "sharding.set_host_result('shared'), one pinned host segment mapped by "
                            "all ranks, each rank streams its own rows into it through the same "
                            "chunk pipeline, halo rows integrated locally)",
                    "private_copy_per_rank_ms_per_step": None if e2e_all_s is None else e2e_all_s * 1e3,
                    "rank0_host_only_ms_per_step": None if e2e_rank0_s is None else e2e_rank0_s * 1e3},
            "gpu_launches": launches,
            "kernels_ms": {"rk45_flowmap": k1_ms, "ftle_stencil": k2_ms},
            "clocks": clocks,
            "roofline": {"bound": "fp64", "achieved": achieved / 1e12, "peak": peak_flops / 1e12,
                         "unit": "TFLOP/s", "frac": achieved / peak_flops,
                         "traffic": dram("rk45_kernel"),
                         "fp64_pipe_active_pct_ncu": (traffic.get("rk45_kernel") or {}).get(
                             "fp64_pipe_active_pct"),
                         "kernel": "rk45_kernel<double_gyre> (K1)",
                         "note": "algorithmic FLOP = attempts*872 + seeds*256 (SURVEY §8d model, "
                                 "attempts counted by the kernel); peak = DFMA chain measured "
                                 "in this run (dlb_dfma_peak), theoretical 37.2 TFLOP/s at "
                                 "1965 MHz"},
            "roofline_hbm": {"bound": "hbm", "achieved": k2_gbs, "peak": hbm_peak, "unit": "GB/s",
                             "frac": k2_gbs / hbm_peak,
                             "traffic": dram("stencil_interior_kernel_ftle"),
                             "kernel": "stencil_interior_kernel<FTLE> + stencil_edge_kernel (K2)",
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_c_port()
            # the reference's own Python path on the same cores (small strided samples):
            # default LSODA integrator and the like-for-like scipy RK45 integrator
            v_lsoda, _, smp, _ = reference_python_port(4096, "lsoda")
            v_rk45, _, _, _ = reference_python_port(4096, "rk45")
            line["cpu_baseline"]["python_reference_port"] = {
                "lsoda_seeds_per_s": v_lsoda, "rk45_seeds_per_s": v_rk45, "sample": smp,
                "what": "oracle/py_oracle.py: scipy per seed through a process pool over all "
                        "cores + np.gradient + np.linalg.eig (what a dynlab user runs)"}
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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
