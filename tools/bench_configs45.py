"""Supplementary measurements for BASELINE configs[3] (bead / duffing: divergent step counts) and
configs[4] (64-slice double-gyre time sweep with LCS ridge extraction).  One JSON object on
stdout; CUDA-event times for the device stages, wall clock for whole API calls.
    python tools/bench_configs45.py [--slices 64]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dynlab_b200.flows as flows  # noqa: E402
from dynlab_b200._engine import Engine  # noqa: E402
from dynlab_b200.diagnostics import FTLE, TrajectoryRepulsionRate  # noqa: E402
from dynlab_b200.sweep import ftle_time_sweep  # noqa: E402

TOL = 1e-8


def timed(fn, reps=3):
    """(best wall ms, per-stage device ms of the best repetition)."""
    eng = Engine.get()
    best = None
    for _ in range(reps):
        eng.profile = {}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        prof = {k: round(sum(a.elapsed_time(b) for a, b in v), 3) for k, v in eng.profile.items()}
        eng.profile = None
        if best is None or wall < best[0]:
            best = (wall, prof)
    return round(best[0], 2), best[1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--slices", type=int, default=64)
    args = ap.parse_args()
    eng = Engine.get()
    out = {"tolerance": TOL, "edge_order": 2}

    # ---- configs[3]: FTLE on flows with strongly varying step counts
    cases = [("bead_on_a_rotating_hoop", np.linspace(-1, 1, 4096) * 3, np.linspace(-1, 1, 4096) * 2.5, (0.0, 2.0)),
             ("duffing_oscillator", np.linspace(-2, 2, 4096), np.linspace(-2, 2, 4096), (0.0, 10.0)),
             ("duffing_oscillator", np.linspace(-2, 2, 4096), np.linspace(-2, 2, 4096), (10.0, 0.0))]
    c4 = []
    for name, x, y, t in cases:
        f = getattr(flows, name)
        d = FTLE()
        d.compute_device(x, y, f, t, edge_order=2, rtol=TOL, atol=TOL)  # warm-up
        wall, prof = timed(lambda: d.compute_device(x, y, f, t, edge_order=2, rtol=TOL, atol=TOL))
        st = eng.read_counters()
        seeds = len(x) * len(y)
        attempts = (st["nfev"] - 2 * st["seeds"]) / 6.0
        e2e, _ = timed(lambda: d.compute(x, y, f, t, edge_order=2, rtol=TOL, atol=TOL), reps=3)
        c4.append({"flow": name, "grid": [len(x), len(y)], "t": t, "device_ms": prof,
                   "seeds_per_s": seeds / (prof["rk45_flowmap"] + prof["ftle_stencil"]) * 1e3,
                   "e2e_ms_host_field": e2e, "mean_nfev_per_seed": st["nfev"] / st["seeds"],
                   "attempts_per_s": attempts / prof["rk45_flowmap"] * 1e3,
                   "failed_seeds": st["failed"]})
    x, y = np.linspace(-1, 1, 4096) * 3, np.linspace(-1, 1, 4096) * 2.5
    r = TrajectoryRepulsionRate()
    r.compute(x, y, f=flows.bead_on_a_rotating_hoop, t=0, edge_order=2)
    wall, prof = timed(lambda: r.compute(x, y, f=flows.bead_on_a_rotating_hoop, t=0, edge_order=2))
    out["config4"] = {"ftle": c4, "trajectory_repulsion_rate_bead": {
        "grid": [4096, 4096], "device_ms": prof, "e2e_ms_host_field_and_mask": wall}}

    # ---- configs[4]: time sweep, T = 20, LCS ridge field + lines per slice, results on device
    x, y = np.linspace(0, 2, 4096), np.linspace(0, 1, 2048)
    spans = [(k * 10.0 / 64 + 20.0, k * 10.0 / 64) for k in range(args.slices)]
    ftle_time_sweep(x, y, flows.double_gyre, spans[:1], edge_order=2, percentile=80, ridge=True,
                    to_host=False, rtol=TOL, atol=TOL)
    res = {}

    def sweep():
        res["r"] = ftle_time_sweep(x, y, flows.double_gyre, spans, edge_order=2, percentile=80,
                                   ridge=True, to_host=False, rtol=TOL, atol=TOL)

    wall, prof = timed(sweep, reps=1)
    seeds = len(x) * len(y) * len(spans)
    dev = sum(prof.values())
    out["config5"] = {"grid": [4096, 2048], "slices": len(spans), "T": 20.0, "percentile": 80,
                      "device_ms": prof, "device_ms_total": round(dev, 2), "wall_ms": wall,
                      "host_linking_and_sync_ms": round(wall - dev, 2),
                      "seeds_per_s_device": seeds / dev * 1e3, "seeds_per_s_wall": seeds / wall * 1e3,
                      "ridge_lines_total": int(sum(len(v[3]) for v in res["r"].values())),
                      "ridge_vertices_total": int(sum(sum(len(l) for l in v[3]) for v in res["r"].values()))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
