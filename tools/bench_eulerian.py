"""Supplementary measurement for BASELINE config 2: AttractionRate / iLES eigen stage on the
bickley_jet velocity field 8192x8192, edge_order 2 (Eulerian stencil + eigen, HBM-bound).
Prints one JSON line (not the driver's bench contract — see bench.py for that)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynlab_b200._engine import Engine
from dynlab_b200._resolve import resolve_flow
from dynlab_b200 import flows, _native as nat
from dynlab_b200.diagnostics import AttractionRate
from oracle import py_oracle as po

N = 8192
eng = Engine.get()
_, fid, _, params = resolve_flow(flows.bickley_jet)
x, y = np.linspace(0, 20000, N), np.linspace(-4000, 4000, N)
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {}
hbm = peaks.get("hbm_gbs", 6650.0)

def ev(fn, n=10, warm=3):
    for _ in range(warm): fn()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

u, v = eng.flow_eval_grid(fid, params, 0.0, x, y)
k0 = ev(lambda: eng.flow_eval_grid(fid, params, 0.0, x, y))
res = {"config": "AttractionRate / iLES eigen stage, bickley_jet 8192x8192, edge_order=2 (BASELINE configs[1])",
       "points": N * N, "K0_flow_eval_ms": k0, "hbm_peak_gbs": hbm, "kernels": {}}
for name, mode, xi, b in (("attraction_rate(s_min)", nat.EULER_S_MIN, 0, 24), ("repulsion_rate(s_max)", nat.EULER_S_MAX, 0, 24),
                          ("trajectory_repulsion_rate", nat.EULER_RATE, 0, 25), ("iles_eigen(s_min+Xi forced)", nat.EULER_S_MIN, 2, 40)):
    ms = ev(lambda: eng.euler_stencil(mode, u, v, 0, x, y, 2, 0, N, False, xi))
    gbs = b * N * N / ms / 1e6
    res["kernels"][name] = {"ms": ms, "algorithmic_bytes_per_point": b, "GB_per_s": gbs, "frac_of_measured_hbm": gbs / hbm,
                            "points_per_s": N * N / ms * 1e3}
# K0 fused into K3 (no u, v in HBM): algorithmic bytes = the 8 B/point result (+ mask / Xi)
res["fused_flow_stencil"] = {}
for name, mode, xi, b in (("attraction_rate(s_min)", nat.EULER_S_MIN, 0, 8), ("trajectory_repulsion_rate", nat.EULER_RATE, 0, 9),
                          ("iles_eigen(s_min+Xi forced)", nat.EULER_S_MIN, 2, 24)):
    ms = ev(lambda: eng.euler_stencil_flow(mode, fid, params, 0.0, x, y, 2, 0, N, False, xi))
    res["fused_flow_stencil"][name] = {"ms": ms, "two_kernel_ms": k0 + res["kernels"][name]["ms"],
                                       "algorithmic_bytes_per_point": b, "GB_per_s": b * N * N / ms / 1e6,
                                       "points_per_s": N * N / ms * 1e3}
# end to end through the public API
uh, vh = u.cpu().numpy(), v.cpu().numpy()
d = AttractionRate()
d.compute(x, y, u=uh, v=vh, edge_order=2)
t0 = time.perf_counter(); d.compute(x, y, u=uh, v=vh, edge_order=2); e2e_uv_second = time.perf_counter() - t0
d.compute(x, y, u=uh, v=vh, edge_order=2)
t0 = time.perf_counter(); d.compute(x, y, u=uh, v=vh, edge_order=2); e2e_uv = time.perf_counter() - t0
d.compute(x, y, f=flows.bickley_jet, t=0.0, edge_order=2)
t0 = time.perf_counter(); d.compute(x, y, f=flows.bickley_jet, t=0.0, edge_order=2); e2e_ft = time.perf_counter() - t0
Engine.PINNED_RESULTS = True  # steady state of a loop over many fields (pinned result blocks cached)
for _ in range(3): d.compute(x, y, u=uh, v=vh, edge_order=2)
t0 = time.perf_counter(); d.compute(x, y, u=uh, v=vh, edge_order=2); e2e_uv_pinned = time.perf_counter() - t0
for _ in range(3): d.compute(x, y, f=flows.bickley_jet, t=0.0, edge_order=2)
t0 = time.perf_counter(); d.compute(x, y, f=flows.bickley_jet, t=0.0, edge_order=2); e2e_ft_pinned = time.perf_counter() - t0
Engine.PINNED_RESULTS = False
res["e2e"] = {"pinned_results_steady_state_ms": {"u_v_host_arrays": e2e_uv_pinned * 1e3, "f_t_route": e2e_ft_pinned * 1e3},
              "u_v_host_arrays_ms": e2e_uv * 1e3, "u_v_host_arrays_second_call_ms": e2e_uv_second * 1e3, "h2d_bytes": 2 * 8 * N * N, "d2h_bytes": 8 * N * N,
              "f_t_route_ms": e2e_ft * 1e3, "points_per_s_u_v": N * N / e2e_uv, "points_per_s_f_t": N * N / e2e_ft}
# CPU baseline: the reference's per-point loop (oracle port, single-threaded like the reference class) on a crop
c = 384
t0 = time.perf_counter()
po.attraction_repulsion_rate(x[:c], y[:c], u=uh[:c, :c], v=vh[:c, :c], edge_order=2)
dt = time.perf_counter() - t0
res["cpu_baseline"] = {"points_per_s": c * c / dt, "cores": 1, "kind": "port",
                       "sample": f"{c}x{c} crop, oracle/py_oracle.py (np.gradient + stacked np.linalg.eig; the reference's class is single-threaded)"}
print(json.dumps(res))
