"""CUDA-event timings of the stencil kernels alone: K2 (FTLE, config-3 grid, random flow map) and
K3 modes (config-2 grid, Bickley jet).  Prints one line per kernel; used for A/B runs of
alternative builds (DLB_LIB_PATH, tools/build_stencil_variant.sh)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynlab_b200 import _native as nat  # noqa: E402
from dynlab_b200 import flows  # noqa: E402
from dynlab_b200._engine import Engine  # noqa: E402
from dynlab_b200._resolve import resolve_flow  # noqa: E402

eng = Engine.get()


def run(fn, n=9, warm=3, inner=20):
    """Median over n batches of `inner` back-to-back launches (the queue stays full, so the
    interval is device time, not host launch latency), per launch."""
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


out = {"lib": os.path.basename(nat.LIB_PATH)}
x, y = np.linspace(0, 2, 8192), np.linspace(0, 1, 4096)
g = torch.Generator(device="cuda").manual_seed(1)
fm = torch.rand(4096, 8192, 2, dtype=torch.float64, device="cuda", generator=g) * 2
res = eng.empty(4096, 8192)
ms = run(lambda: eng.ftle_stencil(fm, 0, x, y, 2, 10.0, 0, 4096, out=res))
out["K2_ftle_ms"] = round(ms, 4)
out["K2_frac_hbm"] = round(24 * 4096 * 8192 / ms / 1e6 / 6553.9, 3)
xi = eng.empty(4096, 8192, 2)
ms = run(lambda: eng.ftle_stencil(fm, 0, x, y, 2, 10.0, 0, 4096, nat.XI_LAPACK, out=res, xi_out=xi))
out["K2_ftle_xi_ms"] = round(ms, 4)
del fm, res, xi
_, fid, _, params = resolve_flow(flows.bickley_jet)
x, y = np.linspace(0, 20000, 8192), np.linspace(-4000, 4000, 8192)
u, v = eng.flow_eval_grid(fid, params, 0.0, x, y)
for name, mode, xim, b in (("s_min", nat.EULER_S_MIN, 0, 24), ("s_max", nat.EULER_S_MAX, 0, 24),
                           ("rate", nat.EULER_RATE, 0, 25), ("s_min_xi", nat.EULER_S_MIN, 2, 40)):
    ms = run(lambda: eng.euler_stencil(mode, u, v, 0, x, y, 2, 0, 8192, False, xim))
    out[f"K3_{name}_ms"] = round(ms, 4)
    out[f"K3_{name}_frac_hbm"] = round(b * 8192 * 8192 / ms / 1e6 / 6553.9, 3)
xu, yu = np.arange(8192) * 2.5, np.arange(8192) * 1.0  # exactly uniform axes
ms = run(lambda: eng.euler_stencil(nat.EULER_S_MIN, u, v, 0, xu, yu, 2, 0, 8192, False, 0))
out["K3_s_min_uniform_axes_ms"] = round(ms, 4)
print(json.dumps(out), flush=True)
