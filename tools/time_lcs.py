"""Stage breakdown of LCS.compute on one slice of BASELINE config 5 (4096 x 2048, T = 10):
flow map, FTLE + xi, percentile, ridge field, contour cells, host linking.  Prints one JSON line.
    python tools/time_lcs.py [xdim ydim]
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dynlab_b200 as dl  # noqa: E402
import dynlab_b200.diagnostics  # noqa: E402,F401
import dynlab_b200.flows  # noqa: E402,F401
from dynlab_b200._engine import Engine  # noqa: E402


def main():
    xdim = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    ydim = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    x, y = np.linspace(0, 2, xdim), np.linspace(0, 1, ydim)
    eng = Engine.get()
    lcs = dl.diagnostics.LCS()
    kw = dict(f=dl.flows.double_gyre, t=(0.0, 10.0), edge_order=2, percentile=90, rtol=1e-8, atol=1e-8)
    lcs.compute(x, y, **kw)  # warm-up
    out = {"grid": [xdim, ydim]}
    for rep in range(2):
        eng.profile = {}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lines = lcs.compute(x, y, **kw)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        prof = {k: round(sum(a.elapsed_time(b) for a, b in v), 3) for k, v in eng.profile.items()}
        eng.profile = None
        out[f"rep{rep}"] = {"wall_ms": round(wall * 1e3, 2), "device_ms": prof, "lines": len(lines),
                            "vertices": int(sum(len(l) for l in lines))}
    # the numpy percentile this replaces, for scale
    f = eng.to_host(lcs._field_dev) if hasattr(lcs, "_field_dev") else None
    if f is None:
        f = np.ma.getdata(dl.diagnostics.FTLE().compute(x, y, dl.flows.double_gyre, (0.0, 10.0), edge_order=2,
                                                        rtol=1e-8, atol=1e-8))
    t0 = time.perf_counter()
    np.percentile(f, 90)
    out["numpy_percentile_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
