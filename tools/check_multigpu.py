"""Run under torchrun with N >= 2 ranks: the row-sharded FTLE / LCS / Eulerian results must be
bit-identical to the unsharded ones computed on the same GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))

from dynlab_b200 import sharding  # noqa: E402
from dynlab_b200.diagnostics import FTLE, LCS, AttractionRate, TrajectoryRepulsionRate, iLES  # noqa: E402
from dynlab_b200.flows import bickley_jet, double_gyre  # noqa: E402

ok = True
for (nx, ny, eo) in ((257, 131, 2), (64, 5, 1), (33, 3, 2), (1024, 512, 2)):
    x, y = np.linspace(0, 2, nx), np.linspace(0, 1, ny)
    sharding.set_distributed(True)
    d = FTLE()
    f_sh = np.ma.getdata(d.compute(x, y, double_gyre, (6, 0), edge_order=eo, rtol=1e-8, atol=1e-8)).copy()
    fm_sh = d.flow_map.copy()
    sharding.set_halo_mode("redundant")  # halo rows integrated locally instead of exchanged
    f_red = np.ma.getdata(FTLE().compute(x, y, double_gyre, (6, 0), edge_order=eo, rtol=1e-8, atol=1e-8)).copy()
    sharding.set_halo_mode("exchange")
    sharding.set_host_result("shared")  # one pinned host segment, each rank copies its own rows
    f_shm = np.ma.getdata(FTLE().compute(x, y, double_gyre, (6, 0), edge_order=eo, rtol=1e-8, atol=1e-8)).copy()
    sharding.set_host_result("all")
    l = LCS()
    l.compute(x, y, f=double_gyre, t=(6, 0), edge_order=eo, percentile=70, debug=True, rtol=1e-8, atol=1e-8)
    xi_sh = np.ma.getdata(l.Xi_max).copy()
    a_sh = np.ma.getdata(AttractionRate().compute(x * 1e4, y * 8e3 - 4e3, f=bickley_jet, t=0.1, edge_order=eo)).copy()
    r = TrajectoryRepulsionRate().compute(x, y, f=double_gyre, t=0.3, edge_order=eo)
    r_sh, rm_sh = np.ma.getdata(r).copy(), np.ma.getmaskarray(r).copy()
    sharding.set_distributed(False)
    d = FTLE()
    f_1 = np.ma.getdata(d.compute(x, y, double_gyre, (6, 0), edge_order=eo, rtol=1e-8, atol=1e-8))
    l = LCS()
    l.compute(x, y, f=double_gyre, t=(6, 0), edge_order=eo, percentile=70, debug=True, rtol=1e-8, atol=1e-8)
    a_1 = np.ma.getdata(AttractionRate().compute(x * 1e4, y * 8e3 - 4e3, f=bickley_jet, t=0.1, edge_order=eo))
    r = TrajectoryRepulsionRate().compute(x, y, f=double_gyre, t=0.3, edge_order=eo)
    same = (np.array_equal(f_sh, f_1) and np.array_equal(f_shm, f_1) and np.array_equal(f_red, f_1) and np.array_equal(fm_sh, d.flow_map)
            and np.array_equal(xi_sh, np.ma.getdata(l.Xi_max)) and np.array_equal(a_sh, a_1)
            and np.array_equal(r_sh, np.ma.getdata(r)) and np.array_equal(rm_sh, np.ma.getmaskarray(r)))
    ok = ok and same
    if rank == 0:
        print(f"grid {nx}x{ny} eo={eo} world={world}: sharded == single: {same}", flush=True)
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
if rank == 0:
    print("MULTIGPU_OK" if t.item() == 1 else "MULTIGPU_MISMATCH", flush=True)
sys.exit(0 if t.item() == 1 else 1)
