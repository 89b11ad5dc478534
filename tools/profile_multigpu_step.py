"""torchrun script: where the time of one sharded FTLE step goes besides K1 (CUDA events on
every rank between the stages; prints rank 0's and the max over ranks)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
from dynlab_b200 import sharding  # noqa: E402
from dynlab_b200._engine import Engine  # noqa: E402
from dynlab_b200._resolve import resolve_flow  # noqa: E402
from dynlab_b200 import flows  # noqa: E402

eng = Engine.get()
x, y = np.linspace(0, 2, 8192), np.linspace(0, 1, 4096)
ydim, xdim = len(y), len(x)
_, fid, _, params = resolve_flow(flows.double_gyre)
slab = sharding.partition(ydim, world, rank)
names = ["K1", "halo", "K2", "gather"]
acc = np.zeros(len(names))
reps = 12
for rep in range(reps + 2):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    dist.barrier(); torch.cuda.synchronize()
    ev[0].record()
    local = eng.empty(slab.nloc, xdim, 2)
    eng.flow_map(fid, params, x, y, slab.row0, slab.rows, 10.0, 0.0, 1e-8, 1e-8,
                 out=local[slab.halo_lo:slab.halo_lo + slab.rows])
    ev[1].record()
    sharding.exchange_halo(local, slab, ydim, rank, world)
    ev[2].record()
    out = torch.empty(slab.rows_max, xdim, dtype=torch.float64, device=eng.device)
    eng.ftle_stencil(local, slab.lo, x, y, 2, 10.0, slab.row0, slab.rows, out=out[:slab.rows])
    ev[3].record()
    field = sharding.gather_rows(out, ydim, world)
    ev[4].record()
    torch.cuda.synchronize()
    if rep >= 2:
        acc += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(len(names))])
acc /= reps
t = torch.tensor(acc, device="cuda")
tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
tmin = t.clone(); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
if rank == 0:
    for n, a, b, c in zip(names, acc, tmin.tolist(), tmax.tolist()):
        print(f"{n:7s} rank0 {a:8.3f} ms   min {b:8.3f}   max {c:8.3f}")
    print("sum rank0", acc.sum())
dist.destroy_process_group()
