"""BASELINE configs[4] on N GPUs (torchrun): 64 double-gyre slices 4096x2048, T = 20, LCS ridge
field + lines per slice, slices dealt round-robin to the ranks (no data-path communication).
Prints one JSON line on rank 0: wall time (max over ranks) for the whole sweep."""
import json, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
import dynlab_b200.flows as flows  # noqa: E402
from dynlab_b200.sweep import ftle_time_sweep  # noqa: E402

x, y = np.linspace(0, 2, 4096), np.linspace(0, 1, 2048)
spans = [(k * 10.0 / 64 + 20.0, k * 10.0 / 64) for k in range(64)]
kw = dict(edge_order=2, percentile=80, ridge=True, to_host=False, rtol=1e-8, atol=1e-8)
ftle_time_sweep(x, y, flows.double_gyre, spans[:world], **kw)  # warm-up: one slice per rank
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
res = ftle_time_sweep(x, y, flows.double_gyre, spans, **kw)
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
nl = torch.tensor([sum(len(v[3]) for v in res.values())], dtype=torch.int64, device="cuda")
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dist.all_reduce(nl)
    dist.destroy_process_group()
if rank == 0:
    seeds = 64 * len(x) * len(y)
    print(json.dumps({"config": "64 slices double_gyre 4096x2048, T=20, LCS ridge lines (percentile 80)",
                      "n_gpus": world, "wall_s": float(dt), "seeds_per_s": seeds / float(dt),
                      "slices_per_rank": len(res), "ridge_lines_total": int(nl)}))
