"""A/B timings of the multi-GPU FTLE step (config 3) inside ONE torchrun session: transport
(peer-memory push gather vs NCCL all-gather), halo mode, chunk count, slab cut.  Prints one JSON
line per variant on rank 0 (ms per step = CUDA events, max over ranks; sha of the field).

    torchrun --nproc-per-node N tools/bench_group_variants.py [--steps 20]
"""
import argparse
import hashlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--e2e", action="store_true", help="also time FTLE().compute (host field)")
ap.add_argument("--trace", action="store_true", help="print an event timeline of one step")
ap.add_argument("--only", default="", help="substring filter on variant names")
args = ap.parse_args()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

from dynlab_b200 import sharding  # noqa: E402
from dynlab_b200._engine import Engine  # noqa: E402
from dynlab_b200.diagnostics import FTLE  # noqa: E402
from dynlab_b200.flows import double_gyre  # noqa: E402

x, y = np.linspace(0, 2, 8192), np.linspace(0, 1, 4096)
kw = dict(edge_order=2, rtol=1e-8, atol=1e-8)
eng = Engine.get()


def run(name, transport="peer", halo="exchange", chunks=4, frac=32, weighted=True, overlap=None,
        fused=1, hchunks=4, side=False, lanes=1):
    if args.only and args.only not in name:
        return
    sharding.set_transport(transport)
    sharding.set_halo_mode(halo)
    d = FTLE()
    d.pipeline_chunks, d.pipeline_last_fraction, d.weighted_slabs = hchunks, frac, weighted
    d.device_chunks = chunks
    if overlap is not None:
        d.overlap_chunks = overlap
    d.fused_gather_chunks = fused
    d.side_stencil = side
    d.push_lanes = lanes
    for _ in range(3):
        d.compute_device(x, y, double_gyre, (10, 0), **kw)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.profile = {}
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        f = d.compute_device(x, y, double_gyre, (10, 0), **kw)
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    prof, eng.profile = eng.profile, None
    k1 = sum(a.elapsed_time(b) for a, b in prof["rk45_flowmap"]) / args.steps
    t = torch.tensor([e0.elapsed_time(e1) / args.steps, k1], dtype=torch.float64, device="cuda")
    tmax, tmin = t.clone(), t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    out = {"variant": name, "transport": transport, "halo": halo, "chunks": list(chunks) if isinstance(chunks, tuple) else chunks, "last_fraction": frac,
           "weighted": weighted, "overlap": overlap, "fused": fused, "side": side, "push_lanes": lanes, "host_chunks": list(hchunks) if isinstance(hchunks, tuple) else hchunks, "ms_per_step": float(tmax[0]),
           "k1_ms_max_rank": float(tmax[1]), "k1_ms_min_rank": float(tmin[1]),
           "sha": hashlib.sha256(f.cpu().numpy().tobytes()).hexdigest()[:16]}
    if args.trace and transport == "peer":
        dist.barrier()
        torch.cuda.synchronize()
        d.trace = []
        d.compute_device(x, y, double_gyre, (10, 0), **kw)
        torch.cuda.synchronize()
        tr, d.trace = d.trace, None
        t0 = tr[0][2]
        out["timeline_ms_rank%d" % rank] = [[lab, round(t0.elapsed_time(ev), 3)] for _, lab, ev in tr]
        if rank not in (0, world - 1, world // 2):
            out.pop("timeline_ms_rank%d" % rank)
        ends = torch.tensor([t0.elapsed_time(tr[-1][2])], dtype=torch.float64, device="cuda")
        allends = [torch.zeros_like(ends) for _ in range(world)]
        dist.all_gather(allends, ends)
        out["end_ms_by_rank"] = [round(float(e[0]), 3) for e in allends]
        if rank in (world - 1, world // 2) and rank != 0:
            print(json.dumps({"variant": name, "rank": rank,
                              "timeline": out.pop("timeline_ms_rank%d" % rank)}), flush=True)
    if args.e2e and transport == "peer":
        import time
        for _ in range(3):
            d.compute(x, y, double_gyre, (10, 0), **kw)
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            d.compute(x, y, double_gyre, (10, 0), **kw)
        torch.cuda.synchronize()
        t = torch.tensor([(time.perf_counter() - t0) / args.steps * 1e3], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["e2e_ms_per_step"] = float(t[0])
    if rank == 0:
        print(json.dumps(out), flush=True)


run("nccl after the kernels (round 1)", transport="nccl")
run("default (.875/.125, 1 push lane)", chunks=(0.875, 0.125), halo="redundant", side=False)
run(".875/.125, 2 push lanes", chunks=(0.875, 0.125), halo="redundant", side=False, lanes=2)
run(".875/.125, 4 push lanes", chunks=(0.875, 0.125), halo="redundant", side=False, lanes=4)
run(".875/.125, 7 push lanes", chunks=(0.875, 0.125), halo="redundant", side=False, lanes=7)
run(".9375/.0625, 1 push lane", chunks=(0.9375, 0.0625), halo="redundant", side=False)
run(".9375/.0625, 4 push lanes", chunks=(0.9375, 0.0625), halo="redundant", side=False, lanes=4)
run(".9375/.0625, 7 push lanes", chunks=(0.9375, 0.0625), halo="redundant", side=False, lanes=7)
run(".92/.08, 4 push lanes", chunks=(0.92, 0.08), halo="redundant", side=False, lanes=4)
run(".9/.1, 4 push lanes", chunks=(0.9, 0.1), halo="redundant", side=False, lanes=4)
dist.destroy_process_group()
