"""Run under torchrun with N >= 2 ranks: the row-sharded FTLE / LCS / Eulerian results must be
bit-identical to the unsharded ones computed on the same GPU, for every transport (peer-memory
path of csrc/peer.cu, NCCL), halo mode (exchange, redundant), host-result mode (shared segment,
private copy), slab cut (by rows, by integrator work) and for the device-resident result.

    torchrun --nproc-per-node N tools/check_multigpu.py [--full] [--backend nccl|gloo] [--same-device]

--full adds BASELINE config 3 (8192x4096, t=(10,0)), also checked against the C oracle on the
every-64th-row/column sample.  --backend gloo --same-device puts every rank on cuda:0 (CUDA IPC
between processes sharing one GPU; torch.distributed only carries the plumbing, so the NCCL
transport is skipped).  HAZARD: the ranks' wait kernels then spin on flags written by other
processes of the SAME GPU, which only works as long as the driver time-slices them; the
profiling guide reports Xid 109 (context-switch timeout) for that pattern, so no test runs it -
it is a manual debugging aid.
"""
import argparse
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--full", action="store_true")
ap.add_argument("--backend", default="nccl", choices=["nccl", "gloo"])
ap.add_argument("--same-device", action="store_true")
args = ap.parse_args()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = 0 if args.same_device else int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
if args.backend == "nccl":
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
else:
    dist.init_process_group("gloo")

from dynlab_b200 import peer, sharding  # noqa: E402
from dynlab_b200.diagnostics import FTLE, LCS, AttractionRate, TrajectoryRepulsionRate  # noqa: E402
from dynlab_b200.flows import bickley_jet, double_gyre  # noqa: E402

ok = True


def say(msg):
    if rank == 0:
        print(msg, flush=True)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


grids = [(257, 131, 2, (6, 0)), (64, 5, 1, (6, 0)), (33, 3, 2, (6, 0)), (1024, 512, 2, (6, 0)),
         (2048, 1030, 2, (3, 0))]
if args.full:
    grids.append((8192, 4096, 2, (10, 0)))
transports = ["peer"] + (["nccl"] if args.backend == "nccl" else [])
for (nx, ny, eo, t) in grids:
    x, y = np.linspace(0, 2, nx), np.linspace(0, 1, ny)
    kw = dict(edge_order=eo, rtol=1e-8, atol=1e-8)
    big = nx * ny >= 1 << 20
    # ---- reference: this rank alone, whole grid
    sharding.set_distributed(False)
    d1 = FTLE()
    f_1 = np.ma.getdata(d1.compute(x, y, double_gyre, t, **kw)).copy()
    fm_1 = d1.flow_map.copy()
    nfev_1 = d1.stats["nfev"]
    sharding.set_distributed(True)
    results = {}
    for tr in transports:
        sharding.set_transport(tr)
        for halo in ("exchange", "redundant"):
            sharding.set_halo_mode(halo)
            for weighted in ((True, False) if (tr == "peer" and big) else (True,)):
                tag = f"{tr}/{halo}/{'work' if weighted else 'rows'}-cut"
                d = FTLE()
                d.weighted_slabs = weighted
                for host in (("shared", "private") if tr == "peer" else ("shared",)):
                    sharding.set_host_result(host)
                    f = np.ma.getdata(d.compute(x, y, double_gyre, t, **kw)).copy()
                    results[f"{tag}/host:{host}"] = np.array_equal(f, f_1)
                sharding.set_host_result("shared")
                # integrator counters: this rank's seeds; summed over ranks = the single run (in
                # both halo modes: redundantly integrated halo rows are not counted)
                tot = torch.tensor([d.stats["nfev"], d.stats["seeds"]], dtype=torch.int64,
                                   device="cuda" if args.backend == "nccl" else "cpu")
                dist.all_reduce(tot)
                results[f"{tag}/nfev"] = tot.tolist() == [nfev_1, nx * ny]
                if not (args.full and nx == 8192 and (tr, halo) != ("peer", "exchange")):
                    results[f"{tag}/flow_map"] = np.array_equal(d.flow_map, fm_1)
                for rep in range(2):  # twice: rotating result buffers
                    fd = d.compute_device(x, y, double_gyre, t, **kw)
                    torch.cuda.synchronize()
                    results[f"{tag}/device{rep}"] = np.array_equal(fd.cpu().numpy(), f_1)
    sharding.set_transport("peer")
    sharding.set_halo_mode("redundant")
    g = peer.current_group(nx * ny)
    g.check_timeouts()
    # ---- LCS / Eulerian: uniform slabs, gathered over peer memory (sharding.gather_slab) or NCCL
    if args.backend == "nccl" and nx * ny < 1 << 22:
        u, v = bickley_jet(0.1, np.meshgrid(x * 1e4, y * 8e3 - 4e3))
        sharding.set_distributed(False)
        l1 = LCS()
        l1.compute(x, y, f=double_gyre, t=t, percentile=70, debug=True, **kw)
        a_1 = np.ma.getdata(AttractionRate().compute(x * 1e4, y * 8e3 - 4e3, f=bickley_jet, t=0.1, edge_order=eo))
        a_uv1 = np.ma.getdata(AttractionRate().compute(x * 1e4, y * 8e3 - 4e3, u=u, v=v, edge_order=eo))
        r1 = TrajectoryRepulsionRate().compute(x, y, f=double_gyre, t=0.3, edge_order=eo)
        sharding.set_distributed(True)
        for tr in transports:
            sharding.set_transport(tr)
            for rep in range(2):  # twice: rotating gather buffers
                lc = LCS()
                lc.compute(x, y, f=double_gyre, t=t, percentile=70, debug=True, **kw)
                a_sh = np.ma.getdata(AttractionRate().compute(x * 1e4, y * 8e3 - 4e3, f=bickley_jet, t=0.1, edge_order=eo))
                a_uv = np.ma.getdata(AttractionRate().compute(x * 1e4, y * 8e3 - 4e3, u=u, v=v, edge_order=eo))
                r = TrajectoryRepulsionRate().compute(x, y, f=double_gyre, t=0.3, edge_order=eo)
                results[f"{tr}/lcs/xi{rep}"] = np.array_equal(np.ma.getdata(lc.Xi_max), np.ma.getdata(l1.Xi_max))
                results[f"{tr}/lcs/ftle{rep}"] = np.array_equal(np.ma.getdata(lc.ftle), np.ma.getdata(l1.ftle))
                results[f"{tr}/lcs/lines{rep}"] = len(lc.lcs) == len(l1.lcs) and all(
                    np.array_equal(p, q) for p, q in zip(lc.lcs, l1.lcs))
                results[f"{tr}/attraction/flow{rep}"] = np.array_equal(a_sh, a_1)
                results[f"{tr}/attraction/uv{rep}"] = np.array_equal(a_uv, a_uv1)
                results[f"{tr}/rate{rep}"] = (np.array_equal(np.ma.getdata(r), np.ma.getdata(r1))
                                             and np.array_equal(np.ma.getmaskarray(r), np.ma.getmaskarray(r1)))
        sharding.set_transport("peer")
        g.check_timeouts()
    extra = ""
    if args.full and nx == 8192:  # the C oracle on the every-64th sample (bench.py's result_check)
        from oracle import c_oracle as co
        from oracle import py_oracle as po

        xs, ys = x[::64], y[::64]
        fm_o, _, _ = co.flow_map("double_gyre", po.flow_params("double_gyre"), xs, ys, float(t[0]),
                                 float(t[1]), 1e-8, 1e-8, nthreads=os.cpu_count() or 1)
        err = float(np.abs(fm_1[::64, ::64] - fm_o).max())
        results["oracle_sample<=1e-10"] = err <= 1e-10
        extra = f" max|flow_map - C oracle| on the /64 sample = {err:.3e}; field sha256 = {sha(f_1)}"
    bad = [k for k, v in results.items() if not v]
    ok = ok and not bad
    say(f"grid {nx}x{ny} eo={eo} t={t} world={world}: {len(results)} checks, "
        f"{'all bit-identical to the single-GPU result' if not bad else 'MISMATCH: ' + ', '.join(bad)}"
        f" [{len(FTLE._bounds_cache)} slab cuts cached]" + extra)
flag = torch.tensor([1 if ok else 0], device="cuda" if args.backend == "nccl" else "cpu")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
say("MULTIGPU_OK" if flag.item() == 1 else "MULTIGPU_MISMATCH")
sys.exit(0 if flag.item() == 1 else 1)
