"""Host->device / device->host transfer timings of one 8192x8192 float64 field (537 MB)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynlab_b200._engine import Engine
eng = Engine.get()
N = 8192
a = np.random.default_rng(0).normal(size=(N, N))
def t(fn, n=3):
    best = 1e9
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3, r
ms, d = t(lambda: torch.from_numpy(a).to(eng.device)); print(f"plain .to(device): {ms:.1f} ms  {a.nbytes/ms/1e6:.1f} GB/s")
for w in (2, 4, 6, 8, 12):
    Engine.H2D_WORKERS = w; eng._h2d = None
    ms, d = t(lambda: eng.to_device(a)); print(f"staged, {w} workers: {ms:.1f} ms  {a.nbytes/ms/1e6:.1f} GB/s")
assert np.array_equal(d.cpu().numpy(), a)
ms, h = t(lambda: eng.to_host(d)); print(f"to_host (staged -> pageable): {ms:.1f} ms  {a.nbytes/ms/1e6:.1f} GB/s")
assert np.array_equal(h, a)
def pinned():
    hh = torch.empty(d.shape, dtype=d.dtype, pin_memory=True); hh.copy_(d, non_blocking=True); torch.cuda.synchronize(); return hh
ms, _ = t(pinned, n=1); print(f"to pinned result, first call (page-locking included): {ms:.1f} ms")
ms, _ = t(pinned); print(f"to pinned result, cached block: {ms:.1f} ms")
b = np.empty_like(a)
ms, _ = t(lambda: np.copyto(b, a)); print(f"np.copyto 537 MB single thread: {ms:.1f} ms  {a.nbytes/ms/1e6:.1f} GB/s")
ms, _ = t(lambda: torch.empty(a.shape, dtype=torch.float64, pin_memory=True)); print(f"pinned alloc: {ms:.1f} ms")
