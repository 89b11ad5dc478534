"""Row-slab sharding of the seed grid over the ranks of one box.

Seeds are independent, so the flow-map kernel shards with no data-path collective.  The
stencil needs the neighbouring slab's boundary row (central difference): one point-to-point
exchange of a single row per neighbour (xdim * 16 B), then an all-gather of the output rows.
Everything here works on CPU tensors with the ``gloo`` backend as well (tests) — it only moves
rows, it never computes.
"""
from __future__ import annotations

import mmap
import os
import socket
import uuid
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

_mode = "auto"  # "auto": shard when torch.distributed is initialised with world_size > 1


def set_distributed(mode) -> None:
    """``True``/``"auto"``: shard across the default process group when it exists;
    ``False``: every rank computes the whole grid."""
    global _mode
    _mode = "auto" if mode in (True, "auto") else "off"


_halo = "redundant"


_barrier_state = None  # (flags ndarray over shared memory, generation)


def host_barrier(timeout_s: float = 120.0) -> None:
    """Barrier between the ranks of the box through a few words of shared memory (a store and a
    spin on the host, ~10 us) instead of a device collective: the shared-result path has nothing
    left to exchange on the device, and an NCCL barrier costs a kernel launch plus a device
    synchronisation on every rank.  Every rank must already have synchronised the stream whose
    copies the others are going to read."""
    global _barrier_state
    rank, size = world()
    if size == 1:
        return
    if _barrier_state is None:
        flags, _ = shared_host_array((max(size, 8),), dtype=np.int64, tag="barrier")
        dist.barrier()  # rank 0 zero-filled the segment before anybody counts on it
        _barrier_state = [flags, 0]
    flags = _barrier_state[0]
    _barrier_state[1] += 1
    gen = _barrier_state[1]
    flags[rank] = gen
    import time

    deadline = time.monotonic() + timeout_s
    spins = 0
    while int(flags[:size].min()) < gen:
        spins += 1
        if spins > 200:  # the others are more than a few us away: stop burning the core
            os.sched_yield() if spins < 20000 else time.sleep(50e-6)
        if spins & 0x3ff == 0 and time.monotonic() > deadline:
            raise RuntimeError("host_barrier: a rank did not arrive (died or never called it)")


def set_halo_mode(mode) -> None:
    """How a rank obtains the flow-map rows just outside its slab that the stencil needs:
    ``"redundant"`` (default): integrate those rows locally too (SURVEY §8e's alternative: 2
    extra rows per slab, no data-path communication before the gather and no waiting for a
    neighbour; bit-identical because seeds are independent; measured 0.5 % faster at 8 GPUs);
    ``"exchange"``: one row from each neighbour (``dlb_halo_exchange`` over peer memory, or
    NCCL point-to-point with ``set_transport("nccl")``)."""
    global _halo
    if mode not in ("exchange", "redundant"):
        raise ValueError('mode must be "exchange" or "redundant"')
    _halo = mode


def halo_mode() -> str:
    return _halo


_host_result = "shared"  # how the ranks of a torchrun job receive the field in host memory


def set_host_result(which) -> None:
    """One process per GPU: how every rank's ``FTLE.compute`` gets the full host field.
    ``"shared"`` (default): the ranks of the box share pinned host segments
    (``SharedResultPool``); each rank streams only its own rows into the segment while it is
    still integrating and every rank returns a view of the whole, so the field crosses PCIe once
    instead of once per rank and nothing is gathered on the devices.  A segment is only handed
    out again when no rank holds a result (or a view of one) in it any more, so results behave
    like the reference's freshly allocated arrays - except that the memory is the same on every
    rank (write into it and the other ranks see it).  ``"private"``: each rank returns its own
    ordinary numpy copy (costs a ydim*xdim*8-byte host memcpy per rank per call)."""
    global _host_result
    if which == "all":  # round-1 name of the private-copy mode
        which = "private"
    if which not in ("shared", "private"):
        raise ValueError('which must be "shared" or "private"')
    _host_result = which


def host_result_mode() -> str:
    return _host_result


_transport = "peer"


def set_transport(which) -> None:
    """One process per GPU: what carries the halo rows and the gathered field between the
    devices.  ``"peer"`` (default): this library's own peer-memory path (``csrc/peer.cu``: copy
    engine writes into CUDA-IPC mapped buffers + generation flags, overlapped with the
    integrator chunk by chunk).  ``"nccl"``: torch.distributed point-to-point + all-gather after
    the kernels (the round-1 path, kept for A/B measurements and as the path of LCS and the
    Eulerian diagnostics)."""
    global _transport
    if which not in ("peer", "nccl"):
        raise ValueError('which must be "peer" or "nccl"')
    _transport = which


def transport() -> str:
    return _transport


def set_devices(spec) -> None:
    """Single process: devices a Lagrangian ``compute`` is sharded over (see ``peer.set_devices``)."""
    from . import peer

    peer.set_devices(spec)


_segments = {}  # (tag, nbytes) -> (mmap object, flat uint8 ndarray)


def shared_host_array(shape, dtype=np.float64, tag="result"):
    """Collective over the default process group: a host array of ``shape`` backed by one
    POSIX shared-memory segment mapped by every rank of the box (page-locked for CUDA when a
    device is in use, so slab copies into it run at PCIe speed).  Segments are cached by size
    and reused by later calls; the name is unlinked as soon as every rank has mapped it, so
    nothing is left in /dev/shm whatever happens to the job afterwards."""
    rank, size = world()
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    key = (tag, nbytes)
    seg = _segments.get(key)
    if seg is None:
        hosts = [None] * size
        dist.all_gather_object(hosts, socket.gethostname())
        if len(set(hosts)) != 1:
            raise RuntimeError('set_host_result("shared") needs all ranks on one box')
        name = [f"/dev/shm/dynlab_b200_{uuid.uuid4().hex}" if rank == 0 else None]
        fd = -1
        if rank == 0:
            fd = os.open(name[0], os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
            os.ftruncate(fd, max(nbytes, 1))
        dist.broadcast_object_list(name, src=0)
        if rank != 0:
            fd = os.open(name[0], os.O_RDWR)
        mm = mmap.mmap(fd, max(nbytes, 1))
        os.close(fd)
        flat = np.frombuffer(mm, dtype=np.uint8)
        if rank == 0:
            flat[:] = 0  # touch the pages once, here, not inside the first timed copy
        dist.barrier()
        if rank == 0:
            os.unlink(name[0])
        if torch.cuda.is_available():
            torch.cuda.check_error(
                torch.cuda.cudart().cudaHostRegister(flat.ctypes.data, max(nbytes, 1), 0))
        seg = _segments[key] = (mm, flat)
    arr = seg[1][:nbytes].view(dtype).reshape(shape)
    return arr, torch.from_numpy(arr)


class SharedResultPool:
    """Shared host segments for the results of one grid shape, handed out so that a result
    array the caller still holds on ANY rank is never overwritten.

    Every ``acquire`` wraps the chosen segment in a fresh ``np.frombuffer`` owner array; numpy
    makes every view derived from the returned result (``field[10:20]``, ``field.T``, the masked
    array's data) keep that owner alive, so a weak reference to it says exactly whether anything
    still looks at the segment.  The ranks exchange their busy bits through shared memory (two
    host barriers, ~20 us) and all pick the first segment free everywhere; if there is none a
    new one is created (collective, ~0.1 s: only when the caller accumulates results).  A plain
    ``f = d.compute(...)`` loop alternates between two segments."""

    _pools: dict = {}

    @classmethod
    def get(cls, shape, dtype=np.float64) -> "SharedResultPool":
        key = (tuple(int(v) for v in shape), np.dtype(dtype).str)
        pool = cls._pools.get(key)
        if pool is None:
            pool = cls._pools[key] = SharedResultPool(*key)
        return pool

    def __init__(self, shape, dtype):
        self.shape, self.dtype = shape, np.dtype(dtype)
        self.owners = []  # per segment: weakref to the owner array of its last hand-out (or None)
        self.calls = 0

    def _busy_bits(self) -> int:
        bits = 0
        for k, ref in enumerate(self.owners):
            if ref is not None and ref() is not None:
                bits |= 1 << k
        return bits

    def acquire(self):
        """Collective.  Returns ``(numpy array, torch tensor)`` over a segment no rank still
        uses; both alias the same page-locked shared memory."""
        import weakref

        rank, size = world()
        board, _ = shared_host_array((max(size, 8),), dtype=np.int64, tag="result-pool-board")
        board[rank] = self._busy_bits()
        host_barrier()
        busy = 0
        for r in range(size):
            busy |= int(board[r])
        host_barrier()  # everybody has read the board before anybody writes it again
        k = 0
        while k < len(self.owners) and (busy >> k) & 1:
            k += 1
        if k == len(self.owners):
            self.owners.append(None)
        # cached by (tag, nbytes): created on first use of index k, collectively
        _, _ = shared_host_array(self.shape, self.dtype, tag=("result", self.shape, k))
        mm, flat = _segments[(("result", self.shape, k), int(np.prod(self.shape)) * self.dtype.itemsize)]
        n = int(np.prod(self.shape))
        owner = np.frombuffer(mm, dtype=self.dtype, count=n)  # fresh object: the liveness token
        self.owners[k] = weakref.ref(owner)
        self.calls += 1
        arr = owner.reshape(self.shape)
        return arr, torch.from_numpy(arr)


def world():
    """(rank, world_size) the diagnostics shard over."""
    if _mode != "off" and dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


@dataclass(frozen=True)
class Slab:
    """Rows a rank owns and the rows it must hold locally (own + halo)."""
    row0: int      # first own row (global)
    rows: int      # number of own rows (0 for idle ranks)
    lo: int        # first local row (global index)
    hi: int        # one past the last local row
    rows_max: int  # largest slab in the partition (gather padding)

    @property
    def nloc(self) -> int:
        return self.hi - self.lo

    @property
    def halo_lo(self) -> int:
        return self.row0 - self.lo

    @property
    def halo_hi(self) -> int:
        return self.hi - (self.row0 + self.rows)


def active_ranks(ydim: int, size: int) -> int:
    """Ranks that get rows: every slab keeps >= 2 rows so a 1-row halo always suffices (the
    one-sided edge_order=2 formulas touch rows 0..2 / ydim-3..ydim-1 only)."""
    return max(1, min(size, ydim // 2))


def partition(ydim: int, size: int, rank: int) -> Slab:
    """Balanced contiguous row slabs with a 1-row halo towards each existing neighbour."""
    n = active_ranks(ydim, size)
    base, extra = divmod(ydim, n)
    rows_max = base + (1 if extra else 0)
    if rank >= n:
        return Slab(ydim, 0, ydim, ydim, rows_max)
    row0 = rank * base + min(rank, extra)
    rows = base + (1 if rank < extra else 0)
    lo = row0 - 1 if rank > 0 else row0
    hi = row0 + rows + (1 if rank < n - 1 else 0)
    return Slab(row0, rows, lo, hi, rows_max)


def uniform_bounds(ydim: int, size: int) -> list:
    """Row boundaries ``b[0..size]`` of ``partition``: rank r owns rows [b[r], b[r+1])."""
    out = [partition(ydim, size, r).row0 for r in range(size)] + [ydim]
    return [min(b, ydim) for b in out]


def weighted_bounds(ydim: int, size: int, row_cost) -> list:
    """Boundaries that give every active rank the same share of ``sum(row_cost)`` (a per-row
    work estimate, e.g. integrator attempts), every slab keeping >= 2 rows.  Deterministic in
    its inputs, so ranks that hold the same estimate agree without talking."""
    n = active_ranks(ydim, size)
    cost = np.maximum(np.asarray(row_cost, dtype=np.float64), 0.0)
    if n == 1 or cost.shape != (ydim,) or not np.isfinite(cost).all() or cost.sum() <= 0:
        return uniform_bounds(ydim, size)
    cum = np.concatenate(([0.0], np.cumsum(cost)))
    b = [0]
    for k in range(1, n):
        target = cum[-1] * k / n
        i = int(np.searchsorted(cum, target, side="left"))
        if i > 0 and target - cum[i - 1] < cum[i] - target:
            i -= 1  # the boundary nearer to the target
        i = max(i, b[-1] + 2)            # >= 2 rows for slab k-1 ...
        i = min(i, ydim - 2 * (n - k))   # ... and for every slab still to come
        b.append(i)
    return b + [ydim] * (size - n + 1)


def slab_from_bounds(bounds, ydim: int, rank: int) -> Slab:
    """``Slab`` of ``rank`` for arbitrary boundaries (1-row halo towards existing neighbours)."""
    row0, row1 = bounds[rank], bounds[rank + 1]
    rows = row1 - row0
    rows_max = max(b1 - b0 for b0, b1 in zip(bounds[:-1], bounds[1:]))
    if rows == 0:
        return Slab(ydim, 0, ydim, ydim, rows_max)
    return Slab(row0, rows, row0 - 1 if row0 > 0 else row0, row1 + 1 if row1 < ydim else row1,
                rows_max)


def exchange_halo(local: torch.Tensor, slab: Slab, ydim: int, rank: int, size: int) -> None:
    """Fill the halo rows of ``local`` ([nloc, ...], own rows already written) from the
    neighbouring ranks.  One row each way per neighbour."""
    n = active_ranks(ydim, size)
    if n == 1 or rank >= n:
        return
    ops, keep = [], []
    first_own, last_own = slab.halo_lo, slab.halo_lo + slab.rows - 1
    if rank > 0:  # neighbour above (smaller rows)
        send = local[first_own].contiguous()
        keep.append(send)
        ops.append(dist.P2POp(dist.isend, send, rank - 1))
        ops.append(dist.P2POp(dist.irecv, local[0], rank - 1))
    if rank < n - 1:
        send = local[last_own].contiguous()
        keep.append(send)
        ops.append(dist.P2POp(dist.isend, send, rank + 1))
        ops.append(dist.P2POp(dist.irecv, local[slab.nloc - 1], rank + 1))
    for req in dist.batch_isend_irecv(ops):
        req.wait()


def gather_slab(part, slab: "Slab", ydim: int, tail, dtype, device) -> torch.Tensor:
    """Rows [slab.row0, +slab.rows) of a ``[ydim, *tail]`` array live on this rank (``part``,
    None for an idle rank); returns the whole array on every rank.  The uniform ``partition``
    is assumed.  CUDA tensors with the ``"peer"`` transport: the rows are pushed into every
    rank's copy over NVLink peer memory (``dlb_gather_rows`` + ``dlb_wait_flags``, csrc/peer.cu;
    the result lives in one of ``peer.RESULT_DEPTH`` rotating symmetric buffers).  Otherwise
    (``"nccl"`` transport, CPU tensors under gloo): padded all-gather through torch.distributed."""
    rank, size = world()
    tail = tuple(int(v) for v in tail)
    G = None
    if torch.device(device).type == "cuda" and _transport == "peer":
        from . import peer

        G = peer.current_group()  # None if the GPUs cannot map each other's memory (-> NCCL)
    if G is not None:
        (m,) = G.local
        gen = G.next_generation()
        item = torch.empty((), dtype=dtype).element_size()
        row_b = int(np.prod(tail, dtype=np.int64)) * item if tail else item
        nbytes = ydim * row_b
        ring = [G.symmetric(("gather", str(dtype), tail, k), nbytes) for k in range(peer.RESULT_DEPTH)]
        buf = ring[gen % peer.RESULT_DEPTH]
        full = buf.view(m, dtype, (ydim,) + tail)
        with torch.cuda.device(m.device):
            main = torch.cuda.current_stream(m.device)
            if slab.rows:
                full[slab.row0:slab.row0 + slab.rows].copy_(part[:slab.rows])
            G.push_rows(m, buf, slab.row0 * row_b if slab.rows else 0, slab.rows * row_b,
                        gen * peer.GEN_STRIDE + peer.DONE, main.cuda_stream)
            G.wait_gathered(m, gen, main.cuda_stream)
        return full
    pad = torch.zeros((slab.rows_max,) + tail, dtype=dtype, device=device)
    if part is not None and slab.rows:
        pad[:slab.rows] = part[:slab.rows]
    return gather_rows(pad, ydim, size)


def gather_rows(own_padded: torch.Tensor, ydim: int, size: int) -> torch.Tensor:
    """All-gather per-rank outputs (``[rows_max, ...]``, first ``rows`` valid) into the full
    ``[ydim, ...]`` array on every rank."""
    rows_max = own_padded.shape[0]
    tail = tuple(own_padded.shape[1:])
    buf = torch.empty((size, rows_max) + tail, dtype=own_padded.dtype, device=own_padded.device)
    if own_padded.is_cuda:
        dist.all_gather_into_tensor(buf, own_padded.contiguous())  # one NCCL all-gather, no copies
    else:  # gloo (CPU tests)
        dist.all_gather(list(buf.unbind(0)), own_padded.contiguous())
    n = active_ranks(ydim, size)
    if n == size and ydim == size * rows_max:
        return buf.view((ydim,) + tail)
    parts = [buf[r, :partition(ydim, size, r).rows] for r in range(n)]
    return torch.cat(parts, dim=0)
