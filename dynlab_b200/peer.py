"""Groups of GPUs of one box that write into each other's memory over NVLink.

A *group* is the set of GPUs one ``compute`` call is sharded over, in slab order.  Two ways to
get one (SURVEY §5 "distributed communication backend", §8e):

* **single process** (the reference's own shape: a plain blocking call that creates and tears
  down its parallelism inside, R:src/dynlab/diagnostics/_base_classes.py:139-148): one Python
  process drives every visible device; ``sharding.set_devices`` picks them.  All members are
  *local*.
* **one process per GPU** (``torchrun``; ``torch.distributed`` initialised with any backend):
  each rank drives one device; exactly one member is local.  torch.distributed is used for
  plumbing only (exchanging 64-byte IPC handles once per buffer, broadcasting the partition).

Either way the data path is the same C-ABI (``csrc/peer.cu``): *symmetric* buffers — one
``dlb_peer_alloc`` block per member that every member can address (peer access / CUDA IPC) —
into which ``dlb_gather_rows`` / ``dlb_halo_exchange`` push rows with a copy engine and release
a generation flag, and ``dlb_wait_flags`` on the receiving stream.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.distributed as dist

from . import _native as nat
from ._engine import Engine

FLAG_WORDS = 32          # uint64 per member: [0, 16) gather flags by source rank, 16/17 halo
FLAG_HALO_UP = 16        # written by the neighbour above (smaller rows)
FLAG_HALO_DOWN = 17      # written by the neighbour below
GEN_STRIDE = 64          # flag value = generation * GEN_STRIDE + (chunk index + 1 | DONE)
DONE = GEN_STRIDE - 1
RESULT_DEPTH = 3         # rotating result buffers: a result survives the next call (see group docs)
WAIT_TIMEOUT_S = 30.0


@dataclass
class Member:
    rank: int            # position in the group = slab index
    engine: Engine       # the device (and its workspace / streams) this member runs on
    # one pinned (device-visible) int: a wait kernel that gave up stores 1 + flag index here, so
    # the host can notice without synchronising the device
    timeout: torch.Tensor = field(default_factory=lambda: torch.zeros(1, dtype=torch.int32,
                                                                      pin_memory=True))

    @property
    def device(self) -> torch.device:
        return self.engine.device


class PeerTransportUnavailable(RuntimeError):
    """The GPUs of this job cannot map each other's memory (CUDA IPC / peer access refused)."""


class _Raw:
    """Lets torch alias a raw device allocation (``torch.as_tensor`` reads this protocol)."""

    def __init__(self, ptr: int, nbytes: int, keep=None):
        self.keep = keep
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1",
                                         "data": (int(ptr), False), "version": 2}


class SymBuf:
    """One equally sized device block per member; ``addr[r][j]`` is the address of member j's
    block as member r's process/device sees it (``r`` local)."""

    def __init__(self, nbytes, addr, own_tensor, closers):
        self.nbytes, self.addr, self._own, self._closers = nbytes, addr, own_tensor, closers

    def own(self, member: Member) -> torch.Tensor:
        """uint8 tensor over the member's own block."""
        return self._own[member.rank]

    def view(self, member: Member, dtype, shape, byte_offset=0) -> torch.Tensor:
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        return self._own[member.rank][byte_offset:byte_offset + n].view(dtype).view(*shape)


class Group:
    """Members in slab order.  ``local`` are the ones this process drives."""

    def __init__(self, size: int, local: list[Member], multiprocess: bool):
        self.size, self.local, self.multiprocess = size, local, multiprocess
        if size > nat.MAX_PEERS:
            raise ValueError(f"a group holds at most {nat.MAX_PEERS} GPUs")
        self.generation = 0
        self._sym: dict = {}
        if not multiprocess:  # let every local device address every other one
            devs = sorted({m.device.index for m in local})
            for a in devs:
                for b in devs:
                    if a != b:
                        nat.check(nat.lib.dlb_peer_enable(a, b))
        self.flags = self.symmetric("flags", FLAG_WORDS * 8)

    # -- symmetric memory ---------------------------------------------------------------
    def symmetric(self, tag, nbytes: int) -> SymBuf:
        """Collective over the group (every process calls it with the same arguments)."""
        key = (tag, int(nbytes))
        buf = self._sym.get(key)
        if buf is not None:
            return buf
        own_ptr, own_tensor = {}, {}
        for m in self.local:
            with torch.cuda.device(m.device):
                p = C.c_void_p()
                nat.check(nat.lib.dlb_peer_alloc(int(nbytes), C.byref(p)))
                own_ptr[m.rank] = p.value
                own_tensor[m.rank] = torch.as_tensor(_Raw(p.value, max(nbytes, 1)), device=m.device)
        closers = []
        if self.multiprocess:
            (m,) = self.local
            with torch.cuda.device(m.device):
                handle = C.create_string_buffer(nat.IPC_HANDLE_BYTES)
                nat.check(nat.lib.dlb_ipc_export(own_ptr[m.rank], handle))
                handles = [None] * self.size
                dist.all_gather_object(handles, handle.raw)
                row, problem = [], None
                for j in range(self.size):
                    if j == m.rank:
                        row.append(own_ptr[m.rank])
                        continue
                    p = C.c_void_p()
                    rc = nat.lib.dlb_ipc_open(handles[j], C.byref(p))
                    if rc != nat.DLB_OK:
                        problem = (nat.lib.dlb_last_error() or b"").decode()
                        break
                    row.append(p.value)
                    closers.append(p.value)
            addr = {m.rank: row}
            # doubles as the barrier "nobody pushes into a block its owner has not finished
            # zeroing"; every rank learns whether every mapping worked
            problems = [None] * self.size
            dist.all_gather_object(problems, problem)
            if any(problems):
                for q in closers:
                    nat.lib.dlb_ipc_close(q)
                dist.barrier()  # every importer has unmapped before an owner frees
                nat.lib.dlb_peer_free(own_ptr[m.rank])
                raise PeerTransportUnavailable(next(q for q in problems if q))
        else:
            row = [own_ptr[r] for r in range(self.size)]
            addr = {m.rank: row for m in self.local}
        buf = self._sym[key] = SymBuf(int(nbytes), addr, own_tensor, closers)
        return buf

    def release(self) -> None:
        """Free every symmetric block of this group (collective in the torchrun case).  Tensors
        that alias them - results of earlier ``compute_device`` calls, ``flow_map`` rows - must
        not be used afterwards.  Blocks are otherwise kept for the life of the process (per grid
        shape: three field-sized blocks and two slab-sized flow-map blocks per member), because
        returned tensors alias them."""
        for m in self.local:
            torch.cuda.synchronize(m.device)
        if self.multiprocess:
            dist.barrier()  # nobody is still pushing into a block that is about to go
        for key, buf in list(self._sym.items()):
            if buf is self.flags:
                continue
            for m in self.local:
                with torch.cuda.device(m.device):
                    for p in buf._closers:
                        nat.check(nat.lib.dlb_ipc_close(p))
                    buf._closers = []
            if self.multiprocess:
                dist.barrier()  # every importer has unmapped before the owner frees
            for m in self.local:
                with torch.cuda.device(m.device):
                    nat.check(nat.lib.dlb_peer_free(buf.addr[m.rank][m.rank]))
            del self._sym[key]

    def next_generation(self) -> int:
        self.generation += 1
        return self.generation

    # -- data path (all enqueue-only) ---------------------------------------------------
    def push_rows(self, m: Member, buf: SymBuf, byte_offset: int, nbytes: int, flag_value: int,
                  stream, dests=None) -> None:
        """``dlb_gather_rows``: bytes [byte_offset, +nbytes) of member m's block go to the same
        place in the block of every member in ``dests`` (default: all); then ``flag_value`` is
        released in those members' gather flag of source ``m.rank`` (own included: it orders
        nothing, it lets the receiver wait on ``size`` flags uniformly)."""
        row, frow = buf.addr[m.rank], self.flags.addr[m.rank]
        n = self.size
        on = set(range(n) if dests is None else dests)
        dst = (C.c_void_p * n)(*[row[j] + byte_offset if j in on else None for j in range(n)])
        # flag_value None: copies only (an earlier chunk; only the member's last departure
        # raises the flags, and the signal kernel needs an SM slot the integrator may not free
        # before its chunk ends)
        flg = (C.c_void_p * n)(*[frow[j] + 8 * m.rank if (j in on and flag_value is not None)
                                 else None for j in range(n)])
        with torch.cuda.device(m.device):
            nat.check(nat.lib.dlb_gather_rows(row[m.rank] + byte_offset, int(nbytes), dst, n, flg,
                                              int(flag_value or 0), stream))
        m.engine.launches += int(flag_value is not None)

    def push_targets(self, m: Member, buf: SymBuf, byte_offset: int, dests=None):
        """(peer addresses, flag addresses) for ``dlb_ftle_stencil_gather``: where member m's rows
        at ``byte_offset`` live in the other destinations' blocks, and every destination's gather
        flag of source ``m.rank`` (own included)."""
        row, frow = buf.addr[m.rank], self.flags.addr[m.rank]
        on = sorted(set(range(self.size) if dests is None else dests))
        return ([row[j] + byte_offset for j in on if j != m.rank],
                [frow[j] + 8 * m.rank for j in on])

    def wait_gathered(self, m: Member, generation: int, stream) -> None:
        """``dlb_wait_flags`` on member m's stream: every member's DONE flag of this generation."""
        base = self.flags.addr[m.rank][m.rank]
        with torch.cuda.device(m.device):
            nat.check(nat.lib.dlb_wait_flags(base, self.size, generation * GEN_STRIDE + DONE,
                                             WAIT_TIMEOUT_S, m.timeout.data_ptr(), stream))
        m.engine.launches += 1

    def halo_push(self, m: Member, buf: SymBuf, row_bytes: int, slabs, generation: int,
                  stream) -> None:
        """``dlb_halo_exchange``.  ``buf`` holds every member's local rows
        ``[slab.nloc][row_bytes]`` (``slabs``: the partition, by rank): the first own row goes
        into the bottom halo slot of the member above, the last own row into the top halo slot
        (local row 0) of the member below, then their halo flags get ``generation``."""
        row, frow = buf.addr[m.rank], self.flags.addr[m.rank]
        r, me = m.rank, slabs[m.rank]
        if not me.rows or not (me.halo_lo or me.halo_hi):
            return
        up = r - 1 if me.halo_lo else None
        down = r + 1 if me.halo_hi else None
        up_slot = row[up] + (slabs[up].nloc - 1) * row_bytes if up is not None else None
        down_slot = row[down] if down is not None else None
        up_flag = frow[up] + 8 * FLAG_HALO_DOWN if up is not None else None
        down_flag = frow[down] + 8 * FLAG_HALO_UP if down is not None else None
        with torch.cuda.device(m.device):
            nat.check(nat.lib.dlb_halo_exchange(row[r], int(row_bytes), me.halo_lo,
                                                me.halo_lo + me.rows - 1, up_slot, down_slot,
                                                up_flag, down_flag, int(generation), stream))
        m.engine.launches += 1

    def halo_wait(self, m: Member, slabs, generation: int, stream) -> None:
        """``dlb_wait_flags`` for the neighbours' rows of ``halo_push``."""
        me, base = slabs[m.rank], self.flags.addr[m.rank][m.rank]
        with torch.cuda.device(m.device):
            for has, word in ((me.halo_lo, FLAG_HALO_UP), (me.halo_hi, FLAG_HALO_DOWN)):
                if me.rows and has:
                    nat.check(nat.lib.dlb_wait_flags(base + 8 * word, 1, int(generation),
                                                     WAIT_TIMEOUT_S, m.timeout.data_ptr(), stream))
                    m.engine.launches += 1

    def pull(self, m: Member, dst: torch.Tensor, buf: SymBuf, src_rank: int, byte_offset: int,
             stream) -> None:
        """One-sided read of another member's block into a tensor of member m (``dlb_peer_copy``)."""
        nbytes = dst.numel() * dst.element_size()
        with torch.cuda.device(m.device):
            nat.check(nat.lib.dlb_peer_copy(dst.data_ptr(), buf.addr[m.rank][src_rank] + byte_offset,
                                            nbytes, stream))

    def check_timeouts(self) -> None:
        """Raise if a wait kernel of an earlier call gave up (a peer died or never called).  Reads
        pinned host words only: no device synchronisation."""
        for m in self.local:
            v = int(m.timeout[0])
            if v:
                m.timeout[0] = 0
                raise RuntimeError(f"dynlab_b200: GPU {m.rank} of the group waited "
                                   f"{WAIT_TIMEOUT_S:.0f} s for rows from a peer (flag {v - 1}): "
                                   "a rank died or did not make the same compute call")


# ---------------------------------------------------------------------------------------
_groups: dict = {}
_devices = "auto"
AUTO_MIN_SEEDS = 1 << 22   # single process: smaller grids stay on one device


def set_devices(spec) -> None:
    """Devices a plain (non-torchrun) process shards a Lagrangian ``compute`` over:
    ``"auto"`` (default) every visible device for grids of at least ``AUTO_MIN_SEEDS`` nodes,
    ``"all"`` every visible device always, an int n the first n devices, a list of device
    indices (an index may repeat: several members on one device, used by the 1-GPU tests),
    ``None`` / 1 the current device only."""
    global _devices
    if spec in ("auto", "all") or spec is None:
        _devices = spec or 1
    elif isinstance(spec, int):
        if spec < 1:
            raise ValueError("need at least one device")
        _devices = spec
    else:
        spec = [int(d) for d in spec]
        if not spec:
            raise ValueError("need at least one device")
        _devices = spec


def devices_spec():
    return _devices


def _single_process_devices(n_seeds):
    spec = _devices
    if spec == "auto":
        import os

        # a rank of a torchrun job owns one device even while sharding is switched off
        if (dist.is_available() and dist.is_initialized()) or int(os.environ.get("WORLD_SIZE", "1")) > 1:
            return None
        n = torch.cuda.device_count()
        return list(range(n)) if (n > 1 and n_seeds is not None and n_seeds >= AUTO_MIN_SEEDS) else None
    if spec == "all":
        n = torch.cuda.device_count()
        return list(range(n)) if n > 1 else None
    if isinstance(spec, int):
        if spec == 1:
            return None
        if spec > torch.cuda.device_count():
            raise ValueError(f"{spec} devices requested, {torch.cuda.device_count()} visible")
        return list(range(spec))
    return list(spec)


def release_all() -> None:
    """``Group.release`` for every group of this process."""
    for g in _groups.values():
        g.release()


def current_group(n_seeds=None):
    """The group a Lagrangian ``compute`` of ``n_seeds`` nodes runs on, or None for the plain
    one-device path."""
    from . import sharding

    rank, size = sharding.world()
    if size > 1:
        key = ("mp", rank, size, torch.cuda.current_device())
        g = _groups.get(key)
        if g is None:
            try:
                g = _groups[key] = Group(size, [Member(rank, Engine.get())], True)
            except PeerTransportUnavailable as exc:
                # decided collectively (Group.symmetric): every rank lands here together
                import warnings

                warnings.warn(f"dynlab_b200: peer-memory transport unavailable ({exc}); the ranks "
                              "fall back to NCCL for halo and gather", RuntimeWarning)
                sharding.set_transport("nccl")
                return None
        return g
    devs = _single_process_devices(n_seeds)
    if devs is None:
        return None
    key = ("sp", tuple(devs))
    g = _groups.get(key)
    if g is None:
        seen, members = set(), []
        for r, d in enumerate(devs):
            eng = Engine.get(torch.device("cuda", d)) if d not in seen else Engine(d)
            seen.add(d)
            members.append(Member(r, eng))
        g = _groups[key] = Group(len(devs), members, False)
    return g
