"""Slab decomposition and halo exchange (one process per GPU).

New work with no counterpart in the reference, which is single-address-space
OpenMP only.  Every statement of the DSL is a point-wise
map with literal relative offsets, so a grid is split into contiguous slabs
along axis 0; rank r owns rows ``slab_range(n0, r, P)`` plus ``ghost`` rows on
either side of every time level.  The ghost rows are exactly the zero rows a
single-GPU level already carries (xgrid_b200/grid.py), so kernels are
unchanged: on a sharded grid the launcher refreshes them from the neighbours
before a sweep reads a level at a non-zero axis-0 offset, and only if that
level was written (or uploaded) since its last exchange.

Transport: NCCL ``ncclSend/ncclRecv`` grouped per exchange and enqueued on the
compute stream through the C ABI (``xgb_halo_exchange``); the NCCL unique id is
distributed through ``torch.distributed``'s store.  The planner and partition
logic are transport-agnostic (tests drive them over gloo on CPU).
"""
from __future__ import annotations

import ctypes
import os

from .log import Logger

_log = Logger("xgrid.dist")


def slab_range(n0: int, rank: int, world: int) -> tuple[int, int]:
    """Rows [lo, hi) of axis 0 owned by ``rank`` (remainder spread over the first ranks)."""
    base, rem = divmod(n0, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class Topology:
    """Ranks along axis 0: an open chain (overstep 'none' / 'limit': the global ends have no
    neighbour, -1), or a ring for overstep 'wrap' (rank 0's lower neighbour is rank P-1, so its lower
    ghost rows hold the LAST rows of the global grid and the periodic wrap needs no special case)."""

    def __init__(self, rank: int, world: int, ring: bool = False) -> None:
        self.rank, self.world, self.ring = rank, world, bool(ring) and world > 1
        if self.ring:
            self.lo_rank, self.hi_rank = (rank - 1) % world, (rank + 1) % world
        else:
            self.lo_rank = rank - 1 if rank > 0 else -1
            self.hi_rank = rank + 1 if rank < world - 1 else -1

    @property
    def sharded(self) -> bool:
        return self.world > 1


_topology: Topology | None = None
_transport = None


def topology() -> Topology:
    """Process topology: from torch.distributed when initialised, else single rank."""
    global _topology, _transport
    from .config import _config
    ring = _config is not None and _config.overstep == "wrap"
    if _topology is not None and _topology.ring != (ring and _topology.world > 1):
        # init() was called again with another overstep mode: same ranks, other neighbours.  The
        # NCCL communicator is kept (it does not depend on the neighbour relation).
        _topology = Topology(_topology.rank, _topology.world, ring)
        if _transport is not None:
            _transport.topo = _topology
    if _topology is None:
        rank, world = 0, 1
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                rank, world = dist.get_rank(), dist.get_world_size()
        except ImportError:
            pass
        _topology = Topology(rank, world, ring)
    return _topology


def reset() -> None:
    global _topology, _transport
    _topology, _transport = None, None


class HaloPlan:
    """Which levels need their ghost rows refreshed before a group runs.

    Pure bookkeeping: ``_Level.halo_rows`` is the number of ghost rows (counted from the
    slab outwards) that hold the neighbours' current rows -- 0 after the level has been
    written or uploaded.  A read that reaches ``h`` rows across the slab boundary finds the
    level stale when ``halo_rows < h``: a level exchanged at depth 1 is NOT fresh for a
    later group that reads it at depth 2."""

    @staticmethod
    def stale(reads: list) -> list:
        """reads: [(grid, level_object, halo_rows)] -> the subset to exchange."""
        want: dict = {}
        for grid, lv, h in reads:
            if h <= 0 or not getattr(grid, "sharded", False):
                continue
            if getattr(lv, "halo_rows", 0) >= h:
                continue
            hit = want.get(id(lv))
            if hit is None or hit[2] < h:          # several reads of one level: the deepest one decides
                want[id(lv)] = (grid, lv, h)
        return list(want.values())


class NcclTransport:
    """ncclSend/ncclRecv neighbour exchange through the C ABI (xgb_halo_exchange)."""

    def __init__(self, topo: Topology) -> None:
        from .runtime import shim
        import torch.distributed as dist
        self.topo = topo
        self.shim = shim
        self.rt = shim.Runtime.get()
        self._comm_stream = None
        self._events: list = []
        self._next = 0
        lib = shim.lib()
        nccl_path = ""
        try:     # use the very libnccl torch already mapped into this process
            with open("/proc/self/maps") as maps:
                for line in maps:
                    if "libnccl.so" in line:
                        nccl_path = line.split()[-1]
                        break
        except OSError:
            pass
        if not nccl_path:
            try:
                import nvidia.nccl
                for base in list(getattr(nvidia.nccl, "__path__", [])):
                    cand = os.path.join(base, "lib", "libnccl.so.2")
                    if os.path.exists(cand):
                        nccl_path = cand
                        break
            except ImportError:
                pass
        shim.check(lib.xgb_nccl_load(nccl_path.encode()))
        store = dist.distributed_c10d._get_default_store()
        key = "xgrid_b200/nccl_id"
        if topo.rank == 0:
            buf = ctypes.create_string_buffer(128)
            shim.check(lib.xgb_nccl_unique_id(ctypes.cast(buf, ctypes.c_void_p)))
            store.set(key, buf.raw)
        raw = store.get(key)
        buf = ctypes.create_string_buffer(bytes(raw), 128)
        shim.check(lib.xgb_nccl_init(ctypes.cast(buf, ctypes.c_void_p), topo.rank, topo.world))
        _log.info(f"NCCL halo transport ready: rank {topo.rank}/{topo.world}")

    def exchange_async(self, items: list) -> None:
        """Exchange on the high-priority comm stream, ordered after everything enqueued on the
        compute stream so far; readers wait on ``level.halo_event`` (launch.py)."""
        if not items:
            return
        rt = self.rt
        if self._comm_stream is None:
            self._comm_stream = rt.stream_create(high_priority=True)
        ready = self._event()
        rt.event_record_raw(ready, 0)
        rt.stream_wait_event(self._comm_stream, ready)
        self.exchange(items, self._comm_stream)
        done = self._event()
        rt.event_record_raw(done, self._comm_stream)
        for _, lv, _ in items:
            lv.halo_event = done

    def all_agree(self, flag: bool) -> bool:
        """Logical AND of `flag` over all ranks (host-side; one tiny all-reduce through torch.distributed's default
        group).  Every decision that changes WHICH exchanges a call issues -- running a solver loop as fused
        pairs, say -- must be the same on every rank, or the neighbours' sends and receives no longer pair up."""
        import torch
        import torch.distributed as dist
        if dist.get_backend() == "nccl":
            t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=torch.device("cuda", self.rt.device))
        else:
            t = torch.tensor([1 if flag else 0], dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(int(t.item()))

    def fence_compute(self) -> None:
        """Order the compute stream behind everything enqueued on the comm stream so far (device-side
        wait only): called before a buffer a halo exchange may still be reading is recycled."""
        if self._comm_stream is not None:
            ev = self._event()
            self.rt.event_record_raw(ev, self._comm_stream)
            self.rt.stream_wait_event(0, ev)

    def _event(self) -> int:
        """Round-robin pool of events (an event is reused long after its waiters ran)."""
        if len(self._events) < 64:
            self._events.append(self.rt.event_create())
            return self._events[-1]
        self._next = (self._next + 1) % len(self._events)
        return self._events[self._next]

    def exchange_bytes(self, data_ptr: int, data_bytes: int, halo_bytes: int, stream: int = 0) -> None:
        """Halo exchange of a raw 1-D byte array laid out [halo | data | halo] (device masks)."""
        d = self.shim.HaloDesc()
        d.bytes = halo_bytes
        d.lo_rank, d.hi_rank = self.topo.lo_rank, self.topo.hi_rank
        d.send_lo, d.recv_lo = data_ptr, data_ptr - halo_bytes
        d.send_hi, d.recv_hi = data_ptr + data_bytes - halo_bytes, data_ptr + data_bytes
        self.shim.check(self.shim.lib().xgb_halo_exchange(self.shim.C.byref(d), 1, stream))

    def exchange(self, items: list, stream: int = 0) -> None:
        """items: [(grid, level, h)] -- refresh h ghost rows on both sides of each level, plus the grid's overhang
        (``Grid._need_halo_over``): the few elements of the row one further out that diagonal taps reach."""
        if not items:
            return
        descs = (self.shim.HaloDesc * len(items))()
        for d, (grid, lv, h) in zip(descs, items):
            row = grid.stride0 * grid.itemsize
            over = getattr(grid, "_halo_over", 0) * grid.itemsize
            n0 = grid.shape[0]
            d.bytes = h * row + over
            d.lo_rank, d.hi_rank = self.topo.lo_rank, self.topo.hi_rank
            d.send_lo = lv.dev                            # my first h rows (+) -> rank-1's upper ghost
            d.recv_lo = lv.dev - h * row - over           # my lower ghost      <- rank-1's last h rows (+)
            d.send_hi = lv.dev + (n0 - h) * row - over    # my last h rows (+)  -> rank+1's lower ghost
            d.recv_hi = lv.dev + n0 * row                 # my upper ghost      <- rank+1's first h rows (+)
        self.shim.check(self.shim.lib().xgb_halo_exchange(descs, len(items), stream))
        for _, lv, h in items:
            lv.halo_rows = h


def transport():
    global _transport
    if _transport is None:
        _transport = NcclTransport(topology())
    return _transport
