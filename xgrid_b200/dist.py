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

Transports (same interface, ``_Transport``): NCCL ``ncclSend/ncclRecv`` grouped per exchange
through the C ABI (``xgb_halo_exchange``; the NCCL unique id is distributed through
``torch.distributed``'s store) -- the default -- and ``PeerTransport`` (``XGB_HALO=peer``): one
kernel per exchange that pushes the edge rows into the neighbours' mailboxes over NVLink, signals,
waits and copies into the ghost rows (``csrc/xgb_peer.cu``).  The planner and partition logic are
transport-agnostic (tests drive them over gloo on CPU).
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


class _Transport:
    """What the launcher asks of a halo transport; the two subclasses differ in `_issue` only."""

    def __init__(self, topo: Topology) -> None:
        from .runtime import shim
        self.topo = topo
        self.shim = shim
        self.rt = shim.Runtime.get()
        self._comm_stream = None
        self._events: list = []
        self._next = 0

    def _comm(self) -> int:
        if self._comm_stream is None:
            self._comm_stream = self.rt.stream_create(high_priority=True)
        return self._comm_stream

    def reserve(self, nbytes: int) -> None:
        """Room for one level's halo of `nbytes` per face (called outside any stream capture)."""

    def _issue(self, descs, n: int, stream: int) -> None:
        raise NotImplementedError

    def exchange_async(self, items: list) -> None:
        """Exchange on the high-priority comm stream, ordered after everything enqueued on the
        compute stream so far; readers wait on ``level.halo_event`` (launch.py)."""
        if not items:
            return
        rt = self.rt
        comm = self._comm()
        ready = self._event()
        rt.event_record_raw(ready, 0)
        rt.stream_wait_event(comm, ready)
        self.exchange(items, comm)
        done = self._event()
        rt.event_record_raw(done, comm)
        for _, lv, _ in items:
            lv.halo_event = done

    def all_agree(self, flag: bool) -> bool:
        """Logical AND of `flag` over all ranks (host-side).  Every decision that changes WHICH exchanges a call
        issues -- running a solver loop as fused pairs, say -- must be the same on every rank, or the neighbours'
        sends and receives no longer pair up."""
        return all_agree(flag, self.rt.device)

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
        self.reserve(halo_bytes)
        self._issue(self.shim.C.byref(d), 1, stream)

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
        self._issue(descs, len(items), stream)
        for _, lv, h in items:
            lv.halo_rows = h


class NcclTransport(_Transport):
    """ncclSend/ncclRecv neighbour exchange through the C ABI (xgb_halo_exchange)."""

    def __init__(self, topo: Topology) -> None:
        super().__init__(topo)
        import torch.distributed as dist
        shim = self.shim
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

    def _issue(self, descs, n: int, stream: int) -> None:
        self.shim.check(self.shim.lib().xgb_halo_exchange(descs, n, stream))


class PeerTransport(_Transport):
    """Neighbour exchange over peer memory: one kernel per exchange pushes the edge rows into the neighbours'
    mailboxes over NVLink, signals, waits for theirs and copies them into the ghost rows (csrc/xgb_peer.cu).

    The kernels pair up by exchange NUMBER, so (a) every exchange of the process runs on the one communication
    stream -- a request from another stream forks into it and joins back with events, which also records
    correctly into a CUDA graph -- and (b) the order of requests is the same on every rank (it is: §5 of DESIGN.md,
    "the exchange sequence must be the same on every rank").  A mailbox is a cuMemCreate allocation shared as a file
    descriptor: every rank serves its descriptor on an abstract Unix socket whose name is gathered once through
    torch.distributed, and a neighbour's mailbox is fetched and mapped on first use, so switching between the open
    chain and the ring (overstep="wrap") needs no new collective.  (Legacy CUDA IPC handles would be simpler, but
    opening one enables device-wide peer access.  Measured, the 3-D sweep runs ~3 % slower beside EITHER kind of
    mapping, which is why this transport is opt-in: DESIGN.md section 5.)"""

    _generation = 0         # mailboxes created by this process so far (part of the socket name)

    def __init__(self, topo: Topology) -> None:
        self._topo = topo
        super().__init__(topo)
        self._boxes: dict = {}
        self._handles: list = []
        self._slot = 0
        self._create(int(float(os.environ.get("XGB_PEER_SLOT_MB", "8")) * (1 << 20)))
        import atexit
        atexit.register(quiesce)
        _log.info(f"peer-memory halo transport ready: rank {topo.rank}/{topo.world}, slot {self._slot >> 10} KiB")

    @property
    def topo(self) -> Topology:
        return self._topo

    @topo.setter
    def topo(self, new: Topology) -> None:
        """Another neighbour relation (init() again with / without overstep="wrap"; every rank does it at the same
        point).  Exchange numbers pair up between neighbours that have exchanged from the start; a pair that has
        not -- the two ends of the chain when the ring closes -- owes each other no credits yet: start over at 0."""
        old, self._topo = self._topo, new
        if (old.lo_rank, old.hi_rank) != (new.lo_rank, new.hi_rank) or old.ring != new.ring:
            import torch.distributed as dist
            self.rt.device_sync()
            dist.barrier()
            self.shim.check(self.shim.lib().xgb_peer_reset())
            dist.barrier()

    def _create(self, slot_bytes: int) -> None:
        """Allocate this rank's mailbox, serve its file descriptor on an abstract Unix socket, and gather every
        rank's ticket.  Collective; the gather doubles as "every rank is listening"."""
        import socket
        import struct
        import threading
        import torch.distributed as dist
        lib = self.shim.lib()
        buf = ctypes.create_string_buffer(64)
        err = None
        PeerTransport._generation += 1
        self._address = "\0xgb_peer_%s_%d_%d_%d" % (os.environ.get("MASTER_PORT", "0"), self.topo.rank, os.getpid(),
                                                    PeerTransport._generation)
        try:
            self.shim.check(lib.xgb_peer_create(slot_bytes, ctypes.cast(buf, ctypes.c_void_p)))
            got = ctypes.c_uint64()
            self.shim.check(lib.xgb_peer_slot_bytes(ctypes.byref(got)))
            self._slot = int(got.value)
            fd = struct.unpack_from("<i", buf.raw, 4)[0]
            server = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            server.bind(self._address)
            server.listen(8)

            def serve() -> None:        # hands the mailbox's descriptor to whoever connects (neighbours, same box)
                while True:
                    try:
                        conn, _ = server.accept()
                    except OSError:     # closed: the mailbox is being replaced or the process exits
                        return
                    with conn:
                        try:
                            socket.send_fds(conn, [b"fd"], [fd])
                        except OSError:
                            pass

            self._server = server
            threading.Thread(target=serve, name="xgb-peer-fd", daemon=True).start()
        except Exception as e:          # noqa: BLE001 -- still join the collective below, then fail everywhere
            err = e
        tickets = [None] * self.topo.world
        dist.all_gather_object(tickets, None if err is not None else (buf.raw, self._address))
        if err is not None:
            raise err
        if any(t is None for t in tickets):
            raise Exception("another rank could not create its halo mailbox")
        self._handles = tickets

    def _stop_server(self) -> None:
        import socket
        try:
            self._server.shutdown(socket.SHUT_RDWR)      # wakes the thread blocked in accept()
        except OSError:
            pass
        self._server.close()

    def _box(self, rank: int):
        """Mapped mailbox of `rank` (None for -1: no neighbour on that side)."""
        if rank < 0:
            return None
        box = self._boxes.get(rank)
        if box is None:
            import socket
            import struct
            ticket, address = self._handles[rank]
            raw = ctypes.create_string_buffer(bytes(ticket), 64)
            fd = -1
            if rank != self.topo.rank:
                with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as c:
                    c.settimeout(30)
                    c.connect(address)
                    _, fds, _, _ = socket.recv_fds(c, 16, 1)
                if not fds:
                    raise Exception(f"rank {rank} did not hand over its mailbox descriptor")
                fd = fds[0]
                struct.pack_into("<i", raw, 4, fd)
            out = ctypes.c_void_p()
            try:
                self.shim.check(self.shim.lib().xgb_peer_open(ctypes.cast(raw, ctypes.c_void_p), ctypes.byref(out)))
            finally:
                if fd >= 0:
                    os.close(fd)            # the import holds its own reference
            box = self._boxes[rank] = out.value
        return box

    def reserve(self, nbytes: int) -> None:
        """A level's halo must fit one mailbox slot.  Growing the mailbox is a collective (every rank reaches this
        point with the same size: halo sizes do not depend on the rank) and drains the device first."""
        need = (int(nbytes) + 255) // 256 * 256
        if need <= self._slot:
            return
        import torch.distributed as dist
        lib = self.shim.lib()
        self.rt.device_sync()
        dist.barrier()                                   # nobody is still pushing into a mailbox
        for box in self._boxes.values():
            self.shim.check(lib.xgb_peer_close(ctypes.c_void_p(box)))
        self._boxes = {}
        dist.barrier()                                   # nobody still maps the mailbox about to be freed
        self._stop_server()
        self.shim.check(lib.xgb_peer_destroy())
        self._create(max(need, 2 * self._slot))
        _log.info(f"peer-memory halo mailbox grown: slot {self._slot >> 10} KiB")

    def _issue(self, descs, n: int, stream: int) -> None:
        rt, comm = self.rt, self._comm()
        if stream != comm:
            ev = self._event()
            rt.event_record_raw(ev, stream)
            rt.stream_wait_event(comm, ev)
        lo, hi = self._box(self.topo.lo_rank), self._box(self.topo.hi_rank)
        self.shim.check(self.shim.lib().xgb_peer_exchange(descs, n, ctypes.c_void_p(lo), ctypes.c_void_p(hi), comm))
        if stream != comm:
            ev = self._event()
            rt.event_record_raw(ev, comm)
            rt.stream_wait_event(stream, ev)


def quiesce() -> None:
    """Drain this rank's device, then wait for every rank to have done so: after that no exchange kernel anywhere
    still stores into a neighbour's mailbox, and the process may free its own (also registered with atexit)."""
    if _transport is None:
        return
    try:
        _transport.rt.device_sync()
    except Exception:       # noqa: BLE001 -- interpreter shutdown: nothing left to protect
        return
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.barrier()
    except Exception:       # noqa: BLE001
        pass


def all_agree(flag: bool, device: int = 0) -> bool:
    """Logical AND of `flag` over all ranks: one tiny host-visible all-reduce through torch.distributed."""
    import torch
    import torch.distributed as dist
    if dist.get_backend() == "nccl":
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=torch.device("cuda", device))
    else:
        t = torch.tensor([1 if flag else 0], dtype=torch.int32)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(int(t.item()))


def transport():
    """The process's halo transport.  ncclSend/ncclRecv by default; XGB_HALO=peer selects the one-kernel exchange
    over peer memory (PeerTransport) -- bit-identical results, a faster exchange of large faces (99 vs 123 us for
    32 MiB per direction on two B200s) but, measured, a 3 % slower 3-D sweep beside it, which is
    why it is not the default yet.  If the mailboxes cannot be set up on EVERY rank (no peer access between the
    GPUs, descriptors cannot be passed) all ranks fall back to NCCL together."""
    global _transport
    if _transport is None:
        topo = topology()
        if os.environ.get("XGB_HALO", "nccl") == "peer":
            from .runtime.shim import Runtime
            peer, why = None, ""
            try:
                peer = PeerTransport(topo)
                for r in {(topo.rank - 1) % topo.world, (topo.rank + 1) % topo.world}:
                    peer._box(r)                            # chain and ring neighbours alike
            except Exception as e:      # noqa: BLE001 -- whatever it is, this box cannot do it
                why = f"{type(e).__name__}: {e}"
            if all_agree(not why, Runtime.get().device):
                _transport = peer
                return _transport
            _log.warn(f"peer-memory halo transport unavailable ({why or 'on another rank'}); using NCCL")
        _transport = NcclTransport(topo)
    return _transport
