"""``xgrid.Grid`` with device-resident time levels.

Public surface and ring semantics follow the reference
(xgrid/xgrid/__init__.py:21-86): ``Grid(shape, dtype)``, ``.now``, ``[]``,
``.boundary`` (a mutable int32 NumPy array), ``.shape``, ``.dimension``,
``fill``; ``_extend_time(depth)`` appends zero levels and truncates to
``depth`` (:43-47); ``_op_invoke`` rotates the oldest buffer to the front
(:52-54).  What changed is where the data lives:

* every level is an HBM allocation laid out C-order with ``ghost`` zero rows
  on both sides of axis 0 (these become the halo rows of a slab when the grid
  is sharded) and a small linear slack, so a stencil tap is one signed linear
  offset and never faults;
* rotation swaps handles, never data;
* the host sees data only through ``.now`` / ``[]`` / ``_data`` which copy
  device -> host on demand and hand the level back to the host (the user may
  write through the returned array); the next kernel call re-uploads it;
* the int32 ``boundary`` array is compiled on upload into a uint8 device mask,
  per-128-point "any non-zero" flags (packed on the host, runtime/maskpack.py) and compacted index lists.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import hostview
from .log import Logger
from .types import Grid as GridT, Structure, Value, parse_annotation

_schedule = None


def _flush() -> None:
    """Execute deferred kernel calls (temporal-blocking queue) before grid state is observed."""
    global _schedule
    if _schedule is None:
        from .lang import schedule
        _schedule = schedule
    if _schedule._PENDING is not None:
        _schedule.flush_pending()


SLACK = 64          # elements of linear slack before / after the padded array
CHUNK = 128         # points per mask flag (XGB_CHUNK in xgb_stencil.cuh)
MASK_GHOST = 64     # bytes of "outside" (255) mask on both sides of the device mask
ALIGN = 256


def parse_numpy_dtype(t: Value):
    if isinstance(t, Structure):
        return np.dtype([(n, parse_numpy_dtype(ft)) for n, ft in t.elements], align=True)
    return np.dtype(t.np_dtype)


PACK_OVERLAP_MIN = 1 << 22   # points; below this the mask is packed in line


class _PackJob:
    """maskpack.pack on a helper thread (the C routine releases the GIL and runs its own worker threads)."""

    def __init__(self, boundary: np.ndarray, n_padded: int) -> None:
        import threading
        from .runtime import maskpack
        self._out, self._err = None, None
        src = np.ascontiguousarray(boundary, np.int32)

        def run() -> None:
            try:
                self._out = maskpack.pack(src, n_padded)
            except BaseException as e:      # noqa: BLE001 -- re-raised on the calling thread
                self._err = e

        self._thread = threading.Thread(target=run, daemon=True)
        self._thread.start()

    def result(self):
        self._thread.join()
        if self._err is not None:
            raise self._err
        return self._out


class _Level:
    """One time level: device allocation + optional host mirror."""
    __slots__ = ("dev", "host", "where", "raw", "halo_rows", "halo_event", "pinned", "xfers", "view")

    def __init__(self, host=None) -> None:
        self.dev = 0            # device pointer of the first *real* element (0 = not allocated)
        self.host = host        # np.ndarray or None (an all-zero level never touched by the host)
        self.where = "host" if host is not None else "zero"
        self.raw = 0            # base of the padded device allocation
        self.halo_rows = 0      # ghost rows (this many, next to the slab) hold the neighbours' current rows; 0 = stale
        self.pinned = 0         # address of the page-locked host mirror (0 = pageable)
        self.halo_event = 0     # event of a halo exchange still in flight on the comm stream
        self.xfers = 0          # address of the mirror that has been transferred once (pin on 2nd use)
        self.view = None        # (mirror, write-tracking view of it) handed out by .now / _data


class Grid:
    _instances = 0

    def __init__(self, shape, dtype) -> None:
        self.logger = Logger(self)
        elem = parse_annotation(dtype)
        if not isinstance(elem, Value):
            self.logger.dead(f"Grid element should be value instead of '{elem}'")
        if isinstance(shape, int):
            shape = (shape,)
        self.element = elem
        self.global_shape = tuple(int(s) for s in shape)
        self.row_range = (0, self.global_shape[0])
        self.sharded = False
        from .config import _config
        if _config is not None and _config.distributed:
            # slab decomposition along axis 0: this process holds rows [lo, hi) (xgrid_b200/dist.py);
            # .now / .boundary / [] address the local slab
            from . import dist
            topo = dist.topology()
            if topo.sharded:
                self.row_range = dist.slab_range(self.global_shape[0], topo.rank, topo.world)
                self.sharded = True
        shape = (self.row_range[1] - self.row_range[0],) + self.global_shape[1:]
        self.shape = tuple(int(s) for s in shape)
        self.numpy_dtype = parse_numpy_dtype(elem)
        self.typing = GridT(elem, len(self.shape))
        self.size = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        self.itemsize = self.numpy_dtype.itemsize
        self.stride0 = self.size // self.shape[0] if self.shape and self.shape[0] else 1

        self._ring: list[_Level] = [_Level(np.zeros(self.shape, self.numpy_dtype))]
        self._scratch: _Level | None = None
        self._spares: list[_Level] = []
        self._boundary = np.zeros(self.shape, dtype=np.int32)
        self._boundary_view = None
        self._boundary_foreign = False  # the mask array was supplied by the program (`g.boundary = arr`)
        self._mask_touched = True
        self._mask_snapshot = None
        self._mask_raw = 0
        self._mask_dev = 0
        self._flags_dev = 0
        self._mask_any = False
        self._lists: dict = {}          # mask value -> (device ptr, count)
        self._mask_version = 0
        self._pair_ok: dict = {}        # (pair id, mask version) -> fused-pair eligibility (lang/jacobi2.py)
        self._mask_hist = None
        self._ghost = 1                 # zero rows on both sides of axis 0
        self._halo_over = 0             # slabs: elements beyond whole ghost rows that every exchange also carries
        self._halo_reserved = 0         # slabs: bytes per face the halo transport has been asked to make room for
        Grid._instances += 1
        self._serial = Grid._instances  # part of every recorded CUDA graph's key: device addresses can repeat
        self._allocs: list[int] = []    # raw device allocations to free
        self._rt = None

    # ------------------------------------------------------------------ reference API
    @property
    def dimension(self) -> int:
        return len(self.shape)

    @property
    def boundary(self) -> np.ndarray:
        _flush()
        self._mask_touched = True
        if self._boundary_view is None or self._boundary_view[0] is not self._boundary:
            self._boundary_view = (self._boundary, hostview.make(self._boundary, self._mask_written))
        return self._boundary_view[1]

    def _mask_written(self, view=None) -> None:
        """A write through a `.boundary` array (possibly one the program kept from earlier) is about to
        happen: deferred calls still see the old mask, the next call re-compiles it if it changed."""
        _flush()
        self._mask_touched = True

    def _host_written(self, view) -> None:
        """A write through a `.now` / `_data` array is about to happen.  If the level it mirrors has newer
        data on the device (the array was kept across kernel calls), bring that back first; the host
        copy is the truth afterwards and the next kernel call uploads it -- like the reference, where
        the array IS the level's storage whatever ring position it has rotated to."""
        addr = view.__array_interface__["data"][0]
        for lv in [*self._ring, *self._spares, *([self._scratch] if self._scratch is not None else [])]:
            host = lv.host
            if host is None or not (host.ctypes.data <= addr < host.ctypes.data + max(host.nbytes, 1)):
                continue
            if lv.where == "device":
                _flush()
                if lv.where == "device":
                    self._download(lv)
            if lv.where != "zero":
                lv.where = "host"
            return

    @boundary.setter
    def boundary(self, value) -> None:
        _flush()
        arr = np.asarray(value, dtype=np.int32)
        if arr.shape != self.shape:
            self.logger.dead("boundary mask has an incompatible shape")
        self._boundary = np.ascontiguousarray(arr)
        # `g.boundary = my_array`: in the reference my_array IS the mask from then on.  If the assigned array could be
        # adopted as it is, later writes through `my_array` (which no view can announce) must still be seen: such a
        # grid re-compares its mask with the snapshot before every call (cost ~ one pass over the mask)
        self._boundary_foreign = isinstance(value, np.ndarray) and np.shares_memory(self._boundary, value)
        self._mask_touched = True

    @property
    def now(self) -> np.ndarray:
        return self._host_view(0)

    def _element_offset(self, key):
        """Linear element offset for a full integer index (one element), else None."""
        if isinstance(key, (int, np.integer)):
            key = (key,)
        if not (isinstance(key, tuple) and len(key) == self.dimension
                and all(isinstance(k, (int, np.integer)) for k in key)):
            return None
        off = 0
        for k, n in zip(key, self.shape):
            k = int(k)
            if k < 0:
                k += n
            if not 0 <= k < n:
                raise IndexError(f"index {key} is out of bounds for grid of shape {self.shape}")
            off = off * n + k
        return off

    def __getitem__(self, key):
        # element indexing on a device-resident level moves ONE element, not the level
        # (the reference forwards to the NumPy array, xgrid/xgrid/__init__.py:82-86)
        _flush()
        lv = self._ring[0]
        off = self._element_offset(key) if lv.where == "device" else None
        if off is None:
            return self.now[key]
        out = np.empty(1, self.numpy_dtype)
        rt = self._runtime()
        rt.d2h(out.ctypes.data, lv.dev + off * self.itemsize, self.itemsize)
        rt.sync()
        return out[0]

    def __setitem__(self, key, value) -> None:
        _flush()
        lv = self._ring[0]
        off = self._element_offset(key) if lv.where == "device" else None
        if off is None:
            self.now[key] = value
            return
        src = np.empty(1, self.numpy_dtype)
        src[0] = value
        rt = self._runtime()
        rt.h2d(lv.dev + off * self.itemsize, src.ctypes.data, self.itemsize)
        rt.sync()
        lv.halo_rows = 0

    def fill(self, data: np.ndarray, time: int = 0) -> None:
        if data.shape != self.shape or data.dtype != self.numpy_dtype:
            self.logger.dead("Unable to fill grid with incompatible shape or data type")
        _flush()
        k = abs(time)
        while len(self._ring) <= k:
            self._ring.append(_Level())
        lv = self._ring[k]
        lv.host = np.ascontiguousarray(data).copy()
        lv.where = "host"

    @property
    def _data(self) -> list:
        """Host copies of every ring level, newest first (reference attribute)."""
        _flush()
        return [self._host_view(k) for k in range(len(self._ring))]

    # ------------------------------------------------------------------ ring (operator.py:37-39)
    def _extend_time(self, depth: int) -> None:
        while len(self._ring) < depth:
            self._ring.append(_Level())
        for lv in self._ring[depth:]:
            self._release(lv)
        del self._ring[depth:]

    def _op_invoke(self, depth: int, tick: bool) -> None:
        _flush()
        self._extend_time(depth)
        if tick:
            self._ring.insert(0, self._ring.pop())

    # ------------------------------------------------------------------ host <-> device
    def _runtime(self):
        if self._rt is None:
            from .runtime.shim import Runtime
            self._rt = Runtime.get()
        return self._rt

    def _layout(self):
        # ghost rows + room for the tiled variant's halo'd bulk copies at the array ends
        extra = SLACK + (2 * self.shape[-1] + 2048 if self.dimension > 1 else 0)
        lead = (extra + self._ghost * self.stride0) * self.itemsize
        lead = (lead + ALIGN - 1) // ALIGN * ALIGN
        tail = (extra + self._ghost * self.stride0) * self.itemsize
        return lead, lead + self.size * self.itemsize + tail

    def _alloc_level(self, lv: _Level) -> None:
        rt = self._runtime()
        lead, total = self._layout()
        raw = rt.alloc(total)          # zero-filled
        self._allocs.append(raw)
        lv.dev = raw + lead
        lv.raw = raw

    def _release(self, lv: _Level) -> None:
        self._unpin(lv)
        if lv.dev and self._rt is not None:
            self._rt.free(lv.raw)
            if lv.raw in self._allocs:
                self._allocs.remove(lv.raw)
        lv.dev = 0

    PIN_MIN_BYTES = 1 << 20

    def _pin(self, lv: _Level) -> None:
        """Page-lock a host mirror so H2D / D2H run at full PCIe speed -- from its SECOND
        transfer on: registering costs about as much as one pageable copy, so one-shot
        uploads (initial conditions) and downloads (final result) stay pageable."""
        if lv.host is None or lv.host.nbytes < self.PIN_MIN_BYTES:
            return
        addr = lv.host.ctypes.data
        if lv.pinned == addr:
            return
        if lv.xfers != addr:            # first transfer of this buffer: remember it, do not pin yet
            lv.xfers = addr
            return
        self._unpin(lv)
        from .runtime import shim
        if shim.lib().xgb_host_register(ctypes.c_void_p(addr), lv.host.nbytes) == 0:
            lv.pinned = addr

    def _unpin(self, lv: _Level) -> None:
        if lv.pinned:
            from .runtime import shim
            shim.lib().xgb_host_unregister(ctypes.c_void_p(lv.pinned))
            lv.pinned = 0

    def _host_view(self, k: int) -> np.ndarray:
        _flush()
        lv = self._ring[k]
        if lv.where == "zero":
            lv.host = np.zeros(self.shape, self.numpy_dtype)
        elif lv.where == "device":
            self._download(lv)
        # the caller may write through the returned array: the host owns the level now
        lv.where = "host"
        if lv.view is None or lv.view[0] is not lv.host:
            lv.view = (lv.host, hostview.make(lv.host, self._host_written))
        return lv.view[1]

    def _download(self, lv: _Level) -> None:
        """Device -> host mirror of one level (blocking)."""
        if lv.host is None:
            lv.host = self._adopt_stale_mirror(lv)
        if lv.host is None:
            lv.host = np.empty(self.shape, self.numpy_dtype)
        rt = self._runtime()
        nbytes = self.size * self.itemsize
        if rt.STAGED and nbytes >= rt.STAGED_MIN and not lv.pinned:
            rt.d2h_staged(lv.host.ctypes.data, lv.dev, nbytes)      # pageable mirror, several host threads
        else:
            self._pin(lv)
            rt.d2h(lv.host.ctypes.data, lv.dev, nbytes)
        rt.sync()

    def _adopt_stale_mirror(self, lv: _Level):
        """A level without a host mirror takes over the mirror of a SPARE level (a buffer that left the
        ring when a several-steps launch wrote into fresh levels; its host copy is stale and no public
        accessor reaches it any more).  Downloading into memory that is already resident -- and possibly
        already page-locked -- avoids first-touch page faults on a fresh array (measured: 28 ms -> pageable
        copy time for a 128 MiB level).  Like the reference's ring, an array handed out by `.now` earlier may
        therefore be written again later."""
        # (only spares: a RING level keeps its own array, so that an array the program kept from `.now` stays the
        #  mirror of the level it was handed out for as that level rotates through the ring -- like the reference)
        for donor in self._spares:
            if donor.host is not None and donor.where == "device" and donor.host.shape == tuple(self.shape):
                host, donor.host = donor.host, None
                if donor.pinned:
                    lv.pinned, donor.pinned = donor.pinned, 0
                donor.xfers = 0
                donor.view = None
                return host
        return None

    def _to_device(self, lv: _Level) -> None:
        if lv.dev == 0:
            self._alloc_level(lv)
            if lv.where == "zero":
                lv.where = "device"
                return
        if lv.where == "host":
            host = np.ascontiguousarray(lv.host)
            lv.host = host
            rt, nbytes = self._runtime(), self.size * self.itemsize
            if rt.STAGED and nbytes >= rt.STAGED_MIN and not lv.pinned:
                rt.h2d_staged(lv.dev, host.ctypes.data, nbytes)         # returns once the mirror has been read
            else:
                self._pin(lv)
                rt.h2d(lv.dev, host.ctypes.data, nbytes)
                if lv.pinned:
                    rt.sync()               # the caller may modify the mirror right after
        elif lv.where == "zero":
            self._runtime().memset(lv.dev, 0, self.size * self.itemsize)
        lv.where = "device"
        lv.halo_rows = 0
        lv.halo_event = 0

    def _ensure_ghost(self, rows: int) -> None:
        """Widen the ghost band of every level to `rows` rows.  Device-resident levels are re-laid-out ON
        THE DEVICE (new allocation from the pool + one device-to-device copy of the real rows, enqueued on
        the compute stream); levels whose truth is on the host just drop their device copy and are uploaded
        into the new layout by the next call.  Scratch and spare levels hold no live data."""
        if rows <= self._ghost:
            return
        self._ghost = rows
        nbytes = self.size * self.itemsize
        for lv in self._ring:
            if not lv.dev:
                continue
            old_raw, old_dev = lv.raw, lv.dev
            if lv.where == "device":
                self._alloc_level(lv)                  # zero-filled, new layout
                self._runtime().d2d(lv.dev, old_dev, nbytes)
            else:
                lv.dev = lv.raw = 0
            self._runtime().free(old_raw)
            if old_raw in self._allocs:
                self._allocs.remove(old_raw)
            lv.halo_rows = 0
            lv.halo_event = 0
        if self._scratch is not None:
            self._release(self._scratch)
            self._scratch = None
        for lv in self._spares:
            self._release(lv)
        self._spares = []

    def _prepare_device(self, ghost_rows: int = 1) -> None:
        """Make every level and the mask resident before launches."""
        if ghost_rows > self._ghost:
            self._ensure_ghost(ghost_rows)
        job = None
        if (self._mask_touched and self._mask_snapshot is None and self.size >= PACK_OVERLAP_MIN
                and any(lv.where == "host" for lv in self._ring)):
            # first upload of a big grid: pack the mask on host threads WHILE the levels cross PCIe
            job = _PackJob(self._boundary, -(-self.size // CHUNK) * CHUNK)
        for lv in self._ring:
            if lv.where != "device":
                self._to_device(lv)
        if self.sharded:
            # the transport may need room for the deepest halo this layout allows (whole ghost rows plus the
            # overhang the padding can take): asked for HERE, before the call may be recorded into a graph -- and
            # after the levels have their device buffers, so that the transport's own allocations do not sit
            # between them
            need = (self._ghost * self.stride0 + SLACK + (2 * self.shape[-1] + 2048 if self.dimension > 1 else 0)) * self.itemsize
            if need > self._halo_reserved:
                from . import dist
                dist.transport().reserve(need)
                self._halo_reserved = need
        if self._mask_touched or self._boundary_foreign:
            self._upload_mask(job.result() if job is not None else None)

    def _scratch_level(self) -> _Level:
        if self._scratch is None or self._scratch.dev == 0:
            self._scratch = _Level()
            self._alloc_level(self._scratch)
            self._scratch.where = "device"
        return self._scratch

    def _spare_levels(self, n: int) -> list:
        """Extra device levels owned by the grid (outputs of the multi-step kernels)."""
        while len(self._spares) < n:
            lv = _Level()
            self._alloc_level(lv)
            lv.where = "device"
            self._spares.append(lv)
        return self._spares[:n]

    def _swap_scratch(self) -> None:
        """Jacobi double buffer: the scratch level becomes level 0."""
        self._ring[0], self._scratch = self._scratch, self._ring[0]
        self._ring[0].where = "device"

    def _arrangement(self) -> tuple:
        """(device pointers of the ring in order, scratch pointer): the buffer state a
        recorded CUDA graph was captured against / leaves behind."""
        return (tuple(lv.dev for lv in self._ring), self._scratch.dev if self._scratch is not None else 0)

    def _halo_state(self) -> tuple:
        """((device pointer, fresh ghost rows) of every ring level and the scratch level): slab grids only."""
        levels = [*self._ring, *([self._scratch] if self._scratch is not None else [])]
        return tuple((lv.dev, lv.halo_rows) for lv in levels) + ((0, self._halo_over),) * bool(self._halo_over)

    def _need_halo_over(self, elements: int) -> None:
        """A sweep reads `elements` linear positions past its deepest ghost row (diagonal taps at the first / last
        column): from now on every halo exchange of this grid carries that many elements more, into the padding
        next to the ghost rows.  Levels exchanged before count as stale."""
        if elements <= self._halo_over:
            return
        room = SLACK + (2 * self.shape[-1] + 2048 if self.dimension > 1 else 0)
        if elements > room:
            raise Exception(f"a stencil tap reaches {elements} elements past the ghost rows of a slab; the layout "
                            f"has room for {room}")
        self._halo_over = elements
        for lv in [*self._ring, *([self._scratch] if self._scratch is not None else [])]:
            lv.halo_rows = 0

    def _restore_halo_state(self, state: tuple) -> None:
        rows = dict(state)          # (an entry with key 0 is the overhang, not a level)
        for lv in [*self._ring, *([self._scratch] if self._scratch is not None else [])]:
            lv.halo_rows = rows.get(lv.dev, 0)
            lv.halo_event = 0

    def _restore_arrangement(self, arrangement: tuple) -> None:
        ring, scratch = arrangement
        pool = {lv.dev: lv for lv in self._ring}
        if self._scratch is not None:
            pool[self._scratch.dev] = self._scratch
        self._ring = [pool[d] for d in ring]
        self._scratch = pool[scratch] if scratch else None
        for lv in self._ring:
            lv.where = "device"

    # ------------------------------------------------------------------ mask compilation
    def _upload_mask(self, prepacked=None) -> None:
        self._mask_touched = False
        b = self._boundary
        if self._mask_snapshot is not None and np.array_equal(b, self._mask_snapshot):
            return                      # touched but unchanged: keep the version (graphs, index lists)
        rt = self._runtime()
        for ptr, _ in self._lists.values():
            if ptr:
                rt.free(ptr)
        self._lists = {}
        self._mask_snapshot = None
        self._mask_hist = None
        self._mask_version += 1
        nchunk = (self.size + CHUNK - 1) // CHUNK
        # pack on the host (several threads): one byte per point, chunk flags, histogram
        from .runtime import maskpack
        packed, flags, hist, bad = prepacked or maskpack.pack(np.ascontiguousarray(b, np.int32), nchunk * CHUNK)
        if bad:
            self.logger.dead("boundary mask values must lie in [0, 254] on the B200 backend")
        # host copy for change detection / index lists / the fused-pair check: the packed bytes
        self._mask_snapshot = packed[:self.size].reshape(self.shape)
        self._mask_hist = hist
        self._mask_any = bool(self.sharded or int(hist[0]) != self.size)
        if not self._mask_any:
            return                      # all-zero class: kernels get null mask pointers, nothing is uploaded
        # device layout: [MASK_GHOST bytes of 255 | mask | zero padding to a whole chunk | MASK_GHOST x 255];
        # 255 = "outside the grid" (never matches a statement); on a sharded 1-D grid the ghost bytes
        # are replaced by the neighbours' edge masks
        if not self._mask_raw:
            self._mask_raw = rt.alloc(MASK_GHOST + nchunk * CHUNK + MASK_GHOST)
            self._mask_dev = self._mask_raw + MASK_GHOST
            self._flags_dev = rt.alloc(nchunk)
            rt.memset(self._mask_raw, 255, MASK_GHOST)
            rt.memset(self._mask_dev + nchunk * CHUNK, 255, MASK_GHOST)
        maskpack.upload(rt, packed, flags, self._mask_dev, self._flags_dev)
        if self.sharded and self.dimension == 1:
            from . import dist
            # trailing ghost sits right after the last real byte on the neighbour's side
            dist.transport().exchange_bytes(self._mask_dev, self.size, MASK_GHOST)

    def _mask_count(self, k: int) -> int:
        """Number of points whose boundary value is k (histogram cached per mask upload)."""
        if self._mask_snapshot is None or not self._mask_any:
            return self.size if k == 0 else 0
        if self._mask_hist is None:
            self._mask_hist = np.bincount(self._mask_snapshot.reshape(-1), minlength=256)
        return int(self._mask_hist[k]) if 0 <= k < len(self._mask_hist) else 0

    def _index_list(self, k: int):
        """Device array of linear indices where boundary == k (cached per mask)."""
        hit = self._lists.get(k)
        if hit is None:
            snap = self._mask_snapshot if self._mask_snapshot is not None else self._boundary
            idx = np.flatnonzero(snap.reshape(-1) == k).astype(np.int64)
            ptr = 0
            if idx.size:
                rt = self._runtime()
                ptr = rt.alloc(idx.nbytes)
                rt.h2d(ptr, idx.ctypes.data, idx.nbytes)
                rt.sync()
            hit = (ptr, int(idx.size))
            self._lists[k] = hit
        return hit

    # ------------------------------------------------------------------ misc
    def device_pointer(self, level: int = 0) -> int:
        _flush()
        return self._ring[level].dev

    def __del__(self) -> None:
        try:
            rt = self._rt
            if rt is None:
                return
            for lv in self._ring:
                self._release(lv)
            if self._scratch is not None:
                self._release(self._scratch)
            for lv in self._spares:
                self._release(lv)
            for ptr, _ in self._lists.values():
                if ptr:
                    rt.free(ptr)
            if self._mask_raw:
                rt.free(self._mask_raw)
                rt.free(self._flags_dev)
        except Exception:
            pass
