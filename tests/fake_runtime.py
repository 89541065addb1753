"""A recording stand-in for xgrid_b200.runtime.shim.Runtime -- TEST INFRASTRUCTURE ONLY.

It executes nothing: device memory is an address counter, copies are dropped, and every launch is
recorded as (kernel name, grid, block, parameter fields).  With it the host side of a kernel call --
ring rotation, scratch swaps, deferral and flushing, variant selection, CUDA-graph bookkeeping, the
memory pool -- runs on a machine without a GPU and can be asserted on.  Kernels are still generated and
compiled for sm_100a by the real NVRTC path.  Numerical results do not exist here; parity is the GPU
suite's job."""
from __future__ import annotations

import ctypes

import numpy as np

from xgrid_b200.runtime import shim


class Info:
    name = b"fake B200"
    sm_count = 148
    cc_major, cc_minor = 10, 0


class FakeRuntime(shim.Runtime):
    def __init__(self) -> None:           # no xgb_init
        self.device, self.info, self.sm_count = 0, Info(), 148
        self._pool, self._pool_bytes, self._sizes = {}, 0, {}
        self.next_ptr = 1 << 40
        self.live: dict = {}
        self.launches: list = []
        self.capturing = None
        self.graphs: dict = {}
        self.functions: dict = {}
        self.modules: dict = {}
        self.handles = 100
        self.copies: list = []
        self.real_allocs = self.real_frees = 0
        self.real_alloc_sizes: list = []

    # ---- memory
    def alloc(self, nbytes: int) -> int:
        nbytes = int(nbytes)
        cached = self._pool.get(nbytes)
        if cached:
            ptr = cached.pop()
            self._pool_bytes -= nbytes
        else:
            self.next_ptr += (max(nbytes, 256) + 4095) // 4096 * 4096 + (1 << 20)
            ptr = self.next_ptr
            self.real_allocs += 1
            self.real_alloc_sizes.append(nbytes)
        self._sizes[ptr] = nbytes
        self.live[ptr] = nbytes
        return ptr

    def free(self, ptr: int) -> None:
        nbytes = self._sizes.pop(ptr, 0)
        self.live.pop(ptr, None)
        if nbytes >= self.POOL_MIN and self._pool_bytes + nbytes <= self.POOL_CAP:
            self._fence_comm()
            self._pool.setdefault(nbytes, []).append(ptr)
            self._pool_bytes += nbytes
        else:
            self.real_frees += 1

    def trim_pool(self) -> None:
        self.real_frees += sum(len(c) for c in self._pool.values())
        self._pool.clear()
        self._pool_bytes = 0

    def memset(self, ptr, byte, nbytes, stream=0):
        pass

    def h2d(self, dst, src, nbytes, stream=0):
        self.copies.append(("h2d", dst, nbytes))

    def d2h(self, dst, src, nbytes, stream=0):
        self.copies.append(("d2h", src, nbytes))

    STAGED = False

    def d2d(self, dst, src, nbytes, stream=0):
        self.copies.append(("d2d", dst, nbytes))

    # ---- streams / events
    def sync(self, stream=0):
        self._flush_deferred()

    def device_sync(self):
        self._flush_deferred()

    def _handle(self) -> int:
        self.handles += 1
        return self.handles

    def stream_create(self, high_priority=False):
        return self._handle()

    def event_create(self):
        return self._handle()

    def event_record(self, ev, stream=0):
        self._flush_deferred()

    def event_record_raw(self, ev, stream=0):
        pass

    def event_sync(self, ev):
        pass

    def event_elapsed_ms(self, a, b):
        return 0.0

    def stream_wait_event(self, stream, ev):
        pass

    # ---- modules
    def module_load(self, image: bytes) -> int:
        assert image[:4] == b"\x7fELF"
        h = self._handle()
        self.modules[h] = image
        return h

    def get_function(self, module: int, name: str) -> int:
        # cuModuleGetFunction fails for a kernel that is not in THIS module: check the cubin's symbol text
        assert (".text." + name).encode() in self.modules[module], f"kernel {name} is not in the module it is looked up in"
        h = self._handle()
        self.functions[h] = name
        return h

    def set_dynamic_smem(self, fn, nbytes):
        assert nbytes <= 227 * 1024

    # ---- launches / graphs
    def launch(self, fn, grid, block, params, smem=0, stream=0):
        fields = {}
        for name, ctype in params._fields_:
            v = getattr(params, name)
            if isinstance(v, ctypes.Array):
                v = list(v)
            elif isinstance(v, ctypes.Structure):
                v = tuple(getattr(v, n) for n, _ in v._fields_)
            fields[name] = v
        rec = (self.functions[fn], tuple(grid), tuple(block), fields)
        assert all(g >= 1 for g in grid) and grid[1] <= 65535 and grid[2] <= 65535, rec[:3]
        assert 1 <= block[0] * block[1] * block[2] <= 1024
        (self.capturing if self.capturing is not None else self.launches).append(rec)

    def launch_count(self) -> int:
        return len(self.launches)

    def graph_begin(self, stream=0):
        assert self.capturing is None
        self.capturing = []

    def graph_end(self, stream=0):
        h = self._handle()
        self.graphs[h], self.capturing = self.capturing, None
        return h, len(self.graphs[h])

    def graph_launch(self, graph, stream=0):
        self.launches.extend(self.graphs[graph])

    def graph_destroy(self, graph):
        self.graphs.pop(graph, None)

    def names(self, since: int = 0) -> list:
        return [r[0] for r in self.launches[since:]]




def install(monkeypatch) -> FakeRuntime:
    """Make FakeRuntime the process's runtime for one test (pytest's monkeypatch undoes it)."""
    rt = FakeRuntime()
    monkeypatch.setattr(shim.Runtime, "_instance", rt)
    from xgrid_b200.lang import launch, schedule
    monkeypatch.setattr(schedule, "_PENDING", None)
    monkeypatch.setattr(launch, "STATS", {})
    return rt


class FakeTransport:
    """Records halo exchanges instead of running NCCL (xgrid_b200/dist.py::NcclTransport's interface)."""

    def __init__(self, topo, rt: FakeRuntime) -> None:
        self.topo, self.rt, self.log = topo, rt, []

    def exchange(self, items, stream=0):
        for grid, lv, h in items:
            self.log.append(("exchange", lv.dev, h, len(self.rt.launches)))
            lv.halo_rows = h

    def exchange_async(self, items):
        for grid, lv, h in items:
            self.log.append(("exchange_async", lv.dev, h, len(self.rt.launches)))
            lv.halo_rows = h
            lv.halo_event = self.rt.event_create()

    def reserve(self, nbytes):
        self.reserved = max(getattr(self, "reserved", 0), nbytes)

    def all_agree(self, flag):
        self.log.append(("all_agree", int(bool(flag)), 0, len(self.rt.launches)))
        return bool(flag)

    def fence_compute(self):
        self.log.append(("fence", 0, 0, len(self.rt.launches)))

    def exchange_bytes(self, data_ptr, data_bytes, halo_bytes, stream=0):
        self.log.append(("exchange_bytes", data_ptr, halo_bytes, len(self.rt.launches)))


def install_sharded(monkeypatch, rank: int, world: int, ring: bool = False):
    """Fake runtime + a fixed slab topology + a recording transport (no torch.distributed needed)."""
    from xgrid_b200 import dist
    rt = install(monkeypatch)
    topo = dist.Topology(rank, world, ring)
    transport = FakeTransport(topo, rt)
    monkeypatch.setattr(dist, "_topology", topo)
    monkeypatch.setattr(dist, "_transport", transport)
    return rt, transport
