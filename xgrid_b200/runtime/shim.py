"""ctypes binding of ``libxgrid_b200.so`` (include/xgrid_b200.h).

The counterpart of the reference's ``Library`` wrapper
(xgrid/util/ffi.py:13-35): every symbol gets explicit ``argtypes`` and a
non-zero status becomes a Python ``Exception`` carrying ``xgb_last_error()``
(reference behaviour: ``Logger.dead`` -> ``Exception``, ffi.py:86-89).  There
is no fallback: if the library is missing the import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from ctypes import POINTER, c_char_p, c_int, c_size_t, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libxgrid_b200.so")

Handle = c_uint64
U3 = c_uint32 * 3


class DeviceInfo(C.Structure):
    _fields_ = [("ordinal", C.c_int32), ("sm_count", C.c_int32), ("cc_major", C.c_int32),
                ("cc_minor", C.c_int32), ("max_smem_per_block_optin", C.c_int32),
                ("l2_bytes", C.c_int32), ("clock_khz", C.c_int32), ("reserved", C.c_int32),
                ("total_mem", c_uint64), ("name", C.c_char * 64)]


class HaloDesc(C.Structure):
    _fields_ = [("send_lo", c_void_p), ("recv_lo", c_void_p), ("send_hi", c_void_p),
                ("recv_hi", c_void_p), ("bytes", c_uint64), ("lo_rank", C.c_int32),
                ("hi_rank", C.c_int32)]


# name -> argtypes; every function returns int status unless listed in _RESTYPE
SIGNATURES = {
    "xgb_abi_version": [],
    "xgb_last_error": [],
    "xgb_init": [c_int],
    "xgb_shutdown": [],
    "xgb_device_count": [POINTER(c_int)],
    "xgb_get_device_info": [POINTER(DeviceInfo)],
    "xgb_device_sync": [],
    "xgb_alloc": [c_size_t, POINTER(c_void_p)],
    "xgb_free": [c_void_p],
    "xgb_memset": [c_void_p, c_int, c_size_t, Handle],
    "xgb_h2d": [c_void_p, c_void_p, c_size_t, Handle],
    "xgb_d2h": [c_void_p, c_void_p, c_size_t, Handle],
    "xgb_d2d": [c_void_p, c_void_p, c_size_t, Handle],
    "xgb_h2d_staged": [c_void_p, c_void_p, c_size_t, Handle],
    "xgb_d2h_staged": [c_void_p, c_void_p, c_size_t, Handle],
    "xgb_mask_pack": [POINTER(C.c_int32), c_size_t, c_size_t, POINTER(C.c_uint8), POINTER(C.c_uint8),
                      POINTER(c_uint64), POINTER(c_int), c_int],
    "xgb_host_alloc": [c_size_t, POINTER(c_void_p)],
    "xgb_host_free": [c_void_p],
    "xgb_host_register": [c_void_p, c_size_t],
    "xgb_host_unregister": [c_void_p],
    "xgb_mem_info": [POINTER(c_uint64), POINTER(c_uint64)],
    "xgb_stream_create": [POINTER(Handle)],
    "xgb_stream_create_ex": [POINTER(Handle), c_int],
    "xgb_stream_destroy": [Handle],
    "xgb_stream_sync": [Handle],
    "xgb_stream_raw": [Handle, POINTER(c_void_p)],
    "xgb_event_create": [POINTER(Handle)],
    "xgb_event_destroy": [Handle],
    "xgb_event_record": [Handle, Handle],
    "xgb_event_sync": [Handle],
    "xgb_event_elapsed_ms": [Handle, Handle, POINTER(C.c_float)],
    "xgb_stream_wait_event": [Handle, Handle],
    "xgb_compile": [c_char_p, c_char_p, POINTER(c_char_p), c_int, POINTER(c_char_p),
                    POINTER(c_char_p), c_int, POINTER(c_void_p), POINTER(c_size_t), POINTER(c_void_p)],
    "xgb_release": [c_void_p],
    "xgb_module_load": [c_void_p, c_size_t, POINTER(Handle)],
    "xgb_module_unload": [Handle],
    "xgb_get_function": [Handle, c_char_p, POINTER(Handle)],
    "xgb_function_info": [Handle, POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int)],
    "xgb_function_set_dynamic_smem": [Handle, c_int],
    "xgb_occupancy": [Handle, c_int, c_int, POINTER(c_int)],
    "xgb_launch": [Handle, POINTER(c_uint32), POINTER(c_uint32), c_uint32, Handle, c_void_p, c_size_t],
    "xgb_launch_count": [POINTER(c_uint64)],
    "xgb_graph_begin": [Handle],
    "xgb_graph_end": [Handle, POINTER(Handle), POINTER(c_int)],
    "xgb_graph_launch": [Handle, Handle],
    "xgb_graph_destroy": [Handle],
    "xgb_nccl_load": [c_char_p],
    "xgb_nccl_unique_id": [c_void_p],
    "xgb_nccl_init": [c_void_p, c_int, c_int],
    "xgb_nccl_shutdown": [],
    "xgb_halo_exchange": [POINTER(HaloDesc), c_int, Handle],
    "xgb_peer_create": [c_uint64, c_void_p],
    "xgb_peer_open": [c_void_p, POINTER(c_void_p)],
    "xgb_peer_close": [c_void_p],
    "xgb_peer_destroy": [],
    "xgb_peer_reset": [],
    "xgb_peer_slot_bytes": [POINTER(c_uint64)],
    "xgb_peer_exchange": [POINTER(HaloDesc), c_int, c_void_p, c_void_p, Handle],
}
_RESTYPE = {"xgb_last_error": c_char_p}

_lib = None


def host_threads(cap: int) -> int:
    """Host threads this process should use for copies / mask packing: the cores it may run on, shared with the
    other ranks of the box (torchrun's LOCAL_WORLD_SIZE), at most `cap`.  Eight ranks that each start eight copy
    threads and sixteen packing threads on a 32-core host only fight each other (measured: the 8-GPU end-to-end
    job moved 138 GB at 46 GB/s aggregate)."""
    try:
        cores = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        cores = os.cpu_count() or 1
    ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
    return max(2, min(cap, cores // ranks))


def lib():
    """The loaded shim; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Exception(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(the B200 backend has no CPU fallback)")
        os.environ.setdefault("XGB_STAGE_LANES", str(host_threads(8)))      # read by the runtime on its first staged copy
        handle = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPE.get(name, c_int)
        _lib = handle
    return _lib


def last_error() -> str:
    return (lib().xgb_last_error() or b"").decode(errors="replace")


def check(status: int) -> None:
    if status != 0:
        raise Exception(f"xgrid_b200 runtime error: {last_error()}")


class Runtime:
    """Process-wide device runtime (one GPU per process)."""

    _instance = None

    def __init__(self, device: int) -> None:
        self.l = lib()
        check(self.l.xgb_init(device))
        self.device = device
        info = DeviceInfo()
        check(self.l.xgb_get_device_info(C.byref(info)))
        self.info = info
        self.sm_count = int(info.sm_count)
        self._pool: dict = {}            # bytes -> [device pointers] (freed, reusable)
        self._pool_bytes = 0
        self._sizes: dict = {}           # live device pointer -> bytes
        if info.cc_major != 10:
            raise Exception(f"xgrid_b200 targets sm_100a (B200); device '{info.name.decode()}' is "
                            f"sm_{info.cc_major}{info.cc_minor}")

    @classmethod
    def get(cls) -> "Runtime":
        if cls._instance is None:
            from ..config import get_config
            cls._instance = Runtime(get_config().device)
        return cls._instance

    # ---- memory
    # Caching pool for large device buffers (time levels, masks).  cudaFree synchronises the device and
    # returns the pages to the driver (measured on B200: ~100 ms per freed 128 MiB level, and the next
    # cudaMalloc of that size pays again), so freed buffers of >= POOL_MIN bytes are kept by exact size and
    # handed out again zero-filled, which is what xgb_alloc guarantees.  Re-use is ordered on the compute
    # stream (the memset is enqueued there); in processes that also run a communication stream (sharded
    # grids) a buffer enters the pool only after the compute stream has been ordered behind the work already
    # enqueued on the communication stream (`fence_compute`: one event record + one stream wait, no host
    # synchronisation), so a halo exchange still reading the buffer finishes before its next owner's memset.
    # The pool is dropped when an allocation fails and at most POOL_CAP bytes are kept.
    POOL_MIN = 1 << 20
    POOL_CAP = int(float(os.environ.get("XGB_POOL_GB", "24")) * (1 << 30))

    def alloc(self, nbytes: int) -> int:
        nbytes = int(nbytes)
        cached = self._pool.get(nbytes)
        if cached:
            ptr = cached.pop()
            self._pool_bytes -= nbytes
            check(self.l.xgb_memset(c_void_p(ptr), 0, nbytes, 0))
        else:
            p = c_void_p()
            if self.l.xgb_alloc(nbytes, C.byref(p)) != 0:
                if not self._pool_bytes:
                    check(1)
                self.trim_pool()                 # out of memory with buffers cached: release them, retry once
                check(self.l.xgb_alloc(nbytes, C.byref(p)))
            ptr = p.value
        self._sizes[ptr] = nbytes
        return ptr

    def free(self, ptr: int) -> None:
        nbytes = self._sizes.pop(ptr, 0)
        if nbytes >= self.POOL_MIN and self._pool_bytes + nbytes <= self.POOL_CAP:
            self._fence_comm()
            self._pool.setdefault(nbytes, []).append(ptr)
            self._pool_bytes += nbytes
            return
        self.l.xgb_free(c_void_p(ptr))

    @staticmethod
    def _fence_comm() -> None:
        from .. import dist
        if dist._transport is not None:
            dist._transport.fence_compute()

    def trim_pool(self) -> None:
        """Return every cached buffer to the driver."""
        for cached in self._pool.values():
            for ptr in cached:
                self.l.xgb_free(c_void_p(ptr))
        self._pool.clear()
        self._pool_bytes = 0

    def memset(self, ptr: int, byte: int, nbytes: int, stream: int = 0) -> None:
        check(self.l.xgb_memset(c_void_p(ptr), byte, nbytes, stream))

    def h2d(self, dst: int, src: int, nbytes: int, stream: int = 0) -> None:
        check(self.l.xgb_h2d(c_void_p(dst), c_void_p(src), nbytes, stream))

    def d2h(self, dst: int, src: int, nbytes: int, stream: int = 0) -> None:
        check(self.l.xgb_d2h(c_void_p(dst), c_void_p(src), nbytes, stream))

    # pageable memory through the runtime's multi-threaded staging path (XGB_STAGED_COPY=0: the driver's
    # single-threaded pageable cudaMemcpy instead)
    STAGED = os.environ.get("XGB_STAGED_COPY", "1") == "1"
    STAGED_MIN = 8 << 20

    def h2d_staged(self, dst: int, src: int, nbytes: int, stream: int = 0) -> None:
        check(self.l.xgb_h2d_staged(c_void_p(dst), c_void_p(src), nbytes, stream))

    def d2h_staged(self, dst: int, src: int, nbytes: int, stream: int = 0) -> None:
        check(self.l.xgb_d2h_staged(c_void_p(dst), c_void_p(src), nbytes, stream))

    def d2d(self, dst: int, src: int, nbytes: int, stream: int = 0) -> None:
        check(self.l.xgb_d2d(c_void_p(dst), c_void_p(src), nbytes, stream))

    def host_alloc(self, nbytes: int) -> int:
        p = c_void_p()
        check(self.l.xgb_host_alloc(nbytes, C.byref(p)))
        return p.value

    def host_free(self, ptr: int) -> None:
        self.l.xgb_host_free(c_void_p(ptr))

    def mem_info(self):
        f, t = c_uint64(), c_uint64()
        check(self.l.xgb_mem_info(C.byref(f), C.byref(t)))
        return f.value, t.value

    # ---- streams / events
    @staticmethod
    def _flush_deferred() -> None:
        from ..lang import schedule
        if schedule._PENDING is not None:
            schedule.flush_pending()

    def sync(self, stream: int = 0) -> None:
        self._flush_deferred()
        check(self.l.xgb_stream_sync(stream))

    def device_sync(self) -> None:
        self._flush_deferred()
        check(self.l.xgb_device_sync())

    def stream_create(self, high_priority: bool = False) -> int:
        h = Handle()
        check(self.l.xgb_stream_create_ex(C.byref(h), 1 if high_priority else 0))
        return h.value

    def side_stream(self):
        """(high-priority side stream, fork event, join event): work that may run beside the compute stream."""
        side = getattr(self, "_side", None)
        if side is None:
            side = self._side = (self.stream_create(high_priority=True), self.event_create(), self.event_create())
        return side

    def stream_raw(self, stream: int = 0) -> int:
        p = c_void_p()
        check(self.l.xgb_stream_raw(stream, C.byref(p)))
        return p.value or 0

    def event_create(self) -> int:
        h = Handle()
        check(self.l.xgb_event_create(C.byref(h)))
        return h.value

    def event_record(self, ev: int, stream: int = 0) -> None:
        self._flush_deferred()      # deferred kernel calls belong before the event
        check(self.l.xgb_event_record(ev, stream))

    def event_record_raw(self, ev: int, stream: int = 0) -> None:
        """Record without flushing the deferred-call queue (internal stream plumbing)."""
        check(self.l.xgb_event_record(ev, stream))

    def event_sync(self, ev: int) -> None:
        check(self.l.xgb_event_sync(ev))

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float()
        check(self.l.xgb_event_elapsed_ms(a, b, C.byref(ms)))
        return float(ms.value)

    def stream_wait_event(self, stream: int, ev: int) -> None:
        check(self.l.xgb_stream_wait_event(stream, ev))

    # ---- modules
    def module_load(self, image: bytes) -> int:
        h = Handle()
        buf = C.create_string_buffer(image, len(image))
        check(self.l.xgb_module_load(C.cast(buf, c_void_p), len(image), C.byref(h)))
        return h.value

    def get_function(self, module: int, name: str) -> int:
        h = Handle()
        check(self.l.xgb_get_function(module, name.encode(), C.byref(h)))
        return h.value

    def function_info(self, fn: int) -> dict:
        r, s, l, m = c_int(), c_int(), c_int(), c_int()
        check(self.l.xgb_function_info(fn, C.byref(r), C.byref(s), C.byref(l), C.byref(m)))
        return {"regs": r.value, "static_smem": s.value, "local_bytes": l.value, "max_threads": m.value}

    def set_dynamic_smem(self, fn: int, nbytes: int) -> None:
        check(self.l.xgb_function_set_dynamic_smem(fn, nbytes))

    def occupancy(self, fn: int, threads: int, smem: int = 0) -> int:
        n = c_int()
        check(self.l.xgb_occupancy(fn, threads, smem, C.byref(n)))
        return n.value

    # ---- launch / graphs
    def launch(self, fn: int, grid, block, params, smem: int = 0, stream: int = 0) -> None:
        check(self.l.xgb_launch(fn, U3(*grid), U3(*block), smem, stream,
                                C.cast(C.byref(params), c_void_p), C.sizeof(params)))

    def launch_count(self) -> int:
        n = c_uint64()
        check(self.l.xgb_launch_count(C.byref(n)))
        return n.value

    def graph_begin(self, stream: int = 0) -> None:
        check(self.l.xgb_graph_begin(stream))

    def graph_end(self, stream: int = 0):
        h, n = Handle(), c_int()
        check(self.l.xgb_graph_end(stream, C.byref(h), C.byref(n)))
        return h.value, n.value

    def graph_launch(self, graph: int, stream: int = 0) -> None:
        check(self.l.xgb_graph_launch(graph, stream))

    def graph_destroy(self, graph: int) -> None:
        self.l.xgb_graph_destroy(graph)


def compile_cuda(source: str, name: str, options: list[str], headers: dict[str, str]) -> tuple[bytes, str]:
    """NVRTC: CUDA C -> cubin bytes.  Works without a GPU (cross-compile)."""
    l = lib()
    opts = (c_char_p * len(options))(*[o.encode() for o in options])
    hn = (c_char_p * len(headers))(*[k.encode() for k in headers])
    hs = (c_char_p * len(headers))(*[v.encode() for v in headers.values()])
    image, size, log = c_void_p(), c_size_t(), c_void_p()
    status = l.xgb_compile(source.encode(), name.encode(), opts, len(options), hn, hs, len(headers),
                           C.byref(image), C.byref(size), C.byref(log))
    log_text = ""
    if log.value:
        log_text = C.string_at(log.value).decode(errors="replace")
        l.xgb_release(log)
    if status != 0:
        raise Exception(f"failed to compile '{name}' due to:", last_error())
    data = C.string_at(image.value, size.value)
    l.xgb_release(image)
    return data, log_text
