"""Host logic of the device-memory pool (xgrid_b200/runtime/shim.py::Runtime.alloc/free) against a fake
C ABI: exact-size reuse, zero-fill on reuse, small buffers bypass, cap, release + retry on allocation failure."""
import ctypes

from xgrid_b200.runtime import shim


class FakeLib:
    def __init__(self, capacity):
        self.capacity, self.used, self.next, self.live = capacity, 0, 0x1000, {}
        self.mallocs = self.frees = self.memsets = 0

    def xgb_alloc(self, nbytes, out):
        if self.used + nbytes > self.capacity:
            return 1
        self.next += 1 << 32
        self.live[self.next] = nbytes
        self.used += nbytes
        self.mallocs += 1
        ctypes.cast(out, ctypes.POINTER(ctypes.c_void_p))[0] = self.next
        return 0

    def xgb_free(self, p):
        self.used -= self.live.pop(p.value)
        self.frees += 1
        return 0

    def xgb_memset(self, p, byte, nbytes, stream):
        assert byte == 0 and stream == 0
        self.memsets += 1
        return 0

    def xgb_last_error(self):
        return b"out of memory (fake)"


def make(capacity=1 << 40, cap=None, monkeypatch=None):
    rt = object.__new__(shim.Runtime)
    rt.l = FakeLib(capacity)
    rt._pool, rt._pool_bytes, rt._sizes = {}, 0, {}
    if cap is not None:
        rt.POOL_CAP = cap
    return rt


def test_exact_size_reuse_is_zero_filled(monkeypatch):
    monkeypatch.setattr(shim, "last_error", lambda: "fake")
    rt = make()
    big = 8 << 20
    a = rt.alloc(big)
    rt.free(a)
    assert rt.l.frees == 0 and rt._pool_bytes == big          # cached, not returned to the driver
    b = rt.alloc(big)
    assert b == a and rt.l.mallocs == 1 and rt.l.memsets == 1 and rt._pool_bytes == 0
    c = rt.alloc(big + 256)                                     # another size: a real allocation
    assert c != a and rt.l.mallocs == 2


def test_small_buffers_and_cap_bypass_the_pool(monkeypatch):
    monkeypatch.setattr(shim, "last_error", lambda: "fake")
    rt = make(cap=16 << 20)
    s = rt.alloc(4096)
    rt.free(s)
    assert rt.l.frees == 1 and rt._pool_bytes == 0
    x, y = rt.alloc(12 << 20), rt.alloc(12 << 20)
    rt.free(x)
    rt.free(y)                                                  # would exceed the cap: really freed
    assert rt._pool_bytes == 12 << 20 and rt.l.frees == 2


def test_pool_is_released_when_an_allocation_fails(monkeypatch):
    monkeypatch.setattr(shim, "last_error", lambda: "fake")
    rt = make(capacity=100 << 20)
    a = rt.alloc(60 << 20)
    rt.free(a)                                                  # cached: the fake device is still 60 % full
    b = rt.alloc(70 << 20)                                      # does not fit until the cache is dropped
    assert b and rt._pool_bytes == 0 and rt.l.used == 70 << 20
    try:
        rt.alloc(90 << 20)
    except Exception as e:
        assert "runtime error" in str(e)
    else:
        raise AssertionError("allocation beyond the device capacity must fail loudly")
    rt.trim_pool()
    assert rt.l.used == 70 << 20


def test_empty_cache_is_public(monkeypatch):
    import xgrid_b200 as xgrid
    monkeypatch.setattr(shim, "last_error", lambda: "fake")
    rt = make()
    monkeypatch.setattr(shim.Runtime, "_instance", rt)
    a = rt.alloc(8 << 20)
    rt.free(a)
    assert rt._pool_bytes
    xgrid.empty_cache()
    assert rt._pool_bytes == 0 and rt.l.frees == 1
