"""Peer-memory halo exchange (csrc/xgb_peer.cu) through the C ABI on ONE GPU: a ring of one rank is its own lo and
hi neighbour, so the very kernel that pushes rows over NVLink between processes pushes them into its own mailbox here
-- the lower ghost rows receive the slab's LAST rows, the upper ghost rows its FIRST rows (a periodic wrap).  Covers the
exchange counter / slot parity over many exchanges, several levels per launch, batches that do not fit one slot,
unaligned byte counts, and replay from a CUDA graph.  The two-process run over real peer memory is tests/test_dist.py."""
import ctypes as C

import numpy as np
import pytest

from xgrid_b200.runtime import shim

pytestmark = pytest.mark.gpu


@pytest.fixture()
def mailbox(tmp_path):
    import xgrid_b200 as xgrid
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"))
    rt = shim.Runtime.get()
    lib = shim.lib()
    shim.check(lib.xgb_peer_destroy())
    handle = C.create_string_buffer(64)
    shim.check(lib.xgb_peer_create(1 << 16, C.cast(handle, C.c_void_p)))
    box = C.c_void_p()
    shim.check(lib.xgb_peer_open(C.cast(handle, C.c_void_p), C.byref(box)))
    yield rt, lib, box
    rt.device_sync()
    shim.check(lib.xgb_peer_close(box))
    shim.check(lib.xgb_peer_destroy())


class Level:
    """[ghost | rows | ghost] bytes on the device with a host mirror."""

    def __init__(self, rt, rows: int, row_bytes: int, h: int, seed: int) -> None:
        self.rt, self.rows, self.row_bytes, self.h = rt, rows, row_bytes, h
        self.total = (rows + 2 * h) * row_bytes
        self.raw = rt.alloc(self.total)
        self.dev = self.raw + h * row_bytes
        self.host = np.zeros(self.total, np.uint8)
        self.fill(seed)

    def fill(self, seed: int) -> None:
        rng = np.random.default_rng(seed)
        self.host[:] = 0
        self.host[self.h * self.row_bytes:(self.h + self.rows) * self.row_bytes] = rng.integers(
            0, 256, self.rows * self.row_bytes, dtype=np.uint8)
        self.rt.h2d(self.raw, self.host.ctypes.data, self.total)
        self.rt.sync()

    def desc(self, d) -> None:
        hb = self.h * self.row_bytes
        d.bytes = hb
        d.lo_rank = d.hi_rank = 0
        d.send_lo, d.recv_lo = self.dev, self.dev - hb
        d.send_hi, d.recv_hi = self.dev + self.rows * self.row_bytes - hb, self.dev + self.rows * self.row_bytes

    def expected(self) -> np.ndarray:
        hb, body = self.h * self.row_bytes, self.rows * self.row_bytes
        want = self.host.copy()
        want[:hb] = self.host[hb + body - hb:hb + body]              # lower ghost <- my last rows (I am my lo neighbour)
        want[hb + body:] = self.host[hb:2 * hb]                      # upper ghost <- my first rows
        return want

    def read(self) -> np.ndarray:
        out = np.empty(self.total, np.uint8)
        self.rt.d2h(out.ctypes.data, self.raw, self.total)
        self.rt.sync()
        return out


def exchange(lib, box, levels, stream=0) -> None:
    descs = (shim.HaloDesc * len(levels))()
    for d, lv in zip(descs, levels):
        lv.desc(d)
    shim.check(lib.xgb_peer_exchange(descs, len(levels), box, box, stream))


def test_ring_of_one_wraps_first_and_last_rows_over_many_exchanges(mailbox):
    rt, lib, box = mailbox
    a = Level(rt, rows=40, row_bytes=4096, h=2, seed=1)
    for n in range(7):                       # exchange numbers 1..7: both slot parities, credits from n - 2
        a.fill(seed=10 + n)
        exchange(lib, box, [a])
        assert np.array_equal(a.read(), a.expected()), f"exchange {n + 1}"


def test_several_levels_per_launch_unaligned_sizes_and_batches_larger_than_a_slot(mailbox):
    rt, lib, box = mailbox
    slot = C.c_uint64()
    shim.check(lib.xgb_peer_slot_bytes(C.byref(slot)))
    assert slot.value == 1 << 16
    # 16-, 8-, 4- and 1-byte copy paths; together 3 * 24 KiB + ... > one 64 KiB slot, and more than 8 levels
    shapes = [(9, 24 * 1024, 1), (9, 24 * 1024, 1), (9, 24 * 1024, 1), (5, 1000, 3), (7, 1004, 1), (6, 333, 2),
              (4, 16, 1), (4, 8, 1), (4, 4, 1), (4, 1, 1), (12, 2048, 2)]
    levels = [Level(rt, r, b, h, seed=100 + n) for n, (r, b, h) in enumerate(shapes)]
    before = shim.Runtime.get().launch_count()
    exchange(lib, box, levels)
    launches = shim.Runtime.get().launch_count() - before
    assert launches >= 2                                              # split: slot capacity and 8 levels per launch
    for n, lv in enumerate(levels):
        assert np.array_equal(lv.read(), lv.expected()), f"level {n} {shapes[n]}"
    # a single level larger than the slot is refused with a message, not truncated
    big = Level(rt, rows=4, row_bytes=(1 << 16) + 16, h=1, seed=5)
    with pytest.raises(Exception, match="larger than the mailbox slot"):
        exchange(lib, box, [big])


def test_exchange_replays_from_a_graph_with_fresh_data_each_time(mailbox):
    rt, lib, box = mailbox
    a = Level(rt, rows=64, row_bytes=8192, h=1, seed=3)
    b = Level(rt, rows=16, row_bytes=512, h=2, seed=4)
    side = rt.stream_create(high_priority=True)
    exchange(lib, box, [a, b], stream=side)                            # exchange 1, directly, on another stream
    rt.sync(side)
    rt.graph_begin(0)
    exchange(lib, box, [a, b])
    graph, nodes = rt.graph_end(0)
    assert nodes == 1
    for n in range(5):                       # the counter lives on the device: every replay is the next exchange
        a.fill(seed=50 + n)
        b.fill(seed=60 + n)
        rt.graph_launch(graph)
        assert np.array_equal(a.read(), a.expected()) and np.array_equal(b.read(), b.expected()), f"replay {n}"
    rt.graph_destroy(graph)
