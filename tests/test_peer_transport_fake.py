"""Host side of the peer-memory halo transport (xgrid_b200/dist.py::PeerTransport) without a GPU: a recording stand-in
for the C ABI, torch.distributed's two collectives patched to a single process, REAL abstract Unix sockets for the
descriptor hand-over.  Checks the stream plumbing (every exchange on the one communication stream, forked from and
joined back into the requesting stream), the ticket / descriptor protocol, collective growth of the mailbox and the
counter reset when the neighbour relation changes."""
import ctypes
import os
import struct

import pytest

from xgrid_b200 import dist as xdist
from xgrid_b200.runtime import shim


class FakeLib:
    def __init__(self):
        self.calls, self.slot, self.pipes, self.opened = [], 0, [], []

    def xgb_peer_create(self, slot_bytes, buf):
        self.slot = (int(slot_bytes) + 255) // 256 * 256
        r, w = os.pipe()
        self.pipes.append((r, w))
        ticket = struct.pack("<IiQqi", 0x58474250, r, 256 + 4 * self.slot, os.getpid(), 0).ljust(64, b"\0")
        ctypes.memmove(buf, ticket, 64)
        self.calls.append(("create", self.slot))
        return 0

    def xgb_peer_slot_bytes(self, out):
        out._obj.value = self.slot
        return 0

    def xgb_peer_open(self, raw, out):
        ticket = ctypes.string_at(raw, 64)
        fd = struct.unpack_from("<i", ticket, 4)[0]
        os.fstat(fd)                                    # must be a live descriptor of THIS process at the time of the call
        self.opened.append(fd)
        out._obj.value = 0x1000 * len(self.opened)
        self.calls.append(("open", fd))
        return 0

    def xgb_peer_close(self, box):
        self.calls.append(("close", box.value))
        return 0

    def xgb_peer_destroy(self):
        self.calls.append(("destroy",))
        return 0

    def xgb_peer_reset(self):
        self.calls.append(("reset",))
        return 0

    def xgb_peer_exchange(self, descs, n, lo, hi, stream):
        self.calls.append(("exchange", n, lo.value, hi.value, stream))
        return 0


class FakeRt:
    device = 0

    def __init__(self, log):
        self.log, self.events, self.streams = log, 0, 0

    def stream_create(self, high_priority=False):
        self.streams += 1
        return 100 + self.streams

    def event_create(self):
        self.events += 1
        return 1000 + self.events

    def event_record_raw(self, ev, stream=0):
        self.log.append(("record", ev, stream))

    def stream_wait_event(self, stream, ev):
        self.log.append(("wait", stream, ev))

    def device_sync(self):
        self.log.append(("device_sync",))


@pytest.fixture()
def peer(monkeypatch):
    import torch.distributed as tdist
    lib = FakeLib()
    rt = FakeRt(lib.calls)
    monkeypatch.setattr(shim, "lib", lambda: lib)
    monkeypatch.setattr(shim, "check", lambda rc: None)
    monkeypatch.setattr(shim.Runtime, "get", classmethod(lambda cls: rt))
    monkeypatch.setenv("MASTER_PORT", str(40000 + os.getpid() % 20000))
    world = 3

    def gather(out, obj):                               # every "rank" publishes this process's ticket and address
        for i in range(len(out)):
            out[i] = obj
        lib.calls.append(("all_gather", len(out)))

    monkeypatch.setattr(tdist, "all_gather_object", gather)
    monkeypatch.setattr(tdist, "barrier", lambda: lib.calls.append(("barrier",)))
    import atexit
    monkeypatch.setattr(atexit, "register", lambda f: f)
    tr = xdist.PeerTransport(xdist.Topology(1, world))
    yield tr, lib, rt
    tr._stop_server()
    for r, w in lib.pipes:
        os.close(r)
        os.close(w)


def test_descriptors_travel_over_the_socket_and_every_exchange_runs_on_the_comm_stream(peer):
    tr, lib, rt = peer
    assert lib.calls[:2] == [("create", 8 << 20), ("all_gather", 3)]
    own_fd = lib.pipes[0][0]
    lo, hi = tr._box(0), tr._box(2)
    assert lo and hi and lo != hi and tr._box(0) == lo                     # mapped once per neighbour
    assert len(lib.opened) == 2 and own_fd not in lib.opened               # duplicates received through SCM_RIGHTS ...
    for fd in lib.opened:
        with pytest.raises(OSError):
            os.fstat(fd)                                                   # ... and closed again after the import
    assert tr._box(1) and lib.opened[-1] == own_fd                         # the rank's own ticket goes in unchanged
    assert tr._box(-1) is None
    # a request from the compute stream: fork into the communication stream, exchange there, join back
    del lib.calls[:]
    descs = (shim.HaloDesc * 2)()
    tr._issue(descs, 2, 0)
    comm = tr._comm_stream
    kinds = [c[0] for c in lib.calls]
    assert kinds == ["record", "wait", "exchange", "record", "wait"], lib.calls
    assert lib.calls[0][2] == 0 and lib.calls[1][1] == comm and lib.calls[2] == ("exchange", 2, lo, hi, comm)
    assert lib.calls[3][2] == comm and lib.calls[4][1] == 0
    # a request that already sits on the communication stream (the asynchronous exchange after the edge bands)
    del lib.calls[:]
    tr._issue(descs, 1, comm)
    assert lib.calls == [("exchange", 1, lo, hi, comm)]
    # a request from a side stream joins back into THAT stream
    del lib.calls[:]
    tr._issue(descs, 1, 77)
    assert lib.calls[0][2] == 77 and lib.calls[-1][1] == 77


def test_growing_the_mailbox_is_a_collective_and_replaces_every_mapping(peer):
    tr, lib, rt = peer
    tr._box(0)
    tr.reserve(1 << 20)                                                    # fits the 8 MiB slot: nothing happens
    assert all(c[0] != "destroy" for c in lib.calls)
    del lib.calls[:]
    tr.reserve(33 << 20)
    kinds = [c[0] for c in lib.calls]
    assert kinds == ["device_sync", "barrier", "close", "barrier", "destroy", "create", "all_gather"], lib.calls
    assert lib.slot >= 33 << 20 and tr._slot == lib.slot and tr._boxes == {}
    assert tr._box(0)                                                      # re-fetched from the NEW server address


def test_a_new_neighbour_relation_zeroes_the_counters_between_two_barriers(peer):
    tr, lib, rt = peer
    del lib.calls[:]
    tr.topo = xdist.Topology(1, 3)                                         # same relation: nothing to do
    assert lib.calls == []
    ring = xdist.Topology(0, 3, ring=True)
    tr.topo = ring
    assert [c[0] for c in lib.calls] == ["device_sync", "barrier", "reset", "barrier"] and tr.topo is ring
