"""One GPU: does the peer exchange kernel disturb the 3-D sweep it is meant to overlap?  The heat3d slab is stepped on
the compute stream while a ring-of-one exchange (push + pull of two 32 MiB faces through the rank's own mailbox) runs
on a high-priority stream, released by an event recorded right before each step's sweep -- the schedule of a sharded
step without the neighbour.  Prints ms/step without and with the concurrent exchange."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xgrid_b200 as xgrid                          # noqa: E402
from examples import workloads as W                 # noqa: E402
from xgrid_b200.runtime import shim                 # noqa: E402


def main():
    xgrid.init(precision="double")
    rt, lib = shim.Runtime.get(), shim.lib()
    k = W.make_kernels()["heat_3d"]
    shape = (256, 2048, 2048)
    u = xgrid.Grid(shape, float)
    u.boundary[0] = 1
    u.boundary[-1] = 1
    for _ in range(6):
        k(u, 0.1)
    rt.device_sync()
    face = 2048 * 2048 * 8
    handle = C.create_string_buffer(64)
    shim.check(lib.xgb_peer_create(face, C.cast(handle, C.c_void_p)))
    box = C.c_void_p()
    shim.check(lib.xgb_peer_open(C.cast(handle, C.c_void_p), C.byref(box)))
    raw = rt.alloc(6 * face)
    d = shim.HaloDesc()
    d.bytes = face
    d.send_lo, d.recv_lo = raw + face, raw
    d.send_hi, d.recv_hi = raw + 4 * face, raw + 5 * face
    side = rt.stream_create(high_priority=True)
    e0, e1, fork = rt.event_create(), rt.event_create(), rt.event_create()

    def run(with_exchange: bool, steps: int = 20) -> float:
        rt.device_sync()
        rt.event_record_raw(e0, 0)
        for _ in range(steps):
            if with_exchange:
                rt.event_record_raw(fork, 0)
                rt.stream_wait_event(side, fork)
                shim.check(lib.xgb_peer_exchange(C.byref(d), 1, box, box, side))
            k(u, 0.1)
        xgrid.flush() if hasattr(xgrid, "flush") else None
        rt.event_record_raw(e1, 0)
        rt.device_sync()
        return rt.event_elapsed_ms(e0, e1) / steps

    for label, on in (("sweep alone", False), ("sweep + exchange", True), ("sweep alone", False), ("sweep + exchange", True)):
        print(f"{label:>18}: {run(on):.4f} ms/step   (XGB_PEER_CTAS={os.environ.get('XGB_PEER_CTAS', 'default')})", flush=True)
    # launched BEHIND the sweep (all SMs hold two sweep CTAs by then): does the exchange run beside them, or wait for one
    # to retire?  device time from "sweep launched" to "exchange done", sweep length ~2.7 ms, first wave ~1.4 ms
    ea, eb = rt.event_create(), rt.event_create()
    lat = []
    for _ in range(8):
        rt.device_sync()
        rt.event_record_raw(fork, 0)
        k(u, 0.1)
        rt.stream_wait_event(side, fork)
        rt.event_record_raw(ea, side)
        shim.check(lib.xgb_peer_exchange(C.byref(d), 1, box, box, side))
        rt.event_record_raw(eb, side)
        rt.device_sync()
        lat.append(rt.event_elapsed_ms(ea, eb) * 1e3)
    print("exchange launched behind the sweep, us:", " ".join(f"{x:.0f}" for x in lat), flush=True)
    # the exchange alone, for scale
    rt.device_sync()
    rt.event_record_raw(e0, side)
    for _ in range(20):
        shim.check(lib.xgb_peer_exchange(C.byref(d), 1, box, box, side))
    rt.event_record_raw(e1, side)
    rt.device_sync()
    print(f"    exchange alone: {rt.event_elapsed_ms(e0, e1) / 20 * 1e3:.1f} us", flush=True)


if __name__ == "__main__":
    main()
