"""xgrid_b200/hostview.py: arrays handed out by `Grid.now` / `Grid.boundary` announce writes before they happen, and
behave like plain ndarrays otherwise."""
import pickle

import numpy as np

from xgrid_b200 import hostview


class Owner:
    def __init__(self):
        self.touched = 0

    def hook(self, view):
        self.touched += 1


def make(n=12):
    o = Owner()
    base = np.arange(n, dtype=np.int32)
    return o, base, hostview.make(base, o.hook)


def test_every_mutating_entry_point_calls_the_hook_first():
    o, base, v = make()
    v[0] = 5
    assert o.touched == 1 and base[0] == 5
    v[2:4] = [7, 8]
    v += 1                                   # in-place operator keeps the view
    assert o.touched == 3 and isinstance(v, hostview.HostView) and base[0] == 6
    np.add(v, 1, out=v)
    np.copyto(v, np.zeros(12, np.int32))
    v.fill(3)
    v.put([1], [9])
    np.put(v, [2], [4])
    np.multiply.at(v, [0], 2)
    assert o.touched >= 9 and list(base[:3]) == [6, 9, 4]          # (np.put announces twice: function and method)
    n = o.touched
    sub = v[4:8]                             # a view of the view reports to the same owner
    sub[0] = 1
    sub.reshape(2, 2)[1, 1] = 2
    assert o.touched == n + 2 and base[4] == 1 and base[7] == 2


def test_reads_do_not_touch_and_results_are_plain_arrays():
    o, base, v = make()
    assert type(v + 1) is np.ndarray and type(v == 3) is np.ndarray and type(np.sum(v)) is not hostview.HostView
    assert int(v.sum()) == int(base.sum()) and np.array_equal(v, base) and v[3] == 3
    assert np.array_equal(np.concatenate([v, v]), np.concatenate([base, base]))
    w = np.empty(12, np.int32)
    np.copyto(w, v)                          # the view as a SOURCE
    assert o.touched == 0 and np.array_equal(w, base)
    assert np.array_equal(pickle.loads(pickle.dumps(v)), base)
    assert v.ctypes.data == base.ctypes.data and v.dtype == base.dtype and v.flags["C_CONTIGUOUS"]


def test_the_hook_is_held_weakly():
    o, base, v = make()
    del o
    v[0] = 1                                 # owner gone: writes still work
    assert base[0] == 1
