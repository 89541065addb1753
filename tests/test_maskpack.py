"""Host half of the mask upload (xgb_mask_pack through the C ABI): packed bytes, 128-point chunk flags and the
per-value histogram against NumPy, single- and multi-threaded, ragged tails, out-of-range values."""
import ctypes as C

import numpy as np
import pytest

from xgrid_b200.runtime import maskpack, shim


def _reference(b, n_padded):
    flat = b.reshape(-1)
    ok = (flat >= 0) & (flat <= 254)
    packed = np.zeros(n_padded, np.uint8)
    packed[:flat.size] = np.where(ok, flat, 255).astype(np.uint8)
    flags = (packed.reshape(-1, 128) != 0).any(axis=1).astype(np.uint8)
    return packed, flags, np.bincount(packed[:flat.size], minlength=256).astype(np.int64), bool((~ok).any())


@pytest.mark.parametrize("shape", [(1,), (127,), (128,), (1000,), (3, 5, 77), (1 << 21,), (300, 4099)])
def test_pack_matches_numpy(shape):
    rng = np.random.default_rng(len(shape) + shape[0])
    b = np.zeros(shape, np.int32)
    hits = rng.integers(0, b.size, max(1, b.size // 500))
    b.reshape(-1)[hits] = rng.integers(1, 255, hits.size)
    b.reshape(-1)[0] = 254
    n_padded = (b.size + 127) // 128 * 128
    packed, flags, hist, bad = maskpack.pack(b, n_padded)
    want = _reference(b, n_padded)
    assert np.array_equal(packed, want[0]) and np.array_equal(flags, want[1])
    assert np.array_equal(hist, want[2]) and bad is False and int(hist.sum()) == b.size


def test_out_of_range_values_are_flagged_and_stored_as_outside():
    b = np.array([0, 5, -1, 300, 255, 254] * 50, np.int32)
    packed, flags, hist, bad = maskpack.pack(b, 384)
    want = _reference(b, 384)
    assert bad is True and np.array_equal(packed, want[0]) and np.array_equal(hist, want[2])
    assert list(packed[:6]) == [0, 5, 255, 255, 255, 254]


def test_thread_count_does_not_change_the_result():
    rng = np.random.default_rng(5)
    b = (rng.random(1 << 22) < 0.001).astype(np.int32) * 3
    n = b.size
    outs = []
    for threads in (1, 3, 8):
        packed, flags, hist = np.empty(n, np.uint8), np.empty(n // 128, np.uint8), np.zeros(256, np.uint64)
        bad = C.c_int(0)
        shim.check(shim.lib().xgb_mask_pack(b.ctypes.data_as(C.POINTER(C.c_int32)), n, n,
                                            packed.ctypes.data_as(C.POINTER(C.c_uint8)),
                                            flags.ctypes.data_as(C.POINTER(C.c_uint8)),
                                            hist.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(bad), threads))
        outs.append((packed, flags, hist))
    for o in outs[1:]:
        assert all(np.array_equal(x, y) for x, y in zip(o, outs[0]))
    assert shim.lib().xgb_mask_pack(b.ctypes.data_as(C.POINTER(C.c_int32)), n, n + 1,
                                    None, None, None, C.byref(bad), 1) != 0          # n_padded % 128
