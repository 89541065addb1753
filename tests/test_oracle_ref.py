"""The C restatement (oracle/xgrid_oracle.c) against the reference's OWN compiled kernels
(oracle/_ref, produced by oracle/make_ref.py from the unmodified reference) on seeded inputs at
sizes well beyond the golden vectors.  Bit-exact on every ring level.  Skipped when oracle/_ref
has not been built (it needs /root/reference, which only the build container has)."""
import numpy as np
import pytest

import oracle
from oracle import HostGrid, ref
from examples import workloads as W

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (python oracle/make_ref.py)")


def pair(shape, ic, mask):
    out = []
    for _ in range(2):
        g = HostGrid(shape)
        g.now[...] = ic
        g.boundary[...] = mask
        out.append(g)
    return out


def same(a: HostGrid, b: HostGrid):
    assert len(a._data) == len(b._data)
    for la, lb in zip(a._data, b._data):
        assert np.array_equal(la, lb, equal_nan=True)


@pytest.mark.parametrize("name,step,scalars", [
    ("convection_1d", oracle.step_conv1d, (1.0, 0.5e-5, 1e-5)),
    ("convection_1d_nonlinear", oracle.step_conv1d_nonlinear, (0.25e-5, 1e-5)),
    ("diffusion_1d", oracle.step_diff1d, (0.01, 2e-9, 1e-5)),
])
def test_1d_kernels(name, step, scalars):
    n = 200_003
    rng = np.random.default_rng(11)
    ic = 1.0 + rng.random(n)
    mask = np.zeros(n, np.int32)
    mask[0] = mask[-1] = 1
    mask[rng.integers(1, n - 1, 50)] = 1
    mask[rng.integers(1, n - 1, 50)] = 7          # no statement for 7: never written (F5)
    a, b = pair((n,), ic, mask)
    for _ in range(9):
        step(a, *scalars)
        ref.call(name, b, *scalars)
    same(a, b)


@pytest.mark.parametrize("name,step,scalars", [
    ("convection_2d", oracle.step_conv2d, (1.0, 0.001, 0.004, 0.004)),
    ("diffusion_2d", oracle.step_diff2d, (0.2,)),
])
def test_2d_kernels_square(name, step, scalars):
    n = 513
    rng = np.random.default_rng(12)
    ic = rng.random((n, n))
    mask = W.shell_mask((n, n))
    mask[rng.integers(1, n - 1, 40), rng.integers(1, n - 1, 40)] = 1
    a, b = pair((n, n), ic, mask)
    for _ in range(7):
        step(a, *scalars)
        ref.call(name, b, *scalars)
    same(a, b)


def test_ewmul():
    n = 10_000
    rng = np.random.default_rng(13)
    grids_a, grids_b = [], []
    for k in range(3):
        x, y = pair((n,), rng.random(n), np.zeros(n, np.int32))
        grids_a.append(x)
        grids_b.append(y)
    for _ in range(2):
        oracle.step_ewmul(*grids_a)
        ref.call("elementwise_mul", *grids_b)
    for x, y in zip(grids_a, grids_b):
        same(x, y)


def test_cavity_257():
    n = 257
    masks = W.cavity_masks(n, n)
    cfg = oracle.Config(1.0, 0.1, 1e-4 * (100.0 / (n - 1)) ** 2, 2.0 / (n - 1), 2.0 / (n - 1))
    rng = np.random.default_rng(14)
    ics = [np.zeros((n, n)), np.zeros((n, n)), 0.01 * rng.random((n, n)), 0.01 * rng.random((n, n))]
    A, B = zip(*[pair((n, n), ic, m) for ic, m in zip(ics, masks)])
    for _ in range(3):
        oracle.step_cavity(*A, cfg)
        ref.call("cavity_kernel", *B, cfg)
    for x, y in zip(A, B):
        same(x, y)
