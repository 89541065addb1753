"""xgb::InvDiv (csrc/templates/xgb_stencil.cuh): fp64 division by a point-independent divisor through a
per-thread reciprocal must be IEEE-exact.  Checked bit-for-bit against the host's division (the
reference's generated C computes `a / c` with SSE2 divsd) over exponent sweeps, special values,
near-exact quotients (the hard cases for rounding) and adversarial divisors, through the DSL:
`r[0, 0] = a[0, 0] / c` on a grid wide enough for the shared-memory pipeline variant, the one
that hoists invariant subexpressions."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import xgrid_b200 as xgrid

SHAPE = (64, 4096)
N = SHAPE[0] * SHAPE[1]


def families(c: float, rng) -> list:
    out = []
    mant = 1.0 + rng.random(N)
    sign = np.where(rng.random(N) < 0.5, -1.0, 1.0)
    out.append(sign * np.ldexp(mant, rng.integers(-1074, 1024, N)))            # every exponent, subnormal to huge
    out.append(sign * np.ldexp(mant, rng.integers(-60, 60, N)))                 # the range PDE data lives in
    with np.errstate(all="ignore"):
        q = sign * np.ldexp(1.0 + rng.random(N), rng.integers(-300, 300, N))
        out.append(q * c)                                                        # x = RN(q c): x / c lands next to q
        out.append(np.nextafter(q * c, np.inf))
        out.append(np.nextafter(q * c, -np.inf))
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, -5e-324, 2.2250738585072014e-308,
                        1.7976931348623157e308, -1.7976931348623157e308, 1.0, -1.0, 3.0, 1 / 3,
                        2.0 ** -900, np.nextafter(2.0 ** -900, 0), 2.0 ** 900, np.nextafter(2.0 ** 900, 0),
                        2.0 ** -1022, 2.0 ** 1023, 4.9e-320, 1e-310])
    tail = np.resize(special, N)
    tail[special.size:] *= np.resize(mant, N)[special.size:]
    out.append(tail)
    return out


DIVISORS = [1.0, 3.0, 1 / 3, 0.1, np.pi, -7.0, np.nextafter(2.0, 0.0), np.nextafter(1.0, 2.0), 2.0 ** -100,
            np.nextafter(2.0 ** -100, 0.0), 2.0 ** 100, 2.0 ** 101, 1e-300, 1e300, 2.0 / 8191, 2.0 * (2.0 / 8191) ** 2,
            6.103515625e-05, 0.5, 1.9999999999999996e-4, -2.4424906541753444e-08]


@pytest.fixture(scope="module")
def divk(tmp_path_factory):
    from xgrid_b200.lang import cudagen
    saved = set(cudagen.INVDIV_VARIANTS)
    cudagen.INVDIV_VARIANTS.add("tiled")          # by default only the instruction-bound variants use it
    xgrid.init(precision="double", cacheroot=str(tmp_path_factory.mktemp("xgdiv")))
    f2 = xgrid.grid[float, 2]

    @xgrid.kernel()
    def divide(r: f2, a: f2, c: float) -> None:
        r[0, 0] = a[0, 0] / c

    assert "InvDiv" in divide.src
    yield divide
    cudagen.INVDIV_VARIANTS.clear()
    cudagen.INVDIV_VARIANTS.update(saved)


@pytest.mark.parametrize("c", DIVISORS)
def test_invariant_divisor_is_exact(divk, c):
    from xgrid_b200.lang.launch import STATS
    rng = np.random.default_rng(abs(hash(float(c))) % (2 ** 32))
    r, a = xgrid.Grid(SHAPE, float), xgrid.Grid(SHAPE, float)
    before = STATS.get("tiled", 0)
    for fam, x in enumerate(families(float(c), rng)):
        a.now[...] = x.reshape(SHAPE)
        divk(r, a, float(c))            # a ticks: the values sit one level back, where the kernel reads them
        got = r.now.reshape(-1)
        with np.errstate(all="ignore"):
            want = x / float(c)
        same = (got.view(np.int64) == want.view(np.int64)) | (np.isnan(got) & np.isnan(want))
        if not same.all():
            bad = np.flatnonzero(~same)[:5]
            raise AssertionError(f"c={c!r} family {fam}: {int((~same).sum())} quotients differ, e.g. "
                                 f"x={x[bad].tolist()} got={got[bad].tolist()} want={want[bad].tolist()}")
    assert STATS.get("tiled", 0) > before          # the variant with hoisted reciprocals ran
