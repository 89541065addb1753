"""Whole-domain NumPy model of the two-steps-per-pass schedule on slabs (lang/schedule.py::_run_batch2): every rank --
first, middle (both neighbours), last -- computes its interior rows from its OWN rows only (the ghost rows are
poisoned while it does) and the rows next to a cut step-at-a-time on the row ranges `two_step_slab_rows` returns, with
one-row-deep halo exchanges plus the overhang of diagonal taps.  The result must equal two plain steps on the whole
grid, for axis-aligned, one-sided and diagonal stencils with linear-address (F10) taps."""
import numpy as np
import pytest

from xgrid_b200.lang.schedule import two_step_slab_rows

STENCILS = {
    "star": [(0, 0), (0, 1), (0, -1), (1, 0), (-1, 0)],
    "upwind": [(0, 0), (-1, 0), (0, -1)],
    "diagonal": [(-1, -1), (-1, 1), (1, -1), (1, 2), (0, 0)],
    "downwind2": [(0, 0), (2, 0), (1, 1)],
    "along_rows": [(0, 0), (0, -1), (0, 2)],               # no axis-0 offset at all: the linear wrap still crosses the cut
}


def step_rows(src_flat, base, cols, rows, taps, coef):
    """One step for the given rows of a slab whose row 0 starts at flat index `base` of `src_flat` (ghost rows and
    padding around it): linear addressing, like the generated kernels."""
    out = {}
    for r in rows:
        at = base + r * cols + np.arange(cols)
        acc = np.zeros(cols)
        for (d0, dk), c in zip(taps, coef):
            acc = acc + c * src_flat[at + d0 * cols + dk]
        out[r] = acc
    return out


def global_two_steps(u, taps, coef):
    n0, cols = u.shape
    pad = 4 * cols
    levels = [u]
    for _ in range(2):
        flat = np.concatenate([np.zeros(pad), levels[-1].ravel(), np.zeros(pad)])
        rows = step_rows(flat, pad, cols, range(n0), taps, coef)
        levels.append(np.stack([rows[r] for r in range(n0)]))
    return levels[1], levels[2]


@pytest.mark.parametrize("name", sorted(STENCILS))
@pytest.mark.parametrize("world", [2, 3, 5])
def test_slab_schedule_equals_two_global_steps(name, world):
    taps = STENCILS[name]
    rng = np.random.default_rng(hash((name, world)) % 1000)
    coef = rng.random(len(taps))
    cols, n0s = 12, [9 + (r % 2) for r in range(world)]           # uneven slabs
    G = sum(n0s)
    u = rng.random((G, cols))
    want1, want2 = global_two_steps(u, taps, coef)
    dmin, dmax = min(t[0] for t in taps), max(t[0] for t in taps)
    h = max(-dmin, dmax, 1)
    over = max([abs(dk) for d0, dk in taps if abs(d0) == h and d0 * dk > 0] + [0])
    starts = np.cumsum([0] + n0s)
    G_H = h + 1                                                   # ghost rows allocated (room for the overhang)
    pad = G_H * cols

    def slab(level):                                              # [ghost | rows | ghost], flat, zero ghosts
        return [np.concatenate([np.zeros(pad), level[starts[r]:starts[r + 1]].ravel(), np.zeros(pad)]) for r in range(world)]

    def exchange(bufs):
        """h rows plus `over` elements across every cut, both directions (dist.exchange's pointer arithmetic)."""
        for r in range(world - 1):
            lo, hi = bufs[r], bufs[r + 1]
            n_lo = n0s[r] * cols
            cnt = h * cols + over
            hi[pad - cnt:pad] = lo[pad + n_lo - cnt:pad + n_lo]              # upper rank's lower ghost <- my last rows (+)
            lo[pad + n_lo:pad + n_lo + cnt] = hi[pad:pad + cnt]              # my upper ghost <- its first rows (+)

    x0 = slab(u)
    d = [np.full_like(b, np.nan) for b in x0]                     # what the bands see: THEIR step-1 rows only (any pass)
    d_last = [np.full_like(b, np.nan) for b in x0]                # plus the interior's middle rows (last pass of a batch)
    x1 = [np.full_like(b, np.nan) for b in x0]
    plans = [two_step_slab_rows(n0s[r], dmin, dmax, r > 0, r < world - 1) for r in range(world)]
    # interior: two steps from the rank's own rows only -- ghosts poisoned to prove it
    for r in range(world):
        r_lo, r_hi, deep, edge = plans[r]
        own = x0[r].copy()
        if r > 0:
            own[:pad] = np.nan
        if r < world - 1:
            own[pad + n0s[r] * cols:] = np.nan
        # middle rows the interior's outputs need, one more on each side for the linear wrap of step 2's taps
        mid_rows = range(max(0, r_lo + dmin - 1), min(n0s[r], r_hi + dmax + 1))
        mid = step_rows(own, pad, cols, mid_rows, taps, coef)
        mid_flat = np.full_like(own, np.nan)
        if r == 0:
            mid_flat[:pad] = 0.0                                   # the global ends read ghost zeros
        if r == world - 1:
            mid_flat[pad + n0s[r] * cols:] = 0.0
        for q, v in mid.items():
            mid_flat[pad + q * cols:pad + (q + 1) * cols] = v
        out = step_rows(mid_flat, pad, cols, range(r_lo, r_hi), taps, coef)
        for o, v in out.items():
            assert not np.isnan(v).any(), f"rank {r}: interior row {o} reads across the cut"
            x1[r][pad + o * cols:pad + (o + 1) * cols] = v
        for q in range(r_lo, r_hi):                               # the LAST pass of a batch stores the middle level too
            d_last[r][pad + q * cols:pad + (q + 1) * cols] = mid[q]
    # bands: step 1 on `deep` (halo of u^n exchanged), halo of the middle level exchanged, step 2 on `edge`
    exchange(x0)
    for r in range(world):
        for lo, hi in plans[r][2]:
            for q, v in step_rows(x0[r], pad, cols, range(lo, hi), taps, coef).items():
                d[r][pad + q * cols:pad + (q + 1) * cols] = v
        d[r][:pad] = 0.0                                           # ghost zeros; the exchange fills them across cuts
        d[r][pad + n0s[r] * cols:] = 0.0
    exchange(d)
    for r in range(world):
        for lo, hi in plans[r][3]:
            for o, v in step_rows(d[r], pad, cols, range(lo, hi), taps, coef).items():
                assert not np.isnan(v).any(), f"rank {r}: band row {o} reads a middle row nobody computed"
                x1[r][pad + o * cols:pad + (o + 1) * cols] = v
    for r in range(world):
        body = slice(pad, pad + n0s[r] * cols)
        assert np.array_equal(x1[r][body].reshape(n0s[r], cols), want2[starts[r]:starts[r + 1]]), f"rank {r}: u^(n+2)"
        full = np.where(np.isnan(d[r][body]), d_last[r][body], d[r][body])          # bands and interior together
        assert np.array_equal(full.reshape(n0s[r], cols), want1[starts[r]:starts[r + 1]]), f"rank {r}: u^(n+1)"
        both = ~np.isnan(d[r][body]) & ~np.isnan(d_last[r][body])                   # rows written by both agree
        assert np.array_equal(d[r][body][both], d_last[r][body][both])
