"""Ring-protocol scenarios: what `Operator.__call__` does to the time ring --
resize to the kernel depth (extend with zero levels AND truncate), rotate once per Grid ARGUMENT, `tick=False`,
level defaults -- observed after every call.  Run by the UNMODIFIED reference
(tests/golden/make_ring_golden.py -> ring_protocol.npz) and replayed through oracle.HostGrid + oracle/interp.py
on CPU (tests/test_interp.py), which the GPU edge-case tests use as their oracle.  The first and last two cells
carry mask value 7 (no statement matches), so no tap leaves the array (undefined in the reference)."""

SOURCE = """\
IMPORT_LINE

f1 = xgrid.grid[float, 1]


@xgrid.kernel()
def deep(u: f1) -> None:
    u[0] = 0.5 * u[0] + 0.25 * u[1][2] + 0.25 * u[-1][-2]


@xgrid.kernel()
def shallow(u: f1) -> None:
    u[0] = u[0] * 0.5 + u[-1]


@xgrid.kernel()
def flat(u: f1) -> None:
    u[0] = 3.0


@xgrid.kernel()
def both(a: f1, b: f1) -> None:
    a[0] = b[0] + 1.0


@xgrid.kernel(tick=False)
def inplace(u: f1, a: float) -> None:
    u[0] = u[0][0] * a + u[0]


@xgrid.kernel()
def older(u: f1, v: f1) -> None:
    u[0] = v[0][1] + 0.5 * v[1][2]
    v[0] = u[0][0] - v[-1]
"""

N = 96
# scenario name -> list of (kernel name, argument spec); "g" / "h" are the scenario's two grids
SCENARIOS = {
    "depth_changes": [("deep", "g"), ("deep", "g"), ("shallow", "g"), ("deep", "g"), ("flat", "g"),
                      ("shallow", "g"), ("deep", "g")],
    "same_grid_twice": [("both", "gg"), ("both", "gg"), ("both", "gg"), ("shallow", "g"), ("both", "gg")],
    "tick_false": [("inplace", "g0.5"), ("inplace", "g0.5"), ("shallow", "g"), ("inplace", "g0.5")],
    "two_grids_mixed_depth": [("older", "gh"), ("older", "gh"), ("shallow", "h"), ("older", "hg"), ("deep", "g"),
                              ("older", "gh")],
}


def initial(np, which: str):
    rng = np.random.default_rng({"g": 21, "h": 22}[which])
    ic = rng.uniform(-1.0, 1.0, N)
    mask = np.zeros(N, np.int32)
    mask[:2] = 7
    mask[-2:] = 7
    mask[N // 3] = 7                    # a never-written interior cell keeps rotating stale values (F5)
    return ic, mask


def arguments(spec: str, grids: dict) -> list:
    out, rest = [], spec
    while rest and rest[0] in "gh":
        out.append(grids[rest[0]])
        rest = rest[1:]
    if rest:
        out.append(float(rest))
    return out
