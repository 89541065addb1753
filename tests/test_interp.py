"""Pin the generic NumPy interpreter (oracle/interp.py) against the golden vectors produced by
the real reference, so that the randomized differential tests that use it as their oracle rest
on something checked.  CPU only."""
import numpy as np
import pytest

import xgrid_b200 as xgrid
from examples import workloads as W
from oracle import HostGrid
from oracle.interp import Interp


def eq(a, b):
    assert a.dtype == b.dtype and a.shape == b.shape
    assert np.array_equal(a, b, equal_nan=True)


def host(arr, mask=None):
    h = HostGrid(arr.shape, arr.dtype)
    h.now[...] = arr
    if mask is not None:
        h.boundary[...] = mask
    return h


@pytest.fixture()
def k64(tmp_path):
    xgrid.init(precision="double", cacheroot=str(tmp_path))
    return W.make_kernels()


@pytest.mark.parametrize("name,kernel", [
    ("conv1d_f64", "convection_1d"), ("conv1d_stale_f64", "convection_1d"),
    ("conv1d_nonlinear_f64", "convection_1d_nonlinear"), ("diff1d_f64", "diffusion_1d"),
    ("conv2d_f64", "convection_2d"), ("diff2d_f64", "diffusion_2d"),
])
def test_single_grid(golden, k64, name, kernel):
    g = golden(name)
    u = host(g["u_in"], g["mask"])
    run = Interp(k64[kernel])
    for _ in range(int(g["steps"])):
        run(u, *[float(x) for x in g["params"]])
    eq(u._data[0], g["u.L0"])
    eq(u._data[1], g["u.L1"])


def test_cavity_41(golden, k64):
    g = golden("cavity_41_f64")
    b, p, u, v = (HostGrid((41, 41)) for _ in range(4))
    b.boundary[...], p.boundary[...], u.boundary[...], v.boundary[...] = g["mb"], g["mp"], g["mu"], g["mv"]
    cfg = W.Config(*[float(x) for x in g["cfg"]])
    run = Interp(k64["cavity_kernel"])
    for _ in range(int(g["steps"])):
        run(b, p, u, v, cfg)
    for tag, grid in (("b", b), ("p", p), ("u", u), ("v", v)):
        eq(grid._data[0], g[f"{tag}.L0"])
        eq(grid._data[1], g[f"{tag}.L1"])


def test_ewmul_ring(golden, k64):
    g = golden("ewmul_f64")
    a, b, r = host(g["a_in"]), host(g["b_in"]), HostGrid((10000,))
    run = Interp(k64["elementwise_mul"])
    run(r, a, b)
    run(r, a, b)
    for tag, grid in (("r2", r), ("a2", a), ("b2", b)):
        eq(grid._data[0], g[f"{tag}.L0"])
        eq(grid._data[1], g[f"{tag}.L1"])


def test_fp32_mixed_precision(golden, tmp_path):
    xgrid.init(cacheroot=str(tmp_path))          # precision="float"
    k = W.make_kernels()
    for name, kern in (("diff1d_f32", "diffusion_1d"), ("conv2d_f32", "convection_2d")):
        g = golden(name)
        u = host(g["u_in"], g["mask"])
        run = Interp(k[kern])
        for _ in range(int(g["steps"])):
            run(u, *[float(x) for x in g["params"]])
        eq(u._data[0], g["u.L0"])
        eq(u._data[1], g["u.L1"])


def _random_cases(name):
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz")
    data = np.load(path)
    return data, sorted({int(k.split(".")[0]) for k in data.files})


_RAND = {mode: _random_cases(name) for mode, name in
         (("none", "randprog"), ("wrap", "randprog_wrap"), ("limit", "randprog_limit"), ("f32", "randprog_f32"))}


@pytest.mark.parametrize("mode,seed", [(m, s) for m, (_, seeds) in _RAND.items() for s in seeds])
def test_random_programs_match_reference(tmp_path, mode, seed):
    """tests/golden/randprog*.npz: the programs of tests/randprog.py run by the UNMODIFIED reference
    (tests/golden/make_random_golden.py), with overstep="none" (cells whose taps could leave the array are
    masked out: undefined in the reference), "wrap" and "limit".  The interpreter must reproduce every ring
    level bit for bit."""
    from randprog import gen_inputs, gen_source, guard_array_ends, load_program
    data = _RAND[mode][0]
    dtype = np.float32 if mode == "f32" else np.float64
    if mode == "f32":       # fp32 grids / scalars, double literals
        xgrid.init(precision="float", cacheroot=str(tmp_path / "xg"))
    else:
        xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"), overstep=mode)
    ndim, ngrids, single, *shape = (int(x) for x in data[f"{seed}.meta"])
    shape = tuple(shape)
    src = gen_source(seed, ndim, ngrids, single_1d=bool(single))
    assert src == str(data[f"{seed}.src"]), "tests/randprog.py changed: regenerate tests/golden/randprog*.npz"
    prog = load_program(src, str(tmp_path), f"randprog_{seed}")
    ics, masks = gen_inputs(seed, shape, ngrids)
    if mode in ("none", "f32"):
        guard_array_ends(masks, shape)
    grids = [host(ic.astype(dtype), m) for ic, m in zip(ics, masks)]
    run = Interp(prog)
    for _ in range(3):
        run(*grids, 0.3, 1.7)
    for n, g in enumerate(grids):
        assert len(g._data) == int(data[f"{seed}.g{n}.depth"])
        for lvl, arr in enumerate(g._data):
            want = data[f"{seed}.g{n}.L{lvl}"]
            assert arr.dtype == want.dtype == dtype
            bad = np.argwhere(~((arr == want) | (np.isnan(arr) & np.isnan(want))))
            assert len(bad) == 0, f"g{n} level {lvl}: {len(bad)} cells differ, first {bad[:5].tolist()}\n{src}"


# ---- ring protocol: resize / truncate / rotate-per-argument / tick=False, after every call
def _ring_scenarios():
    import ring_cases as RC
    return sorted(RC.SCENARIOS)


@pytest.mark.parametrize("scenario", _ring_scenarios())
def test_ring_protocol_matches_reference(tmp_path, scenario):
    """tests/golden/ring_protocol.npz (tests/golden/make_ring_golden.py): the unmodified reference's ring after
    every call of the scenarios in tests/ring_cases.py; HostGrid + the interpreter must agree bit for bit."""
    import os
    import ring_cases as RC
    import importlib.util
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ring_protocol.npz"))
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"))
    path = tmp_path / "ring_kernels.py"
    path.write_text(RC.SOURCE.replace("IMPORT_LINE", "import xgrid_b200 as xgrid"))
    spec = importlib.util.spec_from_file_location("ring_kernels", str(path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    grids = {}
    for which in "gh":
        ic, mask = RC.initial(np, which)
        grids[which] = host(ic, mask)
    runners = {}
    for step, (kernel, spec_) in enumerate(RC.SCENARIOS[scenario]):
        run = runners.setdefault(kernel, Interp(getattr(mod, kernel)))
        run(*RC.arguments(spec_, grids))
        for which, g in grids.items():
            assert len(g._data) == int(gold[f"{scenario}.{step}.{which}.depth"]), (step, kernel, which)
            for lvl, arr in enumerate(g._data):
                want = gold[f"{scenario}.{step}.{which}.L{lvl}"]
                assert np.array_equal(arr, want, equal_nan=True), (scenario, step, kernel, which, lvl)
