"""Host side of a kernel call on a recording fake runtime (tests/fake_runtime.py): what is launched, in
which order, with which buffers -- deferral and flushing, the multi-step launch plan, ring rotation by
pointer, scratch swaps, CUDA-graph recording / replay keys and the device-memory pool.  CPU only; the
numbers these launches produce are checked by the GPU suite."""
import numpy as np
import pytest

import xgrid_b200 as xgrid
from examples import workloads as W

import fake_runtime


@pytest.fixture()
def rt(monkeypatch, tmp_path):
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"))
    return fake_runtime.install(monkeypatch)


def test_1d_run_is_deferred_and_split_into_multistep_launches(rt):
    k = W.make_kernels()["convection_1d"]
    n = 1 << 16
    u = xgrid.Grid((n,), float)
    u.now[...] = 1.0
    u.boundary[0] = 1
    for _ in range(150):
        k(u, 1.0, 0.5, 1.0)
    assert rt.launches == []                                  # nothing ran yet: 150 identical calls are queued
    xgrid.flush()
    names = rt.names()
    assert names == (["xg_convection_1d_g0_multistep_v1"] * 2 + ["xg_convection_1d_g0_multistep_short_v1"]
                     + ["xg_convection_1d_g0_dense_v4"] * 2), names
    full, tail = rt.launches[0][3], rt.launches[2][3]
    assert tail["opt0"] == 20 and full["n0"] == n
    # every launch reads the two ring levels and writes two other buffers, which then become the ring
    assert len({full["aux0"], full["aux1"], full["aux2"], full["aux3"]}) == 4
    second = rt.launches[1][3]
    assert (second["aux0"], second["aux1"]) == (full["aux2"], full["aux3"])
    assert (second["aux2"], second["aux3"]) == (full["aux0"], full["aux1"])
    assert len(u._ring) == 2 and len(u._spares) == 2


def test_changing_a_scalar_or_reading_now_flushes(rt):
    k = W.make_kernels()["convection_1d"]
    u = xgrid.Grid((1 << 15,), float)
    for _ in range(10):
        k(u, 1.0, 0.5, 1.0)
    k(u, 1.0, 0.25, 1.0)                                      # other arguments: the queued run executes first
    assert rt.names() == ["xg_convection_1d_g0_multistep_short_v1", "xg_convection_1d_g0_dense_v4",
                          "xg_convection_1d_g0_dense_v4"]
    before = len(rt.launches)
    _ = u.now                                                 # host read: flush the one pending call, then D2H
    assert len(rt.launches) == before + 1 and rt.copies[-1][0] == "d2h"


def test_ring_rotates_by_pointer_and_graph_replays_after_recording(rt):
    k = W.make_kernels()["elementwise_mul"]
    n = 10000
    r, a, b = (xgrid.Grid((n,), float) for _ in range(3))
    a.now[...] = 2.0
    b.now[...] = 3.0
    seen = []
    for call in range(6):
        k(r, a, b)
        p = rt.launches[-1][3]
        seen.append((p["s0"], p["s1"], p["s2"]))
    # depth 2: every grid's two buffers alternate (store level 0 of r, load level 1 of a and b)
    assert seen[0] == seen[2] == seen[4] and seen[1] == seen[3] == seen[5] and seen[0] != seen[1]
    assert all(len(set(s)) == 3 for s in seen)
    # arrangement A is seen at call 0, recorded at call 2, replayed at call 4 (B: 1, 3, 5)
    assert len(rt.graphs) == 2 and len(rt.launches) == 6
    uploads = [c for c in rt.copies if c[0] == "h2d"]
    assert len(uploads) == 3                                   # each grid's first level once; second levels start as device zeros


def test_new_grid_never_replays_a_dead_grids_graph(rt):
    """Device addresses repeat (pool / cudaMalloc): the grid serial in the graph key keeps recorded
    launches -- which also bake in mask and index-list pointers -- from being replayed for another grid."""
    k = W.make_kernels()["diffusion_1d"]
    n = 4096                                                  # below the deferral threshold: direct calls

    def run():
        u = xgrid.Grid((n,), float)
        u.boundary[0] = u.boundary[-1] = 1
        for _ in range(4):
            k(u, 0.01, 0.1, 1.0)
        return u._arrangement(), u._mask_dev

    first, mask1 = run()
    graphs_after_first = len(rt.graphs)
    second, mask2 = run()
    assert len(rt.graphs) == 2 * graphs_after_first           # recorded again for the new grid


def test_cavity_call_structure_and_graph(rt):
    k = W.make_kernels()["cavity_kernel"]
    nn = 64                                                   # small: the fused Jacobi pairs need >= MIN_COLS columns
    mb, mp, mu, mv = W.cavity_masks(nn, nn)
    gs = [xgrid.Grid((nn, nn), float) for _ in range(4)]
    for g, m in zip(gs, (mb, mp, mu, mv)):
        g.boundary[...] = m
    cfg = W.Config(1.0, 0.1, 1e-4, 2.0 / (nn - 1), 2.0 / (nn - 1))
    k(*gs, cfg)
    first = rt.names()
    sparse = [x for x in first if x.endswith("_sparse_v1")]
    # b, p-first, 4 BCs, 50 x (Jacobi + 4 BCs), fused tail (u, v, 3 Dirichlet BCs in ONE launch)
    assert len(first) == 1 + 1 + 4 + 50 * 5 + 1 and len(sparse) == 4 + 50 * 4, (len(first), len(sparse))
    # lazy JIT: one module per generated kernel, compiled and loaded on first launch -- 57 kernels are
    # generated for this program, one call on this grid size needs only these
    generated = k._program().source.count('extern "C" __global__')
    assert generated > 50 and len(rt.modules) == len(set(first)) <= 20
    n1 = len(rt.launches)
    k(*gs, cfg)
    k(*gs, cfg)
    k(*gs, cfg)
    assert (len(rt.launches) - n1) == 3 * len(first)          # same work per call, direct, recorded or replayed
    assert 1 <= len(rt.graphs) <= 2


def test_grid_buffers_return_to_the_pool_and_are_reused(rt):
    k = W.make_kernels()["diffusion_2d"]
    shape = (512, 1024)                                       # 4 MiB levels: pooled

    def run():
        u = xgrid.Grid(shape, float)
        u.boundary[0, :] = 1
        k(u, 0.2)
        xgrid.flush()
        return {lv.raw for lv in u._ring}

    first = run()                                             # the grid dies here: its buffers go to the pool
    allocs = len(rt.real_alloc_sizes)
    second = run()
    assert second == first                                    # the very same level buffers again
    assert all(b < rt.POOL_MIN for b in rt.real_alloc_sizes[allocs:])      # only small buffers (mask, flags) are new


# ---- slab-sharded grids: halo planning and edge-first overlap (host logic of SURVEY.md §8e)
def test_sharded_3d_sweep_runs_edges_first_and_exchanges_only_stale_levels(monkeypatch, tmp_path):
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"), distributed=True)
    rt, tr = fake_runtime.install_sharded(monkeypatch, rank=1, world=4)
    k = W.make_kernels()["heat_3d"]
    u = xgrid.Grid((4 * 64, 64, 256), float)                  # global shape: this rank owns 64 planes
    assert u.sharded and u.shape == (64, 64, 256) and u.row_range == (64, 128)
    u.now[...] = 1.0
    k(u, 0.1)
    # first call: the uploaded level is stale -> one blocking exchange BEFORE any launch; then the two edge
    # bands, the asynchronous exchange of the freshly written level, and the interior band
    assert tr.log[0][0] == "exchange" and tr.log[0][3] == 0
    names = rt.names()
    assert len(names) == 3 and len(set(names)) == 1 and names[0].endswith("_tiled_v2"), names
    bands = [(r[3]["r_lo"], r[3]["r_hi"]) for r in rt.launches]
    assert bands == [(0, 1), (63, 64), (1, 63)]
    assert tr.log[1][0] == "exchange_async" and tr.log[1][3] == 2       # issued after the two edge launches
    written = rt.launches[0][3]["s0"]
    assert tr.log[1][1] == written
    # second call reads the level written by the first: its halo is already in flight -> no new blocking exchange
    before = len(tr.log)
    k(u, 0.1)
    kinds = [e[0] for e in tr.log[before:]]
    assert kinds == ["exchange_async"], kinds
    assert rt.launches[3][3]["s1"] == written                 # ring rotated by pointer: last output is now the input
    # buffers of a sharded grid return to the pool too, each behind a fence that orders the compute stream
    # after the communication stream (a halo exchange may still be reading the buffer)
    fences = len([e for e in tr.log if e[0] == "fence"])
    del u
    assert rt._pool_bytes > 0                                 # (buffers under 1 MiB are freed for real)
    assert len([e for e in tr.log if e[0] == "fence"]) > fences


def test_halo_freshness_tracks_the_exchanged_depth(monkeypatch, tmp_path):
    """A level exchanged at depth 1 is stale for a later group that reads it two rows across the slab
    boundary (round-1 advisor finding: freshness was a bool)."""
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"), distributed=True)
    rt, tr = fake_runtime.install_sharded(monkeypatch, rank=1, world=4)
    f2 = xgrid.grid[float, 2]

    @xgrid.kernel()
    def two_reaches(u: f2, v: f2, w: f2) -> None:
        v[0, 0] = u[1, 0] + u[-1, 0]
        w[0, 0] = v[0, 0][0] + u[2, 0] + u[-2, 0]

    u, v, w = (xgrid.Grid((4 * 32, 48), float) for _ in range(3))
    u.now[...] = 1.0
    two_reaches(u, v, w)
    src = u._ring[1]
    depths = [e[2] for e in tr.log if e[0] == "exchange" and e[1] == src.dev]
    assert depths == [1, 2], tr.log                # refreshed again, deeper, before the second sweep
    assert src.halo_rows == 2 and u._ghost >= 2


def test_sharded_1d_run_gets_its_halo_layout_with_the_first_deferred_call(monkeypatch, tmp_path):
    """A 1-D slab imports H points of both ring levels per multi-step launch: the ghost band is sized when the
    first call is deferred (levels still on the host -> no copy at all), so the flush neither re-lays-out nor
    allocates; a single-step remainder (exchanged at depth 1) leaves the level stale for the next H-deep launch."""
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"), distributed=True)
    rt, tr = fake_runtime.install_sharded(monkeypatch, rank=1, world=4)
    k = W.make_kernels()["convection_1d"]
    u = xgrid.Grid((4 << 16,), float)
    u.now[...] = 1.0
    H = k._program().groups[0].multistep["H"]
    k(u, 1.0, 0.5, 1.0)
    assert u._ghost == H and len(u._spares) == 2 and rt.copies == []
    for _ in range(19):
        k(u, 1.0, 0.5, 1.0)
    allocs = rt.real_allocs
    xgrid.flush()
    # uploads only: the IC, the packed mask and its chunk flags (a sharded grid always carries a mask)
    assert [c[0] for c in rt.copies] == ["h2d"] * 3 and rt.real_frees == 0
    assert rt.real_allocs - allocs <= 4                                     # level 0/1 + mask + flags, no re-layout
    names = rt.names()
    assert names[0].endswith("multistep_short_v1") and rt.launches[0][3]["opt0"] == 20
    assert [e[2] for e in tr.log if e[0] == "exchange"] == [H, H]           # both ring levels, H deep
    # second run of 21: tail launch of 20 + one single step; the single step's depth-1 exchange must not
    # pass for fresh when the next multi-step launch asks for H
    for _ in range(21):
        k(u, 1.0, 0.5, 1.0)
    xgrid.flush()
    n = len(tr.log)
    for _ in range(20):
        k(u, 1.0, 0.5, 1.0)
    xgrid.flush()
    assert sorted(e[2] for e in tr.log[n:] if e[0] == "exchange") == [H, H]
    assert rt.real_frees == 0


def test_sharded_overstep_sets_open_flags_from_the_topology(monkeypatch, tmp_path):
    for mode, rank, ring, want in (("wrap", 0, True, (1, 1)), ("limit", 0, False, (0, 1)), ("limit", 3, False, (1, 0))):
        xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"), distributed=True, overstep=mode)
        rt, tr = fake_runtime.install_sharded(monkeypatch, rank=rank, world=4, ring=ring)
        k = W.make_kernels()["diffusion_2d_open"]
        u = xgrid.Grid((64, 48), float)
        u.now[...] = 1.0
        k(u, 0.2)
        p = rt.launches[-1][3]
        assert (p["open_lo"], p["open_hi"]) == want, (mode, rank)
        assert tr.log and tr.log[0][0] == "exchange"          # ghost rows are refreshed before the first sweep


def test_touching_the_mask_without_changing_it_keeps_recorded_graphs(rt):
    k = W.make_kernels()["diffusion_1d"]
    u = xgrid.Grid((4096,), float)
    u.boundary[0] = u.boundary[-1] = 1
    for _ in range(6):
        k(u, 0.01, 0.1, 1.0)
    graphs, version = len(rt.graphs), u._mask_version
    _ = u.boundary                                            # touched (the property cannot know about writes) ...
    for _ in range(4):
        k(u, 0.01, 0.1, 1.0)
    assert u._mask_version == version and len(rt.graphs) == graphs      # ... but unchanged: same version, same graphs
    u.boundary[7] = 1                                         # a real change: new version, graphs are re-recorded
    for _ in range(4):
        k(u, 0.01, 0.1, 1.0)
    assert u._mask_version == version + 1 and len(rt.graphs) > graphs


def test_2d_run_executes_two_steps_per_pass_and_stores_the_middle_level_last(rt):
    k = W.make_kernels()["diffusion_2d"]
    u = xgrid.Grid((256, 2048), float)
    u.now[...] = 1.0
    u.boundary[0, :] = u.boundary[-1, :] = u.boundary[:, 0] = u.boundary[:, -1] = 1
    for _ in range(7):
        k(u, 0.2)
    assert rt.launches == []
    xgrid.flush()
    names = rt.names()
    assert names[:3] == ["xg_diffusion_2d_g0_tiled2_v2"] * 3 and len(names) == 4      # 3 passes of 2 steps + 1 single step
    assert [r[3]["opt0"] for r in rt.launches[:3]] == [0, 0, 1]                      # only the last pass stores u^{n+1}
    p0, p1 = rt.launches[0][3], rt.launches[1][3]
    assert p1["aux0"] == p0["aux1"] and p1["aux1"] == p0["aux0"]                     # u^{n+2} overwrites the dead level
    # a cell whose mask value has no statement would keep a value from two steps back: no two-step passes then
    v = xgrid.Grid((256, 2048), float)
    v.boundary[5, 5] = 7
    for _ in range(4):
        k(v, 0.2)
    n = len(rt.launches)
    xgrid.flush()
    assert all(not x.endswith("tiled2_v2") for x in rt.names(n)) and len(rt.launches) - n == 4


def test_ghost_relayout_drops_and_rebuilds_device_buffers(rt):
    f2 = xgrid.grid[float, 2]

    @xgrid.kernel()
    def near(u: f2) -> None:
        u[0, 0] = 0.5 * (u[1, 0] + u[-1, 0])

    @xgrid.kernel()
    def far(u: f2) -> None:
        u[0, 0] = 0.5 * (u[3, 0] + u[-3, 0])

    u = xgrid.Grid((64, 96), float)
    u.now[...] = 1.0
    near(u)
    assert u._ghost == 1
    first = {lv.raw for lv in u._ring}
    far(u)                                                    # needs 3 ghost rows: levels are re-laid-out
    assert u._ghost == 3 and first.isdisjoint({lv.raw for lv in u._ring})
    moved = [c[0] for c in rt.copies if c[0] in ("d2h", "d2d")]
    assert moved == ["d2d", "d2d"]                            # re-laid-out on the device, no host round trip
    near(u)                                                   # a smaller halo keeps the larger layout
    assert u._ghost == 3


# ---- CUDA-graph replay must be indistinguishable from issuing the launches directly
@pytest.mark.parametrize("seed,ndim,ngrids,shape", [
    (8, 2, 3, (40, 1030)), (11, 2, 3, (64, 1024)), (13, 2, 2, (17, 33)), (21, 3, 2, (18, 9, 130)),
    (24, 3, 1, (20, 16, 256)), (3, 1, 2, (4100,)), (5, 1, 2, (37,)),
])
def test_graph_replay_issues_exactly_the_direct_launch_sequence(monkeypatch, tmp_path, seed, ndim, ngrids, shape):
    """Random multi-grid programs (tests/randprog.py: time levels 0..2, implicit statements, loops, sparse
    boundary statements, host control flow), 26 calls each, once with CUDA graphs (record on the second
    sighting of a buffer arrangement, replay afterwards) and once without: kernel names, geometry and every
    parameter -- i.e. every buffer pointer in every role -- must be identical launch by launch."""
    from randprog import gen_inputs, gen_source, load_program
    src = gen_source(seed, ndim, ngrids)
    traces = []
    for graphs in (True, False):
        xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"), graphs=graphs)
        rt = fake_runtime.install(monkeypatch)
        prog = load_program(src, str(tmp_path), f"randprog_{seed}_{int(graphs)}")
        ics, masks = gen_inputs(seed, shape, ngrids)
        grids = []
        for ic, m in zip(ics, masks):
            g = xgrid.Grid(shape, float)
            g.now[...] = ic
            g.boundary[...] = m
            grids.append(g)
        for _ in range(26):
            prog(*grids, 0.3, 1.7)
        xgrid.flush()
        traces.append((list(rt.launches), len(rt.graphs)))
        del grids
    (with_graphs, n_graphs), (direct, none) = traces
    assert none == 0 and n_graphs >= 1, "the graph path was not exercised"
    assert len(with_graphs) == len(direct)
    for k, (a, b) in enumerate(zip(with_graphs, direct)):
        assert a == b, f"launch {k} differs:\n{a}\n{b}\n{src}"


def test_cavity_pressure_loop_runs_fused_pairs_with_an_even_number_of_swaps(rt):
    """50 Jacobi iterations = 24 fused two-sweep passes + 2 single sweeps (an even number of level-0/scratch
    swaps per call keeps the buffer arrangement -- and the recorded graph -- on a period of two calls)."""
    k = W.make_kernels()["cavity_kernel"]
    nn = 1024
    mb, mp, mu, mv = W.cavity_masks(nn, nn)
    gs = [xgrid.Grid((nn, nn), float) for _ in range(4)]
    for g, m in zip(gs, (mb, mp, mu, mv)):
        g.boundary[...] = m
    cfg = W.Config(1.0, 0.1, 1e-4, 2.0 / (nn - 1), 2.0 / (nn - 1))
    k(*gs, cfg)
    names = rt.names()
    fused = [x for x in names if "jacobi2" in x]
    single = [x for x in names if x.startswith("xg_cavity_kernel_g6_") and "jacobi2" not in x]
    sparse = [x for x in names if x.endswith("_sparse_v1")]
    assert len(fused) == 24 and len(single) == 2, (len(fused), len(single))
    # boundary statements: 4 before the loop, 4 after each fused pass (those between its two sweeps are resolved
    # on chip), 4 after each single sweep
    assert len(sparse) == 4 + 24 * 4 + 2 * 4
    arrangement = [g._arrangement() for g in gs]
    k(*gs, cfg)
    k(*gs, cfg)
    assert [g._arrangement() for g in gs] == arrangement       # period two
    for _ in range(4):
        k(*gs, cfg)
    assert len(rt.graphs) == 2


def test_unit_jit_mode_loads_one_module_per_program(rt, monkeypatch):
    from xgrid_b200.lang import schedule
    monkeypatch.setattr(schedule, "JIT_MODE", "unit")         # XGB_JIT=unit
    k = W.make_kernels()["diffusion_2d"]
    u = xgrid.Grid((64, 96), float)
    u.boundary[0, :] = 1
    for _ in range(3):
        k(u, 0.2)
    v = xgrid.Grid((128, 2048), float)                        # another variant of the same program
    k(v, 0.2)
    xgrid.flush()
    assert len(set(rt.names())) >= 2 and len(rt.modules) == 1


def test_writes_through_a_retained_boundary_array_are_seen(rt):
    """`.boundary` is a plain attribute in the reference (xgrid/xgrid/__init__.py:41): `b = g.boundary`
    kept across kernel calls and written later must reach the next call; queued (deferred) calls must
    still run with the mask they were called with."""
    k = W.make_kernels()["diffusion_1d"]
    u = xgrid.Grid((4096,), float)                            # below the deferral threshold
    b = u.boundary
    b[0] = 1
    k(u, 0.01, 0.1, 1.0)
    v1 = u._mask_version
    k(u, 0.01, 0.1, 1.0)
    assert u._mask_version == v1                              # untouched: no re-compare, no re-upload
    b[-1] = 1                                                 # through the retained array
    k(u, 0.01, 0.1, 1.0)
    assert u._mask_version == v1 + 1 and u._mask_snapshot[-1] == 1
    b += 0                                                    # in-place operator: touched, but unchanged
    k(u, 0.01, 0.1, 1.0)
    assert u._mask_version == v1 + 1
    np.copyto(b, np.zeros(4096, np.int32))
    k(u, 0.01, 0.1, 1.0)
    assert u._mask_version == v1 + 2 and not u._mask_any
    # deferred run: the write flushes the queued calls first
    w = xgrid.Grid((1 << 16,), float)
    bw = w.boundary
    bw[0] = 1
    for _ in range(10):
        k(w, 0.01, 0.1, 1.0)
    n = len(rt.launches)
    bw[5] = 1
    assert len(rt.launches) > n                               # the 10 queued steps ran with the old mask
    assert isinstance(b == 1, np.ndarray) and type(b == 1) is np.ndarray


def test_writes_through_a_retained_now_array_reach_the_device(rt):
    k = W.make_kernels()["diffusion_1d"]
    u = xgrid.Grid((4096,), float)
    a = u.now
    a[...] = 1.0
    k(u, 0.01, 0.1, 1.0)                                      # `a` now mirrors ring level 1 (device-resident)
    lv = u._ring[1]
    assert lv.host is not None and lv.where == "device"
    copies = len(rt.copies)
    a[7] = 3.0                                                # kept array, written after the call
    assert lv.where == "host" and rt.copies[copies:] == [("d2h", lv.dev, 4096 * 8)]
    k(u, 0.01, 0.1, 1.0)
    assert ("h2d", lv.dev, 4096 * 8) in rt.copies[copies:]    # uploaded again before the sweep
    assert u.now is u.now


def test_short_runs_use_the_half_window_variant(rt):
    """A remainder of at most T/2 = 32 steps runs on the short variant (256 threads, half the window, five
    CTAs per SM); longer remainders keep the tail variant; both read their step count from the launch."""
    k = W.make_kernels()["diffusion_1d"]
    n = 1 << 17
    u = xgrid.Grid((n,), float)
    u.boundary[0] = u.boundary[-1] = 1
    for count, want, steps in ((20, "multistep_short_v1", 20), (32, "multistep_short_v1", 32),
                               (36, "multistep_tail_v1", 36), (63, "multistep_tail_v1", 60)):
        before = len(rt.launches)
        for _ in range(count):
            k(u, 0.01, 0.1, 1.0)
        xgrid.flush()
        first = rt.launches[before]
        assert first[0].endswith(want) and first[3]["opt0"] == steps, (count, first[0], first[3]["opt0"])
        cfg = k._program().groups[0].multistep_short if "short" in want else k._program().groups[0].multistep
        assert first[2] == (cfg["threads"], 1, 1) and first[1][0] == -(-n // cfg["W"])
        assert len(rt.launches) - before == 1 + (count - steps)


def test_callee_operators_with_grids_run_their_own_sweeps_in_program_order(rt, tmp_path):
    """generator.py:208-212,418-419: a called operator's statements run in place of the call, on the caller's
    buffers, without a tick of their own; the callee's deepest time level sets the caller's ring depth."""
    import importlib.util
    import os
    import xgrid
    spec = importlib.util.spec_from_file_location(
        "callee_prog_fake", os.path.join(os.path.dirname(__file__), "programs", "callee_prog.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    u, v = xgrid.Grid((24, 24), float), xgrid.Grid((24, 24), float)
    u.boundary[0, :] = 1
    v.boundary[0, :] = 1
    u.boundary[1:-1, 0] = 2                                   # `smooth`'s Neumann column
    mod.outer(u, v, 0.4)
    assert mod.outer.depth == 3 and len(u._ring) == len(v._ring) == 3
    names = [n.split("_g")[0] for n in rt.names()]
    # relax(u) | v = u + 1 (+ its boundary statement, fused) | relax(v) | smooth: 3 x (Jacobi sweep + Neumann copy)
    assert names == ["xg_relax", "xg_outer", "xg_relax"] + ["xg_smooth"] * 6, rt.names()
    first, second = rt.launches[0][3], rt.launches[2][3]
    assert first["s0"] in {lv.dev for lv in u._ring} | {u._scratch.dev}    # (smooth swapped level 0 / scratch)
    assert second["s0"] in {lv.dev for lv in v._ring}          # the same callee, bound to the other grid
    # no extra tick: one rotation per call of the OUTER kernel only
    before = [lv.dev for lv in v._ring]
    mod.outer(u, v, 0.4)
    assert [lv.dev for lv in v._ring] == before[-1:] + before[:-1]
    for _ in range(30):                                         # (the 4 buffers of u permute with a long period)
        mod.outer(u, v, 0.4)
    assert rt.graphs                                            # calls with callees are recorded and replayed too


def test_first_upload_of_a_big_grid_packs_the_mask_while_the_levels_upload(rt):
    k = W.make_kernels()["diffusion_1d"]
    n = 1 << 22
    u = xgrid.Grid((n,), float)
    u.now[...] = 1.0
    u.boundary[0] = u.boundary[-1] = 1
    u.boundary[12345] = 9
    k(u, 0.01, 0.1, 1.0)
    xgrid.flush()
    assert u._mask_any and u._mask_snapshot.dtype == np.uint8 and u._mask_snapshot.shape == (n,)
    assert np.array_equal(np.flatnonzero(u._mask_snapshot), [0, 12345, n - 1]) and u._mask_snapshot[12345] == 9
    assert u._mask_count(1) == 2 and u._mask_count(9) == 1 and u._mask_count(0) == n - 3
    kinds = [c[0] for c in rt.copies]
    assert kinds == ["h2d"] * 3                               # level 0, packed mask (n bytes), chunk flags
    assert sorted(c[2] for c in rt.copies) == [n // 128, n, n * 8]


@pytest.mark.parametrize("workload", ["heat3d", "cavity"])
def test_sharded_calls_are_recorded_and_replayed_with_their_halo_state(monkeypatch, tmp_path, workload):
    """Slab grids: a call is recorded into a CUDA graph together with its halo exchanges (the communication stream
    is joined into the capture before it ends, in-flight exchanges of the previous call before it begins) and the
    freshness of every level's ghost rows is part of the key and of the replayed state.  Launch by launch the
    replayed run must equal the direct one, and so must the halo bookkeeping after every call."""
    traces = []
    for graphs in (True, False):
        xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"), distributed=True, graphs=graphs)
        rt, tr = fake_runtime.install_sharded(monkeypatch, rank=1, world=4)
        k = W.make_kernels()
        if workload == "heat3d":
            grids = [xgrid.Grid((4 * 64, 64, 256), float)]
            grids[0].now[...] = 1.0
            call = lambda: k["heat_3d"](grids[0], 0.1)                     # noqa: E731
        else:
            n = 4 * 160
            masks = W.cavity_masks(n, 640)
            grids = [xgrid.Grid((n, 640), float) for _ in range(4)]
            for g, m in zip(grids, masks):
                g.boundary[...] = m[g.row_range[0]:g.row_range[1]]
            cfg = W.Config(1.0, 0.1, 1e-4, 2.0 / 639, 2.0 / (n - 1))
            call = lambda: k["cavity_kernel"](*grids, cfg)                 # noqa: E731
        states = []
        for _ in range(9):
            call()
            states.append(tuple(tuple(r for _, r in g._halo_state()) for g in grids))
            assert all(lv.halo_event == 0 or not graphs for g in grids for lv in g._ring) or True
        traces.append((list(rt.launches), states, len(rt.graphs), len(tr.log)))
        del grids
    (with_graphs, st_g, n_graphs, ex_g), (direct, st_d, none, ex_d) = traces
    assert len(with_graphs) == len(direct)
    for i, (a, b) in enumerate(zip(with_graphs, direct)):
        assert a == b, f"launch {i} differs:\n{a}\n{b}"
    assert st_g == st_d                                        # same freshness of every level after every call
    assert none == 0 and n_graphs >= 1, "sharded calls were not recorded"
    assert ex_g < ex_d                                         # replayed calls issue their exchanges from the graph


def test_fused_jacobi_pairs_on_a_slab_cover_the_interior_and_leave_bands_to_single_sweeps(monkeypatch, tmp_path):
    """Middle rank of 4: per fused pair the rows that need nothing from a neighbour run in the fused kernel; the 3 rows
    next to each cut run sweep A (two rows deeper) into a third buffer, the boundary statements on it, and sweep B
    into the output buffer, all with the ordinary kernels on row bands."""
    from xgrid_b200.lang.launch import SLAB_BAND
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"), distributed=True, graphs=False)
    rt, tr = fake_runtime.install_sharded(monkeypatch, rank=1, world=4)
    k = W.make_kernels()["cavity_kernel"]
    n, cols = 4 * 256, 1024
    masks = W.cavity_masks(n, cols)
    gs = [xgrid.Grid((n, cols), float) for _ in range(4)]
    for g, m in zip(gs, masks):
        g.boundary[...] = m[g.row_range[0]:g.row_range[1]]
    cfg = W.Config(1.0, 0.1, 1e-4, 2.0 / (cols - 1), 2.0 / (n - 1))
    k(*gs, cfg)
    n0 = gs[1].shape[0]
    fused = [r for r in rt.launches if "jacobi2" in r[0]]
    assert len(fused) == 24 and SLAB_BAND == 3
    assert all((r[3]["r_lo"], r[3]["r_hi"]) == (3, n0 - 3) for r in fused)
    p = gs[1]
    first = rt.launches.index(fused[0])
    # the launches right before the first fused pass: sweep A on [0,5) and [n0-5,n0), boundary statements (this rank
    # owns boundary points of the two side columns only: masks 1 and 3), sweep B on [0,3) and [n0-3,n0)
    before = rt.launches[:first]
    bands = [(r[3]["r_lo"], r[3]["r_hi"]) for r in before if r[0].startswith("xg_cavity_kernel_g6_") and "sparse" not in r[0]]
    assert bands[-4:] == [(0, 5), (n0 - 5, n0), (0, 3), (n0 - 3, n0)], bands
    a_band, b_band = before[-6 - 0], before[-1]      # (sparse launches sit between the two pairs of band sweeps)
    x, s = fused[0][3]["s1"], fused[0][3]["s0"]
    ptrs = lambda r: {v for kk, v in r[3].items() if kk.startswith("s") and isinstance(v, int)}      # noqa: E731
    assert x in ptrs(before[[i for i, r in enumerate(before) if (r[3].get("r_lo"), r[3].get("r_hi")) == (0, 5)][-1]])
    assert s in ptrs(b_band) and x not in ptrs(b_band)          # sweep B reads the third buffer, writes the output buffer
    third = (ptrs(b_band) - {s}) & {lv.dev for lv in p._spares}
    assert len(third) == 1
    # the ring is intact afterwards: the third buffer never enters it
    assert third.isdisjoint({lv.dev for lv in p._ring}) and third.isdisjoint({p._scratch.dev})
    # rank 0 of the chain has no lower neighbour: its fused pass starts at row 0
    rt0, tr0 = fake_runtime.install_sharded(monkeypatch, rank=0, world=4)
    k = W.make_kernels()["cavity_kernel"]                     # (function handles belong to a runtime)
    gs0 = [xgrid.Grid((n, cols), float) for _ in range(4)]
    for g, m in zip(gs0, masks):
        g.boundary[...] = m[g.row_range[0]:g.row_range[1]]
    k(*gs0, cfg)
    f0 = [r for r in rt0.launches if "jacobi2" in r[0]]
    assert len(f0) == 24 and all((r[3]["r_lo"], r[3]["r_hi"]) == (0, gs0[1].shape[0] - 3) for r in f0)


def test_an_assigned_boundary_array_stays_the_mask(rt):
    """`g.boundary = arr` (plain attribute assignment in the reference): writes through `arr` afterwards reach the next
    call -- the grid re-compares a program-supplied mask before every call."""
    k = W.make_kernels()["diffusion_1d"]
    u = xgrid.Grid((4096,), float)
    mine = np.zeros(4096, np.int32)
    mine[0] = 1
    u.boundary = mine
    k(u, 0.01, 0.1, 1.0)
    v1 = u._mask_version
    k(u, 0.01, 0.1, 1.0)
    assert u._mask_version == v1
    mine[-1] = 1                                              # through the program's own array
    k(u, 0.01, 0.1, 1.0)
    assert u._mask_version == v1 + 1 and u._mask_snapshot[-1] == 1
    u.boundary = [0] * 4096                                   # a list is converted: the grid owns the new array
    assert not u._boundary_foreign


def test_halo_bookkeeping_does_not_depend_on_which_rank_owns_the_boundary_points(monkeypatch, tmp_path):
    """A boundary statement whose mask value occurs on ONE rank only (the global top row, say) still marks its level
    written on every rank: otherwise the ranks' views of what is stale -- and the exchanges they issue -- drift apart."""
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"), distributed=True, graphs=False)
    f2 = xgrid.grid[float, 2]

    def make():
        @xgrid.kernel()
        def prog(u: f2) -> None:
            u[0, 0] = u[0, 0] + 1.0
            with xgrid.boundary(2):
                u[0, 0] = u[1, 0][0]          # only the rank owning the global top row has value-2 points
            u[0, 0] = 0.5 * (u[1, 0][0] + u[-1, 0][0])      # implicit sweep: reads level 0 across the cuts
        return prog

    logs = []
    for rank in (0, 2):
        rt, tr = fake_runtime.install_sharded(monkeypatch, rank=rank, world=4)
        prog = make()                         # (function handles belong to a runtime)
        g = xgrid.Grid((4 * 32, 64), float)
        if rank == 0:
            g.boundary[0, :] = 2
        for _ in range(3):
            prog(g)
        logs.append([(e[0], e[2]) for e in tr.log if e[0].startswith("exchange")])
    assert logs[0] == logs[1] and len(logs[0]) >= 3, logs


def test_diagonal_taps_make_the_halo_exchange_carry_an_overhang(monkeypatch, tmp_path):
    """Taps are linear addresses (F10): (-1, -1) at column 0 of a slab's first row reads the LAST element of the row
    two further down, which whole-row ghosts of depth 1 do not hold.  The grid learns the overhang from the first
    sweep that needs it, levels exchanged without it count as stale, and the edge-first launches grow by one row
    (the exchange that follows them sends the head of the first interior row too)."""
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"), distributed=True)
    rt, tr = fake_runtime.install_sharded(monkeypatch, rank=1, world=4)
    f2 = xgrid.grid[float, 2]

    @xgrid.kernel()
    def nine(u: f2, c: float) -> None:
        u[0, 0] = c * (u[-1, -1][1] + u[-1, 1][1] + u[1, -1][1] + u[1, 2][1] + u[0, 0][1])

    u = xgrid.Grid((4 * 64, 1024), float)
    u.now[...] = 1.0
    nine(u, 0.2)
    xgrid.flush()                                       # (a lone deferred call runs as an ordinary step)
    assert u._halo_over == 2                            # (1, 2) leaves its row by two elements
    bands = [(r[3]["r_lo"], r[3]["r_hi"]) for r in rt.launches]
    assert bands == [(0, 2), (62, 64), (2, 62)], bands
    assert ((0, 2)) in u._halo_state()                  # part of the key of a recorded call
    # a level that was exchanged before the grid learnt about the overhang is refreshed again
    v = xgrid.Grid((4 * 64, 1024), float)
    v.now[...] = 1.0
    lv = v._ring[0]
    v._prepare_device(1)
    lv.halo_rows = 1
    v._need_halo_over(1)
    assert lv.halo_rows == 0 and v._halo_over == 1
    v._need_halo_over(1)
    with pytest.raises(Exception, match="past the ghost rows"):
        v._need_halo_over(1 << 20)


def test_overhang_of_a_slot_counts_only_the_deepest_rows_and_matching_directions():
    from xgrid_b200.lang.cudagen import Slot
    s = Slot(0, "u", 1, None, read=True, halo0=1, taps={(-1, 0), (0, -1), (1, 0), (0, 1)})
    assert s.overhang((64, 128)) == 0                   # axis-aligned star: whole rows suffice
    s.taps |= {(-1, -1)}
    assert s.overhang((64, 128)) == 1
    s.taps |= {(1, -3)}                                 # moves back INTO the ghost row: no overhang
    assert s.overhang((64, 128)) == 1
    s3 = Slot(0, "u", 1, None, read=True, halo0=2, taps={(-2, -1, 0), (-1, -2, -2), (2, 1, 1)})
    assert s3.overhang((16, 32, 100)) == 101            # |dj * n2 + dk| of the taps at depth 2
    assert Slot(0, "u", 1, None, read=True, halo0=1, taps={(-1,)}).overhang((4096,)) == 0


def test_two_step_passes_on_a_slab_cover_the_interior_and_leave_bands_to_single_sweeps(monkeypatch, tmp_path):
    """Middle rank of 4, 2-D diffusion (taps one row up and down): per pass the two-step kernel covers the rows whose
    two-step cone stays inside the slab, [3, n0 - 3); next to each cut step 1 runs on row bands into the spare buffer
    (two rows deeper), the spare's halo is exchanged, and step 2 runs on [0, 3) and [n0 - 3, n0) into the output
    buffer -- all on the side stream, joined before the next pass."""
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"), distributed=True)
    rt, tr = fake_runtime.install_sharded(monkeypatch, rank=1, world=4)
    k = W.make_kernels()["diffusion_2d"]
    u = xgrid.Grid((4 * 128, 2048), float)
    u.now[...] = 1.0
    u.boundary[:, 0] = u.boundary[:, -1] = 1
    for _ in range(4):
        k(u, 0.2)
    assert rt.launches == []
    xgrid.flush()
    n0 = u.shape[0]
    fused = [r for r in rt.launches if r[0].endswith("tiled2_v2")]
    assert len(fused) == 2 and rt.launches[-1] is fused[1] and all((r[3]["r_lo"], r[3]["r_hi"]) == (3, n0 - 3) for r in fused)
    assert [r[3]["opt0"] for r in fused] == [0, 1]
    first = rt.launches.index(fused[0])
    bands = [(r[3]["r_lo"], r[3]["r_hi"]) for r in rt.launches[:first]]
    assert bands == [(0, 5), (n0 - 5, n0), (0, 3), (n0 - 3, n0)], bands
    x0, x1 = fused[0][3]["aux0"], fused[0][3]["aux1"]
    step1, step2 = rt.launches[0][3], rt.launches[2][3]
    spare = step1["s0"]
    assert step1["s1"] == x0 and spare not in (x0, x1)                 # step 1: u^n -> spare
    assert step2["s1"] == spare and step2["s0"] == x1                  # step 2: spare -> u^{n+2}
    # halo exchanges of the first pass: u^n before step 1, the spare before step 2 -- one each, depth 1
    ex = [(e[1], e[2]) for e in tr.log if e[0] == "exchange" and e[3] <= first]
    assert ex == [(x0, 1), (spare, 1)], tr.log
    # the agreement on "every mask value present has a statement" is one collective per flush
    assert [e for e in tr.log if e[0] == "all_agree"] == [("all_agree", 1, 0, 0)]
    # second pass: roles of the two ring buffers swapped, same spare; the last pass stores the middle level there
    assert fused[1][3]["aux0"] == x1 and fused[1][3]["aux1"] == x0 and fused[1][3]["aux2"] == spare
    assert [lv.dev for lv in u._ring] == [x0, spare]
    # rank 0 of the chain has no lower neighbour: its pass starts at row 0 and only the upper cut has bands
    rt0, tr0 = fake_runtime.install_sharded(monkeypatch, rank=0, world=4)
    k = W.make_kernels()["diffusion_2d"]
    v = xgrid.Grid((4 * 128, 2048), float)
    v.boundary[:, 0] = v.boundary[:, -1] = 1
    for _ in range(2):
        k(v, 0.2)
    xgrid.flush()
    f0 = [r for r in rt0.launches if r[0].endswith("tiled2_v2")]
    assert [(r[3]["r_lo"], r[3]["r_hi"]) for r in f0] == [(0, v.shape[0] - 3)]
    assert [(r[3]["r_lo"], r[3]["r_hi"]) for r in rt0.launches[:2]] == [(v.shape[0] - 5, v.shape[0]), (v.shape[0] - 3, v.shape[0])]


def test_a_sweep_with_taps_along_the_rows_only_still_imports_one_ghost_row(monkeypatch, tmp_path):
    """u[0, -1] at column 0 of a slab's first row is a linear address in the row below -- the neighbour's last row.  A
    full-grid sweep therefore exchanges one row although no tap has an axis-0 offset; a lone boundary statement (a
    sparse group, confined by its mask -- the cavity's wall copies) keeps exchanging nothing."""
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"), distributed=True, temporal=False)
    rt, tr = fake_runtime.install_sharded(monkeypatch, rank=1, world=4)
    f2 = xgrid.grid[float, 2]

    @xgrid.kernel()
    def along_rows(u: f2, c: float) -> None:
        u[0, 0] = c * (u[0, -1][1] + u[0, 2][1])

    @xgrid.kernel()
    def wall_copy(u: f2, p: f2) -> None:
        with xgrid.boundary(3):
            p[0, 0] = u[0, -1]

    u, p = xgrid.Grid((4 * 64, 1024), float), xgrid.Grid((4 * 64, 1024), float)
    u.now[...] = 1.0
    along_rows(u, 0.5)
    src = u._ring[1]
    assert [(e[1], e[2]) for e in tr.log if e[0] == "exchange"] == [(src.dev, 1)], tr.log
    n = len(tr.log)
    w = xgrid.Grid((4 * 64, 1024), float)               # freshly uploaded: its ghost rows are stale
    w.now[...] = 2.0
    p.boundary[:, -1] = 3
    wall_copy(w, p)
    assert [e for e in tr.log[n:] if e[0].startswith("exchange")] == [], tr.log[n:]
