"""Worker for the 2-rank tests (launched by torch.distributed.run).

    mode=cpu : gloo; exercises slab_range / Grid sharding / HaloPlan on the host and checks
               a slab-decomposed oracle run (ghost rows exchanged over gloo) against the
               single-domain oracle.
    mode=gpu : nccl; the sharded CUDA path against the single-domain oracle, bit-exact.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

import oracle
from oracle import HostGrid
import xgrid_b200 as xgrid
from xgrid_b200 import dist as xdist
from examples import workloads as W


def oracle_global(shape, ic, mask, steps, a):
    h = HostGrid(shape)
    h.now[...] = ic
    h.boundary[...] = mask
    for _ in range(steps):
        oracle.step_heat3d(h, a)
    return h


def exchange_host(h: HostGrid, level: int, topo, rows: int = 1):
    """Ghost-row exchange of a HostGrid level over the default (gloo) group: the same
    send/recv pattern NcclTransport.exchange issues (first rows down, last rows up)."""
    arr = h._data[level]
    n0 = arr.shape[0]
    flat = arr.base                      # padded 1-D buffer; ghost rows sit right outside the view
    stride0 = h.size // n0
    start = h._pad
    reqs = []
    # same order as xgb_halo_exchange: sends (lo, hi), then receives (hi, lo) -- with a ring of two both
    # neighbours are the same peer and messages between a pair of ranks match in posting order
    if topo.lo_rank >= 0:
        send = torch.from_numpy(np.ascontiguousarray(arr[:rows]).reshape(-1))
        reqs.append(dist.isend(send, topo.lo_rank))
    if topo.hi_rank >= 0:
        send2 = torch.from_numpy(np.ascontiguousarray(arr[n0 - rows:]).reshape(-1))
        reqs.append(dist.isend(send2, topo.hi_rank))
    if topo.hi_rank >= 0:
        recv2 = torch.empty(rows * stride0, dtype=torch.float64)
        reqs.append(dist.irecv(recv2, topo.hi_rank))
    if topo.lo_rank >= 0:
        recv = torch.empty(rows * stride0, dtype=torch.float64)
        reqs.append(dist.irecv(recv, topo.lo_rank))
    for r in reqs:
        r.wait()
    if topo.lo_rank >= 0:
        flat[start - rows * stride0:start] = recv.numpy()
    if topo.hi_rank >= 0:
        flat[start + h.size:start + h.size + rows * stride0] = recv2.numpy()


def open_diffusion_global(u, a, mode):
    """NumPy restatement of workloads.diffusion_2d_open under overstep="wrap"/"limit" with per-axis
    extents (same operand order as the kernel text, so fp64 results are bit-identical)."""
    if mode == "wrap":
        e, w, s, n = np.roll(u, -1, 1), np.roll(u, 1, 1), np.roll(u, -1, 0), np.roll(u, 1, 0)
    else:
        p = np.pad(u, 1, mode="edge")
        e, w, s, n = p[1:-1, 2:], p[1:-1, :-2], p[2:, 1:-1], p[:-2, 1:-1]
    return u + a * (e + w + s + n - 4.0 * u)


def open_diffusion_slab(h: HostGrid, a, mode, topo):
    """The same step on one slab: axis 1 is wrapped / clamped locally, axis 0 reads the ghost rows where the
    slab has a neighbour (what the generated kernel does with open_lo / open_hi) and clamps otherwise."""
    u = h._data[0]
    n0, n1 = u.shape
    flat, start = u.base, h._pad
    ext = flat[start - n1:start + h.size + n1].reshape(n0 + 2, n1).copy()
    if topo.lo_rank < 0:
        ext[0] = ext[1]              # global lower end (only "limit" has one): clamp
    if topo.hi_rank < 0:
        ext[-1] = ext[-2]
    mid = ext[1:-1]
    if mode == "wrap":
        e, w = np.roll(mid, -1, 1), np.roll(mid, 1, 1)
    else:
        p = np.pad(mid, ((0, 0), (1, 1)), mode="edge")
        e, w = p[:, 2:], p[:, :-2]
    return mid + a * (e + w + ext[2:] + ext[:-2] - 4.0 * mid)


def main():
    mode = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    gshape = (23, 12, 130) if mode == "cpu" else (40, 24, 256)
    steps, a = 5, 0.1
    rng = np.random.default_rng(7)
    ic = rng.random(gshape)
    mask = W.shell_mask(gshape)

    if mode == "cpu":
        dist.init_process_group("gloo")
        xgrid.init(precision="double", distributed=True, cacheroot=os.environ.get("XG_CACHE", ".xgrid"))
        topo = xdist.topology()
        assert (topo.rank, topo.world) == (rank, world)
        lo, hi = xdist.slab_range(gshape[0], rank, world)
        g = xgrid.Grid(gshape, float)                  # product Grid: shards itself, host-only here
        assert g.sharded and g.row_range == (lo, hi) and g.shape == (hi - lo,) + gshape[1:]
        assert np.array_equal(W.shell_mask_slab(gshape, lo, hi), mask[lo:hi])
        # HaloPlan: a stale level of a sharded grid read at axis-0 offset 1 must be exchanged once
        lv = g._ring[0]
        assert xdist.HaloPlan.stale([(g, lv, 1), (g, lv, 1), (g, lv, 0)]) == [(g, lv, 1)]
        lv.halo_rows = 1
        assert xdist.HaloPlan.stale([(g, lv, 1)]) == []
        # slab-decomposed oracle run with gloo ghost exchange == single-domain oracle
        h = HostGrid((hi - lo,) + gshape[1:])
        h.now[...] = ic[lo:hi]
        h.boundary[...] = mask[lo:hi]
        for _ in range(steps):
            exchange_host(h, 0, topo)                  # the level the next step reads (becomes level 1)
            oracle.step_heat3d(h, a)
        ref = oracle_global(gshape, ic, mask, steps, a)
        assert np.array_equal(h.now, ref.now[lo:hi]), "sharded oracle differs from single-domain oracle"
        # overstep="wrap" makes the ranks a ring (rank 0's lower neighbour is the last rank), "limit" keeps
        # the open chain and clamps at the global ends only: slab runs == single-domain restatement
        shape2 = (22, 19)
        ic2 = np.random.default_rng(11).random(shape2)
        for omode in ("wrap", "limit"):
            xgrid.init(precision="double", distributed=True, overstep=omode,
                       cacheroot=os.environ.get("XG_CACHE", ".xgrid"))
            topo2 = xdist.topology()
            if omode == "wrap":
                assert topo2.ring and (topo2.lo_rank, topo2.hi_rank) == ((rank - 1) % world, (rank + 1) % world)
            else:
                assert not topo2.ring and topo2.lo_rank == (rank - 1 if rank else -1)
            lo2, hi2 = xdist.slab_range(shape2[0], rank, world)
            h2 = HostGrid((hi2 - lo2, shape2[1]))
            h2.now[...] = ic2[lo2:hi2]
            want = ic2.copy()
            for _ in range(6):
                exchange_host(h2, 0, topo2)
                h2.now[...] = open_diffusion_slab(h2, 0.2, omode, topo2)
                want = open_diffusion_global(want, 0.2, omode)
            assert np.array_equal(h2.now, want[lo2:hi2]), f"sharded overstep={omode} differs from the single domain"
        dist.barrier()
        if rank == 0:
            print("DIST_CPU_OK")
    else:
        local = int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        xgrid.init(precision="double", distributed=True, device=local,
                   cacheroot=os.environ.get("XG_CACHE", ".xgrid"))
        k = W.make_kernels()
        u = xgrid.Grid(gshape, float)
        lo, hi = u.row_range
        u.now[...] = ic[lo:hi]
        u.boundary[...] = mask[lo:hi]
        for _ in range(steps):
            k["heat_3d"](u, a)
        ref = oracle_global(gshape, ic, mask, steps, a)
        ok = np.array_equal(u.now, ref.now[lo:hi]) and np.array_equal(u._data[1], ref._data[1][lo:hi])
        # 1-D slabs through the deferred multi-step path (64 steps per launch + single-step remainder)
        n1 = 3 * 40000 + 17
        ic1, dx1 = W.ic_1d(n1)
        ic1 = ic1 + 0.01 * np.random.default_rng(3).random(n1)
        m1 = np.zeros(n1, np.int32)
        m1[0] = m1[-1] = 1
        m1[n1 // 2] = 1                      # a Dirichlet point right next to the slab boundary
        m1[n1 // 2 + 3] = 7
        u1 = xgrid.Grid((n1,), float)
        lo1, hi1 = u1.row_range
        u1.now[...] = ic1[lo1:hi1]
        u1.boundary[...] = m1[lo1:hi1]
        h1 = HostGrid((n1,))
        h1.now[...] = ic1
        h1.boundary[...] = m1
        args1 = (0.01, 0.2 * dx1 * dx1 / 0.01, dx1)
        for _ in range(150):
            k["diffusion_1d"](u1, *args1)
            oracle.step_diff1d(h1, *args1)
        ok = ok and np.array_equal(u1.now, h1.now[lo1:hi1]) and np.array_equal(u1._data[1], h1._data[1][lo1:hi1])
        from xgrid_b200.lang.launch import STATS
        ok = ok and STATS.get("multistep", 0) >= 2
        # sharded cavity: implicit Jacobi sweeps + sparse Neumann statements that read level 0
        # across the slab boundary (halo refresh between dependent groups)
        nc = 96
        mb, mp, mu, mv = W.cavity_masks(nc, nc)
        dxc = 2.0 / (nc - 1)
        cfg = W.Config(1.0, 0.1, 1e-4, dxc, dxc)
        gs = [xgrid.Grid((nc, nc), float) for _ in range(4)]
        hs = [HostGrid((nc, nc)) for _ in range(4)]
        loc, hic = gs[0].row_range
        for gg, hh, m in zip(gs, hs, (mb, mp, mu, mv)):
            gg.boundary[...] = m[loc:hic]
            hh.boundary[...] = m
        for _ in range(7):
            k["cavity_kernel"](*gs, cfg)
            oracle.step_cavity(*hs, oracle.Config(cfg.rho, cfg.nu, cfg.dt, cfg.dx, cfg.dy))
        for gg, hh in zip(gs, hs):
            ok = ok and np.array_equal(gg.now, hh.now[loc:hic]) and np.array_equal(gg._data[1], hh._data[1][loc:hic])
        # fused Jacobi pairs on slabs: grid wide enough for the fused pass (>= 512 columns); interior rows fused, the
        # three rows next to the cut on row bands
        n0c, n1c = 128, 640
        mb, mp, mu, mv = W.cavity_masks(n0c, n1c)
        cfgc = W.Config(1.0, 0.1, 1e-4 * (100.0 / (n1c - 1)) ** 2, 2.0 / (n1c - 1), 2.0 / (n0c - 1))
        gs = [xgrid.Grid((n0c, n1c), float) for _ in range(4)]
        hs = [HostGrid((n0c, n1c)) for _ in range(4)]
        loc, hic = gs[0].row_range
        rngc = np.random.default_rng(5)
        for gg, hh, m in zip(gs, hs, (mb, mp, mu, mv)):
            icc = 1e-3 * rngc.random((n0c, n1c))
            gg.now[...] = icc[loc:hic]
            hh.now[...] = icc
            gg.boundary[...] = m[loc:hic]
            hh.boundary[...] = m
        fused0 = STATS.get("jacobi2", 0)
        for _ in range(5):
            k["cavity_kernel"](*gs, cfgc)
            oracle.step_cavity(*hs, oracle.Config(cfgc.rho, cfgc.nu, cfgc.dt, cfgc.dx, cfgc.dy))
        ok = ok and STATS.get("jacobi2", 0) - fused0 == 5 * 24
        for gg, hh in zip(gs, hs):
            ok = ok and np.array_equal(gg.now, hh.now[loc:hic]) and np.array_equal(gg._data[1], hh._data[1][loc:hic])
        # two steps per pass on slabs (deferred runs of a 2-D kernel): interior rows in the two-step kernel, the rows
        # next to a cut step-at-a-time on row bands; uneven slabs (one rank owns one row more)
        n0t, n1t = 80 * world + 1, 512
        t2_0 = STATS.get("tiled2", 0)
        for name, stepf, sargs in (("diffusion_2d", oracle.step_diff2d, (0.2,)),
                                   ("convection_2d", oracle.step_conv2d, (1.0, 0.002, 0.01, 0.01))):
            ict = np.random.default_rng(21).random((n0t, n1t))
            mt = np.zeros((n0t, n1t), np.int32)
            mt[0, :] = mt[-1, :] = mt[:, 0] = mt[:, -1] = 1
            ut, ht = xgrid.Grid((n0t, n1t), float), HostGrid((n0t, n1t))
            lot, hit = ut.row_range
            ut.now[...] = ict[lot:hit]
            ut.boundary[...] = mt[lot:hit]
            ht.now[...] = ict
            ht.boundary[...] = mt
            for _ in range(9):
                k[name](ut, *sargs)
                stepf(ht, *sargs)
            ok = ok and np.array_equal(ut.now, ht.now[lot:hit]) and np.array_equal(ut._data[1], ht._data[1][lot:hit])
        ok = ok and STATS.get("tiled2", 0) - t2_0 == 2 * 4
        # diagonal taps and no boundary statements: at the first / last column a tap leaves its row and -- taps being
        # linear addresses -- reads the row one further out, which the halo exchange carries as an overhang; with and
        # without the several-steps kernels
        f2 = xgrid.grid[float, 2]
        for temporal in (True, False):
            xgrid.init(precision="double", distributed=True, device=local, temporal=temporal,
                       cacheroot=os.environ.get("XG_CACHE", ".xgrid"))

            @xgrid.kernel()
            def nine_point(u: f2, c: float) -> None:
                u[0, 0] = c * (u[-1, -1][1] + u[-1, 1][1] + u[1, -1][1] + u[1, 2][1] + u[0, 0][1])

            n09, n19 = 70 * world + 3, 512
            ic9 = np.random.default_rng(31).random((n09, n19))
            u9 = xgrid.Grid((n09, n19), float)
            lo9, hi9 = u9.row_range
            u9.now[...] = ic9[lo9:hi9]
            want9 = [ic9, ic9]
            pad9 = 2 * n19 + 8
            at9 = np.arange(n09 * n19) + pad9
            for _ in range(7):
                nine_point(u9, 0.2)
                f9 = np.concatenate([np.zeros(pad9), want9[0].ravel(), np.zeros(pad9)])
                t9 = lambda d0, dk: f9[at9 + d0 * n19 + dk]                              # noqa: E731
                new9 = 0.2 * ((((t9(-1, -1) + t9(-1, 1)) + t9(1, -1)) + t9(1, 2)) + t9(0, 0))
                want9 = [new9.reshape(n09, n19), want9[0]]
            ok = ok and np.array_equal(u9.now, want9[0][lo9:hi9]) and np.array_equal(u9._data[1], want9[1][lo9:hi9])
            ok = ok and u9._halo_over == 2
        # overstep modes on slabs: ring ("wrap") / chain with clamping at the global ends ("limit");
        # golden produced by the reference (square 32x32) and a non-square NumPy restatement
        gold_dir = os.path.join(ROOT, "tests", "golden")
        for omode in ("wrap", "limit"):
            xgrid.init(precision="double", distributed=True, device=local, overstep=omode,
                       cacheroot=os.environ.get("XG_CACHE", ".xgrid"))
            ko = W.make_kernels()["diffusion_2d_open"]
            gd = np.load(os.path.join(gold_dir, f"diff2d_{omode}_f64.npz"))
            uo = xgrid.Grid(gd["u_in"].shape, float)
            loo, hio = uo.row_range
            uo.now[...] = gd["u_in"][loo:hio]
            for _ in range(int(gd["steps"])):
                ko(uo, float(gd["params"][0]))
            ok = ok and np.array_equal(uo.now, gd["u.L0"][loo:hio]) and np.array_equal(uo._data[1], gd["u.L1"][loo:hio])
            shape2 = (46, 70)
            ic2 = np.random.default_rng(11).random(shape2)
            un = xgrid.Grid(shape2, float)
            lon, hin = un.row_range
            un.now[...] = ic2[lon:hin]
            want = ic2.copy()
            for _ in range(7):
                ko(un, 0.2)
                want = open_diffusion_global(want, 0.2, omode)
            ok = ok and np.array_equal(un.now, want[lon:hin])
        xdist.quiesce()
        t = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(("DIST_GPU_OK" if int(t.item()) == 1 else "DIST_GPU_MISMATCH")
                  + f" transport={type(xdist.transport()).__name__}")
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
