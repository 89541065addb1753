"""Launch side of a kernel call: marshal one by-value parameter struct per
sweep group, pick the kernel variant and launch geometry, and enqueue on the
backend's compute stream.  This is the counterpart of the reference's per-call
``serialize`` + ctypes foreign call (xgrid/util/ffi.py:30-34,
xgrid/xgrid/__init__.py:60-68) -- asynchronous, and with the descriptors
cached instead of rebuilt every call.
"""
from __future__ import annotations

import os
from dataclasses import astuple

import numpy as np

from ..types import Pointer, Structure
from . import cudagen, jacobi2

SPARSE_FRACTION = 16      # run a mask!=0 statement over its index list when it
                          # covers less than 1/16 of the grid


def _pow2ceil(x: int) -> int:
    p = 1
    while p < x:
        p <<= 1
    return p


def dense_geometry(rows: int, cols: int, V: int):
    """blockDim=(TX,TY) threads, each owning V columns of one row; grid.x walks
    the columns fastest so consecutive CTAs share rows in L2."""
    vec_cols = (cols + V - 1) // V
    tx = min(256, max(32, _pow2ceil(vec_cols)))
    ty = max(1, min(256 // tx, _pow2ceil(rows)))
    gx = (vec_cols + tx - 1) // tx
    rb = (rows + ty - 1) // ty
    gy = min(rb, 65535)
    gz = (rb + gy - 1) // gy
    return (gx, gy, gz), (tx, ty, 1)


# launch-time tunables (overridable from the environment for on-GPU experiments)
TUNE = {
    "vec_bytes": int(os.environ.get("XGB_VEC_BYTES", "32")),
    "march": os.environ.get("XGB_MARCH", "1") != "0",
    "tiled": os.environ.get("XGB_TILED", "1") != "0",
    "overlap": os.environ.get("XGB_OVERLAP", "1") != "0",
    "tx": int(os.environ.get("XGB_TX", "0")),
    "ty": int(os.environ.get("XGB_TY", "0")),
    "chunk0": int(os.environ.get("XGB_CHUNK0", "0")),
    "min_ctas": int(os.environ.get("XGB_MIN_CTAS", "8192")),
}


def march_geometry(shape, V: int, R: int):
    """2-D: blockDim=(TX,1), grid=(col blocks, axis-0 chunks).  3-D: blockDim=(TX,TY)
    over (k, j) so that j+-1 taps hit L1, grid.z walks the axis-0 chunks.  A CTA
    marches `chunk0` points along axis 0; its two halo planes are the only re-read
    traffic, so chunks are made as long as possible while leaving >= MIN_CTAS CTAs."""
    cols = shape[-1]
    vec_cols = (cols + V - 1) // V
    if len(shape) == 2:
        tx = TUNE["tx"] or min(256, max(32, _pow2ceil(vec_cols)))
        ty, gy_mid = 1, 1
    else:
        tx = TUNE["tx"] or min(128, max(32, _pow2ceil(vec_cols)))
        ty = TUNE["ty"] or max(1, min(256 // tx, _pow2ceil(shape[1])))
        gy_mid = (shape[1] + ty - 1) // ty
    gx = (vec_cols + tx - 1) // tx
    want_chunks = max(1, -(-TUNE["min_ctas"] // (gx * gy_mid)))
    chunk0 = max(R, -(-shape[0] // want_chunks))
    chunk0 = -(-chunk0 // R) * R
    if TUNE["chunk0"]:
        chunk0 = TUNE["chunk0"]
    chunks = (shape[0] + chunk0 - 1) // chunk0
    if len(shape) == 2:
        gy = min(chunks, 65535)
        gz = (chunks + gy - 1) // gy
        return (gx, gy, gz), (tx, 1, 1), chunk0
    return (gx, gy_mid, chunks), (tx, ty, 1), chunk0


def tiled_geometry(shape, t: dict):
    """grid = (column tiles, j tiles | axis-0 chunks, axis-0 chunks); one CTA streams
    `chunk0` planes through its shared-memory ring."""
    gx = (shape[-1] + t["W"] - 1) // t["W"]
    gy_mid = 1 if len(shape) == 2 else (shape[1] + t["TJ"] - 1) // t["TJ"]
    # 3-D: twice as many, half as long CTAs (32 instead of 64 planes each on the 256 x 2048^2 slab) measured +1 %
    # in three sessions (profiles/r2b / r2c / r2g): better balance at the end of the launch outweighs the pipeline fills
    want_chunks = max(1, -(-(TUNE["min_ctas"] * (2 if len(shape) == 3 else 1)) // (gx * gy_mid)))
    chunk0 = TUNE["chunk0"] or max(16, -(-shape[0] // want_chunks))
    chunks = (shape[0] + chunk0 - 1) // chunk0
    block = (t["threads"], 1, 1)
    if len(shape) == 2:
        gy = min(chunks, 65535)
        return (gx, gy, (chunks + gy - 1) // gy), block, chunk0
    return (gx, gy_mid, chunks), block, chunk0


STATS: dict = {}          # kernel variant -> launches (diagnostics / tests)
SLAB_BAND = jacobi2.HA + jacobi2.DR + jacobi2.HA      # rows next to a slab cut whose fused-pair value needs the neighbour


def full_grid_variant(g: cudagen.Group, shape) -> tuple:
    """(variant, vector width, dynamic shared memory) of a one-pass sweep of group `g` over the whole grid."""
    cols = shape[-1]
    variant, V, smem = cudagen.VARIANT_DENSE, 1, 0
    vmax = max(1, TUNE["vec_bytes"] // max(s.elem.width_bytes if hasattr(s.elem, "width_bytes") else 16
                                           for s in g.slots))
    for cand in sorted(g.vwidths, reverse=True):
        if cols % cand == 0 and cand <= vmax:
            V = cand
            break
    t = g.tiled
    if (t is not None and TUNE["tiled"] and cols % t["V"] == 0 and cols >= t["W"] and shape[0] >= 16
            and (g.ndim == 2 or shape[1] >= t["TJ"])):
        variant, V = cudagen.VARIANT_TILED, t["V"]
        smem = t["smem"]
    elif g.march and V > 1 and TUNE["march"] and shape[0] >= 4:
        variant = cudagen.VARIANT_MARCH
    return variant, V, smem


class Launcher:
    def __init__(self, program, grids: dict) -> None:
        self.program = program
        self.grids = grids
        self._rt = None
        self.launches = 0
        self.stream = 0             # where sweeps are enqueued (the slab path of the fused pairs uses a side stream)

    @property
    def rt(self):
        if self._rt is None:
            from ..runtime.shim import Runtime
            self._rt = Runtime.get()
        return self._rt

    def _scalar_value(self, t, raw):
        if isinstance(t, Pointer):
            raw = raw.contents.value if hasattr(raw, "contents") else raw.value
            t = t.element
        if isinstance(t, Structure):
            return t.ctype(*astuple(raw))
        return raw.item() if isinstance(raw, np.generic) else raw

    # ---- two solver iterations per pass (lang/jacobi2.py)
    def pair_ok(self, pair) -> bool:
        """Dynamic eligibility of the fused sweep-boundary-sweep pass for this call's grids (cached per mask
        version).  On slabs the fused path issues other exchanges than the step-at-a-time path, so the ranks
        decide TOGETHER: every rank evaluates its local conditions and joins exactly one host all-reduce per
        (pair, mask version) -- whatever its local answer, so that no rank waits in the collective alone."""
        lead = self.grids[pair.sweep.lead]
        key = (id(pair), lead._mask_version)
        ok = lead._pair_ok.get(key)
        if ok is None:
            if len(lead._pair_ok) > 16:
                lead._pair_ok.clear()
            ok = self._pair_ok_local(pair)
            if lead.sharded:
                from .. import dist
                ok = dist.transport().all_agree(ok)
            lead._pair_ok[key] = ok
        return ok

    def _pair_ok_local(self, pair) -> bool:
        g = pair.sweep
        lead = self.grids[g.lead]
        c = pair.config
        if lead.dimension != 2:
            return False
        n0, cols = lead.shape
        if cols % c["V"] or cols < jacobi2.MIN_COLS or n0 < jacobi2.MIN_ROWS:
            return False
        if lead.sharded:
            # a slab keeps SLAB_BAND rows per cut for the step-at-a-time path, run on row bands by a variant
            # that honours them
            if n0 < 2 * (SLAB_BAND + 2) + jacobi2.MIN_ROWS:
                return False
            if full_grid_variant(g, lead.shape)[0] not in (cudagen.VARIANT_TILED, cudagen.VARIANT_MARCH):
                return False
        if any(self.grids[s.grid].shape != lead.shape for s in g.slots):
            return False
        if not lead._mask_any:
            return True
        return bool(jacobi2.chains_fit(pair, lead._mask_snapshot))

    def run_pair(self, pair, env: dict) -> None:
        """Sweep A, boundary statements, sweep B in one pass; then the boundary statements after B.
        On a slab the fused pass covers the rows that need nothing from a neighbour: output rows
        [SLAB_BAND, n0 - SLAB_BAND) read input rows of this slab only (sweep A reaches HA rows, a boundary
        copy chain DR more, sweep B HA again).  The SLAB_BAND rows next to each cut run step-at-a-time on
        row bands with the ordinary kernels -- sweep A into a third buffer T (two rows deeper than the band,
        so that its boundary statements and sweep B find their taps), the boundary statements on T (halo
        refreshed by the usual planner), sweep B from T into the output buffer -- with the buffers' roles
        re-pointed, never copied.  Every value is produced by the same expression text either way."""
        g, c = pair.sweep, pair.config
        lead = self.grids[g.lead]
        n0, cols = lead.shape
        r_lo, r_hi = 0, n0
        if lead.sharded:
            from .. import dist
            topo = dist.topology()
            lo_band = SLAB_BAND if topo.lo_rank >= 0 else 0
            hi_band = SLAB_BAND if topo.hi_rank >= 0 else 0
            r_lo, r_hi = lo_band, n0 - hi_band
            if lo_band or hi_band:
                X, S = lead._ring[0], lead._scratch_level()
                T = lead._spare_levels(1)[0]
                deep = [(0, lo_band + 2)] * bool(lo_band) + [(n0 - hi_band - 2, n0)] * bool(hi_band)
                edge = [(0, lo_band)] * bool(lo_band) + [(n0 - hi_band, n0)] * bool(hi_band)
                # the band chain (a dozen small launches and two halo exchanges) runs on a high-priority side
                # stream BESIDE the fused pass of the interior: it reads X like the fused pass and writes other rows
                side, fork, join = self.rt.side_stream()
                self.rt.event_record_raw(fork, 0)
                self.rt.stream_wait_event(side, fork)
                try:
                    self.stream = side
                    lead._scratch = T
                    self(g, env, bands=deep)                 # sweep A on the bands: X -> T
                    lead._ring[0], lead._scratch = T, S
                    for r in pair.rules:                     # boundary statements A, on T
                        self(r.group, env)
                    self(g, env, bands=edge)                 # sweep B on the bands: T -> S
                finally:
                    self.stream = 0
                    lead._ring[0], lead._scratch = X, S
                self.rt.event_record_raw(join, side)
        P = self._params(g, env)
        gx = (cols + c["W"] - 1) // c["W"]
        want = max(1, -(-TUNE["min_ctas"] // gx))
        chunk0 = jacobi2.CHUNK0 or max(32, -(-(r_hi - r_lo) // want))     # measured: 32 beats 64 / 128 (tail effect)
        chunks = (r_hi - r_lo + chunk0 - 1) // chunk0
        P.chunk0, P.r_lo, P.r_hi = chunk0, r_lo, r_hi
        gy = min(chunks, 65535)
        fn = self.program.function(cudagen.kernel_name(g, jacobi2.VARIANT, c["V"]), c["smem"])
        self.rt.launch(fn, (gx, gy, (chunks + gy - 1) // gy), (c["threads"], 1, 1), P, smem=c["smem"])
        self.launches += 1
        STATS[jacobi2.VARIANT] = STATS.get(jacobi2.VARIANT, 0) + 1
        if lead.sharded and r_lo + (n0 - r_hi):
            self.rt.stream_wait_event(0, join)           # the bands of the output buffer are complete
        self._mark_written(g)
        lead._swap_scratch()
        for r in pair.rules:                 # boundary statements of the second iteration
            self(r.group, env)

    # ---- `with xgrid.c()` (lang/inlinec.py): the text runs as one device thread, in stream order
    def inline(self, stmt, env: dict) -> None:
        ik = self.program.inlines[id(stmt)]
        P = ik.params_cls()
        depth = self.program.depth
        for name in ik.grids:
            grid = self.grids[name]
            levels = getattr(P, f"d_{name}")
            for l in range(depth):
                levels[l] = grid._ring[l].dev if l < len(grid._ring) else None
            shape = getattr(P, f"shape_{name}")
            for a, n in enumerate(grid.shape):
                shape[a] = n
            setattr(P, f"m_{name}", grid._mask_dev if grid._mask_any else None)
            for lv in grid._ring:          # the text may write any level
                lv.halo_rows = 0
                lv.halo_event = 0
        for name, t in ik.scalars:
            if name in env and env[name] is not None:
                setattr(P, f"u_{name}", self._scalar_value(t, env[name]))
        fn = self.program.function(ik.name)
        self.rt.launch(fn, (1, 1, 1), (1, 1, 1), P)
        self.launches += 1
        STATS["inline"] = STATS.get("inline", 0) + 1

    def _params(self, g: cudagen.Group, env: dict):
        """Parameter struct of a group: level pointers, masks, extents, user scalars."""
        lead = self.grids[g.lead]
        P = g.params_cls()
        for s in g.slots:
            grid = self.grids[s.grid]
            if grid.shape != lead.shape:
                raise Exception(f"grid '{s.grid}' has shape {grid.shape} but the sweep over '{g.lead}' "
                                f"runs on {lead.shape}; all grids of one stencil statement must agree")
            lv = grid._scratch_level() if s.level == "scratch" else grid._ring[s.level]
            setattr(P, s.field, lv.dev)
        if lead.sharded:
            if self.program.config.overstep != "none":
                # clamp / wrap along axis 0 happen at the ends of the GLOBAL grid only: where the slab has a
                # neighbour (chain for "limit", ring for "wrap") the kernel reads its ghost rows instead
                from .. import dist
                topo = dist.topology()
                P.open_lo, P.open_hi = int(topo.lo_rank >= 0), int(topo.hi_rank >= 0)
            self._refresh_halos(g)
        for m in g.masks:
            grid = self.grids[m]
            setattr(P, f"m_{m}", grid._mask_dev if grid._mask_any else None)
            setattr(P, f"f_{m}", grid._flags_dev if grid._mask_any else None)
        shape = lead.shape
        cols = shape[-1]
        rows = lead.size // cols if cols else 0
        P.rows, P.cols = rows, cols
        for a, n in enumerate(shape):
            setattr(P, f"n{a}", n)
        for name in g.shapes:
            arr = getattr(P, f"shape_{name}")
            for a, n in enumerate(self.grids[name].shape):
                arr[a] = n
        for name, t in g.scalars.items():
            setattr(P, f"u_{name}", self._scalar_value(t, env[name]))
        return P

    def __call__(self, g: cudagen.Group, env: dict, bands=None) -> None:
        """One sweep group.  `bands`: only these axis-0 row ranges, no edge-first overlap and no scratch swap
        (the slab path of the fused solver pairs drives the buffers itself)."""
        lead = self.grids[g.lead]
        P = self._params(g, env)
        shape = lead.shape
        cols = shape[-1]
        rows = lead.size // cols if cols else 0
        if lead.size == 0:
            return
        variant, V, smem = cudagen.VARIANT_DENSE, 1, 0
        if g.sparse:
            k = g.stmts[0].sweep.mask
            count = lead._mask_count(k)
            if count == 0:
                if lead.sharded:
                    # another rank may own points of this value: the level counts as written HERE too, so that the
                    # halo bookkeeping -- and with it the sequence of exchanges -- stays the same on every rank
                    self._mark_written(g)
                return
            if count * SPARSE_FRACTION <= lead.size:
                variant = cudagen.VARIANT_SPARSE
                ptr, count = lead._index_list(k)
                P.list, P.count = ptr, count
        P.r_lo, P.r_hi = 0, shape[0]
        if variant == cudagen.VARIANT_DENSE:
            variant, V, smem = full_grid_variant(g, shape)
        fn = self.program.function(cudagen.kernel_name(g, variant, V), smem)

        def launch_rows(lo: int, hi: int) -> None:
            """One launch of the chosen variant over axis-0 rows [lo, hi)."""
            if hi <= lo:
                return
            sub = (hi - lo,) + tuple(shape[1:])
            P.r_lo, P.r_hi = lo, hi
            if variant == cudagen.VARIANT_TILED:
                grid_dim, block_dim, P.chunk0 = tiled_geometry(sub, g.tiled)
            elif variant == cudagen.VARIANT_MARCH:
                grid_dim, block_dim, P.chunk0 = march_geometry(sub, V, cudagen.MARCH_ROWS[g.ndim])
            elif variant == cudagen.VARIANT_SPARSE:
                block_dim, grid_dim = (128, 1, 1), ((P.count + 127) // 128, 1, 1)
            else:
                grid_dim, block_dim = dense_geometry(rows, cols, V)
            self.rt.launch(fn, grid_dim, block_dim, P, smem=smem, stream=self.stream)
            self.launches += 1
            STATS[variant] = STATS.get(variant, 0) + 1

        written = [(self.grids[s.grid], s) for s in g.slots if s.written]
        for gr, s in written:
            # a halo exchange of the level about to be overwritten may still be reading its edge rows on
            # the comm stream (nothing read its ghost rows in between): order this sweep behind it
            lv = gr._scratch if s.level == "scratch" else gr._ring[s.level]
            if gr.sharded and lv is not None and lv.halo_event:
                self.rt.stream_wait_event(self.stream, lv.halo_event)
                lv.halo_event = 0
        # rows to ship after the edges are written: what this program's sweeps read across a cut (the ghost band may
        # be wider than that -- widened for another program, or for the several-steps kernels)
        edge = max((min(gr._ghost, self.program.halo0()) for gr, _ in written if gr.sharded), default=0)
        if bands is not None:
            if variant not in (cudagen.VARIANT_TILED, cudagen.VARIANT_MARCH):
                raise Exception(f"internal: row bands need a row-range variant, '{variant}' sweeps the whole grid")
            for lo, hi in bands:
                launch_rows(lo, hi)
            self._mark_written(g)
            return
        if (edge and TUNE["overlap"] and variant in (cudagen.VARIANT_TILED, cudagen.VARIANT_MARCH)
                and shape[0] >= 8 * edge):
            # slab edges first, then ship the fresh rows to the neighbours on the comm stream
            # while the interior of the slab is still being swept
            from .. import dist
            # (an exchange that also carries the overhang of diagonal taps sends the head of one more row)
            first = edge + (1 if any(gr._halo_over for gr, _ in written if gr.sharded) else 0)
            launch_rows(0, first)
            launch_rows(shape[0] - first, shape[0])
            self._mark_written(g)
            items = []
            for gr, s in written:
                if gr.sharded:
                    lv = gr._scratch if s.level == "scratch" else gr._ring[s.level]
                    items.append((gr, lv, edge))
            dist.transport().exchange_async(items)
            launch_rows(first, shape[0] - first)
        else:
            launch_rows(0, shape[0])
            self._mark_written(g)
        if g.implicit:
            self.grids[g.stmts[0].sweep.grid.name]._swap_scratch()

    def _mark_written(self, g: cudagen.Group) -> None:
        for s in g.slots:
            if s.written:
                grid = self.grids[s.grid]
                lv = grid._scratch if s.level == "scratch" else grid._ring[s.level]
                lv.halo_rows = 0
                lv.halo_event = 0

    def _refresh_halos(self, g: cudagen.Group) -> None:
        """Sharded grids: before this group reads a level at a non-zero axis-0 offset its ghost
        rows must hold the neighbours' rows: wait for an exchange already in flight (started
        right after the level's edges were written), or exchange now if the level is stale."""
        from .. import dist
        reads = []
        for s in g.slots:
            if s.read and s.halo0 > 0:
                grid = self.grids[s.grid]
                lv = grid._scratch_level() if s.level == "scratch" else grid._ring[s.level]
                if grid.sharded:
                    grid._need_halo_over(s.overhang(grid.shape))
                    if lv.halo_event:
                        self.rt.stream_wait_event(self.stream, lv.halo_event)
                        lv.halo_event = 0
                reads.append((grid, lv, s.halo0))
        stale = dist.HaloPlan.stale(reads)
        if stale:
            dist.transport().exchange(stale, self.stream)
