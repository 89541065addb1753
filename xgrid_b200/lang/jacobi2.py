"""Two Jacobi sweeps per pass for iterative solver loops.

The reference runs ``for _ in range(n): <implicit statement>; <boundary statements>``
(examples/cavity.py:99-113) as n x (two full-grid traversals + a malloc'ed scratch array +
one full-grid mask scan per boundary statement), xgrid/lang/generator.py:312-352,289-304.
The step-at-a-time B200 path already turns that into one HBM pass (24 B/pt for the cavity's
pressure sweep) plus a few index-list launches per iteration.  This module fuses TWO consecutive
iterations into ONE pass over HBM:

    sweep A (mask-0 points)  ->  boundary statements A  ->  sweep B  [-> boundary statements B]

* input rows of the iterated grid P and of every other grid the statement reads at its centre
  stream once through the bulk-copy pipeline (cp.async.bulk + mbarrier);
* when row r lands the consumer warps compute the *middle* row q = r - 1 (state after sweep A)
  into a small shared-memory ring, meet at a named barrier, and compute the output row
  o = q - 3 (state after sweep B) from that ring;
* the boundary statements between the two sweeps are never executed as sweeps: one row behind
  sweep A every boundary point of the middle row is *resolved* in the ring -- it takes the
  statement's constant, or the middle value of the point at the end of its copy chain (a
  statement that copies from a point an earlier statement wrote sees that statement's result,
  exactly as in the reference's statement-at-a-time order);
* the boundary statements after sweep B run as ordinary index-list launches on the result.

The intermediate state is never stored.  Every shared-memory position is identified with its
LINEAR grid index (like cudagen._emit_tiled2), so taps that leave a row read exactly what the
step-at-a-time sweeps read.  Arithmetic per point is the same expression text
as the step-at-a-time kernels: results are bit-identical (tests/test_jacobi2_gpu.py).

Eligibility is decided in three places: statically on the statements (`match`), per launch on
the grid (shape; on a slab the fused pass covers the interior rows and the rows next to a cut run on row bands,
lang/launch.py::run_pair), and once per boundary-mask version on the host (`chains_fit`:
every copy chain must stay within one row and one column of its start and must end at a point
that no boundary statement writes -- then the statements' program order cannot matter).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

from ..types import Boolean, Structure
from . import ir
from .cudagen import ExprEmitter, Group, ModuleBuilder, _ident, hoist_lines, kernel_name

VARIANT = "jacobi2"
ENABLED = os.environ.get("XGB_JACOBI2", "1") != "0"
VEC = 2                                             # points per thread per row (one 16-byte vector)
WARPS = int(os.environ.get("XGB_J2_NCW", "3"))      # consumer warps per CTA (measured on cavity 8192^2: 2: 16.5 ms,
                                                    # 3: 14.6, 4: 15.0, 6: 15.4, 8: 15.8 -- small barrier groups win)
TILE_W = int(os.environ.get("XGB_J2_W", "0"))       # output columns per CTA; 0 = middle row == one vector per thread
STAGES = int(os.environ.get("XGB_J2_NS", "5"))
MIN_COLS = int(os.environ.get("XGB_J2_MIN_COLS", "512"))   # narrower grids stay step-at-a-time
MIN_ROWS = 16
CHUNK0 = int(os.environ.get("XGB_J2_CHUNK", "0"))      # output rows per CTA (0 = from the CTA target)
SPARSE_LIMIT = 4           # boundary points must be < 1/4 of the grid

HA = 1                     # stencil halo of the iterated grid (rows and columns)
DR = 1                     # rows / columns a boundary copy chain may move away from its start
QS = HA + DR               # the first middle row a CTA needs is i0 - QS
LAG2 = HA + DR + 1         # output row o = middle row q - LAG2 (rows o-1..o+1 finished one barrier earlier)
KEEP = max(2 * HA, LAG2)   # input stages stay live this many rows


@dataclass
class BoundaryRule:
    """One boundary statement of the loop body, in program order."""
    mask: int
    group: Group
    const: object = None            # ir expression (no grid loads) or None
    offset: tuple = (0, 0)          # copy source, relative


@dataclass
class Pair:
    """A `for` body = [implicit sweep group, boundary groups...] that can run two iterations per pass."""
    sweep: Group
    rules: list = field(default_factory=list)
    config: dict | None = None
    extras: list = field(default_factory=list)     # read-only slots other than the iterated grid


def match(body_plan: list, group_of) -> Pair | None:
    """`body_plan` is the plan of a `for` body; `group_of(node)` returns the cudagen.Group of a
    GroupNode or None.  Static conditions only."""
    if not ENABLED or len(body_plan) < 1:
        return None
    groups = [group_of(n) for n in body_plan]
    if any(g is None for g in groups):
        return None
    g = groups[0]
    if not g.implicit or g.ndim != 2 or len(g.stmts) != 1 or g.stmts[0].sweep.mask != 0 or g.shapes:
        return None
    a = g.stmts[0]
    pname = a.sweep.grid.name
    elem = a.sweep.grid.type.element
    if isinstance(elem, (Structure, Boolean)) or elem.width_bytes != 8:
        return None
    own = [ld for ld in a.sweep.loads if ld.variable.name == pname]
    if not own or any(ld.level != 0 or max(abs(ld.space_offset[0]), abs(ld.space_offset[1])) > HA for ld in own):
        return None
    extras = []
    for ld in a.sweep.loads:
        if ld.variable.name == pname:
            continue
        if any(ld.space_offset) or ld.variable.type.element != elem:
            return None
        key = (ld.variable.name, ld.level)
        if key not in extras:
            extras.append(key)
    if len(extras) > 2:
        return None
    pair = Pair(g, extras=extras)
    seen = {0}
    for bg in groups[1:]:
        if bg.implicit or len(bg.stmts) != 1 or bg.shapes:
            return None
        s = bg.stmts[0]
        if s.sweep.grid.name != pname or s.sweep.mask in seen or s.sweep.store.level != 0:
            return None
        seen.add(s.sweep.mask)
        if not s.sweep.loads:
            rule = BoundaryRule(s.sweep.mask, bg, const=s.value)
        elif (isinstance(s.value, ir.Stencil) and s.value.variable.name == pname and s.value.level == 0
              and max(abs(s.value.space_offset[0]), abs(s.value.space_offset[1])) <= 1):
            rule = BoundaryRule(s.sweep.mask, bg, offset=tuple(s.value.space_offset))
        else:
            return None
        pair.rules.append(rule)
    # constants of the boundary statements are evaluated inside the fused kernel: their scalars
    # travel in the sweep group's parameter struct
    for r in pair.rules:
        if r.const is not None:
            for e in ir.walk_expr(r.const):
                if isinstance(e, ir.Identifier):
                    g.scalars.setdefault(e.variable.name, e.variable.type)
    return pair


def configure(pair: Pair) -> dict:
    V, NCW = VEC, WARPS
    hkm = -(-(HA + DR) // V) * V
    hk0 = -(-(hkm + HA) // V) * V
    # default tile: the middle row (W + 2*HKM columns) is exactly one V-vector per consumer thread
    W = TILE_W or (NCW * 32 * V - 2 * hkm)
    assert W % V == 0 and W > 0
    wp0, wpm = W + 2 * hk0, W + 2 * hkm
    nx = len(pair.extras)
    stg = wp0 + nx * wpm
    ns = max(KEEP + 2, STAGES)
    mr = LAG2 + HA + 2            # rows o-1 .. q, plus the one being written
    nb = -(-wpm // 64)
    wpmb = -(-wpm // 16) * 16
    smem = 256 + (ns * stg + mr * wpm) * 8 + mr * wpmb + ((mr * nb + 15) // 16) * 16
    pair.config = {"V": V, "NCW": NCW, "W": W, "HKM": hkm, "HK0": hk0, "WP0": wp0, "WPM": wpm, "WPMB": wpmb, "NX": nx,
                   "STG": stg, "NS": ns, "MR": mr, "NB": nb, "smem": smem, "threads": (NCW + 1) * 32}
    return pair.config


def emit(pair: Pair, module: ModuleBuilder) -> str:
    """CUDA C of the fused kernel (+ its boundary-resolution helper)."""
    g, c = pair.sweep, pair.config
    a = g.stmts[0]
    pname = a.sweep.grid.name
    T = module.ctype(a.sweep.grid.type.element)
    V = c["V"]
    name = kernel_name(g, VARIANT, V)
    pslot = g.slot(pname, 0)
    oslot = g.slot(pname, "scratch")
    xslots = [g.slot(gn, lv) for gn, lv in pair.extras]
    xindex = {(s.grid, s.level): n for n, s in enumerate(xslots)}

    # row windows of the iterated grid per axis-0 offset: [lo, hi] over the contiguous-axis offsets
    win: dict = {0: (0, 0)}
    for ld in a.sweep.loads:
        if ld.variable.name == pname:
            lo, hi = win.get(ld.space_offset[0], (0, 0))
            win[ld.space_offset[0]] = (min(lo, ld.space_offset[1]), max(hi, ld.space_offset[1]))

    def tap(e: ir.Stencil) -> str:
        if e.variable.name == pname:
            lo, _ = win[e.space_offset[0]]
            return f"w{e.space_offset[0] + HA}[v + {e.space_offset[1] - lo}]"
        return f"x{xindex[(e.variable.name, e.level)]}[v]"

    hoist: dict = {}
    rhs = ExprEmitter(module, _ident, tap, hoist, VARIANT)(a.value)
    lo0, _ = win[0]
    centre = f"w{HA}[v + {-lo0}]"

    # ---- boundary resolution: the value a boundary point takes from the boundary statements that
    #      follow sweep A.  Walks the copy chain through the boundary values kept in shared memory and
    #      returns a constant or the middle-ring value of the chain's end.  `chains_fit` (host)
    #      guarantees that every chain ends at a point no boundary statement writes and stays within
    #      DR rows / columns, so the in-place update of the ring is race-free and order-independent.
    cemit = ExprEmitter(module, _ident, None, None)
    R = [f"__device__ __noinline__ {T} {name}_res(const {T} *mids, const uint8_t *mm, int msx, int pos, "
         f"const {g.name}_P &p)", "{",
         f"    constexpr int WPM = {c['WPM']}, WPMB = {c['WPMB']}, MR = {c['MR']}, REACH = {DR};",
         f"    int limit = {len(pair.rules)}, rr = 0;",
         "    int slot, q;",
         "#pragma unroll 1",
         "    for (;;) {",
         "        const int r2 = rr < -REACH ? -REACH : (rr > REACH ? REACH : rr);",
         "        q = pos < 0 ? 0 : (pos >= WPM ? WPM - 1 : pos);",
         "        slot = msx + r2;",
         "        if (slot < 0) slot += MR;",
         "        if (slot >= MR) slot -= MR;",
         "        const int m = mm[slot * WPMB + q];",
         "        int j = -1, dr = 0, dc = 0;",
         "        switch (m) {"]
    for j, r in enumerate(pair.rules):
        R.append(f"        case {r.mask}: j = {j}; dr = {r.offset[0]}; dc = {r.offset[1]}; break;")
    R += ["        default: break;", "        }",
          "        if (j < 0 || j >= limit) break;"]
    for j, r in enumerate(pair.rules):
        if r.const is not None:
            R.append(f"        if (j == {j}) return ({T})({cemit(r.const)});")
    R += ["        rr += dr; pos += dc; limit = j;",
          "    }",
          "    return mids[slot * WPM + q];",
          "}", ""]

    def windows(prefix: str, ind: str) -> list:
        return [f"{ind}T w{d0 + HA}[V + {hi - lo}]; xgb::lds_window<T, V, {lo}, {hi}>({prefix}{d0 + HA} + kk, w{d0 + HA});"
                for d0, (lo, hi) in win.items()]

    def xloads(prefix: str, ind: str) -> list:
        return [f"{ind}T x{n}[V]; xgb::ld_vec<T, V>({prefix}{n} + kk, x{n});" for n in range(len(xslots))]

    nit1 = -(-c["WPM"] // (c["NCW"] * 32 * V))
    nit2 = -(-c["W"] // (c["NCW"] * 32 * V))
    L = [f'extern "C" __global__ void __launch_bounds__({c["threads"]}) {name}(const __grid_constant__ {g.name}_P p)', "{"]
    L.append(f"    constexpr int V = {V}, NCW = {c['NCW']}, NT = NCW * 32, W = {c['W']}, HKM = {c['HKM']}, HK0 = {c['HK0']}, "
             f"WP0 = {c['WP0']}, WPM = {c['WPM']}, NX = {c['NX']}, STG = {c['STG']}, NS = {c['NS']}, MR = {c['MR']}, NB = {c['NB']}, "
             f"WPMB = {c['WPMB']}, HA = {HA}, QS = {QS}, LAG2 = {LAG2}, KEEP = {KEEP}, NIT1 = {nit1}, NIT2 = {nit2};")
    L.append(f"    typedef {T} T;")
    L.append("    typedef typename xgb::Pack<V>::type MB;                    // the V boundary values of one vector, one byte each")
    L.append("    extern __shared__ __align__(128) unsigned char xgb_smem[];")
    L.append("    uint64_t *full = reinterpret_cast<uint64_t *>(xgb_smem);")
    L.append("    uint64_t *empty = full + NS;")
    L.append("    T *stages = reinterpret_cast<T *>(xgb_smem + 256);        // [NS][STG]: input row r | centre rows r-HA of the other grids")
    L.append("    T *mids = stages + NS * STG;                              // [MR][WPM]: middle rows (state after sweep A + boundary statements)")
    L.append("    uint8_t *mm = reinterpret_cast<uint8_t *>(mids + MR * WPM);   // [MR][WPMB]: boundary values of the middle rows")
    L.append("    uint8_t *mnz = mm + MR * WPMB;                            // [MR][NB]: 64-column block holds a boundary point")
    L.append(f"    const T *src = p.{pslot.field};")
    L.append(f"    T *out = p.{oslot.field};")
    L.append("    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;")
    L.append("    const int64_t c0 = (int64_t)blockIdx.x * W;")
    L.append("    const int64_t i0 = p.r_lo + ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * p.chunk0;")
    L.append("    if (i0 >= p.r_hi) return;")
    L.append("    const int64_t iend = (i0 + p.chunk0 < p.r_hi) ? (i0 + p.chunk0) : p.r_hi;")
    L.append("    const int64_t S0 = p.cols, total = p.n0 * p.cols;")
    L.append("    const int nrows = (int)(iend - i0) + 2 * HA + QS + LAG2;     // input rows this CTA streams")
    L.append("    if (threadIdx.x == 0) {")
    L.append("        for (int s = 0; s < NS; ++s) { xgb::pipe::mbar_init(&full[s], 1); xgb::pipe::mbar_init(&empty[s], NCW); }")
    L.append("        xgb::pipe::fence_barrier_init();")
    L.append("    }")
    L.append("    __syncthreads();")
    L.append("    if (warp == NCW) {                                         // producer: bulk copies of one input row per grid")
    L.append("        int s = 0, eph = 1;")
    L.append("        const int64_t r0 = i0 - (HA + QS);")
    L.append("        const T *row = src + (r0 * S0 + (c0 - HK0));")
    for n, xs in enumerate(xslots):
        # rows r0-HA .. r0+HA-1 of the other grids are never used (the first middle row is r0+HA): start at the
        # first needed row and hold it for the first 2*HA copies, so nothing before row i0-QS is touched
        L.append(f"        const T *xrow{n} = p.{xs.field} + ((r0 + HA) * S0 + (c0 - HKM));")
    L.append("        for (int t = 0; t < nrows; ++t, ++s) {")
    L.append("            if (s == NS) { s = 0; eph ^= 1; }")
    L.append("            if (t >= NS) xgb::pipe::mbar_wait(&empty[s], eph);")
    L.append("            if (lane == 0) {")
    L.append("                xgb::pipe::mbar_expect_tx(&full[s], (uint32_t)(STG * sizeof(T)));")
    L.append("                // the last row a CTA streams (iend + HA + QS) only feeds a middle row nothing reads; past the")
    L.append("                // padded end of the level it is replaced by the row before it")
    L.append("                const T *rsrc = (r0 + t <= p.n0 + QS) ? row : (row - S0);")
    L.append("                xgb::pipe::bulk_g2s(stages + s * STG, rsrc, (uint32_t)(WP0 * sizeof(T)), &full[s]);")
    for n in range(len(xslots)):
        L.append(f"                xgb::pipe::bulk_g2s(stages + s * STG + WP0 + {n} * WPM, xrow{n}, (uint32_t)(WPM * sizeof(T)), &full[s]);")
    L.append("            }")
    L.append("            __syncwarp();")
    L.append("            row += S0;")
    for n in range(len(xslots)):
        L.append(f"            if (t >= 2 * HA) xrow{n} += S0;")
    L.append("        }")
    L.append("        return;")
    L.append("    }")
    L.extend(hoist_lines(hoist))
    L.append("    const int tid = warp * 32 + lane;                           // consumer thread 0..255")
    L.append("    int fs = 0, fph = 0, rs = 0, ms = 0;                        // newest stage / parity; oldest live stage; middle slot")
    L.append("    int64_t lin_mid = (i0 - QS) * S0 + (c0 - HKM);              // linear index of the next middle row's first column")
    L.append("    int64_t lin_out = i0 * S0 + c0;")
    L.append("    int fnx[NIT1];                                              // chunk flags of the NEXT middle row (global loads")
    L.append("#pragma unroll")
    L.append("    for (int i = 0; i < NIT1; ++i) {                            // issued one row ahead of their use)")
    L.append("        const int kk = (tid + i * NT) * V;")
    L.append("        const int64_t lin = lin_mid + kk;")
    L.append("        fnx[i] = -1;")
    L.append(f"        if (kk < WPM && lin >= 0 && lin + V <= total) fnx[i] = xgb::ld_flag(p.m_{pname}, p.f_{pname}, lin);")
    L.append("    }")
    L.append("    for (int t = 0; t < nrows; ++t) {")
    L.append("        xgb::pipe::mbar_wait(&full[fs], fph);")
    L.append("        if (t >= 2 * HA) {")
    L.append("            // ---- sweep A: middle row q = i0 - QS + (t - 2*HA) over columns [c0 - HKM, c0 + W + HKM)")
    for d0 in win:
        k = HA - d0                      # input row q + d0 sits k stages behind the newest (row q + HA)
        L.append(f"            const T *in{d0 + HA} = stages + ((fs >= {k}) ? (fs - {k}) : (fs - {k} + NS)) * STG + (HK0 - HKM);")
    for n in range(len(xslots)):
        L.append(f"            const T *xa{n} = stages + fs * STG + WP0 + {n} * WPM;")
    L.append("            T *mrow = mids + ms * WPM;")
    L.append("            int fls[NIT1];")
    L.append("#pragma unroll")
    L.append("            for (int i = 0; i < NIT1; ++i) {")
    L.append("                const int kk = (tid + i * NT) * V;")
    L.append("                const int64_t lin = lin_mid + S0 + kk;")
    L.append("                fls[i] = fnx[i];")
    L.append("                fnx[i] = -1;")
    L.append(f"                if (kk < WPM && lin >= 0 && lin + V <= total && t + 1 < nrows) fnx[i] = xgb::ld_flag(p.m_{pname}, p.f_{pname}, lin);")
    L.append("            }")
    L.append("#pragma unroll")
    L.append("            for (int i = 0; i < NIT1; ++i) {")
    L.append("                const int kk = (tid + i * NT) * V;")
    L.append("                bool nz = false;")
    L.append("                if (kk < WPM) {")
    L.append("                    const int fl = fls[i];")
    L.append("                    const int64_t lin = lin_mid + kk;")
    L.append("                    T val[V];")
    L.append("                    unsigned mb = 0u;                                   // the V boundary values, one byte each")
    L.append("                    if (fl == 0) {")
    L.extend(windows("in", "                        "))
    L.extend(xloads("xa", "                        "))
    L.append("#pragma unroll")
    L.append(f"                        for (int v = 0; v < V; ++v) val[v] = {rhs};")
    L.append("                    } else if (fl > 0) {")
    L.append(f"                        int m[V]; xgb::ld_mask_flagged<V>(p.m_{pname}, fl, lin, m);")
    L.extend(windows("in", "                        "))
    L.extend(xloads("xa", "                        "))
    L.append("#pragma unroll")
    L.append("                        for (int v = 0; v < V; ++v) {")
    L.append(f"                            if (m[v] == 0) val[v] = {rhs}; else {{ val[v] = {centre}; nz = true; }}")
    L.append("                            mb |= (unsigned)m[v] << (8 * v);")
    L.append("                        }")
    L.append("                    } else {")
    L.append("#pragma unroll")
    L.append("                        for (int v = 0; v < V; ++v) val[v] = T(0);       // outside the array: ghost zeros,")
    L.append("                        mb = 0xffffffffu;                                 // boundary value 255 = no statement")
    L.append("                    }")
    L.append("                    xgb::st_vec<T, V>(mrow + kk, val);")
    L.append("                    *reinterpret_cast<MB *>(mm + ms * WPMB + kk) = (MB)mb;")
    L.append("                }")
    L.append("                const bool any = __any_sync(0xffffffffu, nz);")
    L.append("                if (lane == 0 && kk < WPM) mnz[ms * NB + (kk >> 6)] = any ? 1 : 0;")
    L.append("            }")
    L.append("            lin_mid += S0;")
    L.append("            asm volatile(\"bar.sync 1, %0;\" :: \"n\"(NT) : \"memory\");      // middle row q complete")
    L.append("        }")
    L.append("        if (t >= 2 * HA + 1) {")
    L.append("            // ---- boundary statements on middle row q-1 (rows q-2 .. q are complete): each boundary point takes")
    L.append("            //      its resolved value in place; chain ends are never written, so no ordering is needed")
    L.append("            const int msx = (ms >= 1) ? (ms - 1) : (MR - 1);")
    L.append("#pragma unroll")
    L.append("            for (int i = 0; i < NIT1; ++i) {")
    L.append("                const int kk = (tid + i * NT) * V;")
    L.append("                if (kk < WPM && mnz[msx * NB + (kk >> 6)]) {")
    L.append("                    const unsigned mb = *reinterpret_cast<const MB *>(mm + msx * WPMB + kk);")
    L.append("#pragma unroll 1")
    L.append("                    for (int v = 0; v < V; ++v)")
    L.append("                        if ((mb >> (8 * v)) & 0xffu)")
    L.append(f"                            mids[msx * WPM + kk + v] = {name}_res(mids, mm, msx, kk + v, p);")
    L.append("                }")
    L.append("            }")
    L.append("        }")
    L.append("        if (t >= 2 * HA + QS + LAG2) {")
    L.append("            // ---- sweep B: output row o = q - LAG2 = i0 + (t - 2*HA - QS - LAG2); its middle rows o-1 .. o+1 had")
    L.append("            //      their boundary statements applied before this iteration's barrier")
    for d0 in win:
        k = LAG2 - d0
        L.append(f"            const T *md{d0 + HA} = mids + ((ms >= {k}) ? (ms - {k}) : (ms - {k} + MR)) * WPM + HKM;")
    L.append("            const int mso = (ms >= LAG2) ? (ms - LAG2) : (ms - LAG2 + MR);")
    L.append("            const int xst = (fs >= LAG2) ? (fs - LAG2) : (fs - LAG2 + NS);     // stage that carries the other grids' row o")
    for n in range(len(xslots)):
        L.append(f"            const T *xb{n} = stages + xst * STG + WP0 + {n} * WPM + HKM;")
    L.append("#pragma unroll")
    L.append("            for (int i = 0; i < NIT2; ++i) {")
    L.append("                const int kk = (tid + i * NT) * V;")
    L.append("                if (kk < W && c0 + kk < p.cols) {")
    L.append("                    const int64_t lin = lin_out + kk;")
    L.append("                    T val[V];")
    L.extend(windows("md", "                    "))
    L.extend(xloads("xb", "                    "))
    L.append("                    unsigned mb = 0u;")
    L.append("                    if (mnz[mso * NB + ((kk + HKM) >> 6)]) mb = *reinterpret_cast<const MB *>(mm + mso * WPMB + kk + HKM);")
    L.append("#pragma unroll")
    L.append("                    for (int v = 0; v < V; ++v) {")
    L.append(f"                        val[v] = {rhs};")
    L.append(f"                        if ((mb >> (8 * v)) & 0xffu) val[v] = {centre};     // boundary point: carried / resolved value")
    L.append("                    }")
    L.append("                    xgb::st_vec<T, V>(out + lin, val);")
    L.append("                }")
    L.append("            }")
    L.append("            lin_out += S0;")
    L.append("        }")
    L.append("        if (t >= KEEP) {                                            // the stage of time t - KEEP is dead")
    L.append("            __syncwarp();")
    L.append("            if (lane == 0) xgb::pipe::mbar_arrive(&empty[rs]);")
    L.append("            rs = (rs + 1 == NS) ? 0 : rs + 1;")
    L.append("        }")
    L.append("        if (t >= 2 * HA) ms = (ms + 1 == MR) ? 0 : ms + 1;")
    L.append("        if (++fs == NS) { fs = 0; fph ^= 1; }")
    L.append("    }")
    L.append("}")
    return "\n".join(R + L) + "\n"


def chains_fit(pair: Pair, mask: np.ndarray) -> bool:
    """Host-side check, once per boundary-mask version: simulate the in-kernel boundary resolution
    for every boundary point that has a copy rule and require each chain to stay within DR rows and
    DR columns of its start (the kernel keeps exactly that much of the middle state on chip)."""
    flat = mask.reshape(-1)
    total = flat.size
    cols = mask.shape[-1]
    pts = np.flatnonzero(flat)
    if pts.size * SPARSE_LIMIT > total:
        return False
    nr = len(pair.rules)
    lut = np.full(256, -1, np.int64)
    dr = np.zeros(nr + 1, np.int64)
    dc = np.zeros(nr + 1, np.int64)
    is_const = np.zeros(nr + 1, bool)
    for j, r in enumerate(pair.rules):
        if not 0 < r.mask < 255:
            return False
        lut[r.mask] = j
        dr[j], dc[j] = r.offset
        is_const[j] = r.const is not None
    y = pts.astype(np.int64)
    limit = np.full(y.size, nr, np.int64)
    rr = np.zeros(y.size, np.int64)
    cc = np.zeros(y.size, np.int64)
    live = np.ones(y.size, bool)
    for _ in range(nr + 1):
        inside = live & (y >= 0) & (y < total)
        m = np.zeros(y.size, np.int64)
        m[inside] = flat[y[inside]]
        m[live & ~inside] = 255
        if m.max(initial=0) > 255 or m.min(initial=0) < 0:
            return False
        j = lut[m]
        if (live & (j >= 0) & (j >= limit)).any():
            return False            # a chain stops at a point a LATER statement writes: in-place update unsafe
        hop = live & (j >= 0) & (j < limit) & ~is_const[j]
        live = hop
        if not hop.any():
            return True
        rr[hop] += dr[j[hop]]
        cc[hop] += dc[j[hop]]
        y[hop] += dr[j[hop]] * cols + dc[j[hop]]
        limit[hop] = j[hop]
        if np.abs(rr[hop]).max() > DR or np.abs(cc[hop]).max() > DR:
            return False
    return not live.any()
