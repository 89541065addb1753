"""Back end: sweep groups -> CUDA C for sm_100a.

Replaces the reference's C99/OpenMP generator (xgrid/lang/generator.py:79-442).
The reference emits, per stencil statement, a full-grid loop nest with a mask
test (generator.py:289-304,353-361).  Here a *group* of statements is lowered
to one ``__global__`` function assembled from the hand-written primitives in
``csrc/templates/xgb_stencil.cuh``; this module only decides *what* is
computed per point (expression text, which row windows are read, which slots
are written) -- the ``how`` lives in the template header.

Expression lowering follows generator.py:377-442: fully parenthesised C in
source order, unsuffixed float literals (so fp32 programs keep the reference's
mixed precision), ``**`` -> ``pow``/``powf`` by result type, with
the literal exponent ``2.0`` emitted as a product because gcc folds it at the
reference's optimisation levels (F7).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field

from ..types import Boolean, Floating, Grid as GridT, Integer, Pointer, Structure, Value, Void
from . import ir

VARIANT_DENSE = "dense"
VARIANT_SPARSE = "sparse"
VARIANT_MARCH = "march"
VARIANT_TILED = "tiled"
VARIANT_MULTISTEP = "multistep"
VARIANT_MULTISTEP_TAIL = "multistep_tail"
VARIANT_MULTISTEP_SHORT = "multistep_short"     # run-time step count <= T/2 on half-size windows (short runs)
VARIANT_TILED2 = "tiled2"
MARCH_ROWS = {2: 16, 3: 8}      # unroll depth of the marching loop (axis-0 points per trip)
import os as _os
# rows loaded ahead of use per chain; 0 = off (measured: register prefetch costs occupancy, the
# tiled variant prefetches through shared memory instead)
MARCH_PREFETCH = int(_os.environ.get("XGB_PF", "0"))
# exact division by point-independent divisors through a per-thread reciprocal (xgb::InvDiv), per kernel
# variant: it pays where a kernel is instruction-bound (the temporal-blocking variants); in the HBM-bound
# one-pass kernels the divide is hidden anyway and the extra live registers cost more than they save
# (measured: conv1d_nl multistep 655 -> 956 Gpoint-updates/s; cavity with it in the one-pass / fused kernels 15.8 -> 16.1-16.3 ms)
INVDIV_VARIANTS = set(filter(None, _os.environ.get("XGB_INVDIV", "multistep").split(",")))


@dataclass
class Slot:
    """One (grid argument, ring level) buffer touched by a group.  ``level`` is
    an int ring index, or "scratch" for the Jacobi double buffer that replaces
    the reference's per-statement ``malloc`` (generator.py:321,351)."""
    index: int
    grid: str
    level: object
    elem: Value
    read: bool = False
    written: bool = False
    halo0: int = 0          # max |axis-0 offset| this slot is read at (rows a slab must import)
    taps: set = field(default_factory=set)      # every space offset this slot is read at

    def overhang(self, shape: tuple) -> int:
        """Elements BEYOND `halo0` whole rows that a slab must import: a tap at the deepest axis-0 offset that also
        moves along the trailing axes in the same direction leaves its row at the first / last column and -- taps
        being linear addresses (F10) -- lands in the tail of the row one further out."""
        if not self.halo0 or len(shape) < 2:
            return 0
        strides = [1]
        for n in reversed(shape[2:]):
            strides.insert(0, strides[0] * n)
        over = 0
        for off in self.taps:
            if abs(off[0]) == self.halo0:
                rest = sum(o * st for o, st in zip(off[1:], strides))
                if rest * off[0] > 0:
                    over = max(over, abs(rest))
        return over

    @property
    def field(self) -> str:
        return f"s{self.index}"


@dataclass
class Group:
    gid: int
    ndim: int
    stmts: list                                   # ir.Assignment with .sweep
    slots: list = field(default_factory=list)     # Slot
    masks: list = field(default_factory=list)     # grid names whose mask is consulted
    scalars: dict = field(default_factory=dict)   # variable name -> type (by-value params)
    shapes: list = field(default_factory=list)    # grid names whose shape() is queried
    implicit: bool = False
    sparse: bool = False                          # every statement runs on a mask value != 0
    lead: str = ""                                # grid whose shape drives the sweep
    name: str = ""
    params_cls: type | None = None
    vwidths: tuple = (1,)
    halo0: int = 0                                # max |axis-0 offset| over all loads
    halo_last: int = 0
    march: bool = False                           # has axis-0 marching variants
    tiled: dict | None = None                     # geometry of the async shared-memory pipeline variant
    multistep: dict | None = None                 # temporal-blocking variant (1-D, whole-kernel groups)
    multistep_short: dict | None = None           # its short-run sibling (half the halo, half the window)
    tiled2: dict | None = None                    # two time steps per pass, 2-D (time-skewed tiled pipeline)

    def slot(self, grid: str, level) -> Slot:
        for s in self.slots:
            if s.grid == grid and s.level == level:
                return s
        raise KeyError((grid, level))


class CodegenError(Exception):
    pass


# --------------------------------------------------------------------------- expressions
class ExprEmitter:
    """IR expression -> C text.  ``ident`` maps a scalar variable to its C
    spelling; ``tap`` maps a Stencil load to its C spelling."""

    def __init__(self, module: "ModuleBuilder", ident, tap=None, hoist: dict | None = None,
                 variant: str = "") -> None:
        self.module, self.ident, self.tap = module, ident, tap
        self.hoist = hoist          # C text -> name of a kernel-prologue constant (None = off)
        self.invdiv = variant in INVDIV_VARIANTS

    def __call__(self, e) -> str:
        text = getattr(self, "x_" + type(e).__name__)(e)
        if self.hoist is not None and isinstance(e, (ir.Binary, ir.Unary, ir.Condition, ir.Cast, ir.Call)) \
                and _invariant(e):
            # point-independent subexpression: evaluate once per thread instead of once per point
            # (`auto` keeps the C type of the expression, so mixed-precision programs are unchanged)
            name = self.hoist.get(text)
            if name is None:
                name = self.hoist[text] = f"h{len(self.hoist)}"
            return name
        return text

    def x_Constant(self, e: ir.Constant) -> str:
        return repr(e.value).lower()            # generator.py:393-394

    def x_Identifier(self, e: ir.Identifier) -> str:
        return self.ident(e.variable)

    def x_Access(self, e: ir.Access) -> str:
        return f"({self(e.value)}).{e.attribute}"

    def x_Unary(self, e: ir.Unary) -> str:
        return f"({e.operator} {self(e.right)})"

    def x_Binary(self, e: ir.Binary) -> str:
        if e.operator == "^":
            wide = isinstance(e.type, Floating) and e.type.width_bits == 64
            base = self(e.left)
            if isinstance(e.right, ir.Constant) and e.right.value == 2.0:
                ctype = "double" if wide else "float"
                return f"xgb::sq<{ctype}>({base})"
            return f"{'pow' if wide else 'powf'}({base}, {self(e.right)})"
        if e.operator == "/" and isinstance(e.left.type, Floating) and self.tap is not None:
            if (self.invdiv and self.hoist is not None and _invariant(e.right) and not _invariant(e.left)
                    and all(isinstance(t, Floating) and t.width_bits == 64 for t in (e.left.type, e.right.type))):
                # point-independent divisor: reciprocal once per thread, exact quotient per point
                divisor = self(e.right)
                text = f"xgb::InvDiv({divisor})"
                name = self.hoist.get(text)
                if name is None:
                    name = self.hoist[text] = f"h{len(self.hoist)}"
                return f"{name}.div({self(e.left)})"
            return f"xgb::fdiv({self(e.left)}, {self(e.right)})"     # exact; see xgb_stencil.cuh
        return f"({self(e.left)} {e.operator} {self(e.right)})"

    def x_Condition(self, e: ir.Condition) -> str:
        return f"({self(e.condition)} ? {self(e.body)} : {self(e.orelse)})"

    def x_Cast(self, e: ir.Cast) -> str:
        return f"(({self.module.ctype(e.type)})({self(e.value)}))"

    def x_Stencil(self, e: ir.Stencil) -> str:
        if self.tap is None:
            raise CodegenError(f"grid access to '{e.variable.name}' outside a stencil statement")
        return self.tap(e)

    def x_GridInfo(self, e: ir.GridInfo) -> str:
        if e.info == "dimension":
            return str(e.variable.type.dimension)
        return self.module.shape_ref(e.variable.name, self(e.dimension))

    def x_Call(self, e: ir.Call) -> str:
        args = ", ".join(self(a) for a in e.arguments)
        if isinstance(e.operator, ir.Constructor):
            return f"({self.module.ctype(e.operator.type)}{{{args}}})"
        return f"{self.module.device_function(e.operator)}({args})"


def _invariant(e) -> bool:
    """True when the expression reads no grid and calls nothing that could write memory."""
    for n in ir.walk_expr(e):
        if isinstance(n, ir.Stencil):
            return False
        if isinstance(n, ir.Call) and not isinstance(n.operator, ir.Constructor):
            if any(isinstance(t, Pointer) for _, t in n.operator.signature.arguments):
                return False
    return True


def hoist_lines(hoist: dict, indent: str = "    ") -> list:
    return [f"{indent}const auto {name} = {text};" for text, name in hoist.items()]


# --------------------------------------------------------------------------- module
_KERNEL_NAME = __import__("re").compile(r'extern "C" __global__ void (?:__launch_bounds__\([^)]*\) )?(\w+)\(')


class ModuleBuilder:
    """Accumulates one CUDA translation unit: struct definitions, ``__device__``
    helpers for called operators, and one or more sweep kernels."""

    def __init__(self, overstep: str = "none", comment: bool = False) -> None:
        self.overstep = overstep
        self.comment = comment
        self.structs: dict[str, str] = {}
        self.functions: dict[str, str] = {}
        self.kernels: list[str] = []
        self.includes: list[str] = []          # user headers (`includes=` of the kernel and of every operator it calls)
        self._shape_ref = None

    def include(self, names) -> None:
        for n in names:
            if n not in self.includes:
                self.includes.append(n)

    def _preamble(self) -> list:
        return ['#include "xgb_stencil.cuh"\n', *[f'#include "{n}"\n' for n in self.includes]]

    # ---- types
    def ctype(self, t) -> str:
        if isinstance(t, Void):
            return "void"
        if isinstance(t, Structure):
            if t.name not in self.structs:
                self.structs[t.name] = ""      # reserve (recursive fields)
                body = "".join(f"    {self.ctype(ft)} {fn};\n" for fn, ft in t.elements)
                self.structs[t.name] = f"struct {t.name} {{\n{body}}};\n"
            return t.name
        if isinstance(t, (Boolean, Integer, Floating)):
            return t.cname
        if isinstance(t, Pointer):
            return self.ctype(t.element) + "*"
        raise CodegenError(f"type '{t}' has no device representation")

    def shape_ref(self, grid: str, dim_text: str) -> str:
        if self._shape_ref is None:
            raise CodegenError("shape() is only available inside kernels")
        return self._shape_ref(grid, dim_text)

    # ---- callee operators become __device__ functions (generator.py:208-212,418-419)
    def device_function(self, op) -> str:
        self.include(op.includes)
        if op.mode == "external":
            # the reference emits an `extern` prototype and leaves the definition to the C files named in
            # `includes=` (generator.py:208-212); here the definition is a `__device__` function in a CUDA
            # header named the same way -- the call is emitted by name and the header must declare it
            if not op.includes:
                raise CodegenError(f"external operator '{op.name}' needs `includes=[...]` naming the CUDA header "
                                   "that defines it as a __device__ function")
            return op.name
        cname = op.name.replace(".", "_")
        if cname in self.functions:
            return cname
        self.functions[cname] = ""             # reserve (recursion)
        d = op.ir
        for _, t in d.signature.arguments:
            if isinstance(t, GridT):
                raise CodegenError(f"operator '{op.name}' takes a grid and cannot be inlined into a sweep")
        args = ", ".join(f"{self.ctype(t)} {n}" for n, t in d.signature.arguments)
        lines = [f"__device__ __forceinline__ {self.ctype(d.signature.return_type)} {cname}({args}) {{"]
        for n, v in d.scope.items():
            if n not in d.signature.argnames_map:
                lines.append(f"    {self.ctype(v.type)} {n};")

        def ident(var):
            return f"(*{var.name})" if isinstance(var.type, Pointer) else var.name

        emit = ExprEmitter(self, ident)
        self._scalar_block(d.body, emit, lines, 1)
        lines.append("}")
        self.functions[cname] = "\n".join(lines) + "\n"
        return cname

    def _scalar_block(self, stmts, emit, out, ind) -> None:
        pad = "    " * ind
        for s in stmts:
            if isinstance(s, ir.Assignment):
                if s.sweep is not None:
                    raise CodegenError("stencil statements inside called operators are not supported")
                out.append(f"{pad}{emit(s.terminal)} = {emit(s.value)};")
            elif isinstance(s, ir.Return):
                out.append(f"{pad}return{'' if s.value is None else ' ' + emit(s.value)};")
            elif isinstance(s, ir.Break):
                out.append(f"{pad}break;")
            elif isinstance(s, ir.Continue):
                out.append(f"{pad}continue;")
            elif isinstance(s, ir.If):
                out.append(f"{pad}if ({emit(s.condition)}) {{")
                self._scalar_block(s.body, emit, out, ind + 1)
                if s.orelse:
                    out.append(f"{pad}}} else {{")
                    self._scalar_block(s.orelse, emit, out, ind + 1)
                out.append(f"{pad}}}")
            elif isinstance(s, ir.While):
                out.append(f"{pad}while ({emit(s.condition)}) {{")
                self._scalar_block(s.body, emit, out, ind + 1)
                out.append(f"{pad}}}")
            elif isinstance(s, ir.For):
                v = s.variable.name
                out.append(f"{pad}for ({v} = {emit(s.start)}; {v} < {emit(s.end)}; {v} += {emit(s.step)}) {{")
                self._scalar_block(s.body, emit, out, ind + 1)
                out.append(f"{pad}}}")
            elif isinstance(s, ir.Evaluation):
                out.append(f"{pad}{emit(s.value)};")
            elif isinstance(s, ir.Inline):
                out.append(f"{pad}{s.source}")
            else:
                raise CodegenError(f"unsupported statement {type(s).__name__} in device function")

    # ---- final text
    def source(self) -> str:
        parts = self._preamble()
        parts.extend(self.structs.values())
        parts.extend(self.functions.values())
        parts.extend(self.kernels)
        return "\n".join(parts)

    def chunks(self, n: int) -> list:
        """The translation unit split into `n` independent ones for parallel compilation:
        [(kernel names, source)].  Every chunk carries all declarations (user structs, callee
        functions, parameter structs, inline-block prelude) in their original order and a share of
        the kernels, balanced by text length -- kernels only depend on declarations, never on each
        other."""
        shared = [*self._preamble(), *self.structs.values(), *self.functions.values()]
        kernels = []
        for text in self.kernels:
            names = _KERNEL_NAME.findall(text)
            if names:
                kernels.append((names, text))
            else:
                shared.append(text)
        n = max(1, min(n, len(kernels)))
        bins = [([], [], 0) for _ in range(n)]
        for names, text in sorted(kernels, key=lambda k: -len(k[1])):
            k = min(range(n), key=lambda i: bins[i][2])
            bins[k] = (bins[k][0] + names, bins[k][1] + [text], bins[k][2] + len(text))
        return [(names, "\n".join(shared + texts)) for names, texts, _ in bins if names]


# --------------------------------------------------------------------------- group analysis
def analyse_group(g: Group, scope: dict) -> None:
    """Fill slots / masks / scalars / halo extents of a group from its statements."""
    def slot(grid, level, elem):
        for s in g.slots:
            if s.grid == grid and s.level == level:
                return s
        s = Slot(len(g.slots), grid, level, elem)
        g.slots.append(s)
        return s

    for a in g.stmts:
        sw = a.sweep
        if sw.grid.name not in g.masks:
            g.masks.append(sw.grid.name)
        store_level = "scratch" if g.implicit else sw.store.level
        slot(sw.grid.name, store_level, sw.grid.type.element).written = True
        if g.implicit:
            # unwritten points carry the old value through the double buffer
            slot(sw.grid.name, 0, sw.grid.type.element).read = True
        for ld in sw.loads:
            sl = slot(ld.variable.name, ld.level, ld.variable.type.element)
            sl.read = True
            sl.halo0 = max(sl.halo0, abs(ld.space_offset[0]))
            sl.taps.add(tuple(ld.space_offset))
            g.halo0 = max(g.halo0, abs(ld.space_offset[0]) if g.ndim > 1 else 0)
            g.halo_last = max(g.halo_last, abs(ld.space_offset[-1]))
        for e in ir.walk_expr(a.value):
            if isinstance(e, ir.Identifier):
                v = e.variable
                if isinstance(v.type, GridT):
                    continue
                g.scalars.setdefault(v.name, v.type)
            elif isinstance(e, ir.GridInfo) and e.info == "shape":
                if e.variable.name not in g.shapes:
                    g.shapes.append(e.variable.name)
    g.lead = g.stmts[0].sweep.grid.name
    g.sparse = all(a.sweep.mask != 0 for a in g.stmts) and len(g.stmts) == 1
    if g.ndim > 1 and not g.sparse:
        # a slot read only at axis-0 offset 0 but off its column / plane position -- u[0, -1] -- still leaves its row
        # at the first / last column (taps are linear addresses, F10) and then reads the adjacent row: on a slab that
        # is a ghost row, so a sweep over the whole grid imports one.  (A lone boundary statement -- a sparse group --
        # is confined by its mask; the cavity's wall copies are of that kind and keep exchanging nothing.)
        for sl in g.slots:
            if sl.read and sl.halo0 == 0 and any(any(o != 0 for o in off[1:]) for off in sl.taps):
                sl.halo0 = 1
                g.halo0 = max(g.halo0, 1)


def vector_widths(g: Group) -> tuple:
    sizes = []
    for s in g.slots:
        if isinstance(s.elem, Structure):
            return (1,)
        sizes.append(s.elem.width_bytes)
    big = max(sizes) if sizes else 8
    base = max(1, 16 // big)
    out = [1]
    if base > 1:
        out.append(base)
    out.append(base * 2)
    return tuple(out)


# --------------------------------------------------------------------------- params struct
def build_params(g: Group, module: ModuleBuilder, scope_types: dict, grid_ndims: dict) -> tuple[str, type]:
    """C text of the by-value parameter struct + its ctypes twin (same field
    order; natural alignment on both sides)."""
    c_fields, py_fields = [], []

    def add(ctype_text, name, pytype):
        c_fields.append(f"    {ctype_text} {name};")
        py_fields.append((name, pytype))

    for s in g.slots:
        t = module.ctype(s.elem)
        qual = "" if s.written else "const "
        add(f"{qual}{t}* __restrict__" if not (s.read and s.written) else f"{t}*", s.field, ctypes.c_void_p)
    for m in g.masks:
        add("const uint8_t* __restrict__", f"m_{m}", ctypes.c_void_p)
        add("const uint8_t* __restrict__", f"f_{m}", ctypes.c_void_p)
    for a in range(4):
        add("void*", f"aux{a}", ctypes.c_void_p)  # extra level pointers of the multi-step variant
    add("const int64_t* __restrict__", "list", ctypes.c_void_p)
    add("int64_t", "count", ctypes.c_int64)
    add("int64_t", "chunk0", ctypes.c_int64)     # axis-0 points per CTA in the marching variant
    add("int64_t", "opt0", ctypes.c_int64)       # variant-specific flag (tiled2: also write the middle level)
    add("int64_t", "open_lo", ctypes.c_int64)    # slab has a neighbour below / above: its ghost points
    add("int64_t", "open_hi", ctypes.c_int64)    # are real grid points (multi-step variant)
    add("int64_t", "r_lo", ctypes.c_int64)       # axis-0 range [r_lo, r_hi) swept by this launch
    add("int64_t", "r_hi", ctypes.c_int64)       # (march / tiled variants; edge-first launches of a slab)
    add("int64_t", "rows", ctypes.c_int64)
    add("int64_t", "cols", ctypes.c_int64)
    for a in range(g.ndim):
        add("int64_t", f"n{a}", ctypes.c_int64)
    for name in g.shapes:
        nd = grid_ndims[name]
        add(f"int32_t", f"shape_{name}[{nd}]", ctypes.c_int32 * nd)
    # user scalars / structs, widest first to avoid padding surprises
    def width(t):
        return ctypes.sizeof(t.ctype) if isinstance(t, Structure) else t.width_bytes
    for name, t in sorted(g.scalars.items(), key=lambda kv: -min(8, width(kv[1].element if isinstance(kv[1], Pointer) else kv[1]))):
        vt = t.element if isinstance(t, Pointer) else t
        add(module.ctype(vt), f"u_{name}", vt.ctype)
    # ctypes cannot declare "name[N]" -- strip the suffix for the python twin
    py_clean = [(n.split("[")[0], t) for n, t in py_fields]
    cls = type(f"{g.name}_P", (ctypes.Structure,), {"_fields_": py_clean})
    text = f"struct {g.name}_P {{\n" + "\n".join(c_fields) + "\n};\n"
    return text, cls


# --------------------------------------------------------------------------- kernels
def _outer_offset_text(g: Group, outer: tuple) -> str:
    """Linear element offset of an outer-axis tap (all axes but the last)."""
    terms = []
    for a, d in enumerate(outer):
        if d == 0:
            continue
        # stride of axis a = product of extents of the axes after it
        stride = " * ".join(f"p.n{b}" for b in range(a + 1, g.ndim))
        terms.append(f"({d}LL) * {stride}")
    return " + ".join(terms) if terms else "0"


def emit_group(g: Group, module: ModuleBuilder, scope: dict, grid_ndims: dict) -> None:
    """Append the parameter struct and every variant of the group's kernel."""
    g.vwidths = vector_widths(g)
    ptext, g.params_cls = build_params(g, module, scope, grid_ndims)
    module.kernels.append(ptext)
    module._shape_ref = lambda grid, dim: f"p.shape_{grid}[{dim}]"
    try:
        if module.overstep == "none":
            for v in g.vwidths:
                module.kernels.append(_emit_windowed(g, module, v, VARIANT_DENSE))
            if g.sparse:
                module.kernels.append(_emit_windowed(g, module, 1, VARIANT_SPARSE))
            g.march = march_supported(g)
            if g.march:
                for v in g.vwidths:
                    if v > 1:
                        module.kernels.append(_emit_march(g, module, v, MARCH_ROWS[g.ndim]))
                g.tiled = tiled_config(g)
                if g.tiled is not None:
                    module.kernels.append(_emit_tiled(g, module, g.tiled))
            g.multistep = multistep_config(g)
            if g.multistep is not None:
                module.kernels.append(_emit_multistep(g, module, g.multistep))
                module.kernels.append(_emit_multistep(g, module, g.multistep, tail=True))
                g.multistep_short = multistep_config(g, short=True)
                if g.multistep_short is not None:
                    module.kernels.append(_emit_multistep(g, module, g.multistep_short, tail=True, short=True))
            g.tiled2 = tiled2_config(g)
            if g.tiled2 is not None:
                module.kernels.append((_emit_tiled2_3d if g.ndim == 3 else _emit_tiled2)(g, module, g.tiled2))
        else:
            module.kernels.append(_emit_general(g, module, VARIANT_DENSE))
            module.kernels.append(_emit_general(g, module, VARIANT_SPARSE))
            g.vwidths = (1,)
    finally:
        module._shape_ref = None


def kernel_name(g: Group, variant: str, v: int) -> str:
    return f"{g.name}_{variant}_v{v}"


def _ident(var) -> str:
    return f"p.u_{var.name}"


def _emit_statements(g: Group, module: ModuleBuilder, tap, lines: list, vexpr: str,
                     hoist: dict | None = None) -> None:
    """Per-point body: statement-at-a-time in program order, each predicated on
    the stored grid's mask value (generator.py:297-298)."""
    emit = ExprEmitter(module, _ident, tap, hoist)
    for a in g.stmts:
        sw = a.sweep
        store_level = "scratch" if g.implicit else sw.store.level
        s = g.slot(sw.grid.name, store_level)
        rhs = emit(a.value)
        if module.comment:
            # init(comment=True): map device code back to the user's Python source
            # (xgrid/lang/generator.py:242-244 emits the same #line directives into its C)
            lines.append(f'#line {a.location.line} "{a.location.file}"')
        lines.append(f"        if (m_{sw.grid.name}[{vexpr}] == {sw.mask}) {{ "
                     f"o_{s.field}[{vexpr}] = {rhs}; wr_{s.field} |= 1u << {vexpr}; }}")
        if g.implicit:
            old = tap(ir.Stencil(a.location, sw.grid.type.element, "load", sw.grid, 0,
                                 (0,) * g.ndim, 0))
            lines.append(f"        else {{ o_{s.field}[{vexpr}] = {old}; wr_{s.field} |= 1u << {vexpr}; }}")


def _emit_windowed(g: Group, module: ModuleBuilder, V: int, variant: str) -> str:
    # windows: (slot index, outer offsets) -> [lo, hi] over the contiguous axis
    windows: dict = {}

    def need(slot: Slot, offsets: tuple) -> None:
        key = (slot.index, tuple(offsets[:-1]))
        lo, hi = windows.get(key, (0, 0))
        windows[key] = (min(lo, offsets[-1]), max(hi, offsets[-1]))

    for a in g.stmts:
        for ld in a.sweep.loads:
            need(g.slot(ld.variable.name, ld.level), ld.space_offset)
        if g.implicit:
            need(g.slot(a.sweep.grid.name, 0), (0,) * g.ndim)
    wnames = {key: f"w{n}" for n, key in enumerate(windows)}

    def tap(e: ir.Stencil) -> str:
        slot = g.slot(e.variable.name, e.level)
        key = (slot.index, tuple(e.space_offset[:-1]))
        lo, _ = windows[key]
        return f"{wnames[key]}[v + {e.space_offset[-1] - lo}]"

    name = kernel_name(g, variant, V)
    L = [f'extern "C" __global__ void __launch_bounds__(256) {name}(const __grid_constant__ {g.name}_P p)', "{"]
    L.append(f"    constexpr int V = {V};")
    if variant == VARIANT_DENSE:
        L.append("    const xgb::Tile t = xgb::dense_tile<V>(p.rows, p.cols);")
        L.append("    if (!t.active) return;")
        L.append("    const int64_t base = t.row * p.cols + t.col;")
        for m in g.masks:
            L.append(f"    int m_{m}[V]; xgb::ld_mask<V>(p.m_{m}, p.f_{m}, base, m_{m});")
    else:
        L.append("    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;")
        L.append("    if (tid >= p.count) return;")
        L.append("    const int64_t base = p.list[tid];")
        for m in g.masks:
            L.append(f"    int m_{m}[V] = {{{g.stmts[0].sweep.mask}}};")
    for key, (lo, hi) in windows.items():
        slot = g.slots[key[0]]
        t = module.ctype(slot.elem)
        off = _outer_offset_text(g, key[1])
        L.append(f"    {t} {wnames[key]}[V + {hi - lo}]; "
                 f"xgb::ld_window<{t}, V, {lo}, {hi}>(p.{slot.field} + (base + {off}), {wnames[key]});")
    for s in g.slots:
        if s.written:
            L.append(f"    {module.ctype(s.elem)} o_{s.field}[V]; unsigned wr_{s.field} = 0u;")
    L.append("#pragma unroll")
    L.append("    for (int v = 0; v < V; ++v) {")
    hoist: dict = {}
    _emit_statements(g, module, tap, L, "v", hoist)
    L.append("    }")
    for s in g.slots:
        if s.written:
            L.append(f"    xgb::st_pred<{module.ctype(s.elem)}, V>(p.{s.field} + base, o_{s.field}, wr_{s.field});")
    L.append("}")
    L[3:3] = hoist_lines(hoist)
    return "\n".join(L) + "\n"


def _emit_general(g: Group, module: ModuleBuilder, variant: str) -> str:
    """overstep = "limit" / "wrap": every tap is addressed through clamped /
    wrapped per-axis coordinates (generator.py:172-177, with correct extents)."""
    adj = "xgb::clamp_idx" if module.overstep == "limit" else "xgb::wrap_idx"

    def tap(e: ir.Stencil) -> str:
        slot = g.slot(e.variable.name, e.level)
        coords = [f"{adj}(i{a} + ({d}), p.n{a})" for a, d in enumerate(e.space_offset)]
        # axis 0 of a slab reads its ghost rows where a neighbour exists (open_lo / open_hi; both 0 unsharded)
        coords[0] = f"{adj}0(i0 + ({e.space_offset[0]}), p.n0, p.open_lo, p.open_hi)"
        lin = coords[0]
        for a in range(1, g.ndim):
            lin = f"({lin}) * p.n{a} + {coords[a]}"
        return f"p.{slot.field}[{lin}]"

    name = kernel_name(g, variant, 1)
    L = [f'extern "C" __global__ void __launch_bounds__(256) {name}(const __grid_constant__ {g.name}_P p)', "{"]
    L.append("    constexpr int V = 1;")
    if variant == VARIANT_DENSE:
        L.append("    const xgb::Tile t = xgb::dense_tile<V>(p.rows, p.cols);")
        L.append("    if (!t.active) return;")
        L.append("    const int64_t base = t.row * p.cols + t.col;")
        for m in g.masks:
            L.append(f"    int m_{m}[V]; xgb::ld_mask<V>(p.m_{m}, p.f_{m}, base, m_{m});")
    else:
        L.append("    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;")
        L.append("    if (tid >= p.count) return;")
        L.append("    const int64_t base = p.list[tid];")
        for m in g.masks:
            L.append(f"    int m_{m}[V] = {{{g.stmts[0].sweep.mask}}};")
    L.append("    int64_t rem = base;")
    for a in range(g.ndim - 1, -1, -1):
        L.append(f"    const int64_t i{a} = rem % p.n{a}; rem /= p.n{a};")
    for s in g.slots:
        if s.written:
            L.append(f"    {module.ctype(s.elem)} o_{s.field}[V]; unsigned wr_{s.field} = 0u;")
    L.append("    { const int v = 0;")
    _emit_statements(g, module, tap, L, "v")
    L.append("    }")
    for s in g.slots:
        if s.written:
            L.append(f"    xgb::st_pred<{module.ctype(s.elem)}, V>(p.{s.field} + base, o_{s.field}, wr_{s.field});")
    L.append("}")
    return "\n".join(L) + "\n"


# --------------------------------------------------------------------------- marching variant
def march_supported(g: Group) -> bool:
    """Axis-0 marching needs 2-D / 3-D number grids and no slot that is both read
    and written by the group (in-place boundary statements stay on the direct path)."""
    if g.ndim not in (2, 3) or g.sparse:
        return False
    for s in g.slots:
        if s.read and s.written:
            return False
        if isinstance(s.elem, (Structure, Boolean)):
            return False
    return True


def _emit_march(g: Group, module: ModuleBuilder, V: int, R: int) -> str:
    """Each thread owns V contiguous columns (and one axis-1 index in 3-D) and
    marches R points along axis 0.  Taps that differ only in their axis-0 offset
    form a *chain* whose bodies stay in registers and rotate, so every input
    element is loaded once per thread; contiguous-axis neighbours come from the
    adjacent lanes by shuffle (xgb::window_from_body)."""
    nd = g.ndim
    chains: dict = {}      # (slot index, mid offsets) -> {"dmin","dmax","fringe": {d0: [lo, hi]}}

    def need(slot: Slot, off: tuple) -> None:
        key = (slot.index, tuple(off[1:-1]))
        ch = chains.setdefault(key, {"dmin": off[0], "dmax": off[0], "fringe": {}})
        ch["dmin"], ch["dmax"] = min(ch["dmin"], off[0]), max(ch["dmax"], off[0])
        lo, hi = ch["fringe"].get(off[0], (0, 0))
        ch["fringe"][off[0]] = (min(lo, off[-1]), max(hi, off[-1]))

    for a in g.stmts:
        for ld in a.sweep.loads:
            need(g.slot(ld.variable.name, ld.level), ld.space_offset)
        if g.implicit:
            need(g.slot(a.sweep.grid.name, 0), (0,) * nd)
    for n, ch in enumerate(chains.values()):
        ch["id"] = n
        for lo, hi in ch["fringe"].values():
            if -lo > V or hi > V:
                raise CodegenError("contiguous-axis offset wider than the vector body")

    def reg(ch, d0) -> str:
        return f"c{ch['id']}_{d0 - ch['dmin']}"

    def tap(e: ir.Stencil) -> str:
        slot = g.slot(e.variable.name, e.level)
        ch = chains[(slot.index, tuple(e.space_offset[1:-1]))]
        d0, dk = e.space_offset[0], e.space_offset[-1]
        lo, hi = ch["fringe"][d0]
        if (lo, hi) == (0, 0):
            return f"{reg(ch, d0)}[v]"
        return f"w{ch['id']}_{d0 - ch['dmin']}[v + {dk - lo}]"

    def mid_text(mid: tuple) -> str:
        terms = []
        for a, d in enumerate(mid, start=1):
            if d:
                stride = " * ".join(f"p.n{b}" for b in range(a + 1, nd))
                terms.append(f"({d}LL) * {stride}")
        return " + ".join(terms) if terms else "0"

    name = kernel_name(g, VARIANT_MARCH, V)
    L = [f'extern "C" __global__ void __launch_bounds__(256) {name}(const __grid_constant__ {g.name}_P p)', "{"]
    L.append(f"    constexpr int V = {V}, R = {R};")
    L.append("    const int lane = threadIdx.x & 31;")
    L.append("    const int64_t col_raw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;")
    L.append("    const bool act_c = col_raw < p.cols;")
    L.append("    const int64_t col = act_c ? col_raw : (p.cols - V);")
    L.append("    const bool edge_l = (lane == 0);")
    L.append("    const bool edge_r = (lane == 31) || (col_raw + V >= p.cols);")
    if nd == 3:
        L.append("    const int64_t j_raw = (int64_t)blockIdx.y * blockDim.y + threadIdx.y;")
        L.append("    const bool act = act_c && (j_raw < p.n1);")
        L.append("    const int64_t j = j_raw < p.n1 ? j_raw : (p.n1 - 1);")
        L.append("    const int64_t i0 = p.r_lo + (int64_t)blockIdx.z * p.chunk0;")
        L.append("    const int64_t S0 = p.n1 * p.n2;")
        L.append("    int64_t base = i0 * S0 + j * p.cols + col;")
    else:
        L.append("    const bool act = act_c;")
        L.append("    const int64_t i0 = p.r_lo + ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * p.chunk0;")
        L.append("    const int64_t S0 = p.cols;")
        L.append("    int64_t base = i0 * S0 + col;")
    L.append("    if (i0 >= p.r_hi) return;")
    L.append("    const int64_t iend = (i0 + p.chunk0 < p.r_hi) ? (i0 + p.chunk0) : p.r_hi;")
    # chain registers + prologue (all but the leading row of every chain)
    PF = MARCH_PREFETCH
    for ch in chains.values():
        ch["top"] = ch["dmax"] + PF          # furthest row held in registers (prefetched, not yet used)
    # rows past the slab's last ghost row must not be touched: clamp the prefetch distance
    L.append("    #define XGB_ROW(d, dmax) ((ib + r + (d) <= p.n0 - 1 + (dmax)) ? (int64_t)(d) : (p.n0 - 1 + (dmax) - (ib + r)))")
    for key, ch in chains.items():
        slot = g.slots[key[0]]
        t = module.ctype(slot.elem)
        for d0 in range(ch["dmin"], ch["top"] + 1):
            L.append(f"    {t} {reg(ch, d0)}[V];")
    L.append("    { const int64_t ib = i0; const int r = 0;")
    for key, ch in chains.items():
        slot = g.slots[key[0]]
        t = module.ctype(slot.elem)
        for d0 in range(ch["dmin"], ch["top"]):
            L.append(f"    xgb::ld_vec<{t}, V>(p.{slot.field} + (base + XGB_ROW({d0}, {ch['dmax']}) * S0 + {mid_text(key[1])}), {reg(ch, d0)});")
    L.append("    }")
    L.append("    for (int64_t ib = i0; ib < iend; ib += R) {")
    L.append("#pragma unroll")
    L.append("    for (int r = 0; r < R; ++r) {")
    L.append("        if (ib + r >= iend) break;")
    for key, ch in chains.items():
        slot = g.slots[key[0]]
        t = module.ctype(slot.elem)
        d0 = ch["top"]
        L.append(f"        xgb::ld_vec<{t}, V>(p.{slot.field} + (base + XGB_ROW({d0}, {ch['dmax']}) * S0 + {mid_text(key[1])}), {reg(ch, d0)});")
    for m in g.masks:
        L.append(f"        int m_{m}[V]; xgb::ld_mask<V>(p.m_{m}, p.f_{m}, base, m_{m});")
    for key, ch in chains.items():
        slot = g.slots[key[0]]
        t = module.ctype(slot.elem)
        for d0, (lo, hi) in ch["fringe"].items():
            if (lo, hi) != (0, 0):
                wn = f"w{ch['id']}_{d0 - ch['dmin']}"
                L.append(f"        {t} {wn}[V + {hi - lo}]; xgb::window_from_body<{t}, V, {lo}, {hi}>("
                         f"p.{slot.field} + (base + ({d0}LL) * S0 + {mid_text(key[1])}), {reg(ch, d0)}, edge_l, edge_r, {wn});")
    for s in g.slots:
        if s.written:
            L.append(f"        {module.ctype(s.elem)} o_{s.field}[V]; unsigned wr_{s.field} = 0u;")
    L.append("#pragma unroll")
    L.append("        for (int v = 0; v < V; ++v) {")
    body: list = []
    hoist: dict = {}
    _emit_statements(g, module, tap, body, "v", hoist)
    L.extend("    " + b for b in body)
    L.append("        }")
    for s in g.slots:
        if s.written:
            L.append(f"        if (act) xgb::st_pred<{module.ctype(s.elem)}, V>(p.{s.field} + base, o_{s.field}, wr_{s.field});")
    for ch in chains.values():
        for d0 in range(ch["dmin"], ch["top"]):
            L.append(f"#pragma unroll\n        for (int v = 0; v < V; ++v) {reg(ch, d0)}[v] = {reg(ch, d0 + 1)}[v];")
    L.append("        base += S0;")
    L.append("    }")
    L.append("    }")
    L.append("    #undef XGB_ROW")
    L.append("}")
    L[3:3] = hoist_lines(hoist)
    return "\n".join(L) + "\n"


# --------------------------------------------------------------------------- tiled (async pipeline) variant
TILED_SMEM_BUDGET = int(_os.environ.get("XGB_SMEM", "0"))      # 0 = per-dimension default below
TILED_TJ = int(_os.environ.get("XGB_TJ", "8"))
TILED_WX3 = int(_os.environ.get("XGB_WX3", "1"))       # consumer warps side by side along k in 3-D
TILED_NSV = int(_os.environ.get("XGB_NSV", "0"))         # 0 = per-dimension default below
TILED_BATCH = _os.environ.get("XGB_TILED_BATCH", "1") != "0"   # all-interior iterations: loads of every vector first
TILED_BATCH_MAX = int(_os.environ.get("XGB_TILED_BATCH_MAX", "56"))  # ... if the windows of all vectors fit this many fp64 registers
TILED_MINB = int(_os.environ.get("XGB_TILED_MINB", "0"))       # __launch_bounds__ min CTAs per SM (0 = unset)
TILED_WX2 = int(_os.environ.get("XGB_WX2", "0"))         # consumer warps per CTA in 2-D; 0 = by stream count
# measured on B200 (profiles/r1_experiments.md): 2-D likes 2 vectors per thread and ~56 KB rings
# (4 CTAs/SM); 3-D amortises the per-plane bookkeeping better with 4 vectors per thread (256-column
# tiles) and deeper rings (2 CTAs/SM)
_TILED_DEFAULTS = {2: (2, 56 * 1024), 3: (4, 110 * 1024)}


def tiled_config(g: Group):
    """Geometry of the bulk-copy pipeline variant, or None when the group does not
    qualify (needs 2-D/3-D, one element width across all slots, halo that fits)."""
    widths = {s.elem.width_bytes for s in g.slots}
    if len(widths) != 1:
        return None
    esize = widths.pop()
    if esize not in (4, 8):
        return None
    V = 16 // esize
    NSV = TILED_NSV or _TILED_DEFAULTS[g.ndim][0]
    budget = TILED_SMEM_BUDGET or _TILED_DEFAULTS[g.ndim][1]
    dmin = dmax = hj = hk = 0
    for a in g.stmts:
        for ld in a.sweep.loads:
            off = ld.space_offset
            dmin, dmax = min(dmin, off[0]), max(dmax, off[0])
            if g.ndim == 3:
                hj = max(hj, abs(off[1]))
            hk = max(hk, abs(off[-1]))
    hk = -(-hk // V) * V                       # keep shared rows 16-byte aligned
    if g.ndim == 2:
        # consumer warps side by side.  Measured on 8192^2 / 16384^2 fp64: kernels that stream two or more
        # levels (cavity's pressure / velocity sweeps) run 8 % faster with 4 warps per CTA (more, smaller
        # barrier groups per SM); the single-stream 5-point sweep is 6 % faster with 8
        wx2 = TILED_WX2 or (4 if sum(1 for s in g.slots if s.read) >= 2 else 8)
        ncw, tj, wx = wx2, 1, wx2
    else:
        tj, wx = TILED_TJ, TILED_WX3           # tj rows x wx warps per row
        ncw = tj * wx
    W = wx * 32 * V * NSV
    wp, rp = W + 2 * hk, tj + 2 * hj
    nread = sum(1 for s in g.slots if s.read)
    if nread == 0:
        return None
    stage_bytes = nread * rp * wp * esize
    ns = min(8, max((dmax - dmin) + 3, (budget - 256) // stage_bytes))
    if 256 + ns * stage_bytes > 224 * 1024:            # 227 KB per CTA on sm_100a, 1 KB of it static
        return None
    return {"V": V, "NSV": NSV, "NCW": ncw, "TJ": tj, "WX": wx, "W": W, "HJ": hj, "HK": hk, "WP": wp,
            "RP": rp, "NS": ns, "DMIN": dmin, "DMAX": dmax, "NREAD": nread, "ESIZE": esize,
            "smem": 256 + ns * stage_bytes, "threads": (ncw + 1) * 32}


def _emit_tiled(g: Group, module: ModuleBuilder, c: dict) -> str:
    nd = g.ndim
    V, NSV = c["V"], c["NSV"]
    read_slots = [s for s in g.slots if s.read]
    ridx = {s.index: n for n, s in enumerate(read_slots)}
    T = module.ctype(read_slots[0].elem)

    # windows per (read slot, di, dj): [lo, hi] over dk
    windows: dict = {}

    def need(slot: Slot, off: tuple) -> None:
        key = (slot.index, off[0], off[1] if nd == 3 else 0)
        lo, hi = windows.get(key, (0, 0))
        windows[key] = (min(lo, off[-1]), max(hi, off[-1]))

    for a in g.stmts:
        for ld in a.sweep.loads:
            need(g.slot(ld.variable.name, ld.level), ld.space_offset)
        if g.implicit:
            need(g.slot(a.sweep.grid.name, 0), (0,) * nd)
    wnames = {key: f"w{n}" for n, key in enumerate(windows)}

    def tap(e: ir.Stencil) -> str:
        slot = g.slot(e.variable.name, e.level)
        key = (slot.index, e.space_offset[0], e.space_offset[1] if nd == 3 else 0)
        lo, _ = windows[key]
        return f"{wnames[key]}[v + {e.space_offset[-1] - lo}]"

    name = kernel_name(g, VARIANT_TILED, V)
    bounds = f'{c["threads"]}, {TILED_MINB}' if TILED_MINB else f'{c["threads"]}'
    L = [f'extern "C" __global__ void __launch_bounds__({bounds}) {name}(const __grid_constant__ {g.name}_P p)', "{"]
    L.append(f"    constexpr int V = {V}, NSV = {NSV}, NCW = {c['NCW']}, TJ = {c['TJ']}, WX = {c['WX']}, W = {c['W']};")
    L.append(f"    constexpr int HJ = {c['HJ']}, HK = {c['HK']}, WP = {c['WP']}, RP = {c['RP']}, NS = {c['NS']};")
    L.append(f"    constexpr int DMIN = {c['DMIN']}, DMAX = {c['DMAX']}, DSPAN = DMAX - DMIN, NREAD = {c['NREAD']};")
    L.append(f"    typedef {T} T;")
    L.append("    constexpr int STAGE = RP * WP;                       // elements per (slot, plane)")
    L.append("    extern __shared__ __align__(128) unsigned char xgb_smem[];")
    L.append("    uint64_t *full = reinterpret_cast<uint64_t *>(xgb_smem);")
    L.append("    uint64_t *empty = full + NS;")
    L.append("    T *stages = reinterpret_cast<T *>(xgb_smem + 256);   // [NS][NREAD][RP][WP]")
    L.append("    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;")
    L.append("    const int64_t c0 = (int64_t)blockIdx.x * W;")
    if nd == 3:
        L.append("    const int64_t j0 = (int64_t)blockIdx.y * TJ;")
        L.append("    const int64_t i0 = p.r_lo + (int64_t)blockIdx.z * p.chunk0;")
        L.append("    const int64_t S0 = p.n1 * p.n2;")
    else:
        L.append("    const int64_t j0 = 0;")
        L.append("    const int64_t i0 = p.r_lo + ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * p.chunk0;")
        L.append("    const int64_t S0 = p.cols;")
    L.append("    if (i0 >= p.r_hi) return;")
    L.append("    const int64_t iend = (i0 + p.chunk0 < p.r_hi) ? (i0 + p.chunk0) : p.r_hi;")
    L.append("    const int nout = (int)(iend - i0);                    // output planes of this CTA")
    L.append("    const int planes = nout + DSPAN;                      // input planes this CTA streams")
    L.append("    if (threadIdx.x == 0) {")
    L.append("        for (int s = 0; s < NS; ++s) { xgb::pipe::mbar_init(&full[s], 1); xgb::pipe::mbar_init(&empty[s], NCW); }")
    L.append("        xgb::pipe::fence_barrier_init();")
    L.append("    }")
    L.append("    __syncthreads();")
    # ---------------- producer warp
    L.append("    if (warp == NCW) {")
    L.append("        const T *src[NREAD] = {" + ", ".join(f"p.{s.field}" for s in read_slots) + "};")
    L.append("        int s = 0, eph = 1;                                // stage / parity of its previous release")
    L.append("        for (int t = 0; t < planes; ++t, ++s) {")
    L.append("            if (s == NS) { s = 0; eph ^= 1; }")
    L.append("            if (t >= NS) xgb::pipe::mbar_wait(&empty[s], eph);")
    L.append("            if (lane == 0) xgb::pipe::mbar_expect_tx(&full[s], (uint32_t)(NREAD * RP * WP * sizeof(T)));")
    L.append("            __syncwarp();")
    L.append("            const int64_t plane = i0 + DMIN + t;")
    L.append("            for (int q = lane; q < NREAD * RP; q += 32) {")
    L.append("                const int r = q / RP, jj = q % RP;")
    if nd == 3:
        L.append("                int64_t j = j0 - HJ + jj;")
        L.append("                if (j > p.n1 - 1 + HJ) j = p.n1 - 1 + HJ;")
        L.append("                const T *g = src[r] + (plane * S0 + j * p.n2 + (c0 - HK));")
    else:
        L.append("                const T *g = src[r] + (plane * S0 + (c0 - HK));")
    L.append("                xgb::pipe::bulk_g2s(stages + ((int64_t)(s * NREAD + r) * RP + jj) * WP, g, (uint32_t)(WP * sizeof(T)), &full[s]);")
    L.append("            }")
    L.append("        }")
    L.append("        return;")
    L.append("    }")
    # ---------------- consumer warps
    L.append("    const int ty = warp / WX, wx = warp % WX;")
    L.append("    const int64_t j = j0 + ty;")
    L.append("    int kk[NSV]; bool act[NSV]; int64_t base[NSV];")
    L.append("#pragma unroll")
    L.append("    for (int sv = 0; sv < NSV; ++sv) {")
    L.append("        kk[sv] = (wx * NSV + sv) * (32 * V) + lane * V;")
    L.append("        const int64_t col = c0 + kk[sv];")
    L.append("        act[sv] = (col < p.cols)" + (" && (j < p.n1);" if nd == 3 else ";"))
    L.append("        base[sv] = act[sv] ? (i0 * S0 + " + ("j * p.n2 + " if nd == 3 else "") + "col) : (i0 * S0);")
    L.append("    }")
    # chunk flags ("any boundary point near this warp's span of the row?"), fetched 32 planes at a time: lane l
    # asks for plane it0 + l, the loop broadcasts one lane's answer per plane with a shuffle.  The round trip
    # to global memory is paid once per 32 planes (the first one while the pipeline fills) instead of once per
    # plane inside the dependent chain of every output vector (ncu, round 1: 41 % of all stall samples).
    for m in g.masks:
        L.append(f"    int fb_{m}[NSV];")
    L.append("    const T *srow = stages + (int64_t)(ty + HJ) * WP + HK;   // this warp's row inside a plane")
    L.append("    int fs = 0, fph = 0;                                  // stage / parity of the newest plane")
    L.append("    int ps = 0;                                           // stage of plane o + DMIN")
    L.append("    for (int it = 0; it < nout; ++it) {")
    if g.masks:
        L.append("        if ((it & 31) == 0) {                             // (warp-uniform) flags of the next 32 planes")
        L.append("#pragma unroll")
        L.append("            for (int sv = 0; sv < NSV; ++sv) {")
        L.append("                const int64_t col0 = c0 + (int64_t)(wx * NSV + sv) * (32 * V);")
        L.append("                const bool wact = (col0 < p.cols)" + (" && (j < p.n1)" if nd == 3 else "") + " && (it + lane < nout);")
        L.append("                const int64_t row0 = (i0 + it + lane) * S0" + (" + j * p.n2;" if nd == 3 else ";"))
        L.append("                const int64_t col1 = (col0 + 32 * V < p.cols) ? (col0 + 32 * V) : p.cols;")
        for m in g.masks:
            L.append(f"                fb_{m}[sv] = wact ? xgb::ld_flag_span(p.m_{m}, p.f_{m}, row0 + col0, row0 + col1 - 1) : 0;")
        L.append("            }")
        L.append("        }")
    L.append("        if (it == 0) for (int t = 0; t < DSPAN; ++t) { xgb::pipe::mbar_wait(&full[fs], fph); if (++fs == NS) { fs = 0; fph ^= 1; } }")
    for m in g.masks:
        L.append(f"        int fc_{m}[NSV];")
        L.append(f"#pragma unroll\n        for (int sv = 0; sv < NSV; ++sv) fc_{m}[sv] = __shfl_sync(0xffffffffu, fb_{m}[sv], it & 31);")
    L.append("        xgb::pipe::mbar_wait(&full[fs], fph);")
    L.append("        if (++fs == NS) { fs = 0; fph ^= 1; }")
    # stage base of every input plane this output plane reads (once per plane, not per tap)
    dis = sorted({key[1] for key in windows})
    for di in dis:
        L.append(f"        const T *pl{di - c['DMIN']} = srow + (int64_t)((ps + {di - c['DMIN']} >= NS) ? "
                 f"(ps + {di - c['DMIN']} - NS) : (ps + {di - c['DMIN']})) * (NREAD * RP * WP);")

    def load_windows(ind: str) -> list:
        out = []
        for key, (lo, hi) in windows.items():
            si, di, dj = key
            out.append(f"{ind}T {wnames[key]}[V + {hi - lo}]; xgb::lds_window<T, V, {lo}, {hi}>("
                       f"pl{di - c['DMIN']} + ({ridx[si]} * RP + ({dj})) * WP + kk[sv], {wnames[key]});")
        return out

    fast_body: list = []
    slow_body: list = []
    hoist: dict = {}
    emit = ExprEmitter(module, _ident, tap, hoist, VARIANT_TILED)
    fast_written = []
    for a in g.stmts:
        sw = a.sweep
        sl = g.slot(sw.grid.name, "scratch" if g.implicit else sw.store.level)
        if sw.mask == 0:
            fast_body.append(f"                    o_{sl.field}[v] = {emit(a.value)};")
            if sl not in fast_written:
                fast_written.append(sl)
    _emit_statements(g, module, tap, slow_body, "v", hoist)
    flags_zero = " && ".join(f"fc_{m}[sv] == 0" for m in g.masks) or "true"
    batch_doubles = NSV * sum(V + hi - lo for lo, hi in windows.values()) * (read_slots[0].elem.width_bytes / 8.0)
    # (3-D only: the 2-D one-pass kernels run at 0.96-0.97 of the copy peak as they are, and the multi-statement
    # cavity groups would go from 128 to 194 registers)
    if TILED_BATCH and nd == 3 and NSV > 1 and fast_body and batch_doubles <= TILED_BATCH_MAX:
        # The common iteration -- every vector of this thread active and no boundary point near the warp's spans --
        # runs WITHOUT a branch per vector: all shared-memory windows are requested first, then all vectors are
        # computed, then all are stored, so the loads of NSV vectors overlap instead of each vector paying its own
        # load -> fp64 chain -> store latency in turn (ncu, round 2: consumers were never starved for data -- 1.5 % of
        # stall samples on the "full" barrier -- but only 38 % of issue slots were used at 16 consumer warps per SM).
        def tapb(e: ir.Stencil) -> str:
            slot = g.slot(e.variable.name, e.level)
            key = (slot.index, e.space_offset[0], e.space_offset[1] if nd == 3 else 0)
            lo, _ = windows[key]
            return f"{wnames[key]}b[sv][v + {e.space_offset[-1] - lo}]"

        emitb = ExprEmitter(module, _ident, tapb, hoist, VARIANT_TILED)
        L.append("        bool allfast = true;")
        L.append("#pragma unroll")
        L.append(f"        for (int sv = 0; sv < NSV; ++sv) allfast = allfast && act[sv] && ({flags_zero});")
        L.append("        if (allfast) {")
        for key, (lo, hi) in windows.items():
            L.append(f"            T {wnames[key]}b[NSV][V + {hi - lo}];")
        L.append("#pragma unroll")
        L.append("            for (int sv = 0; sv < NSV; ++sv) {")
        for key, (lo, hi) in windows.items():
            si, di, dj = key
            L.append(f"                xgb::lds_window<T, V, {lo}, {hi}>(pl{di - c['DMIN']} + ({ridx[si]} * RP + ({dj})) * WP + kk[sv], {wnames[key]}b[sv]);")
        L.append("            }")
        for sl in fast_written:
            L.append(f"            {module.ctype(sl.elem)} o_{sl.field}b[NSV][V];")
        L.append("#pragma unroll")
        L.append("            for (int sv = 0; sv < NSV; ++sv) {")
        L.append("#pragma unroll")
        L.append("                for (int v = 0; v < V; ++v) {")
        for a in g.stmts:
            if a.sweep.mask == 0:
                sl = g.slot(a.sweep.grid.name, "scratch" if g.implicit else a.sweep.store.level)
                L.append(f"                    o_{sl.field}b[sv][v] = {emitb(a.value)};")
        L.append("                }")
        L.append("            }")
        L.append("#pragma unroll")
        L.append("            for (int sv = 0; sv < NSV; ++sv) {")
        for sl in fast_written:
            L.append(f"                xgb::st_vec<{module.ctype(sl.elem)}, V>(p.{sl.field} + base[sv], o_{sl.field}b[sv]);")
        L.append("                base[sv] += S0;")
        L.append("            }")
        L.append("        } else")
    L.append("#pragma unroll")
    L.append("        for (int sv = 0; sv < NSV; ++sv) {")
    L.append("            if (act[sv]) {")
    L.append(f"                if ({flags_zero}) {{            // no boundary point in this 128-point chunk")
    L.extend(load_windows("                    "))
    for sl in fast_written:
        L.append(f"                    {module.ctype(sl.elem)} o_{sl.field}[V];")
    L.append("#pragma unroll")
    L.append("                    for (int v = 0; v < V; ++v) {")
    L.extend(fast_body)
    L.append("                    }")
    for sl in fast_written:
        L.append(f"                    xgb::st_vec<{module.ctype(sl.elem)}, V>(p.{sl.field} + base[sv], o_{sl.field});")
    L.append("                } else {")
    for m in g.masks:
        L.append(f"                    int m_{m}[V]; xgb::ld_mask_flagged<V>(p.m_{m}, fc_{m}[sv], base[sv], m_{m});")
    L.extend(load_windows("                    "))
    for s_ in g.slots:
        if s_.written:
            L.append(f"                    {module.ctype(s_.elem)} o_{s_.field}[V]; unsigned wr_{s_.field} = 0u;")
    L.append("#pragma unroll")
    L.append("                    for (int v = 0; v < V; ++v) {")
    L.extend("            " + b for b in slow_body)
    L.append("                    }")
    for s_ in g.slots:
        if s_.written:
            L.append(f"                    xgb::st_pred<{module.ctype(s_.elem)}, V>(p.{s_.field} + base[sv], o_{s_.field}, wr_{s_.field});")
    L.append("                }")
    L.append("            }")
    L.append("            base[sv] += S0;")
    L.append("        }")
    L.append("        __syncwarp();")
    L.append("        if (lane == 0) xgb::pipe::mbar_arrive(&empty[ps]);       // plane o + DMIN is dead")
    L.append("        ps = (ps + 1 == NS) ? 0 : ps + 1;")
    L.append("    }")
    L.append("}")
    at = next(n for n, line in enumerate(L) if line.startswith("    const int ty = warp / WX"))
    L[at:at] = hoist_lines(hoist)          # consumer warps only; the producer never needs them
    return "\n".join(L) + "\n"


# --------------------------------------------------------------------------- multi-step (temporal blocking) variant
MULTISTEP_T = int(_os.environ.get("XGB_MS_T", "64"))
MULTISTEP_S = int(_os.environ.get("XGB_MS_S", "4"))      # register sub-steps per shared-memory round trip
MULTISTEP_P = int(_os.environ.get("XGB_MS_P", "8"))      # points per thread in the register path


def multistep_config(g: Group, short: bool = False):
    """1-D groups that read only the previous level of the ONE grid they update can run T
    time steps per launch from shared memory.  ``short``: the variant for
    runs of at most T/2 steps -- half the halo and half the window (256 threads, ~43 KB of shared memory, so
    five CTAs per SM instead of two): a launch's fixed cost is its load and store phases, which only
    overlap ACROSS resident CTAs."""
    if g.ndim != 1 or g.implicit or g.sparse:
        return None
    if not any(st.sweep.mask == 0 for st in g.stmts):
        # the register path of all-zero-mask tiles advances x{s} from x{s-1}; without a mask-0 statement
        # those points are never written and must keep the value from TWO steps back
        return None
    grids = {s.grid for s in g.slots}
    if len(grids) != 1:
        return None
    levels = {(s.level, s.read, s.written) for s in g.slots}
    if levels != {(0, False, True), (1, True, False)}:
        return None
    elem = g.slots[0].elem
    if isinstance(elem, (Structure, Boolean)) or elem.width_bytes not in (4, 8):
        return None
    h = max(1, g.halo_last)
    T = MULTISTEP_T
    while T * h > 64:                 # halo must stay inside the level's zero slack
        T //= 2
    if short:
        T //= 2
    S = MULTISTEP_S                   # time steps advanced in registers per shared-memory round trip
    T -= T % (2 * S) if T >= 2 * S else T % 2
    if T < 4:
        return None
    if T % S or S % 2:
        S = 2
    if (S * h) % (16 // elem.width_bytes):
        return None                   # register windows must start on a 16-byte boundary
    NT, P = (256 if short else 512), MULTISTEP_P
    H = T * h
    L = NT * P                        # window = one P-point body per thread
    if L <= 4 * H:
        return None
    W = L - 2 * H
    PAD = S * h + (-(S * h)) % (16 // elem.width_bytes)      # keeps bodies 16-byte aligned
    V = 16 // elem.width_bytes
    if P & (P - 1) or P % V:
        return None
    swlen, marg = L + (L // P) * V, 2 * PAD + 2 * V
    return {"T": T, "S": S, "P": P, "W": W, "H": H, "h": h, "L": L, "PAD": PAD, "threads": NT, "V": V,
            "smem": 2 * (swlen + 2 * marg) * elem.width_bytes + L + 64}


def _emit_multistep(g: Group, module: ModuleBuilder, c: dict, tail: bool = False, short: bool = False) -> str:
    """T steps per launch (``tail``: the same kernel with the step count read from ``p.opt0`` --
    a multiple of S below T -- for the remainder of a deferred run; the window keeps its T*h halo).  Two shared-memory buffers start as copies of the two ring levels
    (now / previous) of an L = W+2H window; every step writes the *older* buffer where a
    statement's mask matches -- exactly what T ticks + T sweeps do to the ring (unwritten points
    keep the value from two steps back) -- and the window of valid points shrinks by
    h per step.  Tiles without boundary points take the register path: each thread loads its P
    points plus an S*h fringe, advances S steps in registers and writes back the last two.
    Shared memory is laid out with one 16-byte pad per P-point body (XSW) so that the 128-bit
    accesses of consecutive threads fall on different banks."""
    elem = g.slots[0].elem
    T_ = module.ctype(elem)
    gname = g.slots[0].grid
    S, P, h, V = c["S"], c["P"], c["h"], c["V"]
    psh = P.bit_length() - 1

    def slow_tap(e: ir.Stencil) -> str:
        return f"cur[XSW(q + ({e.space_offset[-1]}))]"

    hoist: dict = {}
    slow_emit = ExprEmitter(module, _ident, slow_tap, hoist, VARIANT_MULTISTEP)
    slow = [f"if (m == {a.sweep.mask}) nxt[XSW(q)] = {slow_emit(a.value)};" for a in g.stmts]
    mask0 = [a for a in g.stmts if a.sweep.mask == 0]

    # register path: level s (1..S) holds N0 - 2*s*h values; x{s}[i] is point (first - (S-s)*h + i)
    N0 = P + 2 * S * h
    reg_lines = []
    for s_ in range(1, S + 1):
        n_out = N0 - 2 * s_ * h
        reg_lines.append(f"            E x{s_}[{n_out}];")
        reg_lines.append("#pragma unroll")
        reg_lines.append(f"            for (int i = 0; i < {n_out}; ++i) {{")

        def reg_tap(e: ir.Stencil, lvl=s_ - 1) -> str:
            return f"x{lvl}[i + ({h + e.space_offset[-1]})]"

        emit = ExprEmitter(module, _ident, reg_tap, hoist, VARIANT_MULTISTEP)
        if mask0:
            reg_lines.append(f"                x{s_}[i] = {emit(mask0[-1].value)};")
        else:
            reg_lines.append(f"                x{s_}[i] = x{s_ - 1}[i + {h}];")
        reg_lines.append("            }")

    name = kernel_name(g, VARIANT_MULTISTEP_SHORT if short else VARIANT_MULTISTEP_TAIL if tail else VARIANT_MULTISTEP, 1)
    nsteps = "(int)p.opt0" if tail else "T"
    L = [f'extern "C" __global__ void __launch_bounds__({c["threads"]}) {name}(const __grid_constant__ {g.name}_P p)', "{"]
    L.append(f"    constexpr int T = {c['T']}, S = {S}, P = {P}, PSH = {psh}, HS = {h}, H = {c['H']}, W = {c['W']}, L = {c['L']}, "
             f"PAD = {c['PAD']}, NT = {c['threads']}, V = {V};")
    L.append("    #define XSW(i) ((i) + (((i) >> PSH) * V))          /* one V-element pad per P-point body */")
    L.append("    constexpr int MARG = 2 * PAD + 2 * V, SWLEN = L + (L >> PSH) * V;")
    L.append(f"    typedef {T_} E;")
    L.extend(hoist_lines(hoist))
    L.append("    extern __shared__ __align__(16) unsigned char xgb_smem[];")
    L.append("    E *b0 = reinterpret_cast<E *>(xgb_smem) + MARG;      // starts as the current level  (u^n)")
    L.append("    E *b1 = b0 + SWLEN + 2 * MARG;                       // starts as the previous level (u^{n-1})")
    L.append("    uint8_t *sm = reinterpret_cast<uint8_t *>(b1 + SWLEN + MARG);")
    L.append("    const E *now = static_cast<const E *>(p.aux0), *prev = static_cast<const E *>(p.aux1);")
    L.append("    E *out0 = static_cast<E *>(p.aux2), *out1 = static_cast<E *>(p.aux3);")
    L.append("    const int64_t g0 = (int64_t)blockIdx.x * W - H;   // global index of window element 0")
    L.append("    if (threadIdx.x < 2 * PAD) {                       // fringes outside the window: never valid, keep finite")
    L.append("        const int k = threadIdx.x < PAD ? (int)threadIdx.x - PAD : L + (int)threadIdx.x - PAD;")
    L.append("        b0[XSW(k)] = E(0); b1[XSW(k)] = E(0);")
    L.append("    }")
    L.append("    for (int q = threadIdx.x * V; q < L; q += NT * V) {")
    L.append("        E a[V], b[V];")
    L.append("        const int64_t gi = g0 + q;                 // stay inside the level's zero slack")
    L.append("        // p.count != 0: every mask value present has a statement, so each point is rewritten every step and the")
    L.append("        // previous level is never observed (points outside the grid are zero in both levels): it is not loaded")
    L.append("        for (int v = 0; v < V; ++v) b[v] = E(0);")
    L.append("        if (gi >= -64 && gi + V <= p.n0 + 64) { xgb::ld_vec<E, V>(now + gi, a); if (p.count == 0) xgb::ld_vec<E, V>(prev + gi, b); }")
    L.append("        else { for (int v = 0; v < V; ++v) a[v] = E(0); }")
    L.append("        xgb::st_vec<E, V>(b0 + XSW(q), a); xgb::st_vec<E, V>(b1 + XSW(q), b);")
    L.append("    }")
    L.append("    // A window that lies inside the grid and whose chunk flags are all clear has no boundary point: it")
    L.append("    // takes the register path without fetching a mask byte (L/128 + 1 flag bytes instead of L mask bytes).")
    L.append("    int any = 1;")
    L.append("    if (g0 >= 0 && g0 + L <= p.n0) {")
    L.append("        any = 0;")
    L.append(f"        if (p.m_{gname} != nullptr) {{")
    L.append(f"            if (p.f_{gname} == nullptr) any = 1;")
    L.append(f"            else if ((int)threadIdx.x <= (L >> XGB_CHUNK_SHIFT)) {{")
    L.append(f"                const int64_t ch = (g0 >> XGB_CHUNK_SHIFT) + threadIdx.x;")
    L.append(f"                if ((ch << XGB_CHUNK_SHIFT) < g0 + L) any = p.f_{gname}[ch];")
    L.append("            }")
    L.append("        }")
    L.append("    }")
    L.append("    int masked = __syncthreads_or(any);")
    L.append("    if (masked) {")
    L.append("        any = 0;")
    L.append("        for (int q = threadIdx.x; q < L; q += NT) {")
    L.append("            const int64_t gi = g0 + q;")
    L.append("            int m = 255;                              // outside the grid: never updated")
    L.append(f"            if ((gi >= 0 || p.open_lo) && (gi < p.n0 || p.open_hi)) m = (p.m_{gname} != nullptr) ? p.m_{gname}[gi] : 0;")
    L.append("            sm[q] = (uint8_t)m; any |= m;")
    L.append("        }")
    L.append("        masked = __syncthreads_or(any);")
    L.append("    }")
    L.append("    if (!masked) {")
    L.append("        // register path: S steps per round; b0 always receives the newest level (S is even)")
    L.append("        const int first = threadIdx.x * P;")
    L.append(f"        for (int round = 0; round < {nsteps} / S; ++round) {{")
    L.append(f"            E x0[{N0}];")
    L.append("#pragma unroll")
    L.append(f"            for (int i = 0; i < {N0}; i += V) xgb::ld_vec<E, V>(b0 + XSW(first - S * HS + i), *reinterpret_cast<E (*)[V]>(&x0[i]));")
    L.extend(reg_lines)
    L.append("            __syncthreads();                          // everyone has read b0 before it is overwritten")
    L.append("#pragma unroll")
    L.append(f"            for (int i = 0; i < P; i += V) {{")
    L.append(f"                E lo_[V], hi_[V];")
    L.append("#pragma unroll")
    L.append(f"                for (int v = 0; v < V; ++v) {{ hi_[v] = x{S}[i + v]; lo_[v] = x{S - 1}[i + v + HS]; }}")
    L.append(f"                xgb::st_vec<E, V>(b0 + XSW(first + i), hi_);")
    L.append(f"                xgb::st_vec<E, V>(b1 + XSW(first + i), lo_);")
    L.append("            }")
    L.append("            __syncthreads();")
    L.append("        }")
    L.append("    } else {")
    L.append("        E *cur = b0, *nxt = b1;")
    L.append(f"        for (int s = 1; s <= {nsteps}; ++s) {{")
    L.append("            for (int q = s * HS + threadIdx.x; q < L - s * HS; q += NT) {")
    L.append("                const int m = sm[q];")
    L.extend("                " + x for x in slow)
    L.append("            }")
    L.append("            __syncthreads();")
    L.append("            E *t = cur; cur = nxt; nxt = t;")
    L.append("        }")
    L.append("    }")
    L.append("    // the step count is even: b0 holds the newest level, b1 the one before it")
    L.append("    for (int q = H + threadIdx.x * V; q < H + W; q += NT * V) {")
    L.append("        const int64_t gi = g0 + q;")
    L.append("        E a[V], b[V];")
    L.append("        xgb::ld_vec<E, V>(b0 + XSW(q), a); xgb::ld_vec<E, V>(b1 + XSW(q), b);")
    L.append("        if (gi + V <= p.n0) {")
    L.append("            xgb::st_vec<E, V>(out0 + gi, a); xgb::st_vec<E, V>(out1 + gi, b);")
    L.append("        } else {")
    L.append("            for (int v = 0; v < V; ++v) if (gi + v < p.n0) { out0[gi + v] = a[v]; out1[gi + v] = b[v]; }")
    L.append("        }")
    L.append("    }")
    L.append("    #undef XSW")
    L.append("}")
    return "\n".join(L) + "\n"


# --------------------------------------------------------------------------- tiled2: two time steps per pass (2-D)
def tiled2_config(g: Group):
    """2-D groups that read only the previous level of the ONE grid they update can advance two
    time steps per pass: rows stream through the bulk-copy pipeline once, the first step's rows
    live in a small shared-memory ring, the second step is written to HBM."""
    if g.ndim == 3:
        return tiled2_config_3d(g)
    if g.ndim != 2 or g.implicit or g.sparse or g.tiled is None:
        return None
    if {s.grid for s in g.slots} != {g.slots[0].grid}:
        return None
    if {(s.level, s.read, s.written) for s in g.slots} != {(0, False, True), (1, True, False)}:
        return None
    elem = g.slots[0].elem
    if isinstance(elem, (Structure, Boolean)) or elem.width_bytes != 8:
        return None
    V, NSV, NCW = 2, int(_os.environ.get("XGB_T2_NSV", "2")), int(_os.environ.get("XGB_T2_NCW", "4"))
    dmin = dmax = hk = 0
    for a in g.stmts:
        for ld in a.sweep.loads:
            dmin, dmax = min(dmin, ld.space_offset[0]), max(dmax, ld.space_offset[0])
            hk = max(hk, abs(ld.space_offset[-1]))
    if dmax - dmin > 4 or hk > 4:
        return None
    W = NCW * 32 * V * NSV
    hkm = -(-hk // V) * V                 # halo of the middle rows (rounded up to whole vectors)
    # halo of the input rows: EVERY middle element that is computed -- the rounded-up ones included -- must find
    # its taps inside the loaded row.  (Round 1 used ceil(2*hk / V) * V, which equals hkm for hk = 1: the outermost
    # middle vector then read one element past its input row -- into the next ring slot, or for the last slot into
    # the first element of the middle ring.  The value was never used, but compute-sanitizer racecheck rightly
    # reported the stray read against the write of that element, profiles/r2_sanitize.md.)
    hk0 = -(-(hkm + hk) // V) * V
    wp0, wpm = W + 2 * hk0, W + 2 * hkm
    dspan = dmax - dmin
    mr = dspan + 2                        # middle-row ring
    ns = max(dspan + 2, int(_os.environ.get("XGB_T2_NS", "5")))      # input-row ring (3 CTAs/SM at 5)
    smem = 256 + ns * wp0 * 8 + mr * wpm * 8
    return {"V": V, "NSV": NSV, "NCW": NCW, "W": W, "HK": hk, "HKM": hkm, "HK0": hk0, "WP0": wp0, "WPM": wpm,
            "DMIN": dmin, "DMAX": dmax, "MR": mr, "NS": ns, "smem": smem, "threads": (NCW + 1) * 32,
            "ghost": 2 * max(abs(dmin), abs(dmax), 1)}


def _emit_tiled2(g: Group, module: ModuleBuilder, c: dict) -> str:
    """Time-skewed two-step sweep.  For output row o (step 2) the middle rows o+DMIN..o+DMAX
    (step 1) are needed, each of which needs input rows q+DMIN..q+DMAX.  Input rows stream in
    order; when row r lands the consumers compute middle row q = r - DMAX into the middle ring,
    meet at a named barrier, and then compute output row o = q - DMAX from the middle ring.
    Every shared-memory position is identified with its LINEAR grid index, so columns outside a
    row read / produce exactly what the step-at-a-time sweeps do (linear addressing, F10);
    positions outside the array are ghost zeros.  The launcher guarantees that every mask value
    present in the grid has a statement, so every cell is rewritten each step and the level two
    steps back is never needed."""
    T = module.ctype(g.slots[0].elem)
    gname = g.slots[0].grid
    V = c["V"]
    hoist: dict = {}
    # row windows per axis-0 offset: [lo, hi] over the contiguous-axis offsets
    win: dict = {}
    for a in g.stmts:
        for ld in a.sweep.loads:
            lo, hi = win.get(ld.space_offset[0], (0, 0))
            win[ld.space_offset[0]] = (min(lo, ld.space_offset[-1]), max(hi, ld.space_offset[-1]))

    def tap(e: ir.Stencil) -> str:
        lo, _ = win[e.space_offset[0]]
        return f"w{e.space_offset[0] - c['DMIN']}[v + {e.space_offset[-1] - lo}]"

    emit = ExprEmitter(module, _ident, tap, hoist, VARIANT_TILED2)
    slow = [f"if (m[v] == {a.sweep.mask}) val[v] = {emit(a.value)};" for a in g.stmts]
    fast = [f"val[v] = {emit(a.value)};" for a in g.stmts if a.sweep.mask == 0][-1:]

    def windows(prefix: str, ind: str) -> list:
        return [f"{ind}T w{d0 - c['DMIN']}[V + {hi - lo}]; xgb::lds_window<T, V, {lo}, {hi}>({prefix}{d0 - c['DMIN']} + kk, w{d0 - c['DMIN']});"
                for d0, (lo, hi) in win.items()]

    def body(prefix: str, ind: str) -> list:
        out = [f"{ind}T val[V];", f"{ind}if (fl == 0) {{"]
        out += windows(prefix, ind + "    ")
        out += ["#pragma unroll", f"{ind}    for (int v = 0; v < V; ++v) {{ " + " ".join(fast) + " }", f"{ind}}} else {{"]
        out += [f"{ind}    int m[V]; xgb::ld_mask_flagged<V>(p.m_{gname}, fl, lin, m);"]
        out += windows(prefix, ind + "    ")
        out += ["#pragma unroll", f"{ind}    for (int v = 0; v < V; ++v) {{ val[v] = T(0); " + " ".join(slow) + " }", f"{ind}}}"]
        return out

    name = kernel_name(g, VARIANT_TILED2, V)
    nit1 = -(-c["WPM"] // (c["NCW"] * 32 * V))
    nit2 = c["W"] // (c["NCW"] * 32 * V)
    L = [f'extern "C" __global__ void __launch_bounds__({c["threads"]}) {name}(const __grid_constant__ {g.name}_P p)', "{"]
    L.append(f"    constexpr int V = {V}, NCW = {c['NCW']}, NT = NCW * 32, W = {c['W']}, HKM = {c['HKM']}, HK0 = {c['HK0']}, WP0 = {c['WP0']}, "
             f"WPM = {c['WPM']}, NS = {c['NS']}, MR = {c['MR']}, DMIN = {c['DMIN']}, DMAX = {c['DMAX']}, DSPAN = DMAX - DMIN, "
             f"NIT1 = {nit1}, NIT2 = {nit2};")
    L.append(f"    typedef {T} T;")
    L.append("    extern __shared__ __align__(128) unsigned char xgb_smem[];")
    L.append("    uint64_t *full = reinterpret_cast<uint64_t *>(xgb_smem);")
    L.append("    uint64_t *empty = full + NS;")
    L.append("    T *stages = reinterpret_cast<T *>(xgb_smem + 256);        // [NS][WP0]   input rows (u^n)")
    L.append("    T *mids = stages + NS * WP0;                              // [MR][WPM]   middle rows (u^{n+1})")
    L.append("    const T *src = static_cast<const T *>(p.aux0);")
    L.append("    T *out2 = static_cast<T *>(p.aux1), *out1 = static_cast<T *>(p.aux2);")
    L.append("    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;")
    L.append("    const int64_t c0 = (int64_t)blockIdx.x * W;")
    L.append("    const int64_t i0 = p.r_lo + ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * p.chunk0;")
    L.append("    if (i0 >= p.r_hi) return;")
    L.append("    const int64_t iend = (i0 + p.chunk0 < p.r_hi) ? (i0 + p.chunk0) : p.r_hi;")
    L.append("    const int64_t S0 = p.cols, total = p.n0 * p.cols;")
    L.append("    const int nrows = (int)(iend - i0) + 2 * DSPAN;              // input rows this CTA streams")
    L.append("    if (threadIdx.x == 0) {")
    L.append("        for (int s = 0; s < NS; ++s) { xgb::pipe::mbar_init(&full[s], 1); xgb::pipe::mbar_init(&empty[s], NCW); }")
    L.append("        xgb::pipe::fence_barrier_init();")
    L.append("    }")
    L.append("    __syncthreads();")
    L.append("    if (warp == NCW) {                                         // producer: one bulk copy per input row")
    L.append("        int s = 0, eph = 1;")
    L.append("        const T *row = src + ((i0 + 2 * DMIN) * S0 + (c0 - HK0));")
    L.append("        for (int t = 0; t < nrows; ++t, ++s, row += S0) {")
    L.append("            if (s == NS) { s = 0; eph ^= 1; }")
    L.append("            if (t >= NS) xgb::pipe::mbar_wait(&empty[s], eph);")
    L.append("            if (lane == 0) {")
    L.append("                xgb::pipe::mbar_expect_tx(&full[s], (uint32_t)(WP0 * sizeof(T)));")
    L.append("                xgb::pipe::bulk_g2s(stages + s * WP0, row, (uint32_t)(WP0 * sizeof(T)), &full[s]);")
    L.append("            }")
    L.append("            __syncwarp();")
    L.append("        }")
    L.append("        return;")
    L.append("    }")
    L.extend(hoist_lines(hoist))
    L.append("    const int tid = warp * 32 + lane;                           // consumer thread 0..255")
    L.append("    int fs = 0, fph = 0, rs = 0, ms = 0;                        // newest input stage / parity; oldest live stage; middle slot")
    L.append("    int64_t lin_mid = (i0 + DMIN) * S0 + (c0 - HKM);            // linear index of the next middle row's first column")
    L.append("    int64_t lin_out = i0 * S0 + c0;")
    L.append("    for (int t = 0; t < nrows; ++t) {")
    L.append("        xgb::pipe::mbar_wait(&full[fs], fph);")
    L.append("        if (t >= DSPAN) {")
    L.append("            // ---- step 1: middle row q = i0 + DMIN + (t - DSPAN) over columns [c0 - HKM, c0 + W + HKM)")
    for d0 in win:
        k = c["DMAX"] - d0
        L.append(f"            const T *in{d0 - c['DMIN']} = stages + ((fs >= {k}) ? (fs - {k}) : (fs - {k} + NS)) * WP0 + (HK0 - HKM);")
    L.append("            T *mrow = mids + ms * WPM;")
    L.append("            const bool row_in = (lin_mid >= 0) && (lin_mid + WPM <= total);")
    L.append("            int fls[NIT1];")
    L.append("#pragma unroll")
    L.append("            for (int i = 0; i < NIT1; ++i) {                       // flag bytes first: their latency overlaps")
    L.append("                const int kk = (tid + i * NT) * V;")
    L.append("                const int64_t lin = lin_mid + kk;")
    L.append("                fls[i] = -1;")
    L.append(f"                if (kk < WPM && (row_in || (lin >= 0 && lin + V <= total))) fls[i] = xgb::ld_flag(p.m_{gname}, p.f_{gname}, lin);")
    L.append("            }")
    L.append("#pragma unroll")
    L.append("            for (int i = 0; i < NIT1; ++i) {")
    L.append("                const int kk = (tid + i * NT) * V;")
    L.append("                if (kk < WPM) {")
    L.append("                    const int fl = fls[i];")
    L.append("                    const int64_t lin = lin_mid + kk;")
    L.append("                    if (fl >= 0) {")
    L.extend(body("in", "                        "))
    L.append("                        xgb::st_vec<T, V>(mrow + kk, val);")
    L.append("                    } else {")
    L.append("                        T z[V];")
    L.append("#pragma unroll")
    L.append("                        for (int v = 0; v < V; ++v) z[v] = T(0);         // outside the array: ghost zeros")
    L.append("                        xgb::st_vec<T, V>(mrow + kk, z);")
    L.append("                    }")
    L.append("                }")
    L.append("            }")
    L.append("            lin_mid += S0;")
    L.append("            asm volatile(\"bar.sync 1, %0;\" :: \"n\"(NT) : \"memory\");      // middle row complete")
    L.append("        }")
    L.append("        if (t >= 2 * DSPAN) {")
    L.append("            // ---- step 2: output row o = i0 + (t - 2*DSPAN) over the CTA's own columns; middle row o+d0 sits")
    L.append("            //      (DMAX - d0) slots behind the one just written")
    for d0 in win:
        k = c["DMAX"] - d0
        L.append(f"            const T *md{d0 - c['DMIN']} = mids + ((ms >= {k}) ? (ms - {k}) : (ms - {k} + MR)) * WPM + HKM;")
    if 0 not in win:
        k = c["DMAX"]
        L.append(f"            const T *md{0 - c['DMIN']} = mids + ((ms >= {k}) ? (ms - {k}) : (ms - {k} + MR)) * WPM + HKM;")
    L.append("            int fls[NIT2];")
    L.append("#pragma unroll")
    L.append("            for (int i = 0; i < NIT2; ++i) {")
    L.append("                const int kk = (tid + i * NT) * V;")
    L.append("                fls[i] = -1;")
    L.append(f"                if (c0 + kk < p.cols) fls[i] = xgb::ld_flag(p.m_{gname}, p.f_{gname}, lin_out + kk);")
    L.append("            }")
    L.append("#pragma unroll")
    L.append("            for (int i = 0; i < NIT2; ++i) {")
    L.append("                const int kk = (tid + i * NT) * V;")
    L.append("                const int fl = fls[i];")
    L.append("                if (fl >= 0) {")
    L.append("                    const int64_t lin = lin_out + kk;")
    L.extend(body("md", "                    "))
    L.append("                    xgb::st_vec<T, V>(out2 + lin, val);")
    L.append("                    if (p.opt0) {                                     // last pass of a batch: u^{n+1} is observable")
    L.append(f"                        T mv[V]; xgb::ld_vec<T, V>(md{0 - c['DMIN']} + kk, mv); xgb::st_vec<T, V>(out1 + lin, mv);")
    L.append("                    }")
    L.append("                }")
    L.append("            }")
    L.append("            lin_out += S0;")
    L.append("        }")
    L.append("        if (t >= DSPAN) {                                           // input row q + DMIN is dead")
    L.append("            __syncwarp();")
    L.append("            if (lane == 0) xgb::pipe::mbar_arrive(&empty[rs]);")
    L.append("            rs = (rs + 1 == NS) ? 0 : rs + 1;")
    L.append("            ms = (ms + 1 == MR) ? 0 : ms + 1;")
    L.append("        }")
    L.append("        if (++fs == NS) { fs = 0; fph ^= 1; }")
    L.append("    }")
    L.append("}")
    return "\n".join(L) + "\n"


# --------------------------------------------------------------------------- tiled2 in 3-D
# opt-in: measured slower than the single-step tiled kernel on heat3d (instruction-bound, 0.67 vs
# 0.83 of the HBM roofline; profiles/r1_experiments.md)
TILED2_3D = _os.environ.get("XGB_TILED2_3D", "0") != "0"


def tiled2_config_3d(g: Group):
    """Same two-steps-per-pass scheme with (j, k) planes: a CTA owns a TJ x W tile, the middle
    planes cover the tile plus one stencil halo, the input planes plus two."""
    if not TILED2_3D or g.implicit or g.sparse or g.tiled is None:
        return None
    if {s.grid for s in g.slots} != {g.slots[0].grid}:
        return None
    if {(s.level, s.read, s.written) for s in g.slots} != {(0, False, True), (1, True, False)}:
        return None
    elem = g.slots[0].elem
    if isinstance(elem, (Structure, Boolean)) or elem.width_bytes != 8:
        return None
    V, NCW = 2, 8
    dmin = dmax = hj = hk = 0
    for a in g.stmts:
        for ld in a.sweep.loads:
            off = ld.space_offset
            dmin, dmax = min(dmin, off[0]), max(dmax, off[0])
            hj, hk = max(hj, abs(off[1])), max(hk, abs(off[2]))
    if dmax - dmin > 2 or hj > 2 or hk > 2:
        return None
    TJ, W = int(_os.environ.get("XGB_T2_TJ", "8")), int(_os.environ.get("XGB_T2_W", "128"))
    hkm = -(-hk // V) * V
    hk0 = -(-(hkm + hk) // V) * V
    wpm, wp0 = W + 2 * hkm, W + 2 * hk0
    rpm, rp0 = TJ + 2 * hj, TJ + 4 * hj
    dspan = dmax - dmin
    mr = dspan + 2
    ns = max(dspan + 2, int(_os.environ.get("XGB_T2_NS3", "5")))
    smem = 256 + ns * rp0 * wp0 * 8 + mr * rpm * wpm * 8
    if smem > 200 * 1024:
        return None
    return {"V": V, "NCW": NCW, "TJ": TJ, "W": W, "HJ": hj, "HK": hk, "HKM": hkm, "HK0": hk0, "WPM": wpm, "WP0": wp0,
            "RPM": rpm, "RP0": rp0, "DMIN": dmin, "DMAX": dmax, "MR": mr, "NS": ns, "smem": smem,
            "threads": (NCW + 1) * 32, "ghost": 2 * max(abs(dmin), abs(dmax), 1)}


def _emit_tiled2_3d(g: Group, module: ModuleBuilder, c: dict) -> str:
    """3-D twin of _emit_tiled2: stages and middle slots hold (j, k) planes; the consumers walk
    the plane regions as flattened vector lists.  Shared-memory positions keep their linear grid
    index (plane * S0 + j * n2 + k with j, k possibly outside their ranges), so wrap reads behave
    exactly as in the step-at-a-time sweeps."""
    T = module.ctype(g.slots[0].elem)
    gname = g.slots[0].grid
    V = c["V"]
    hoist: dict = {}
    win: dict = {}          # (d0, dj) -> [lo, hi] over dk
    for a in g.stmts:
        for ld in a.sweep.loads:
            key = (ld.space_offset[0], ld.space_offset[1])
            lo, hi = win.get(key, (0, 0))
            win[key] = (min(lo, ld.space_offset[2]), max(hi, ld.space_offset[2]))
    wname = {key: f"w{n}" for n, key in enumerate(win)}

    def tap(e: ir.Stencil) -> str:
        key = (e.space_offset[0], e.space_offset[1])
        return f"{wname[key]}[v + {e.space_offset[2] - win[key][0]}]"

    emit = ExprEmitter(module, _ident, tap, hoist, VARIANT_TILED2)
    slow = [f"if (m[v] == {a.sweep.mask}) val[v] = {emit(a.value)};" for a in g.stmts]
    fast = [f"val[v] = {emit(a.value)};" for a in g.stmts if a.sweep.mask == 0][-1:]

    def windows(prefix: str, pitch: str, ind: str) -> list:
        return [f"{ind}T {wname[k]}[V + {hi - lo}]; xgb::lds_window<T, V, {lo}, {hi}>({prefix}{k[0] - c['DMIN']} + (jj + ({k[1]})) * {pitch} + kk, {wname[k]});"
                for k, (lo, hi) in win.items()]

    def body(prefix: str, pitch: str, ind: str) -> list:
        out = [f"{ind}T val[V];", f"{ind}if (fl == 0) {{"]
        out += windows(prefix, pitch, ind + "    ")
        out += ["#pragma unroll", f"{ind}    for (int v = 0; v < V; ++v) {{ " + " ".join(fast) + " }", f"{ind}}} else {{"]
        out += [f"{ind}    int m[V]; xgb::ld_mask_flagged<V>(p.m_{gname}, fl, lin, m);"]
        out += windows(prefix, pitch, ind + "    ")
        out += ["#pragma unroll", f"{ind}    for (int v = 0; v < V; ++v) {{ val[v] = T(0); " + " ".join(slow) + " }", f"{ind}}}"]
        return out

    d0s = sorted({k[0] for k in win} | {0})
    name = kernel_name(g, VARIANT_TILED2, V)
    nv1 = c["RPM"] * (c["WPM"] // V)
    nv2 = c["TJ"] * (c["W"] // V)
    NT = c["NCW"] * 32
    L = [f'extern "C" __global__ void __launch_bounds__({c["threads"]}) {name}(const __grid_constant__ {g.name}_P p)', "{"]
    L.append(f"    constexpr int V = {V}, NCW = {c['NCW']}, NT = NCW * 32, TJ = {c['TJ']}, W = {c['W']}, HJ = {c['HJ']}, HKM = {c['HKM']}, "
             f"HK0 = {c['HK0']}, WPM = {c['WPM']}, WP0 = {c['WP0']}, RPM = {c['RPM']}, RP0 = {c['RP0']}, NS = {c['NS']}, MR = {c['MR']}, "
             f"DMIN = {c['DMIN']}, DMAX = {c['DMAX']}, DSPAN = DMAX - DMIN, NV1 = {nv1}, NV2 = {nv2}, "
             f"NIT1 = {-(-nv1 // NT)}, NIT2 = {-(-nv2 // NT)};")
    L.append(f"    typedef {T} T;")
    L.append("    extern __shared__ __align__(128) unsigned char xgb_smem[];")
    L.append("    uint64_t *full = reinterpret_cast<uint64_t *>(xgb_smem);")
    L.append("    uint64_t *empty = full + NS;")
    L.append("    T *stages = reinterpret_cast<T *>(xgb_smem + 256);        // [NS][RP0][WP0] input planes (u^n)")
    L.append("    T *mids = stages + NS * RP0 * WP0;                        // [MR][RPM][WPM] middle planes (u^{n+1})")
    L.append("    const T *src = static_cast<const T *>(p.aux0);")
    L.append("    T *out2 = static_cast<T *>(p.aux1), *out1 = static_cast<T *>(p.aux2);")
    L.append("    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;")
    L.append("    const int64_t c0 = (int64_t)blockIdx.x * W, j0 = (int64_t)blockIdx.y * TJ;")
    L.append("    const int64_t i0 = p.r_lo + (int64_t)blockIdx.z * p.chunk0;")
    L.append("    if (i0 >= p.r_hi) return;")
    L.append("    const int64_t iend = (i0 + p.chunk0 < p.r_hi) ? (i0 + p.chunk0) : p.r_hi;")
    L.append("    const int64_t S1 = p.n2, S0 = p.n1 * p.n2, total = p.n0 * S0;")
    L.append("    const int nplanes = (int)(iend - i0) + 2 * DSPAN;            // input planes this CTA streams")
    L.append("    if (threadIdx.x == 0) {")
    L.append("        for (int s = 0; s < NS; ++s) { xgb::pipe::mbar_init(&full[s], 1); xgb::pipe::mbar_init(&empty[s], NCW); }")
    L.append("        xgb::pipe::fence_barrier_init();")
    L.append("    }")
    L.append("    __syncthreads();")
    L.append("    if (warp == NCW) {                                         // producer: RP0 row copies per input plane")
    L.append("        int s = 0, eph = 1;")
    L.append("        const T *plane = src + ((i0 + 2 * DMIN) * S0 + (j0 - 2 * HJ) * S1 + (c0 - HK0));")
    L.append("        for (int t = 0; t < nplanes; ++t, ++s, plane += S0) {")
    L.append("            if (s == NS) { s = 0; eph ^= 1; }")
    L.append("            if (t >= NS) xgb::pipe::mbar_wait(&empty[s], eph);")
    L.append("            if (lane == 0) xgb::pipe::mbar_expect_tx(&full[s], (uint32_t)(RP0 * WP0 * sizeof(T)));")
    L.append("            __syncwarp();")
    L.append("            for (int jj = lane; jj < RP0; jj += 32)")
    L.append("                xgb::pipe::bulk_g2s(stages + ((int64_t)s * RP0 + jj) * WP0, plane + jj * S1, (uint32_t)(WP0 * sizeof(T)), &full[s]);")
    L.append("        }")
    L.append("        return;")
    L.append("    }")
    L.extend(hoist_lines(hoist))
    L.append("    const int tid = warp * 32 + lane;")
    L.append("    int fs = 0, fph = 0, rs = 0, ms = 0;")
    L.append("    int64_t lin_mid = (i0 + DMIN) * S0 + (j0 - HJ) * S1 + (c0 - HKM);   // first cell of the next middle plane's region")
    L.append("    int64_t lin_out = i0 * S0 + j0 * S1 + c0;")
    L.append("    for (int t = 0; t < nplanes; ++t) {")
    L.append("        xgb::pipe::mbar_wait(&full[fs], fph);")
    L.append("        if (t >= DSPAN) {")
    L.append("            // ---- step 1: middle plane over rows [j0 - HJ, j0 + TJ + HJ) x columns [c0 - HKM, c0 + W + HKM)")
    for d0 in d0s:
        k = c["DMAX"] - d0
        L.append(f"            const T *in{d0 - c['DMIN']} = stages + (int64_t)((fs >= {k}) ? (fs - {k}) : (fs - {k} + NS)) * (RP0 * WP0) + HJ * WP0 + (HK0 - HKM);")
    L.append("            T *mpl = mids + (int64_t)ms * (RPM * WPM);")
    L.append("            int fls[NIT1];")
    L.append("#pragma unroll")
    L.append("            for (int i = 0; i < NIT1; ++i) {")
    L.append("                const int idx = tid + i * NT;")
    L.append("                const int jj = idx / (WPM / V), kk = (idx % (WPM / V)) * V;")
    L.append("                const int64_t lin = lin_mid + jj * S1 + kk;")
    L.append("                fls[i] = -1;")
    L.append(f"                if (idx < NV1 && lin >= 0 && lin + V <= total) fls[i] = xgb::ld_flag(p.m_{gname}, p.f_{gname}, lin);")
    L.append("            }")
    L.append("#pragma unroll")
    L.append("            for (int i = 0; i < NIT1; ++i) {")
    L.append("                const int idx = tid + i * NT;")
    L.append("                if (idx < NV1) {")
    L.append("                    const int jj = idx / (WPM / V), kk = (idx % (WPM / V)) * V;")
    L.append("                    const int fl = fls[i];")
    L.append("                    const int64_t lin = lin_mid + jj * S1 + kk;")
    L.append("                    if (fl >= 0) {")
    L.extend(body("in", "WP0", "                        "))
    L.append("                        xgb::st_vec<T, V>(mpl + jj * WPM + kk, val);")
    L.append("                    } else {")
    L.append("                        T z[V];")
    L.append("#pragma unroll")
    L.append("                        for (int v = 0; v < V; ++v) z[v] = T(0);")
    L.append("                        xgb::st_vec<T, V>(mpl + jj * WPM + kk, z);")
    L.append("                    }")
    L.append("                }")
    L.append("            }")
    L.append("            lin_mid += S0;")
    L.append("            asm volatile(\"bar.sync 1, %0;\" :: \"n\"(NT) : \"memory\");")
    L.append("        }")
    L.append("        if (t >= 2 * DSPAN) {")
    L.append("            // ---- step 2: output plane over the CTA's own TJ x W tile")
    for d0 in d0s:
        k = c["DMAX"] - d0
        L.append(f"            const T *md{d0 - c['DMIN']} = mids + (int64_t)((ms >= {k}) ? (ms - {k}) : (ms - {k} + MR)) * (RPM * WPM) + HJ * WPM + HKM;")
    L.append("            int fls[NIT2];")
    L.append("#pragma unroll")
    L.append("            for (int i = 0; i < NIT2; ++i) {")
    L.append("                const int idx = tid + i * NT;")
    L.append("                const int jj = idx / (W / V), kk = (idx % (W / V)) * V;")
    L.append("                fls[i] = -1;")
    L.append(f"                if (idx < NV2 && j0 + jj < p.n1 && c0 + kk < p.n2) fls[i] = xgb::ld_flag(p.m_{gname}, p.f_{gname}, lin_out + jj * S1 + kk);")
    L.append("            }")
    L.append("#pragma unroll")
    L.append("            for (int i = 0; i < NIT2; ++i) {")
    L.append("                const int idx = tid + i * NT;")
    L.append("                const int fl = fls[i];")
    L.append("                if (fl >= 0) {")
    L.append("                    const int jj = idx / (W / V), kk = (idx % (W / V)) * V;")
    L.append("                    const int64_t lin = lin_out + jj * S1 + kk;")
    L.extend(body("md", "WPM", "                    "))
    L.append("                    xgb::st_vec<T, V>(out2 + lin, val);")
    L.append("                    if (p.opt0) {")
    L.append(f"                        T mv[V]; xgb::ld_vec<T, V>(md{0 - c['DMIN']} + jj * WPM + kk, mv); xgb::st_vec<T, V>(out1 + lin, mv);")
    L.append("                    }")
    L.append("                }")
    L.append("            }")
    L.append("            lin_out += S0;")
    L.append("        }")
    L.append("        if (t >= DSPAN) {")
    L.append("            __syncwarp();")
    L.append("            if (lane == 0) xgb::pipe::mbar_arrive(&empty[rs]);")
    L.append("            rs = (rs + 1 == NS) ? 0 : rs + 1;")
    L.append("            ms = (ms + 1 == MR) ? 0 : ms + 1;")
    L.append("        }")
    L.append("        if (++fs == NS) { fs = 0; fph ^= 1; }")
    L.append("    }")
    L.append("}")
    return "\n".join(L) + "\n"
