"""Kernel program: grouping, host control flow, launches.

The reference turns a kernel into ONE C function: scalar prologue and control
flow run serially on the calling thread and each stencil statement is an OpenMP
loop nest in program order (xgrid/lang/generator.py:216-225,285-364;
SURVEY.md §3d).  Here the same program order is kept, but

* scalar statements and control flow are evaluated on the host by ``HostEval``
  with C semantics (declared-width integers/floats, truncating ``/``,
  dividend-signed ``%``, unsuffixed-literal doubles);
* consecutive stencil statements that are provably independent point-wise are
  fused into one *sweep group* = one launch of a generated sm_100a kernel;
* statements that only touch a sparse boundary set (mask value != 0) run over
  a compacted index list instead of scanning the whole grid;
* Jacobi-style "implicit" statements (generator.py:312-352) write a per-grid
  scratch level that is then pointer-swapped with level 0 -- no malloc/free and
  no second copy sweep.

A whole ``Operator.__call__`` is recorded once per (scalar arguments, buffer
identities) into a CUDA graph and replayed afterwards.
"""
from __future__ import annotations

import copy
import hashlib
import math
import os
import threading
from concurrent.futures import ThreadPoolExecutor
from dataclasses import astuple, is_dataclass

import numpy as np

from ..config import get_config
from ..log import Logger
from ..types import Boolean, Floating, Grid as GridT, Integer, Pointer, Structure, Void
from . import cudagen, inlinec, ir, jacobi2

_TEMPLATE_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "csrc", "templates")


def template_headers() -> dict:
    out = {}
    for fn in sorted(os.listdir(_TEMPLATE_DIR)):
        if fn.endswith(".cuh"):
            with open(os.path.join(_TEMPLATE_DIR, fn)) as f:
                out[fn] = f.read()
    return out


# --------------------------------------------------------------------------- plan nodes
class GroupNode:
    def __init__(self, group: cudagen.Group) -> None:
        self.group = group


def _reads_level0_of(stmt, grids: set) -> bool:
    return any(ld.level == 0 and ld.variable.name in grids for ld in stmt.sweep.loads)


def build_plan(stmts: list, groups: list, ndim_of) -> list:
    """Replace runs of fusable stencil statements by GroupNodes (recursively)."""
    plan: list = []
    cur: list = []

    def flush():
        nonlocal cur
        if cur:
            g = cudagen.Group(len(groups), cur[0].sweep.grid.type.dimension, cur,
                              implicit=cur[0].sweep.implicit)
            groups.append(g)
            plan.append(GroupNode(g))
            cur = []

    def can_join(s) -> bool:
        if not cur:
            return True
        if s.sweep.implicit or cur[0].sweep.implicit:
            return False
        if s.sweep.grid.type.dimension != cur[0].sweep.grid.type.dimension:
            return False
        stored = {a.sweep.grid.name for a in cur}
        if any((a.sweep.grid.name, a.sweep.mask) == (s.sweep.grid.name, s.sweep.mask) for a in cur):
            return False
        if s.sweep.store.level != 0 or any(a.sweep.store.level != 0 for a in cur):
            return False
        # flow dependence: s reads a level-0 buffer some earlier member writes
        if _reads_level0_of(s, stored):
            return False
        # anti dependence: an earlier member reads level 0 of what s writes
        if any(_reads_level0_of(a, {s.sweep.grid.name}) for a in cur):
            return False
        # a sparse (mask != 0) statement that reads its own level 0 stays alone
        if _reads_level0_of(s, {s.sweep.grid.name}):
            return False
        return True

    for s in stmts:
        if isinstance(s, ir.Assignment) and s.sweep is not None:
            if not can_join(s):
                flush()
            cur.append(s)
            if s.sweep.implicit or _reads_level0_of(s, {s.sweep.grid.name}):
                flush()
            continue
        flush()
        if isinstance(s, ir.If):
            plan.append(("if", s, build_plan(s.body, groups, ndim_of), build_plan(s.orelse, groups, ndim_of)))
        elif isinstance(s, ir.While):
            plan.append(("while", s, build_plan(s.body, groups, ndim_of)))
        elif isinstance(s, ir.For):
            plan.append(("for", s, build_plan(s.body, groups, ndim_of)))
        else:
            plan.append(("stmt", s))
    flush()
    return plan


# --------------------------------------------------------------------------- host evaluator
class _Return(Exception):
    def __init__(self, value) -> None:
        self.value = value


class _Break(Exception):
    pass


class _Continue(Exception):
    pass


def np_type(t):
    if isinstance(t, Boolean):
        return np.bool_
    return t.np_dtype


def coerce(t, value):
    """Convert a Python / numpy value to the C object type ``t``."""
    if isinstance(t, Structure):
        # C passes, assigns and returns structs BY VALUE (the reference emits `struct S q = p;`,
        # generator.py:216-225): `q = p; q.x = 5.0` must leave p -- and the caller's dataclass -- untouched
        return copy.deepcopy(value)
    if isinstance(t, Integer):
        half = 1 << (t.width_bits - 1)
        return np_type(t)((int(value) + half) % (2 * half) - half)   # two's-complement wrap
    return np_type(t)(value)


class HostEval:
    """Evaluates scalar IR with C semantics (the host half of a kernel)."""

    def __init__(self, env: dict, grids: dict, launcher=None) -> None:
        self.env = env          # name -> numpy scalar | dataclass | ctypes pointer target
        self.grids = grids
        self.launcher = launcher    # the calling kernel's launcher (None: scalar-only context)

    def __call__(self, e):
        return getattr(self, "v_" + type(e).__name__)(e)

    def v_Constant(self, e):
        v = e.value
        if type(v) is bool:
            return np.bool_(v)
        if type(v) is int:
            return np.int32(v)
        return np.float64(v)        # unsuffixed literal: a C double

    def v_Identifier(self, e):
        val = self.env[e.variable.name]
        if isinstance(e.variable.type, Pointer):
            return coerce(e.variable.type.element, _deref(val))
        return val

    def v_Access(self, e):
        base = self(e.value)
        return coerce(e.type, getattr(base, e.attribute))

    def v_Unary(self, e):
        r = self(e.right)
        if e.operator == "!":
            return np.bool_(not bool(r))
        if e.operator == "-":
            with np.errstate(all="ignore"):
                return -r
        return +r

    def v_Binary(self, e):
        op = e.operator
        if op == "&&":
            return np.bool_(bool(self(e.left)) and bool(self(e.right)))
        if op == "||":
            return np.bool_(bool(self(e.left)) or bool(self(e.right)))
        a, b = self(e.left), self(e.right)
        with np.errstate(all="ignore"):
            if op == "+":
                return a + b
            if op == "-":
                return a - b
            if op == "*":
                return a * b
            if op == "/":
                if isinstance(a, np.integer) and isinstance(b, np.integer):
                    if int(b) == 0:
                        raise ZeroDivisionError("integer division by zero in kernel scalar code")
                    q = abs(int(a)) // abs(int(b))
                    return type(a)(q if (int(a) < 0) == (int(b) < 0) else -q)
                return a / b
            if op == "%":
                if isinstance(a, np.integer) and isinstance(b, np.integer):
                    return type(a)(math.fmod(int(a), int(b)))
                raise Exception("operator % on floating operands is not valid C")
            if op == "^":
                wide = isinstance(e.type, Floating) and e.type.width_bits == 64
                if isinstance(e.right, ir.Constant) and e.right.value == 2.0:
                    x = np.float64(a) if wide else np.float32(a)
                    return x * x
                if wide:
                    return np.float64(math.pow(float(a), float(b)))
                return np.float32(np.float32(a) ** np.float32(b))
            return np.bool_({"==": a == b, "!=": a != b, ">": a > b, ">=": a >= b,
                             "<": a < b, "<=": a <= b}[op])

    def v_Condition(self, e):
        return self(e.body) if bool(self(e.condition)) else self(e.orelse)

    def v_Cast(self, e):
        v = self(e.value)
        if isinstance(e.type, Integer) and isinstance(v, np.floating):
            v = math.trunc(float(v))
        return coerce(e.type, v)

    def v_GridInfo(self, e):
        g = self.grids[e.variable.name]
        if e.info == "dimension":
            return np.int32(len(g.shape))
        return np.int32(g.shape[int(self(e.dimension))])

    def v_Stencil(self, e):
        raise Exception(f"grid access to '{e.variable.name}' outside a stencil statement")

    def v_Call(self, e):
        # a grid argument is the Grid object itself (the reference passes its struct by value: same buffers)
        args = [self.grids[a.variable.name] if isinstance(a.type, GridT) and isinstance(a, ir.Identifier) else self(a)
                for a in e.arguments]
        if isinstance(e.operator, ir.Constructor):
            st = e.operator.type
            return st.dataclass(*[_to_py(a) for a in args])
        return call_operator(e.operator, args, self.launcher)


def _to_py(v):
    return v.item() if isinstance(v, np.generic) else v


def _deref(p):
    if hasattr(p, "contents"):
        return p.contents.value
    return p.value


def _store_ptr(p, v):
    if hasattr(p, "contents"):
        p.contents.value = _to_py(v)
    else:
        p.value = _to_py(v)


def call_operator(op, args, caller_launcher=None):
    """A call of another ``@kernel`` / ``@function`` operator from a kernel's host-side code.  The
    reference emits the callee into the same translation unit and calls it like any C function
    (xgrid/lang/generator.py:208-212,418-419): no tick, no ring resize -- the callee's statements simply
    run, in order, on the caller's buffers.  Here a scalar-only callee is interpreted on the host; a callee
    with grid parameters runs ITS program's plan (its own generated sweep kernels, fused groups, solver
    pairs) through a launcher bound to the caller's Grid objects, on the same stream."""
    if op.mode == "external":
        raise Exception(f"external operator '{op.name}' is defined by a CUDA header (includes=[...]) and can only "
                        "be called from stencil statements, not from a kernel's scalar code")
    d = op.ir
    env, grids = {}, {}
    for (n, t), v in zip(d.signature.arguments, args):
        if isinstance(t, GridT):
            grids[n] = v
        else:
            env[n] = coerce(t, v) if isinstance(t, Structure) else v  # by-value struct parameters
    if not grids:
        return _Interpreter(d, env, {}, launcher=None).run(build_plan(d.body, [], None))
    if caller_launcher is None:
        raise Exception(f"operator '{op.name}' takes grids and can only be called from a kernel")
    if _Launcher is None:
        _late_imports()
    prog = op._program()
    return _Interpreter(prog.ir, env, grids, _Launcher(prog, grids), prog.pairs).run(prog.plan)


def grid_callees(body) -> list:
    """Operators with grid parameters called (directly) from these statements."""
    out = []
    for st in ir.walk_stmts(body):
        for root in (getattr(st, "value", None), getattr(st, "condition", None)):
            if root is None or isinstance(root, (list, str)):
                continue
            for e in ir.walk_expr(root):
                if isinstance(e, ir.Call) and not isinstance(e.operator, ir.Constructor) \
                        and e.operator.mode != "external" \
                        and any(isinstance(t, GridT) for _, t in e.operator.signature.arguments) \
                        and e.operator not in out:
                    out.append(e.operator)
    return out


class _Interpreter:
    def __init__(self, definition, env, grids, launcher, pairs=None) -> None:
        self.d, self.env, self.grids, self.launcher = definition, env, grids, launcher
        self.ev = HostEval(env, grids, launcher if callable(getattr(launcher, "_params", None)) else None)
        self.pairs = pairs or {}        # id(ir.For) -> jacobi2.Pair (two iterations per pass)

    def run(self, plan):
        try:
            self.block(plan)
        except _Return as r:
            return r.value
        return None

    def block(self, plan) -> None:
        for node in plan:
            if isinstance(node, GroupNode):
                if self.launcher is None:
                    raise Exception("stencil statements need a kernel context (grids bound to a launcher)")
                self.launcher(node.group, self.env)
                continue
            kind = node[0]
            if kind == "stmt":
                self.statement(node[1])
            elif kind == "if":
                self.block(node[2] if bool(self.ev(node[1].condition)) else node[3])
            elif kind == "while":
                while bool(self.ev(node[1].condition)):
                    try:
                        self.block(node[2])
                    except _Break:
                        break
                    except _Continue:
                        continue
            elif kind == "for":
                s = node[1]
                name, t = s.variable.name, s.variable.type
                self.env[name] = coerce(t, self.ev(s.start))
                pair = self.pairs.get(id(s))
                if pair is not None and not self.launcher.pair_ok(pair):
                    pair = None
                fused_left = 0
                if pair is not None:
                    # the body holds only sweeps, so the trip count is known here.  Every fused pass swaps
                    # the iterated grid's level 0 with its scratch buffer ONCE (two single sweeps swap twice);
                    # keep the number of swaps per call even, so that the buffer arrangement -- and with it
                    # the recorded CUDA graph -- repeats every second call instead of every sixth
                    with np.errstate(all="ignore"):
                        step, lo, hi = int(self.ev(s.step)), int(self.env[name]), int(self.ev(s.end))
                    trips = max(0, -(-(hi - lo) // step)) if step > 0 else 0
                    fused_left = trips // 2
                    if fused_left % 2 == 1 and trips % 2 == 0:
                        fused_left -= 1
                while bool(self.env[name] < self.ev(s.end)):
                    if fused_left > 0:
                        # sweep, boundary statements, sweep in ONE pass (jacobi2), then the boundary
                        # statements of the second iteration
                        fused_left -= 1
                        self.launcher.run_pair(pair, self.env)
                        with np.errstate(all="ignore"):
                            self.env[name] = coerce(t, coerce(t, self.env[name] + step) + step)
                        continue
                    try:
                        self.block(node[2])
                    except _Break:
                        break
                    except _Continue:
                        pass
                    with np.errstate(all="ignore"):
                        self.env[name] = coerce(t, self.env[name] + self.ev(s.step))

    def statement(self, s) -> None:
        if isinstance(s, ir.Assignment):
            value = self.ev(s.value)
            self.assign(s.terminal, value)
        elif isinstance(s, ir.Return):
            raise _Return(None if s.value is None else self.ev(s.value))
        elif isinstance(s, ir.Break):
            raise _Break()
        elif isinstance(s, ir.Continue):
            raise _Continue()
        elif isinstance(s, ir.Evaluation):
            self.ev(s.value)
        elif isinstance(s, ir.Inline):
            if not hasattr(self.launcher, "inline"):
                raise Exception("`with xgrid.c()` blocks are only supported in kernels")
            self.launcher.inline(s, self.env)
        else:
            raise Exception(f"unsupported statement {type(s).__name__}")

    def _designated(self, e):
        """The struct object an lvalue expression names (`q`, `q.inner`), not a by-value copy of it."""
        if isinstance(e, ir.Identifier) and not isinstance(e.variable.type, Pointer):
            return self.env[e.variable.name]
        if isinstance(e, ir.Access):
            return getattr(self._designated(e.value), e.attribute)
        return self.ev(e)

    def assign(self, target, value) -> None:
        if isinstance(target, ir.Identifier):
            var = target.variable
            if isinstance(var.type, Pointer):
                _store_ptr(self.env[var.name], coerce(var.type.element, value))
            else:
                self.env[var.name] = coerce(var.type, value)
        elif isinstance(target, ir.Access):
            base = self._designated(target.value)
            setattr(base, target.attribute, _to_py(coerce(target.type, value)))
        else:
            raise Exception("unsupported assignment target")


# --------------------------------------------------------------------------- deferred calls (temporal blocking)
_PENDING = None          # {"program", "args", "grid", "key", "count"}: identical calls not yet executed
PENDING_LIMIT = 4096
# Slab calls are recorded like any other.  The graph is instantiated with per-node priorities (xgb_graph_end), so a
# recorded halo exchange still overtakes the interior sweep it overlaps with: 8 GPUs, 256x2048^2 slabs, 3.14 ms / step
# recorded, 3.15 ms direct -- and 3.49 ms when the graph ran every node at the launch stream's priority.
SHARDED_GRAPH_MIN_LAUNCHES = int(os.environ.get("XGB_SHARDED_GRAPH_MIN", "0"))
MULTISTEP_MIN_POINTS = int(os.environ.get("XGB_MS_MIN", "16384"))
MULTISTEP_TAIL = os.environ.get("XGB_MS_TAIL", "1") != "0"     # remainders of a run: tail variant, not single steps
TILED2_ENABLED = os.environ.get("XGB_TILED2", "1") != "0"


def two_step_slab_rows(n0: int, dmin: int, dmax: int, has_lo: bool, has_hi: bool) -> tuple:
    """Row ranges of a two-steps-per-pass sweep over a slab of `n0` rows whose taps reach axis-0 offsets
    dmin <= 0 <= dmax: (r_lo, r_hi, deep, edge).  Output rows [r_lo, r_hi) need nothing from a neighbour -- their
    two-step cone, widened by one row because a tap that leaves its row at the first / last column lands in the
    adjacent row (F10), stays inside the slab.  The others run step-at-a-time: step 1 on the `deep` ranges (far enough
    into the slab for step 2's taps, again plus the one row), step 2 on the `edge` ranges.  Checked against a
    whole-domain NumPy model in tests/test_slab_two_step_model.py."""
    lo_band = 2 * -dmin + 1 if has_lo else 0
    hi_band = 2 * dmax + 1 if has_hi else 0
    deep, edge = [], []
    if lo_band:
        deep.append((0, min(n0, lo_band + dmax + 1)))
        edge.append((0, lo_band))
    if hi_band:
        deep.append((max(0, n0 - hi_band + dmin - 1), n0))
        edge.append((n0 - hi_band, n0))
    return lo_band, n0 - hi_band, deep, edge


def _world() -> int:
    from .. import dist
    return dist.topology().world


def multistep_launches(count: int, T: int, S: int, tail: bool = True) -> tuple[list, int]:
    """Split a run of `count` identical deferred 1-D calls into multi-step launches: full launches of
    T steps, then ONE tail launch of the largest multiple of S that is left if that is at least 2*S
    (S is even, so every launch advances an even number of steps and the ring order is preserved).
    Returns (steps per launch, calls left for step-at-a-time execution)."""
    steps = [T] * (count // T)
    left = count % T
    rest = left - left % S
    if tail and rest >= 2 * S:
        steps.append(rest)
        left -= rest
    return steps, left


JIT_MODE = os.environ.get("XGB_JIT", "lazy")                    # "lazy": one module per generated kernel, on first launch
JIT_THREADS = int(os.environ.get("XGB_JIT_THREADS", "8"))      # parallel NVRTC compilations in Program.images()
_nvrtc_warm = False


def _warm_nvrtc() -> None:
    """The runtime resolves libnvrtc on its first compilation; do that once before threads race for it."""
    global _nvrtc_warm
    if not _nvrtc_warm:
        from ..runtime import shim
        shim.compile_cuda('extern "C" __global__ void xgb_warm() {}', "warm.cu", ["--gpu-architecture=sm_100a"], {})
        _nvrtc_warm = True


def flush_pending() -> None:
    """Execute the deferred run of identical kernel calls, T time steps per launch."""
    global _PENDING
    p, _PENDING = _PENDING, None
    if p is not None:
        p["program"]._run_batch(p["args"], p["grid"], p["count"])


_Grid = None
_Launcher = None
_Runtime = None


def _late_imports() -> None:
    """grid / launch / shim import this module's siblings; resolve them once, not per call."""
    global _Grid, _Launcher, _Runtime
    from ..grid import Grid
    from ..runtime.shim import Runtime
    from .launch import Launcher
    _Grid, _Launcher, _Runtime = Grid, Launcher, Runtime


# --------------------------------------------------------------------------- Program
class Program:
    """A compiled kernel: plan + generated module + launch logic."""

    _serial = 0

    def __init__(self, op) -> None:
        self.op = op
        self.logger = Logger(self)
        self.config = get_config()
        self.ir = op.ir
        self.depth = self.ir.depth
        self.groups: list[cudagen.Group] = []
        self.plan = build_plan(self.ir.body, self.groups, None)
        self._sig = self.ir.signature.arguments
        self._typed_ok: dict = {}
        self.grid_args = [(n, t) for n, t in self.ir.signature.arguments if isinstance(t, GridT)]
        self.grid_ndims = {n: t.dimension for n, t in self.grid_args}
        Program._serial += 1
        tag = "".join(ch if ch.isalnum() else "_" for ch in op.name)
        self.module_builder = cudagen.ModuleBuilder(self.config.overstep, self.config.comment)
        self.module_builder.include(op.includes)
        # operators with grid parameters called from this kernel run their own programs on our buffers
        self.callees = grid_callees(self.ir.body)
        scope_types = {n: v.type for n, v in self.ir.scope.items()}
        for g in self.groups:
            g.name = f"xg_{tag}_g{g.gid}"
            cudagen.analyse_group(g, scope_types)
        # solver loops whose body is [implicit sweep, boundary statements...]: two iterations per pass
        self.pairs: dict = {}
        if self.config.overstep == "none":
            self._match_pairs(self.plan)
        for g in self.groups:
            cudagen.emit_group(g, self.module_builder, scope_types, self.grid_ndims)
        for pair in self.pairs.values():
            jacobi2.configure(pair)
            self.module_builder.kernels.append(jacobi2.emit(pair, self.module_builder))
        # `with xgrid.c()` blocks: one single-thread device kernel each (lang/inlinec.py)
        self.inlines: dict = {}
        for k, st in enumerate(inlinec.collect(self.ir.body)):
            text, ik = inlinec.emit(tag, k, st, self.ir.scope, self.depth, self.module_builder)
            if not self.inlines:
                self.module_builder.kernels.append(inlinec.PRELUDE + "\n".join(op.macro) + "\n")
            self.module_builder.kernels.append(text)
            self.inlines[id(st)] = ik
        self.source = self.module_builder.source() if (self.groups or self.inlines) else ""
        self._modules: dict = {}          # unit index -> loaded module
        self._where = None                # kernel name -> unit index
        self._unit_list = None            # [(kernel names, source)]
        self._cubins: list = []
        self._prefetched = False
        self._functions: dict = {}
        # temporal blocking: the whole kernel is scalar prologue + ONE 1-D group on one grid
        self.batchable = False
        self._grid_pos = -1
        if (len(self.groups) == 1 and self.groups[0].multistep is not None and self.depth == 2
                and not self.callees and len(self.grid_args) == 1 and op.tick and self.config.overstep == "none"
                and isinstance(self.ir.signature.return_type, Void)
                and isinstance(self.plan[-1], GroupNode)
                and all(isinstance(n, tuple) and n[0] == "stmt" and isinstance(n[1], ir.Assignment)
                        for n in self.plan[:-1])
                and not any(isinstance(t, (Pointer, Structure)) for _, t in self.ir.signature.arguments)):
            self.batchable = True
            self._grid_pos = [n for n, (_, t) in enumerate(self.ir.signature.arguments)
                              if isinstance(t, GridT)][0]
        # two-steps-per-pass variant for 2-D kernels (same program shape, one 2-D group)
        self.batchable2 = False
        if (len(self.groups) == 1 and self.groups[0].tiled2 is not None and self.depth == 2
                and not self.callees and len(self.grid_args) == 1 and op.tick and self.config.overstep == "none"
                and isinstance(self.ir.signature.return_type, Void)
                and isinstance(self.plan[-1], GroupNode)
                and all(isinstance(n, tuple) and n[0] == "stmt" and isinstance(n[1], ir.Assignment)
                        for n in self.plan[:-1])
                and not any(isinstance(t, (Pointer, Structure)) for _, t in self.ir.signature.arguments)):
            self.batchable2 = True
            self._grid_pos = [n for n, (_, t) in enumerate(self.ir.signature.arguments)
                              if isinstance(t, GridT)][0]
        self._graphs: dict = {}
        self._halo0 = None
        self._batch_params: dict = {}     # (scalars, grid, mask) -> marshalled parameter struct of a deferred 1-D run
        self._launches_per_call = 0       # launches of the last directly executed call
        self._seen: set = set()
        self._image = None
        self._preloaded = False
        # calls that write through ptr[T] arguments have host-visible side effects: never replayed
        self.cacheable = not any(
            isinstance(st, ir.Assignment) and isinstance(st.terminal, ir.Identifier)
            and isinstance(st.terminal.variable.type, Pointer) for st in ir.walk_stmts(self.ir.body))

    def halo0(self, _seen=None) -> int:
        """Largest axis-0 reach of any sweep this call can launch, callee programs included."""
        if _seen is None and self._halo0 is not None:
            return self._halo0
        seen = _seen if _seen is not None else set()
        seen.add(id(self.op))
        h = max([1] + [g.halo0 for g in self.groups])
        for op in self.callees:
            if id(op) not in seen:
                h = max(h, op._program().halo0(seen))
        if _seen is None:
            self._halo0 = h             # (groups and callees are fixed once the program is built)
        return h

    def replayable(self, _seen=None) -> bool:
        seen = _seen if _seen is not None else set()
        seen.add(id(self.op))
        return self.cacheable and all(id(op) in seen or op._program().replayable(seen) for op in self.callees)

    def _match_pairs(self, plan: list) -> None:
        for node in plan:
            if isinstance(node, GroupNode):
                continue
            if node[0] == "for":
                body = node[2]
                pair = jacobi2.match(body, lambda n: n.group if isinstance(n, GroupNode) else None)
                if pair is not None:
                    used = set(pair.sweep.scalars)
                    for r in pair.rules:
                        used |= set(r.group.scalars)
                    if node[1].variable.name not in used:      # both iterations see the same scalars
                        self.pairs[id(node[1])] = pair
                        continue
                self._match_pairs(body)
            elif node[0] == "while":
                self._match_pairs(node[2])
            elif node[0] == "if":
                self._match_pairs(node[2])
                self._match_pairs(node[3])

    # ---- JIT (replaces Compiler.compile's md5 cache, xgrid/util/ffi.py:67-85)
    def _compile_cached(self, source: str, label: str) -> bytes:
        """One NVRTC compilation through the on-disk cache (sha256 of source + flags + template headers)."""
        from ..runtime import shim
        headers = dict(template_headers())
        headers.update(self._user_headers())
        flags = self.config.nvrtc_flags
        key = hashlib.sha256("\0".join([source, *flags, *headers.values()]).encode()).hexdigest()[:32]
        root = os.path.join(".", self.config.cacheroot)
        os.makedirs(root, exist_ok=True)
        cu, cubin = os.path.join(root, key + ".cu"), os.path.join(root, key + ".cubin")
        if os.path.exists(cubin) and os.path.exists(cu):
            with open(cubin, "rb") as f:
                image = f.read()
            self.logger.info(f"jit loaded '{cubin}' from cache")
            return image
        image, log = shim.compile_cuda(source, label + ".cu", flags, headers)
        tmp = f"{cu}.{os.getpid()}.{threading.get_ident()}.tmp"
        with open(tmp, "w") as f:
            f.write("// nvrtc " + " ".join(flags) + "\n" + source)
        os.replace(tmp, cu)
        tmp = f"{cubin}.{os.getpid()}.{threading.get_ident()}.tmp"
        with open(tmp, "wb") as f:
            f.write(image)
        os.replace(tmp, cubin)
        self.logger.info(f"jit compiled '{cu}' to '{cubin}'")
        return image

    def _user_headers(self) -> dict:
        """CUDA headers named by `includes=` / `import a.b` (reference: `#include "a/b.h"` resolved by gcc
        relative to the working directory, generator.py:96-97): read here and handed to NVRTC by name.
        Looked up in the working directory, then next to the kernel's source file."""
        out = {}
        if not self.module_builder.includes:
            return out
        roots = ["."]
        try:
            import inspect
            roots.append(os.path.dirname(os.path.abspath(inspect.getsourcefile(self.op.func))))
        except (TypeError, OSError):
            pass
        for name in self.module_builder.includes:
            for root in roots:
                path = os.path.join(root, name)
                if os.path.isfile(path):
                    with open(path) as f:
                        out[name] = f.read()
                    break
            else:
                self.logger.dead(f"header '{name}' (includes= / import) was not found in {roots}")
        return out

    def image(self) -> bytes:
        """The whole translation unit as ONE sm_100a cubin (tools, tests, the build check)."""
        if self._image is None:
            self._image = self._compile_cached(self.source, self.op.name)
        return self._image

    def _units(self) -> list:
        """[(kernel names, source)]: the compilation units of the launch path.  JIT mode "lazy" (default): one
        unit per kernel, compiled when the kernel is first launched -- a kernel with many statements has dozens
        of generated variants (cavity: 57 kernels, 20 s in one NVRTC call, 6 s of it for each of two variants
        that large grids never use), while one call needs about a dozen of them.  Kernels depend only on the
        shared declarations, never on each other, and are compiled function by function either way, so the
        machine code is the same.  Mode "unit" (XGB_JIT=unit): the whole kernel as one module."""
        if self._unit_list is None:
            n = self.source.count('extern "C" __global__') if JIT_MODE == "lazy" else 1
            self._unit_list = self.module_builder.chunks(max(1, n))
            self._where = {k: i for i, (names, _) in enumerate(self._unit_list) for k in names}
            self._cubins = [None] * len(self._unit_list)
        return self._unit_list

    def _unit_image(self, index: int) -> bytes:
        units = self._units()
        if self._cubins[index] is None:
            names, source = units[index]
            whole = len(units) == 1
            self._cubins[index] = self.image() if whole else self._compile_cached(source, f"{self.op.name}.{names[0]}")
        return self._cubins[index]

    def images(self) -> list:
        """[(kernel names, cubin)] of EVERY unit, compiled now, in parallel threads (NVRTC compiles distinct
        programs concurrently; ctypes releases the GIL).  The build step uses it to fill the on-disk cache."""
        units = self._units()
        todo = [i for i in range(len(units)) if self._cubins[i] is None]
        threads = min(JIT_THREADS, os.cpu_count() or 1, len(todo))
        if threads > 1:
            _warm_nvrtc()
            with ThreadPoolExecutor(max_workers=threads) as pool:
                list(pool.map(self._unit_image, todo))
        return [(names, self._unit_image(i)) for i, (names, _) in enumerate(units)]

    def _prefetch(self, grids: dict) -> None:
        """First call of a program: compile the units this call is about to launch in PARALLEL threads
        (only compilation -- modules are still loaded by `function` on first use).  The prediction repeats
        the launcher's variant choice; a miss merely falls back to compiling that kernel when it is needed."""
        self._prefetched = True
        if JIT_MODE != "lazy" or not self.source:
            return
        from .launch import full_grid_variant
        names = set()
        for g in self.groups:
            lead = grids.get(g.lead)
            if lead is None or lead.size == 0:
                continue
            if g.sparse:
                names.add(cudagen.kernel_name(g, cudagen.VARIANT_SPARSE, 1))
            else:
                variant, V, _ = full_grid_variant(g, lead.shape)
                names.add(cudagen.kernel_name(g, variant, V))
        for pair in self.pairs.values():
            names.add(cudagen.kernel_name(pair.sweep, jacobi2.VARIANT, pair.config["V"]))
        names.update(ik.name for ik in self.inlines.values())
        self._compile_named(names)

    def _compile_named(self, names) -> None:
        """Compile the units of the given kernels in parallel threads (no-op for units already compiled)."""
        units = self._units()
        todo = sorted({self._where[n] for n in names if n in self._where and self._cubins[self._where[n]] is None})
        threads = min(JIT_THREADS, os.cpu_count() or 1, len(todo))
        if threads > 1 and len(units) > 1:
            _warm_nvrtc()
            with ThreadPoolExecutor(max_workers=threads) as pool:
                list(pool.map(self._unit_image, todo))

    def function(self, name: str, dynamic_smem: int = 0) -> int:
        fn = self._functions.get(name)
        if fn is None:
            from ..runtime.shim import Runtime
            rt = Runtime.get()
            self._units()
            unit = self._where[name]
            module = self._modules.get(unit)
            if module is None:                  # compiled (or read from the cache) and loaded on first use
                module = self._modules[unit] = rt.module_load(self._unit_image(unit))
            fn = rt.get_function(module, name)
            if dynamic_smem > 48 * 1024:
                rt.set_dynamic_smem(fn, dynamic_smem)
            self._functions[name] = fn
        return fn

    # ---- call
    def __call__(self, *args):
        global _PENDING
        p = _PENDING
        if p is not None and p["program"] is self and p["count"] < PENDING_LIMIT and args == p["args"]:
            p["count"] += 1         # one more identical deferred call (grids compare by identity)
            return None
        sig = self.ir.signature.arguments
        if len(args) != len(sig):
            # xgrid/util/ffi.py:31-33
            raise TypeError(f"this function takes {len(sig)} argument ({len(args)} given)")
        if (self.batchable or self.batchable2) and self.config.temporal:
            # defer: a run of identical calls is executed several steps per launch on flush
            grid = args[self._grid_pos]
            if self.batchable:
                defer = getattr(grid, "size", 0) >= MULTISTEP_MIN_POINTS and grid.dimension == 1
            else:
                t2 = self.groups[0].tiled2
                nd = getattr(grid, "dimension", 0)
                # (a slab's own row count may differ by one between ranks: the smallest one decides for all)
                rows = grid.global_shape[0] // _world() if grid.sharded else grid.shape[0] if nd else 0
                defer = (nd == self.groups[0].ndim and TILED2_ENABLED
                         and grid.shape[-1] >= t2["W"] and grid.shape[-1] % t2["V"] == 0 and rows >= 64
                         and (nd == 2 or grid.shape[1] >= t2["TJ"]))
                if defer and grid.sharded:      # the rows next to a cut run on row bands: needs a row-range kernel
                    from .launch import full_grid_variant
                    defer = full_grid_variant(self.groups[0], (rows,) + tuple(grid.shape[1:]))[0] in (
                        cudagen.VARIANT_TILED, cudagen.VARIANT_MARCH)
            if defer:
                key = tuple(a for n, a in enumerate(args) if n != self._grid_pos)
                p = _PENDING
                if p is not None and p["program"] is self and p["grid"] is grid and p["key"] == key \
                        and p["count"] < PENDING_LIMIT:
                    p["count"] += 1
                    return None
                flush_pending()
                self._bind(args)        # type / arity errors surface at the call site
                if not self._preloaded:
                    self._preload_batch_kernels(grid)
                # output levels of the several-steps kernels: allocate with the first deferred call, in
                # the ghost layout the flush will ask for (a 1-D slab imports H points of both ring levels
                # per launch) -- a flush inside a timed region should only launch, never re-lay-out
                if self.batchable:
                    need, spares = (self.groups[0].multistep["H"] if grid.sharded else 1), 2
                else:
                    need, spares = self.groups[0].tiled2["ghost"], 1
                if grid._ghost < need or len(grid._spares) < spares:
                    grid._ensure_ghost(need)
                    grid._spare_levels(spares)
                _PENDING = {"program": self, "args": args, "grid": grid, "key": key, "count": 1}
                return None
        if _PENDING is not None:
            flush_pending()
        return self._call_now(args)

    def _preload_batch_kernels(self, grid) -> None:
        """Resolve the kernels a deferred run can launch when its FIRST call is queued: the several-steps
        variants and the one-pass kernel that runs the remainder.  With the lazy JIT each of them is compiled
        (or read from the cache) and loaded on first use; a flush inside a timed region should only launch."""
        self._preloaded = True
        g = self.groups[0]
        from .launch import full_grid_variant
        variant, V, smem = full_grid_variant(g, grid.shape)
        wanted = [(cudagen.kernel_name(g, variant, V), smem)]
        if self.batchable:
            wanted += [(cudagen.kernel_name(g, v, 1), g.multistep["smem"])
                       for v in (cudagen.VARIANT_MULTISTEP, cudagen.VARIANT_MULTISTEP_TAIL)]
            if g.multistep_short is not None:
                wanted.append((cudagen.kernel_name(g, cudagen.VARIANT_MULTISTEP_SHORT, 1), g.multistep_short["smem"]))
        else:
            wanted.append((cudagen.kernel_name(g, cudagen.VARIANT_TILED2, g.tiled2["V"]), g.tiled2["smem"]))
        if JIT_MODE == "lazy":
            self._compile_named([name for name, _ in wanted])
        for name, dynamic_smem in wanted:
            self.function(name, dynamic_smem)

    def _bind(self, args):
        if _Grid is None:
            _late_imports()
        Grid = _Grid
        sig = self._sig
        env, grids = {}, {}
        for (name, t), a in zip(sig, args):
            if isinstance(t, GridT):
                if not isinstance(a, Grid):
                    raise TypeError(f"argument '{name}' must be an xgrid.Grid")
                if (id(a.typing), name) not in self._typed_ok:
                    if a.dimension != t.dimension or a.element != t.element:
                        raise TypeError(f"argument '{name}' expects {t!r}, got Grid({a.dimension}) of {a.element!r}")
                    self._typed_ok[(id(a.typing), name)] = a.typing      # keeps the id alive
                grids[name] = a
            elif isinstance(t, Pointer):
                env[name] = a
            elif isinstance(t, Structure):
                if not is_dataclass(a):
                    raise TypeError(f"argument '{name}' expects dataclass {t.name}")
                env[name] = copy.deepcopy(a)          # by value: the kernel works on its own copy
            else:
                env[name] = coerce(t, a)
        return env, grids

    def _call_now(self, args):
        sig = self._sig
        env, grids = self._bind(args)
        if not self._prefetched:
            self._prefetch(grids)
        # tick the field and resize the time ring (xgrid/lang/operator.py:37-39)
        for (name, t), a in zip(sig, args):
            if isinstance(t, GridT):
                a._op_invoke(self.depth, self.op.tick)   # once per argument, like the reference
        ghost = self.halo0()
        for g in grids.values():
            g._prepare_device(ghost)

        Launcher = _Launcher
        sharded = any(g.sharded for g in grids.values())
        if sharded and self.config.graphs:
            # a recorded call must be self-contained: no dependency may cross the capture boundary, so halo
            # exchanges still in flight from the previous call are joined into the compute stream first (the
            # next reader would wait for them anyway); the halo freshness of every level is part of the key
            self._join_halo_events(grids)
        key = self._graph_key(env, grids) if (self.config.graphs and self.replayable()
                                              and (self.groups or self.callees)) else None
        if sharded and key is not None and self._launches_per_call < SHARDED_GRAPH_MIN_LAUNCHES:
            key = None                      # (diagnostic knob, off by default: XGB_SHARDED_GRAPH_MIN)
        hit = self._graphs.get(key) if key is not None else None
        if hit is not None:
            # steady state: replay the recorded launches (on slabs: halo exchanges included), then apply the
            # recorded buffer permutation and halo state
            self._runtime().graph_launch(hit["exec"])
            for name, g in grids.items():
                g._restore_arrangement(hit["final"][name])
                if g.sharded:
                    g._restore_halo_state(hit["halo"][name])
            return hit["result"]

        launcher = Launcher(self, grids)
        record = key is not None and key in self._seen
        rt = self._runtime() if (self.groups or self.callees) else None
        if record:
            rt.graph_begin()
        try:
            result = _Interpreter(self.ir, env, grids, launcher, self.pairs).run(self.plan)
            if record and sharded:
                self._join_halo_events(grids)          # the communication stream re-joins the captured stream
        finally:
            if record:
                graph, nodes = rt.graph_end()
        self._launches_per_call = launcher.launches
        rtype = self.ir.signature.return_type
        if isinstance(rtype, Void) or result is None:
            result = None
        elif not isinstance(rtype, Structure):
            result = _to_py(coerce(rtype, result))
        if record:
            if len(self._graphs) >= 64:
                old_key, old = next(iter(self._graphs.items()))
                rt.graph_destroy(old["exec"])
                del self._graphs[old_key]
            self._graphs[key] = {"exec": graph, "nodes": nodes, "result": result,
                                 "final": {n: g._arrangement() for n, g in grids.items()},
                                 "halo": {n: g._halo_state() for n, g in grids.items() if g.sharded}}
            rt.graph_launch(graph)          # the capture only recorded the work
        elif key is not None:
            self._seen.add(key)
            if len(self._seen) > 4096:
                self._seen.clear()
        return result

    def _runtime(self):
        if _Runtime is None:
            _late_imports()
        return _Runtime.get()

    def _join_halo_events(self, grids: dict) -> None:
        """Order the compute stream behind every halo exchange still in flight on the communication stream."""
        rt = None
        for g in grids.values():
            if not g.sharded:
                continue
            for lv in [*g._ring, *([g._scratch] if g._scratch is not None else [])]:
                if lv.halo_event:
                    rt = rt or self._runtime()
                    rt.stream_wait_event(0, lv.halo_event)
                    lv.halo_event = 0

    def _run_batch2(self, args, grid, count: int) -> None:
        """2-D: `count` deferred identical calls, two time steps per pass (cudagen._emit_tiled2).
        A pass reads level 0 and writes u^{n+2} into the buffer of the dead level 1; only the last
        pass of the batch also stores the middle level u^{n+1} (into a spare buffer), because the
        intermediate ones are overwritten before anything can observe them."""
        g = self.groups[0]
        cfg = g.tiled2
        passes = count // 2
        done = 0
        if passes:
            env, grids = self._bind(args)
            captured = []
            _Interpreter(self.ir, env, grids, lambda grp, e: captured.append(dict(e))).run(self.plan)
            env = captured[0]
            rt = self._runtime()
            grid._extend_time(2)
            grid._prepare_device(cfg["ghost"])
            # every cell must be rewritten each step: all mask values present need a statement
            handled = {a.sweep.mask for a in g.stmts}
            present = {k for k in range(255) if grid._mask_count(k) > 0}
            ok = present <= handled
            if grid.sharded:
                from .. import dist
                ok = dist.transport().all_agree(ok)         # another slab may hold a value this one lacks
                self._join_halo_events(grids)
            if not ok:
                passes = 0
        if passes:
            fn = self.function(cudagen.kernel_name(g, cudagen.VARIANT_TILED2, cfg["V"]), cfg["smem"])
            P = g.params_cls()
            n0, cols = grid.shape[0], grid.shape[-1]
            for a, n in enumerate(grid.shape):
                setattr(P, f"n{a}", n)
            P.rows, P.cols = grid.size // cols, cols
            r_lo, r_hi, deep, edge = 0, n0, [], []
            if grid.sharded:
                # slab: the two-step pass covers the rows whose two-step cone stays inside the slab (one row
                # more than the taps say, because a tap that leaves its row at the first / last column lands in the
                # neighbouring row, F10); the rows next to a cut run step-at-a-time on row bands with the ordinary
                # kernels -- step 1 into the spare buffer (deep enough for step 2's taps), its halo exchanged by the
                # usual planner, step 2 into the output buffer -- on a side stream, beside the pass of the interior
                topo = dist.topology()
                r_lo, r_hi, deep, edge = two_step_slab_rows(n0, cfg["DMIN"], cfg["DMAX"], topo.lo_rank >= 0,
                                                            topo.hi_rank >= 0)
            P.r_lo, P.r_hi = r_lo, r_hi
            gname = g.slots[0].grid
            setattr(P, f"m_{gname}", grid._mask_dev if grid._mask_any else None)
            setattr(P, f"f_{gname}", grid._flags_dev if grid._mask_any else None)
            from .launch import Launcher, STATS, TUNE
            marshal = Launcher(self, grids)
            for name, t in g.scalars.items():
                setattr(P, f"u_{name}", marshal._scalar_value(t, env[name]))
            gx = (cols + cfg["W"] - 1) // cfg["W"]
            gj = 1 if grid.dimension == 2 else (grid.shape[1] + cfg["TJ"] - 1) // cfg["TJ"]
            want = max(1, -(-TUNE["min_ctas"] // (gx * gj)))
            chunk0 = max(64, -(-(r_hi - r_lo) // want))
            chunks = (r_hi - r_lo + chunk0 - 1) // chunk0
            P.chunk0 = chunk0
            if grid.dimension == 2:
                gy = min(chunks, 65535)
                geometry = ((gx, gy, (chunks + gy - 1) // gy), (cfg["threads"], 1, 1))
            else:
                geometry = ((gx, gj, chunks), (cfg["threads"], 1, 1))
            for pi in range(passes):
                x0, x1 = grid._ring[0], grid._ring[1]
                last = pi == passes - 1
                P.aux0, P.aux1 = x0.dev, x1.dev
                if last or deep:
                    d = grid._spare_levels(1)[0]
                if last:
                    P.aux2, P.opt0 = d.dev, 1
                else:
                    P.aux2, P.opt0 = None, 0
                if deep:
                    side, fork, join = rt.side_stream()
                    rt.event_record_raw(fork, 0)
                    rt.stream_wait_event(side, fork)
                    try:
                        marshal.stream = side
                        grid._ring = [d, x0]
                        marshal(g, env, bands=deep)             # step 1 on the bands: u^n -> spare
                        grid._ring = [x1, d]
                        marshal(g, env, bands=edge)             # step 2 on the bands: spare -> u^{n+2}
                    finally:
                        marshal.stream = 0
                        grid._ring = [x0, x1]
                    rt.event_record_raw(join, side)
                rt.launch(fn, geometry[0], geometry[1], P, smem=cfg["smem"])
                STATS["tiled2"] = STATS.get("tiled2", 0) + 1
                if deep:
                    rt.stream_wait_event(0, join)               # the bands of the output buffer are complete
                if last:
                    grid._ring, grid._spares = [x1, d], [x0] + grid._spares[1:]
                else:
                    grid._ring = [x1, x0]
                for lv in grid._ring:
                    lv.where = "device"
                    lv.halo_rows = 0
                done += 2
        for _ in range(count - done):
            self._call_now(args)

    def _run_batch(self, args, grid, count: int) -> None:
        """`count` deferred identical calls.  While at least T remain, one launch of the
        multi-step kernel advances T time steps: it reads the two ring levels, iterates in
        shared memory and writes the two newest levels into spare buffers that then become
        the ring (T is even, so the ring order equals the order after T single ticks).  A
        remainder of at least 2*S steps runs through the tail variant of the same kernel
        (step count = the largest multiple of S, passed in ``opt0``); what is left after
        that (< S calls, or a run shorter than 2*S) runs step-at-a-time."""
        if grid.dimension >= 2:
            return self._run_batch2(args, grid, count)
        g = self.groups[0]
        cfg = g.multistep
        T = cfg["T"]
        launches, _ = multistep_launches(count, T, cfg["S"], MULTISTEP_TAIL)
        done = 0
        if launches:
            rt = self._runtime()
            grid._extend_time(2)
            # a slab needs the neighbours' next H points of BOTH ring levels (and of the mask)
            grid._prepare_device(cfg["H"] if grid.sharded else 1)
            from .launch import Launcher, STATS
            if grid.sharded:
                from .. import dist
            # the marshalled parameter struct of a run depends on the scalar arguments, the grid and its mask only:
            # a program that flushes short runs again and again (observing the field every few steps) re-uses it
            # instead of re-binding, re-evaluating the scalar prologue and re-marshalling (~50 us of host time that
            # the device would otherwise spend idle in front of a 150 us launch)
            try:
                ckey = (tuple(a for n, a in enumerate(args) if n != self._grid_pos), grid._serial, grid._mask_version,
                        grid._mask_dev, grid.shape)
                P = self._batch_params.get(ckey)
            except TypeError:                       # unhashable argument
                ckey, P = None, None
            if P is None:
                env, grids = self._bind(args)
                captured = []
                _Interpreter(self.ir, env, grids, lambda grp, e: captured.append(dict(e))).run(self.plan)
                env = captured[0]
                P = g.params_cls()
                P.n0 = grid.shape[0]
                P.rows, P.cols = 1, grid.shape[0]
                gname = g.slots[0].grid
                setattr(P, f"m_{gname}", grid._mask_dev if grid._mask_any else None)
                setattr(P, f"f_{gname}", grid._flags_dev if grid._mask_any else None)
                marshal = Launcher(self, grids)
                for name, t in g.scalars.items():
                    setattr(P, f"u_{name}", marshal._scalar_value(t, env[name]))
                if grid.sharded:
                    topo = dist.topology()
                    P.open_lo, P.open_hi = int(topo.lo_rank >= 0), int(topo.hi_rank >= 0)
                else:
                    # the level two steps back is only ever observed through points no statement writes
                    # (SURVEY.md F5); when every mask value present has a statement the kernel does not load it
                    # (a quarter of a launch's HBM traffic).  Slabs keep loading it: a neighbour may own such points.
                    handled = {a.sweep.mask for a in g.stmts}
                    present = {k for k in range(255) if grid._mask_count(k) > 0}
                    P.count = int(present <= handled)
                if ckey is not None:
                    if len(self._batch_params) > 64:
                        self._batch_params.clear()
                    self._batch_params[ckey] = P
            short = g.multistep_short
            for steps in launches:
                # a remainder of at most T/2 steps runs on the short variant's half-size windows
                mc = short if (steps != T and short is not None and steps <= short["T"]) else cfg
                variant = (cudagen.VARIANT_MULTISTEP if steps == T else
                           cudagen.VARIANT_MULTISTEP_SHORT if mc is short else cudagen.VARIANT_MULTISTEP_TAIL)
                fn = self.function(cudagen.kernel_name(g, variant, 1), mc["smem"])
                blocks = (grid.shape[0] + mc["W"] - 1) // mc["W"]
                x0, x1 = grid._ring[0], grid._ring[1]
                if grid.sharded:
                    stale = [(grid, lv, cfg["H"]) for lv in (x0, x1) if lv.halo_rows < cfg["H"]]
                    if stale:
                        dist.transport().exchange(stale)
                c, d = grid._spare_levels(2)
                P.aux0, P.aux1, P.aux2, P.aux3 = x0.dev, x1.dev, c.dev, d.dev
                P.opt0 = steps
                rt.launch(fn, (blocks, 1, 1), (mc["threads"], 1, 1), P, smem=mc["smem"])
                STATS["multistep"] = STATS.get("multistep", 0) + 1
                grid._ring, grid._spares = [c, d], [x0, x1]
                c.where = d.where = "device"
                c.halo_rows = d.halo_rows = 0
                done += steps
        for _ in range(count - done):
            self._call_now(args)

    def _graph_key(self, env: dict, grids: dict):
        """Everything a recorded call depends on: scalar argument values, the
        identity (device pointers) and order of every ring level and scratch
        buffer, the mask contents version and the ghost layout."""
        parts = []
        for name, t in self.ir.signature.arguments:
            if isinstance(t, GridT):
                # the grid's serial: a recorded graph also bakes in mask / index-list pointers, and a new
                # grid can get the very addresses a dead one had (caching pool, or cudaMalloc itself)
                parts.append(grids[name]._arrangement() + (grids[name]._mask_version, grids[name].shape,
                                                           grids[name]._serial))
                if grids[name].sharded:             # which exchanges a call issues depends on what is stale
                    parts.append(grids[name]._halo_state())
            elif isinstance(t, Pointer):
                parts.append(_deref(env[name]))
            elif isinstance(t, Structure):
                parts.append(astuple(env[name]))
            else:
                parts.append(env[name].item())
        return tuple(parts)
