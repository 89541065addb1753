"""Generic NumPy interpreter of xgrid kernels -- TEST INFRASTRUCTURE ONLY.

Executes ANY kernel the front end accepts, with the reference's execution
model (xgrid/lang/generator.py:285-364, xgrid/lang/operator.py:37-41):

* every ``HostGrid`` argument is resized to the kernel depth and ticked, once
  per argument;
* scalar statements / ``for`` / ``if`` / ``while`` run in program order;
* each stencil statement is ONE traversal of the whole grid, predicated on the
  stored grid's int32 mask; the right-hand side is evaluated for all points
  first and then assigned where the mask matches.  For "implicit" statements
  that is exactly the reference's scratch + barrier + copy-back; for explicit
  ones it is equivalent because they never read the level they write, except
  boundary statements reading their own level 0 at a *different* mask value;
* a tap is a linear offset into the padded level buffer (correct C-order
  strides).

It shares the parser / IR with the product (``xgrid_b200.lang``) but none of
its scheduling or code generation, and is itself pinned against the golden
vectors produced by the real reference (tests/test_interp.py).  Used by the
randomized differential tests (tests/test_random_gpu.py).
"""
from __future__ import annotations

import numpy as np

from xgrid_b200.lang import ir
from xgrid_b200.types import Boolean, Floating, Grid as GridT, Integer, Pointer, Structure

from . import HostGrid


def _np_dtype(t):
    if isinstance(t, Boolean):
        return np.bool_
    return np.dtype(t.np_dtype)


class _Return(Exception):
    def __init__(self, v):
        self.v = v


class _Break(Exception):
    pass


class _Continue(Exception):
    pass


class Interp:
    def __init__(self, operator) -> None:
        self.op = operator
        self.d = operator.ir
        self.depth = self.d.depth
        from xgrid_b200.config import get_config
        self.overstep = get_config().overstep             # "none" | "limit" | "wrap" (generator.py:172-177)

    # ------------------------------------------------------------------ call
    def __call__(self, *args):
        sig = self.d.signature.arguments
        assert len(args) == len(sig)
        self.env, self.grids = {}, {}
        for (name, t), a in zip(sig, args):
            if isinstance(t, GridT):
                assert isinstance(a, HostGrid)
                self.grids[name] = a
                a._op_invoke(self.depth, self.op.tick)        # once per argument (operator.py:37-39)
            elif isinstance(t, (Structure, Pointer)):
                self.env[name] = a
            else:
                self.env[name] = _np_dtype(t).type(a)
        try:
            self.block(self.d.body)
        except _Return as r:
            return r.v
        return None

    # ------------------------------------------------------------------ statements
    def block(self, stmts) -> None:
        for s in stmts:
            if isinstance(s, ir.Assignment):
                if s.sweep is not None:
                    self.sweep(s)
                else:
                    self.assign(s.terminal, self.ev(s.value, None))
            elif isinstance(s, ir.If):
                self.block(s.body if bool(self.ev(s.condition, None)) else s.orelse)
            elif isinstance(s, ir.While):
                while bool(self.ev(s.condition, None)):
                    try:
                        self.block(s.body)
                    except _Break:
                        break
                    except _Continue:
                        continue
            elif isinstance(s, ir.For):
                name, t = s.variable.name, _np_dtype(s.variable.type).type
                self.env[name] = t(self.ev(s.start, None))
                while bool(self.env[name] < self.ev(s.end, None)):
                    try:
                        self.block(s.body)
                    except _Break:
                        break
                    except _Continue:
                        pass
                    self.env[name] = t(self.env[name] + self.ev(s.step, None))
            elif isinstance(s, ir.Return):
                raise _Return(None if s.value is None else self.ev(s.value, None))
            elif isinstance(s, ir.Break):
                raise _Break()
            elif isinstance(s, ir.Continue):
                raise _Continue()
            elif isinstance(s, ir.Evaluation):
                self.ev(s.value, None)
            else:
                raise NotImplementedError(type(s).__name__)

    def assign(self, target, value) -> None:
        if isinstance(target, ir.Identifier):
            t = target.variable.type
            if isinstance(t, Pointer):
                self.env[target.variable.name].contents.value = value.item()
            elif isinstance(t, Structure):
                self.env[target.variable.name] = value
            else:
                self.env[target.variable.name] = _np_dtype(t).type(value)
        elif isinstance(target, ir.Access):
            setattr(self.ev(target.value, None), target.attribute, value.item())
        else:
            raise NotImplementedError

    def sweep(self, s: ir.Assignment) -> None:
        sw = s.sweep
        g: HostGrid = self.grids[sw.grid.name]
        value = self.ev(s.value, g)
        out = g._data[sw.store.level]
        sel = g.boundary == sw.mask
        if np.ndim(value) == 0:
            out[sel] = value
        else:
            out[sel] = np.asarray(value).astype(out.dtype, copy=False)[sel]

    # ------------------------------------------------------------------ expressions
    def tap(self, e: ir.Stencil, lead: HostGrid) -> np.ndarray:
        g = self.grids[e.variable.name]
        assert g.shape == lead.shape, "all grids of one statement must have the same shape"
        level = g._data[e.level]
        if self.overstep != "none":
            # per-axis clamped / wrapped coordinates, extents paired correctly
            idx = []
            for n, d in zip(g.shape, e.space_offset):
                i = np.arange(n) + d
                idx.append(np.clip(i, 0, n - 1) if self.overstep == "limit" else np.mod(i, n))
            return level[np.ix_(*idx)]
        raw = level.base                                  # the padded 1-D buffer
        off, stride = 0, 1
        for n, d in zip(reversed(g.shape), reversed(e.space_offset)):
            off += d * stride
            stride *= n
        assert abs(off) <= g._pad
        start = g._pad + off
        return raw[start:start + g.size].reshape(g.shape)

    def ev(self, e, lead):
        if isinstance(e, ir.Constant):
            v = e.value
            if type(v) is bool:
                return np.bool_(v)
            if type(v) is int:
                return np.int32(v)
            return np.float64(v)                           # unsuffixed literal = C double (F6)
        if isinstance(e, ir.Identifier):
            v = self.env[e.variable.name]
            if isinstance(e.variable.type, Pointer):
                return _np_dtype(e.variable.type.element).type(v.contents.value)
            return v
        if isinstance(e, ir.Access):
            return _np_dtype(e.type).type(getattr(self.ev(e.value, lead), e.attribute))
        if isinstance(e, ir.Stencil):
            assert lead is not None
            return self.tap(e, lead)
        if isinstance(e, ir.Unary):
            r = self.ev(e.right, lead)
            return np.logical_not(r) if e.operator == "!" else (-r if e.operator == "-" else +r)
        if isinstance(e, ir.Condition):
            return np.where(self.ev(e.condition, lead), self.ev(e.body, lead), self.ev(e.orelse, lead))
        if isinstance(e, ir.Cast):
            v = self.ev(e.value, lead)
            if isinstance(e.type, Integer):
                v = np.trunc(v) if np.issubdtype(np.asarray(v).dtype, np.floating) else v
            return np.asarray(v).astype(_np_dtype(e.type))[()]
        if isinstance(e, ir.GridInfo):
            g = self.grids[e.variable.name]
            if e.info == "dimension":
                return np.int32(len(g.shape))
            return np.int32(g.shape[int(self.ev(e.dimension, lead))])
        if isinstance(e, ir.Call):
            args = [self.ev(a, lead) for a in e.arguments]
            if isinstance(e.operator, ir.Constructor):
                return e.operator.type.dataclass(*[a.item() if isinstance(a, np.generic) else a for a in args])
            sub = Interp(e.operator)
            sub.env = {n: a for (n, _), a in zip(e.operator.ir.signature.arguments, args)}
            sub.grids = {}
            try:
                sub.block(e.operator.ir.body)
            except _Return as r:
                return r.v
            return None
        if isinstance(e, ir.Binary):
            op = e.operator
            a, b = self.ev(e.left, lead), self.ev(e.right, lead)
            with np.errstate(all="ignore"):
                if op == "+":
                    return a + b
                if op == "-":
                    return a - b
                if op == "*":
                    return a * b
                if op == "/":
                    if np.issubdtype(np.asarray(a).dtype, np.integer):
                        return np.fix(np.asarray(a, np.float64) / b).astype(np.asarray(a).dtype)[()]
                    return a / b
                if op == "%":
                    return np.fmod(a, b)
                if op == "^":
                    wide = isinstance(e.type, Floating) and e.type.width_bits == 64
                    x = np.asarray(a).astype(np.float64 if wide else np.float32)[()]
                    if isinstance(e.right, ir.Constant) and e.right.value == 2.0:
                        return x * x                     # gcc folds pow(x, 2.0) (F7)
                    return np.power(x, np.asarray(b).astype(x.dtype))
                if op == "&&":
                    return np.logical_and(a, b)
                if op == "||":
                    return np.logical_or(a, b)
                return {"==": np.equal, "!=": np.not_equal, ">": np.greater, ">=": np.greater_equal,
                        "<": np.less, "<=": np.less_equal}[op](a, b)
        raise NotImplementedError(type(e).__name__)
