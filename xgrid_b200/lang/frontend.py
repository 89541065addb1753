"""Front end: Python ``ast`` of a decorated function -> typed IR.

Accepts exactly the DSL surface of the reference parser
(xgrid/lang/parser.py:124-641) and reports errors
the same way (``Logger.dead`` -> ``Exception``).  Differences, all deliberate:

* stencil statements are tagged *here* with their ``Sweep`` record (the
  reference does it in a separate pass, xgrid/lang/generator.py:23-76) and the
  ring depth is computed here (generator.py:428);
* ``with xgrid.boundary(u, k)`` -- the stale two-argument form used by the
  reference's own test.py:217 -- is accepted as an alias of ``boundary(k)``;
* free variables are also looked up in the function's closure, not only in its
  module globals.
"""
from __future__ import annotations

import ast
import inspect
import textwrap
import typing
from functools import reduce

from ..config import get_config
from ..log import Logger
from ..types import (BaseType, Boolean, C_INT, Floating, Grid as GridT, Integer, Number, Pointer,
                     Reference, Structure, Value, Void, parse_annotation)
from . import ir

_UNARY = {ast.UAdd: "+", ast.USub: "-", ast.Not: "!"}
_BINARY = {
    ast.Add: "+", ast.Sub: "-", ast.Mult: "*", ast.Div: "/", ast.Pow: "^", ast.Mod: "%",
    ast.Eq: "==", ast.NotEq: "!=", ast.Gt: ">", ast.GtE: ">=", ast.Lt: "<", ast.LtE: "<=",
    ast.And: "&&", ast.Or: "||",
}


def _ctx(node) -> str:
    return "load" if isinstance(node.ctx, ast.Load) else "store"


def _int_literal(node):
    """``3`` or ``-3`` -> int, anything else -> None (parser.py:483-503)."""
    if isinstance(node, ast.Constant) and type(node.value) is int:
        return node.value
    if (isinstance(node, ast.UnaryOp) and isinstance(node.op, ast.USub)
            and isinstance(node.operand, ast.Constant) and type(node.operand.value) is int):
        return -node.operand.value
    return None


class Parser:
    def __init__(self, func, name: str, mode: str, self_type: BaseType | None) -> None:
        self.logger = Logger(self)
        self.func, self.name, self.mode, self.self_type = func, name, mode, self_type
        self.file = inspect.getsourcefile(func) or "<unknown>"

        lines, first_line = inspect.getsourcelines(func)
        source = textwrap.dedent("".join(line.expandtabs(4) for line in lines))
        tree = ast.parse(source, self.file).body[0]
        ast.increment_lineno(tree, first_line)

        self.scope: dict[str, ir.Variable] = {}
        self.args: list = []
        self.includes: list[str] = []
        self.mask = 0               # value of the enclosing ``with boundary(k)``
        self.in_c = False           # inside ``with xgrid.c()``
        self.depth = 0              # max |time offset| seen
        self.loads: list | None = None   # Stencil loads of the statement being parsed

        self.lookup = dict(func.__globals__)
        try:
            self.lookup.update(inspect.getclosurevars(func).nonlocals)
        except Exception:
            pass
        self.lookup.update({"int": int, "float": float, "bool": bool})

        if not isinstance(tree, ast.FunctionDef):
            self.error(tree, "Only plain function definitions can be compiled")
        self.result = self.function(tree)

    # ------------------------------------------------------------------ infrastructure
    def error(self, node, message: str) -> typing.NoReturn:
        line = getattr(node, "lineno", 1) - 1
        self.logger.dead(f"File {self.file}, line {line}, in {self.name}",
                         f"  Syntax error: {message}")

    def loc(self, node) -> ir.Location:
        return ir.Location(self.file, self.name, getattr(node, "lineno", 1) - 1)

    def stmt(self, node):
        handler = getattr(self, "s_" + type(node).__name__, None)
        if handler is None:
            self.error(node, f"Python syntax '{type(node).__name__}' is currently unsupported")
        return handler(node)

    def block(self, nodes) -> list:
        out = []
        for n in nodes:
            r = self.stmt(n)
            if isinstance(r, list):
                out.extend(r)
            elif r is not None:
                out.append(r)
        return out

    def expr(self, node) -> ir.Expression:
        handler = getattr(self, "e_" + type(node).__name__, None)
        if handler is None:
            self.error(node, f"Python syntax '{type(node).__name__}' is currently unsupported")
        return handler(node)

    # ------------------------------------------------------------------ definition
    def function(self, node: ast.FunctionDef) -> ir.Definition:
        sig = inspect.signature(self.func)
        for pname, p in sig.parameters.items():
            if p.kind != inspect.Parameter.POSITIONAL_OR_KEYWORD:
                self.error(node, f"Argument '{pname}' of kind '{p.kind}' is not supported")
            if p.annotation is inspect.Parameter.empty:
                if self.self_type is None:
                    self.error(node, f"Argument '{pname}' requires type annotation")
                ptype = self.self_type
            else:
                ptype = parse_annotation(p.annotation, self.lookup)
            if ptype is None or isinstance(ptype, Void):
                self.error(node, f"Argument '{pname}' requires non-void type annotation ({p.annotation})")
            self.scope[pname] = ir.Variable(pname, ptype)
            self.args.append((pname, ptype))

        ret = None if sig.return_annotation is inspect.Signature.empty \
            else parse_annotation(sig.return_annotation, self.lookup)
        if ret is None or isinstance(ret, Reference):
            self.error(node, f"Invalid return type '{ret}'")
        self.return_type = ret

        body = [] if self.mode == "external" else self.block(node.body)
        return ir.Definition(self.loc(node), self.name, self.mode, ir.Signature(self.args, ret),
                             self.scope, body, depth=self.depth + 1)

    # ------------------------------------------------------------------ statements
    def s_Return(self, node: ast.Return):
        value = None if node.value is None else self.scalar_expr(node.value)
        if value is not None and value.type != self.return_type:
            self.error(node, f"Incompatible return type '{value.type}' with '{self.return_type}'")
        return ir.Return(self.loc(node), value)

    def s_Pass(self, node):
        return []

    def s_Break(self, node):
        return ir.Break(self.loc(node))

    def s_Continue(self, node):
        return ir.Continue(self.loc(node))

    def s_If(self, node: ast.If):
        cond = self.scalar_expr(node.test)
        return ir.If(self.loc(node), cond, self.block(node.body), self.block(node.orelse))

    def s_While(self, node: ast.While):
        cond = self.scalar_expr(node.test)
        if node.orelse:
            self.error(node, "While statement does not support else clause")
        return ir.While(self.loc(node), cond, self.block(node.body))

    def s_For(self, node: ast.For):
        if not isinstance(node.target, ast.Name):
            self.error(node, "For loop variable should be a name")
        it = node.iter
        if not (isinstance(it, ast.Call) and isinstance(it.func, ast.Name) and it.func.id == "range"):
            self.error(node, "For loop only supports range")
        if len(it.args) not in (2, 3):
            self.error(node, "For loop requires start:end:step or start:end")
        bounds = [self.scalar_expr(a) for a in it.args]
        if len(bounds) == 2:
            bounds.append(ir.Constant(self.loc(node), Integer(C_INT), 1))
        want = bounds[2].type
        for b in bounds:
            if b.type != want or not isinstance(b.type, Number):
                self.error(node, f"Incompatible loop range type '{b.type}'")
        var = self.scope.setdefault(node.target.id, ir.Variable(node.target.id, want))
        if var.type != want:
            self.error(node, f"Incompatible loop variable type '{var.type}' with range type '{want}'")
        return ir.For(self.loc(node), var, bounds[0], bounds[1], bounds[2], self.block(node.body))

    def s_Expr(self, node: ast.Expr):
        if self.in_c:
            if isinstance(node.value, ast.Constant) and isinstance(node.value.value, str):
                return ir.Inline(self.loc(node), node.value.value)
            return None
        return ir.Evaluation(self.loc(node), self.scalar_expr(node.value))

    def s_Import(self, node: ast.Import):
        for alias in node.names:
            if alias.asname is not None:
                self.error(node, f"Using import as include requires no alias for '{alias.name}'")
            self.includes.append(alias.name.replace(".", "/") + ".h")
        return []

    def s_With(self, node: ast.With):
        from .. import lang
        if len(node.items) != 1:
            self.error(node, "Only one pragma switch at once")
        call = node.items[0].context_expr
        if not isinstance(call, ast.Call):
            self.error(node, f"Invalid pragma switch '{call}'")
        target = self.resolve_global(call.func)
        if target is lang.c:
            prev, self.in_c = self.in_c, True
            try:
                body = self.block(node.body)
            finally:
                self.in_c = prev
            # the reference pastes every string into ONE C scope: consecutive strings form one block
            merged: list = []
            for st in body:
                if isinstance(st, ir.Inline) and merged and isinstance(merged[-1], ir.Inline):
                    merged[-1] = ir.Inline(merged[-1].location, merged[-1].source + "\n" + st.source)
                else:
                    merged.append(st)
            return merged
        if target is lang.boundary:
            args = call.args
            if len(args) == 2:      # stale ``boundary(u, k)`` form (test.py:217) -> alias
                args = args[1:]
            if len(args) != 1 or not isinstance(args[0], ast.Constant) or type(args[0].value) is not int:
                self.error(node, "Invalid pragram switch 'boundary")
            self.mask = args[0].value
            try:
                return self.block(node.body)
            finally:
                self.mask = 0
        self.error(node, f"Unknown pragma switch '{target}'")

    def s_Assign(self, node: ast.Assign):
        where = self.loc(node)
        self.loads = []
        try:
            value = self.expr(node.value)
            if isinstance(value.type, GridT):
                self.error(node, "Incompatible assignment to grid type")
            if len(node.targets) != 1:
                self.error(node, "Multiple assignment is not supported")
            tnode = node.targets[0]
            target = self.resolve_local(tnode)
            loads = self.loads
        finally:
            self.loads = None
        if target is None:
            if not isinstance(tnode, ast.Name):
                self.error(node, f"Undefined identifier {target}")
            var = ir.Variable(tnode.id, value.type)
            self.scope[tnode.id] = var
            target = ir.Identifier(where, value.type, "store", var)
        if target.type != value.type:
            self.error(node, f"Incompatible assignment from type {value.type} to {target.type}")

        sweep = None
        if isinstance(target, ir.Stencil):
            implicit = target.boundary_mask == 0 and any(
                l.time_offset == 0 and l.variable is target.variable and l.boundary_mask == 0 for l in loads)
            sweep = ir.Sweep(target.variable, target.boundary_mask, implicit, loads, target)
        elif loads:
            self.error(node, f"Unable to perform load operation to grid '{loads[0].variable.name}' "
                             "without stencil context")
        return ir.Assignment(where, target, value, sweep)

    def s_AugAssign(self, node: ast.AugAssign):
        value = self.scalar_expr(node.value)
        target = self.resolve_local(node.target)
        if target is None:
            self.error(node, f"Undefined identifier {target}")
        if isinstance(target, ir.Stencil):
            # the reference trips an assert here
            self.error(node, "Augmented assignment to a grid is not supported")
        op = _BINARY.get(type(node.op))
        if op is None:
            self.error(node, f"Unsupported binary operator '{type(node.op).__name__}'")
        if not isinstance(target.type, Number) or not isinstance(value.type, Number) or target.type != value.type:
            self.error(node, f"Incompatible binary operator '{op}' with type '{target.type}' and '{value.type}'")
        if op == "^":
            rtype = Floating(8) if isinstance(target.type, Floating) and value.type.width_bits == 64 \
                else Floating(4)
        else:
            rtype = target.type
        where = self.loc(node)
        load = ir.Identifier(where, target.type, "load", target.variable) \
            if isinstance(target, ir.Identifier) else target
        return ir.Assignment(where, target, ir.Binary(where, rtype, op, load, value))

    # ------------------------------------------------------------------ expressions
    def scalar_expr(self, node) -> ir.Expression:
        """Expression outside a stencil statement: grid loads are illegal
        (xgrid/lang/generator.py:63-65)."""
        outer, self.loads = self.loads, []
        try:
            e = self.expr(node)
            if self.loads:
                self.error(node, f"Unable to perform load operation to grid "
                                 f"'{self.loads[0].variable.name}' without stencil context")
            return e
        finally:
            self.loads = outer

    def constant(self, node, value) -> ir.Constant:
        t = type(value)
        if t is bool:
            return ir.Constant(self.loc(node), Boolean(), value)
        if t is int:
            return ir.Constant(self.loc(node), Integer(C_INT), value)
        if t is float:
            return ir.Constant(self.loc(node), Floating(get_config().fsize), value)
        self.error(node, f"Incompatible constant '{value}' of type '{t}'")

    def e_Constant(self, node: ast.Constant):
        return self.constant(node, node.value)

    def e_UnaryOp(self, node: ast.UnaryOp):
        op = _UNARY.get(type(node.op))
        if op is None:
            self.error(node, f"Unsupported unary operator '{type(node.op).__name__}'")
        right = self.expr(node.operand)
        ok = isinstance(right.type, Boolean) if op == "!" else isinstance(right.type, Number)
        if not ok:
            self.error(node, f"Incompatible unary operator '{op}' with type '{right.type}'")
        return ir.Unary(self.loc(node), right.type, op, right)

    def e_BinOp(self, node: ast.BinOp):
        op = _BINARY.get(type(node.op))
        if op is None:
            self.error(node, f"Unsupported binary operator '{type(node.op).__name__}'")
        left, right = self.expr(node.left), self.expr(node.right)
        if not isinstance(left.type, Number) or not isinstance(right.type, Number) or left.type != right.type:
            self.error(node, f"Incompatible binary operator '{op}' with type '{left.type}' and '{right.type}'")
        if op == "^":   # parser.py:378-382
            rtype = Floating(8) if isinstance(left.type, Floating) and left.type.width_bits == 64 \
                else Floating(get_config().fsize)
        else:
            rtype = left.type
        return ir.Binary(self.loc(node), rtype, op, left, right)

    def e_BoolOp(self, node: ast.BoolOp):
        op = _BINARY[type(node.op)]
        where = self.loc(node)
        vals = [self.expr(v) for v in node.values]
        for v in vals:
            if not isinstance(v.type, Boolean):
                self.error(node, f"Incompatible boolean operator '{op}' with '{v.type}'")
        return reduce(lambda a, b: ir.Binary(where, Boolean(), op, a, b), vals)

    def e_Compare(self, node: ast.Compare):
        where = self.loc(node)
        left = self.expr(node.left)
        if not isinstance(left.type, Number):
            self.error(node, f"Incompatible compare expression with type '{left.type}'")
        rights = [self.expr(c) for c in node.comparators]
        ops = [_BINARY[type(o)] for o in node.ops]
        for o, r in zip(ops, rights):
            if r.type != left.type:
                self.error(node, f"Incompatible compare operator '{o}' with type '{left.type}' and '{r.type}'")
        # parser.py:409-411: every comparison is against the *first* operand
        parts = [ir.Binary(where, Boolean(), o, left, r) for o, r in zip(ops, rights)]
        return reduce(lambda a, b: ir.Binary(where, Boolean(), "&&", a, b), parts)

    def e_IfExp(self, node: ast.IfExp):
        cond = self.expr(node.test)
        if not isinstance(cond.type, Boolean):
            self.error(node, f"Incompatible condition type '{cond.type}' of if expression")
        a, b = self.expr(node.body), self.expr(node.orelse)
        if a.type != b.type or not isinstance(a.type, Value):
            self.error(node, f"Incompatible type '{a.type}' and '{b.type}' of if expression")
        return ir.Condition(self.loc(node), a.type, cond, a, b)

    def e_Name(self, node):
        local = self.resolve_local(node)
        return local if local is not None else self.constant(node, self.resolve_global(node))

    e_Attribute = e_Name

    def e_Subscript(self, node):
        return self.resolve_local(node)

    # ------------------------------------------------------------------ name resolution
    def resolve_global(self, node):
        chain = []
        cur = node
        while True:
            if isinstance(cur, ast.Name):
                chain.append(cur.id)
                break
            if isinstance(cur, ast.Attribute):
                chain.append(cur.attr)
                cur = cur.value
            else:
                self.error(cur, f"Python syntax '{type(cur).__name__}' is currently unsupported")
        obj, seen = self.lookup, ["globals"]
        for attr in reversed(chain):
            try:
                obj = obj[attr] if isinstance(obj, dict) else getattr(obj, attr)
            except (KeyError, AttributeError):
                self.error(node, f"Undefined attribute '{attr}' of '{'.'.join(seen)}'")
            seen.append(attr)
        return obj

    def stencil(self, sub: ast.Subscript, time_offset: int) -> ir.Stencil:
        if not isinstance(sub.value, ast.Name):
            self.error(sub, f"Incompatible subscript to '{type(sub.value).__name__}'")
        var = self.scope.get(sub.value.id)
        if var is None:
            self.error(sub, f"Undefined identifier '{sub.value.id}'")
        if not isinstance(var.type, GridT):
            self.error(sub, f"Incompatible subscript to type '{var.type}'")
        parts = sub.slice.elts if isinstance(sub.slice, ast.Tuple) else [sub.slice]
        offsets = []
        for p in parts:
            v = _int_literal(p)
            if v is None:
                if isinstance(p, (ast.UnaryOp, ast.Constant)):
                    self.error(sub, f"Incompatible subscript '{ast.dump(p)}'")
                # parser.py:483-503 only inspects unary-minus and constant nodes; any other expression
                # (a name, ``1 + 1`` ...) falls through with offset 0.  Same result here, but say so.
                self.logger.warn(f"File {self.file}, line {getattr(sub, 'lineno', 1) - 1}, in {self.name}",
                                 f"  subscript '{ast.unparse(p)}' of grid '{var.name}' is not an integer literal; "
                                 "like the reference, it is taken as offset 0")
                v = 0
            offsets.append(v)
        if len(offsets) != var.type.dimension:
            self.error(sub, f"Incompatible subscript length '{len(offsets)}' with dimension {var.type.dimension}")
        node = ir.Stencil(self.loc(sub), var.type.element, _ctx(sub), var, time_offset,
                          tuple(offsets), self.mask)
        self.depth = max(self.depth, abs(time_offset))
        if node.context == "load":
            if self.loads is None:
                self.error(sub, f"Unable to perform load operation to grid '{var.name}' without stencil context")
            self.loads.append(node)
        return node

    def resolve_local(self, node):
        if isinstance(node, ast.Subscript):
            if isinstance(node.value, ast.Subscript):     # g[...][t]
                t = _int_literal(node.slice)
                if t is None:
                    self.error(node.value, "Invalid time dimension subscript")
                if _ctx(node) == "store":
                    # the inner subscript of a store target has a *load* context in the Python AST, so the
                    # reference ends up with a load outside a stencil statement (generator.py:56-65): a store
                    # cannot name its time level -- it always writes level 0
                    name = node.value.value.id if isinstance(node.value.value, ast.Name) else "?"
                    self.error(node, f"Unable to perform load operation to grid '{name}' without stencil context")
                return self.stencil(node.value, t)
            # default time level: store -> 0, load -> -1 (parser.py:523)
            return self.stencil(node, 0 if _ctx(node) == "store" else -1)
        if isinstance(node, ast.Attribute):
            base = self.resolve_local(node.value)
            if base is None or not isinstance(base.type, Structure) or node.attr not in base.type.elements_map:
                return None
            return ir.Access(self.loc(node), base.type.elements_map[node.attr], _ctx(node), base, node.attr)
        if isinstance(node, ast.Name):
            var = self.scope.get(node.id)
            if var is None:
                return None
            t = var.type.element if isinstance(var.type, Pointer) else var.type
            return ir.Identifier(self.loc(node), t, _ctx(node), var)
        self.error(node, f"Python syntax '{type(node).__name__}' is currently unsupported")

    # ------------------------------------------------------------------ calls
    def e_Call(self, node: ast.Call):
        from .operator import Operator

        func = None
        fname = None
        self_type = None
        args: list = []
        recv = None
        if isinstance(node.func, ast.Attribute):
            recv = self.resolve_local(node.func.value)
        if recv is not None:
            if not isinstance(recv.type, Structure):
                self.error(node, f"Invalid method call on type {recv.type}")
            func = getattr(recv.type.dataclass, node.func.attr, None)
            fname = f"{recv.type.dataclass.__qualname__}.{node.func.attr}"
            self_type = recv.type
            args.append(recv)
        else:
            func = self.resolve_global(node.func)

        if func is typing.cast:      # cast(T, e) -> C cast (parser.py:582-589)
            if len(node.args) != 2:
                self.error(node, f"Cast requires 2 arguments, got {len(node.args)}")
            ttype = parse_annotation(self.resolve_global(node.args[0]), self.lookup)
            if ttype is None:
                self.error(node, "Invalid cast type")
            return ir.Cast(self.loc(node), ttype, self.expr(node.args[1]))

        args.extend(self.expr(a) for a in node.args)

        if isinstance(func, type):   # dataclass constructor (parser.py:594-602)
            st = parse_annotation(func, self.lookup)
            if not isinstance(st, Structure):
                self.error(node, f"Invalid type constructor '{func.__name__}'")
            func = ir.Constructor(st, ir.Signature(list(st.elements), st))
            fname = f"{st.name}.constructor"

        if not isinstance(func, (Operator, ir.Constructor)):
            flag = getattr(func, "__xgrid_method", None)
            if flag is not None and self_type is not None:
                func = Operator(func, "function", flag[0], flag[1], self_type)
            else:
                self.error(node, f"Invalid call to object '{func}', it is not an operator")
        if fname is None:
            fname = func.name
        if isinstance(func, Operator) and func.mode == "external" and func.name == "tick" \
                and func.func.__module__.split(".")[0] == __name__.split(".")[0]:
            # the reference parses this call but its generated C (`extern void tick(None grid);`)
            # never compiles, so no working program contains it.  Say so instead of emitting dead code.
            self.error(node, "xgrid.tick inside a kernel is a dead feature of the reference (its generated C does "
                             "not compile); every Grid argument is ticked once by the kernel call itself")

        want = func.signature.arguments
        if len(want) != len(args):
            self.error(node, f"Operator '{fname}' requires {len(want)} arguments, but got {len(args)}")

        if isinstance(func, Operator) and func.mode == "external" and func.typecheck_override is not None:
            try:
                rtype = func.typecheck_override([a.type for a in args])
            except Exception as exc:
                self.error(node, exc.args[0])
            if func.name in ("shape", "dimension"):
                if not isinstance(args[0], ir.Identifier):
                    self.error(node, f"Incompatible '{func.name}' to argument '{args[0]}'")
                return ir.GridInfo(self.loc(node), Integer(C_INT), func.name, args[0].variable,
                                   args[1] if func.name == "shape" else None)
        else:
            for (aname, atype), given in zip(want, args):
                if atype != given.type:
                    self.error(node, f"Incompatible type '{given.type}' with '{atype}' of argument "
                                     f"'{aname}, operator '{fname}'")
            rtype = func.signature.return_type
            if isinstance(func, Operator):
                self.depth = max(self.depth, func.ir.depth - 1)
        return ir.Call(self.loc(node), rtype, func, args)
