"""Typed IR for xgrid kernels.

The node set carries the same information as the reference IR
(xgrid/lang/ir/{expression,statement}.py) -- in particular ``Stencil`` keeps
``variable / time_offset / space_offset / boundary_mask``
(expression.py:116-127) -- but it is a single flat module of slotted
dataclasses, and every stencil *statement* is annotated in place with the
``Sweep`` record the B200 scheduler works from (the counterpart of the
reference's ``StencilFlag``, xgrid/lang/generator.py:23-76).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Optional

from ..types import BaseType


@dataclass(frozen=True)
class Location:
    file: str
    func: str
    line: int

    def __repr__(self) -> str:
        return f"File {self.file}, line {self.line}, at {self.func}"


@dataclass
class Variable:
    name: str
    type: BaseType


# --------------------------------------------------------------------------- expressions
@dataclass
class Expression:
    location: Location
    type: BaseType


@dataclass
class Constant(Expression):
    value: Any


@dataclass
class Identifier(Expression):
    context: str  # "load" | "store"
    variable: Variable


@dataclass
class Access(Expression):
    context: str
    value: Expression
    attribute: str


@dataclass
class Stencil(Expression):
    """``g[d0, .., dn-1][t]`` -- a relative-offset access to a grid."""
    context: str
    variable: Variable
    time_offset: int
    space_offset: tuple
    boundary_mask: int = 0

    @property
    def level(self) -> int:
        # sign is dropped by the reference code generator
        # (xgrid/lang/generator.py:428,432): [-1] and [1] both mean "one back"
        return abs(self.time_offset)


@dataclass
class Unary(Expression):
    operator: str  # "+", "-", "!"
    right: Expression


@dataclass
class Binary(Expression):
    operator: str  # + - * / ^ % == != > >= < <= && ||
    left: Expression
    right: Expression


@dataclass
class Condition(Expression):
    condition: Expression
    body: Expression
    orelse: Expression


@dataclass
class Cast(Expression):
    value: Expression


@dataclass
class Signature:
    arguments: list
    return_type: BaseType

    def __post_init__(self) -> None:
        self.argnames_map = dict(self.arguments)


@dataclass
class Constructor:
    type: BaseType
    signature: Signature


@dataclass
class Call(Expression):
    operator: Any  # Operator | Constructor
    arguments: list


@dataclass
class GridInfo(Expression):
    info: str  # "shape" | "dimension"
    variable: Variable
    dimension: Optional[Expression]


# --------------------------------------------------------------------------- statements
@dataclass
class Statement:
    location: Location


@dataclass
class Sweep:
    """Scheduling facts about one stencil statement (one full-grid traversal
    in the reference, xgrid/lang/generator.py:285-364)."""
    grid: Variable               # the stored grid; its shape and mask drive the sweep
    mask: int                    # statement runs where grid.boundary == mask
    implicit: bool               # reads level 0 of the grid it stores (Jacobi, generator.py:67-68)
    loads: list = field(default_factory=list)   # every Stencil load in the RHS
    store: Optional[Stencil] = None


@dataclass
class Assignment(Statement):
    terminal: Expression
    value: Expression
    sweep: Optional[Sweep] = None


@dataclass
class Return(Statement):
    value: Optional[Expression]


@dataclass
class Break(Statement):
    pass


@dataclass
class Continue(Statement):
    pass


@dataclass
class If(Statement):
    condition: Expression
    body: list
    orelse: list


@dataclass
class While(Statement):
    condition: Expression
    body: list


@dataclass
class For(Statement):
    variable: Variable
    start: Expression
    end: Expression
    step: Expression
    body: list


@dataclass
class Evaluation(Statement):
    value: Expression


@dataclass
class Inline(Statement):
    source: str


@dataclass
class Definition(Statement):
    name: str
    mode: str
    signature: Signature
    scope: dict
    body: list
    depth: int = 1            # ring levels the kernel needs = max|t| + 1 (generator.py:108,428)

    def show(self, device=None) -> None:
        import sys
        (device or sys.stdout).write(dump(self))


# --------------------------------------------------------------------------- traversal helpers
def walk_expr(e):
    """Pre-order walk over an expression tree."""
    yield e
    if isinstance(e, Binary):
        yield from walk_expr(e.left)
        yield from walk_expr(e.right)
    elif isinstance(e, Unary):
        yield from walk_expr(e.right)
    elif isinstance(e, Condition):
        yield from walk_expr(e.condition)
        yield from walk_expr(e.body)
        yield from walk_expr(e.orelse)
    elif isinstance(e, Cast):
        yield from walk_expr(e.value)
    elif isinstance(e, Access):
        yield from walk_expr(e.value)
    elif isinstance(e, Call):
        for a in e.arguments:
            yield from walk_expr(a)
    elif isinstance(e, GridInfo) and e.dimension is not None:
        yield from walk_expr(e.dimension)


def walk_stmts(stmts):
    """Pre-order walk over statements, descending into control flow."""
    for s in stmts:
        yield s
        if isinstance(s, If):
            yield from walk_stmts(s.body)
            yield from walk_stmts(s.orelse)
        elif isinstance(s, (While, For)):
            yield from walk_stmts(s.body)


# --------------------------------------------------------------------------- pretty printer
def fmt_expr(e) -> str:
    if isinstance(e, Constant):
        return repr(e.value)
    if isinstance(e, Identifier):
        return "%" + e.variable.name
    if isinstance(e, Access):
        return f"{fmt_expr(e.value)}.{e.attribute}"
    if isinstance(e, Stencil):
        return f"%{e.variable.name}[{', '.join(map(str, e.space_offset))}][{e.time_offset}]"
    if isinstance(e, Unary):
        return f"({e.operator}{fmt_expr(e.right)})"
    if isinstance(e, Binary):
        return f"({fmt_expr(e.left)} {e.operator} {fmt_expr(e.right)})"
    if isinstance(e, Condition):
        return f"({fmt_expr(e.condition)} ? {fmt_expr(e.body)} : {fmt_expr(e.orelse)})"
    if isinstance(e, Cast):
        return f"({fmt_expr(e.value)} as {e.type!r})"
    if isinstance(e, Call):
        name = repr(e.operator.type) if isinstance(e.operator, Constructor) else e.operator.name
        return f"{name}({', '.join(fmt_expr(a) for a in e.arguments)})"
    if isinstance(e, GridInfo):
        extra = "" if e.dimension is None else ", " + fmt_expr(e.dimension)
        return f"{e.info}(%{e.variable.name}{extra})"
    return repr(e)


def _dump_block(stmts, out, ind):
    pad = "  " * ind
    for s in stmts:
        if isinstance(s, Assignment):
            tag = ""
            if s.sweep is not None:
                tag = f"   ; sweep over %{s.sweep.grid.name} where mask == {s.sweep.mask}" + \
                      (" (implicit)" if s.sweep.implicit else "")
            out.append(f"{pad}{fmt_expr(s.terminal)} : {s.terminal.type!r} = {fmt_expr(s.value)}{tag}")
        elif isinstance(s, Return):
            out.append(f"{pad}return {'' if s.value is None else fmt_expr(s.value)}")
        elif isinstance(s, Break):
            out.append(f"{pad}break")
        elif isinstance(s, Continue):
            out.append(f"{pad}continue")
        elif isinstance(s, If):
            out.append(f"{pad}if {fmt_expr(s.condition)} do")
            _dump_block(s.body, out, ind + 1)
            if s.orelse:
                out.append(f"{pad}else")
                _dump_block(s.orelse, out, ind + 1)
            out.append(f"{pad}end")
        elif isinstance(s, While):
            out.append(f"{pad}while {fmt_expr(s.condition)} do")
            _dump_block(s.body, out, ind + 1)
            out.append(f"{pad}end")
        elif isinstance(s, For):
            out.append(f"{pad}for %{s.variable.name} in {fmt_expr(s.start)} : {fmt_expr(s.end)} : {fmt_expr(s.step)}")
            _dump_block(s.body, out, ind + 1)
            out.append(f"{pad}end")
        elif isinstance(s, Evaluation):
            out.append(f"{pad}evaluate {fmt_expr(s.value)}")
        elif isinstance(s, Inline):
            out.append(f"{pad}inline begin")
            out.append(f"{pad}  {s.source}")
            out.append(f"{pad}end")


def dump(d: Definition) -> str:
    out = []
    args = ", ".join(f"{t!r} %{n}" for n, t in d.signature.arguments)
    out.append(f"{d.mode} {d.signature.return_type!r} {d.name}({args}) requires")
    for n, v in d.scope.items():
        if n not in d.signature.argnames_map:
            out.append(f"  %{n} : {v.type!r}")
    out.append("begin")
    _dump_block(d.body, out, 1)
    out.append("end")
    return "\n".join(out) + "\n"
