"""``@xgrid.kernel`` / ``@xgrid.function`` / ``@xgrid.external``.

Same decorator surface and call contract as the reference
(xgrid/lang/operator.py:13-88): a kernel is parsed and compiled lazily on the
first call, every ``Grid`` argument is resized to the kernel's ring depth and
ticked (operator.py:37-39), then the body runs.  What runs is different: the
scalar prologue / control flow is evaluated on the host with C semantics and
every stencil statement becomes a launch of a generated sm_100a kernel
(see ``xgrid_b200.lang.schedule``).
"""
from __future__ import annotations

from typing import Any, Callable

from ..log import Logger
from ..types import BaseType

CustomTypecheck = Callable[[list], BaseType]


class Operator:
    def __init__(self, func, mode: str, name: str | None = None, includes: list | None = None,
                 self_type: BaseType | None = None, typecheck_override: CustomTypecheck | None = None,
                 tick: bool = True, macro: list | None = None) -> None:
        self.func = func
        self.mode = mode
        self.logger = Logger(self)
        self.name = func.__name__ if name is None else name
        self.includes = [] if includes is None else includes
        self.macro = [] if macro is None else macro
        self.self_type = self_type
        self.typecheck_override = typecheck_override
        self.tick = tick
        self.native = None      # the compiled Program (kept under the reference's attribute name)
        self.depth = 1
        self._ir = None
        self._epoch = -1

    def __call__(self, *args: Any) -> Any:
        if self.mode == "kernel":
            from ..config import config_epoch
            if self.native is None or self._epoch != config_epoch():
                from .schedule import Program
                self._ir = None
                self.native = Program(self)
                self.depth = self.native.depth
                self._epoch = config_epoch()
            return self.native(*args)
        if self.mode == "function":
            return self.func(*args)
        self.logger.dead(f"Invalid call to non-kernel or non-function ({self.mode}) operator '{self.name}'")

    @property
    def ir(self):
        if self._ir is None:
            from .frontend import Parser
            parser = Parser(self.func, self.name, self.mode, self.self_type)
            self._ir = parser.result
            self.includes.extend(parser.includes)
        return self._ir

    @property
    def src(self) -> str:
        """CUDA C text of every device kernel generated for this operator."""
        from .schedule import Program
        prog = self.native if self.native is not None else Program(self)
        return prog.source

    @property
    def signature(self):
        return self.ir.signature


def kernel(*, name: str | None = None, includes: list | None = None, tick: bool = True,
           macro: list | None = None):
    def wrap(func):
        return Operator(func, "kernel", name, includes, tick=tick, macro=macro)
    return wrap


def function(*, method: bool = False, name: str | None = None, includes: list | None = None,
             macro: list | None = None):
    if method:
        def mark(func):
            setattr(func, "__xgrid_method", (name, includes))
            return func
        return mark

    def wrap(func):
        return Operator(func, "function", name, includes, macro=macro)
    return wrap


def external(*, name: str | None = None, includes: list | None = None,
             typecheck_override: CustomTypecheck):
    def wrap(func):
        return Operator(func, "external", name, includes, typecheck_override=typecheck_override)
    return wrap
