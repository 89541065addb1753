"""``@xgrid.kernel`` / ``@xgrid.function`` / ``@xgrid.external``.

Decorator surface and call contract of the reference (xgrid/lang/operator.py:13-88): a kernel
is parsed and compiled lazily on its first call, every ``Grid`` argument is resized to the
kernel's ring depth and ticked (operator.py:37-39), then the body runs.  What runs differs: the
compiled object is a ``schedule.Program`` -- host-evaluated scalar code plus generated sm_100a
sweep kernels -- and calls are asynchronous (1-D time loops are even deferred and batched).
"""
from __future__ import annotations

from typing import Any, Callable

from .. import config as _config
from ..log import Logger
from ..types import BaseType

CustomTypecheck = Callable[[list], BaseType]
_MODES = ("kernel", "function", "external")


class Operator:
    """A decorated DSL function.  ``mode`` selects how a call is served:
    kernel -> compiled Program, function -> the Python body itself, external -> declaration only."""

    def __init__(self, func, mode: str, name: str | None = None, includes: list | None = None,
                 self_type: BaseType | None = None, typecheck_override: CustomTypecheck | None = None,
                 tick: bool = True, macro: list | None = None) -> None:
        assert mode in _MODES
        self.func, self.mode = func, mode
        self.name = name or func.__name__
        self.includes = list(includes or [])
        self.macro = list(macro or [])
        self.self_type, self.typecheck_override, self.tick = self_type, typecheck_override, tick
        self.logger = Logger(self)
        self.native = None          # compiled Program (attribute name kept from the reference)
        self.depth = 1
        self._ir = None
        self._epoch = -1            # xgrid.init() generation the Program was built under

    # ---- introspection -------------------------------------------------------------------
    @property
    def ir(self):
        if self._ir is None:
            from .frontend import Parser
            parsed = Parser(self.func, self.name, self.mode, self.self_type)
            self._ir = parsed.result
            self.includes.extend(h for h in parsed.includes if h not in self.includes)
        return self._ir

    @property
    def signature(self):
        return self.ir.signature

    @property
    def src(self) -> str:
        """CUDA C text of every device kernel generated for this operator."""
        return self._program().source

    # ---- execution -----------------------------------------------------------------------
    def _program(self):
        from ..config import config_epoch
        if self.native is None or self._epoch != config_epoch():
            from .schedule import Program
            self._ir = None                      # types depend on init(precision=...)
            self.native = Program(self)
            self.depth = self.native.depth
            self._epoch = config_epoch()
        return self.native

    def __call__(self, *args: Any) -> Any:
        if self.mode == "kernel":
            prog = self.native
            if prog is None or self._epoch != _config._epoch:
                prog = self._program()
            return prog(*args)
        if self.mode == "function":
            return self.func(*args)
        self.logger.dead(f"Invalid call to non-kernel or non-function ({self.mode}) operator '{self.name}'")


def _wrap(mode: str, **options):
    def decorate(func):
        return Operator(func, mode, **options)
    return decorate


def kernel(*, name: str | None = None, includes: list | None = None, tick: bool = True,
           macro: list | None = None):
    return _wrap("kernel", name=name, includes=includes, tick=tick, macro=macro)


def function(*, method: bool = False, name: str | None = None, includes: list | None = None,
             macro: list | None = None):
    if not method:
        return _wrap("function", name=name, includes=includes, macro=macro)

    def mark(func):
        # methods of dataclasses stay plain Python callables; the front end builds the
        # Operator when it meets a call on a struct-typed receiver (frontend.e_Call)
        setattr(func, "__xgrid_method", (name, includes))
        return func
    return mark


def external(*, name: str | None = None, includes: list | None = None,
             typecheck_override: CustomTypecheck):
    return _wrap("external", name=name, includes=includes, typecheck_override=typecheck_override)
