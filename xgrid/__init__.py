"""``import xgrid`` -- the reference's public module name, served by the B200 backend.

Existing programs (`README.md:17`, `test.py:10-14`, `examples/cavity.py:4`) say ``import xgrid`` and reach a few
sub-modules by their dotted names; this package makes both resolve to ``xgrid_b200`` so such programs run
unchanged.  Nothing is implemented here: every name is the very object the backend defines (the front end
recognises ``xgrid.boundary`` / ``xgrid.c`` / ``xgrid.shape`` by identity), and the reference's module tree is
registered in ``sys.modules`` as thin namespaces over the backend's modules:

    xgrid.lang, xgrid.lang.operator, xgrid.lang.ir          -> xgrid_b200.lang[.operator|.ir]
    xgrid.lang.parser                                       -> xgrid_b200.lang.frontend  (Parser)
    xgrid.xgrid                                             -> xgrid_b200.grid           (Grid)
    xgrid.util.init / .logging / .console / .typing[.*]     -> config / log / a Console sink / types
    xgrid.util.ffi                                          -> Compiler / Library raise: the gcc JIT they
                                                               wrapped is replaced by the C ABI (no CPU path)
"""
import sys as _sys
import types as _types

import xgrid_b200 as _backend
from xgrid_b200 import *            # noqa: F401,F403  (kernel, function, init, ptr, grid, boundary, c, ...)
from xgrid_b200 import __all__ as _names, config as _config, log as _log, types as _t
from xgrid_b200.grid import Grid as _GridClass
from xgrid_b200.lang import frontend as _frontend, ir as _ir, operator as _operator
import xgrid_b200.lang as _lang

__all__ = list(_names)
__version__ = _backend.__version__


class Console:
    """File-like sink with the reference console's `print` / `println` (xgrid/util/console.py:25-49, no colours)."""

    def __init__(self, textio) -> None:
        self.io = textio
        self.tty = bool(getattr(textio, "isatty", lambda: False)())

    def isatty(self) -> bool:
        return self.tty

    def print(self, msg: str, style=None, foreground=None) -> "Console":
        self.io.write(msg)
        return self

    def println(self, msg: str, style=None, foreground=None) -> "Console":
        return self.print(msg + "\n")


def _gone(name: str):
    class _Gone:
        def __init__(self, *a, **k) -> None:
            raise Exception(f"xgrid.util.ffi.{name} wrapped the reference's gcc JIT; the B200 backend compiles with "
                            "NVRTC through libxgrid_b200.so (include/xgrid_b200.h) and has no CPU path")
    _Gone.__name__ = name
    return _Gone


def _namespace(name: str, **members):
    mod = _types.ModuleType(name)
    mod.__dict__.update(members)
    _sys.modules[name] = mod
    parent, _, leaf = name.rpartition(".")
    setattr(_sys.modules[parent], leaf, mod)
    return mod


_sys.modules[__name__ + ".lang"] = _lang
lang = _lang
_sys.modules[__name__ + ".lang.operator"] = _operator
_sys.modules[__name__ + ".lang.ir"] = _ir
_sys.modules[__name__ + ".lang.parser"] = _frontend
_namespace(__name__ + ".xgrid", Grid=_GridClass)
_namespace(__name__ + ".util")
_namespace(__name__ + ".util.init", init=_config.init, get_config=_config.get_config,
           Configuration=_config.Configuration)
_namespace(__name__ + ".util.logging", Logger=_log.Logger, LogLevel=_log.LogLevel)
_namespace(__name__ + ".util.console", Console=Console)
_namespace(__name__ + ".util.ffi", Compiler=_gone("Compiler"), Library=_gone("Library"))
_typing = {k: getattr(_t, k) for k in ("BaseType", "Void", "Value", "Boolean", "Number", "Integer", "Floating",
                                        "Structure", "Reference", "Pointer", "Grid", "ptr", "grid", "parse_annotation")}
_namespace(__name__ + ".util.typing", **_typing)
for _leaf in ("value", "reference", "annotation"):
    _namespace(f"{__name__}.util.typing.{_leaf}", **_typing)
