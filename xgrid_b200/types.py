"""DSL type system and annotation surface.

Same vocabulary as the reference (xgrid/util/typing/{value,reference,
annotation}.py): ``Boolean``, ``Integer(width_bytes)``,
``Floating(width_bytes)``, ``Structure``, ``Pointer``, ``Grid`` plus the
``grid[T, N]`` / ``ptr[T]`` annotation markers -- but one flat module with
value-equality classes.  Every type knows three projections:

* ``cname``   -- the CUDA C spelling used by the code generator,
* ``np_dtype``-- host mirror dtype (xgrid/xgrid/__init__.py:10-18),
* ``ctype``   -- ctypes twin used to marshal kernel parameters; struct layout is
                 declaration order with natural C alignment
                 (xgrid/util/typing/value.py:94-97) which is also what nvcc
                 uses for the generated ``struct``.
"""
from __future__ import annotations

import ctypes
import struct as _struct
from dataclasses import fields as _dc_fields, is_dataclass
from typing import Any, Generic, TypeVar, get_args, get_origin

import numpy as np

from .config import get_config


class BaseType:
    """Root of the DSL type lattice."""

    cname = "void"

    def __eq__(self, other: object) -> bool:
        return type(self) is type(other) and self._key() == other._key()  # type: ignore[attr-defined]

    def __hash__(self) -> int:
        return hash((type(self).__name__, self._key()))

    def _key(self):
        return ()


class Void(BaseType):
    def __repr__(self) -> str:
        return "Void"


class Ignore(BaseType):
    """``typing.Any`` in an annotation: compares equal to everything
    (xgrid/util/typing/__init__.py:26-29)."""

    def __eq__(self, other: object) -> bool:
        return True

    __hash__ = BaseType.__hash__


class Value(BaseType):
    abbr = ""


class Boolean(Value):
    cname = "bool"
    abbr = "b"
    np_dtype = np.bool_
    ctype = ctypes.c_bool
    width_bytes = 1

    def __repr__(self) -> str:
        return "Boolean"


class Number(Value):
    def __init__(self, width_bytes: int) -> None:
        self.width_bytes = int(width_bytes)

    @property
    def width_bits(self) -> int:
        return self.width_bytes * 8

    def _key(self):
        return (self.width_bytes,)


class Integer(Number):
    _CT = {1: ctypes.c_int8, 2: ctypes.c_int16, 4: ctypes.c_int32, 8: ctypes.c_int64}
    _NP = {1: np.int8, 2: np.int16, 4: np.int32, 8: np.int64}

    def __init__(self, width_bytes: int) -> None:
        super().__init__(width_bytes)
        assert self.width_bytes in self._CT

    cname = property(lambda self: f"int{self.width_bits}_t")
    abbr = property(lambda self: f"i{self.width_bits}")
    np_dtype = property(lambda self: self._NP[self.width_bytes])
    ctype = property(lambda self: self._CT[self.width_bytes])

    def __repr__(self) -> str:
        return f"Integer({self.width_bits})"


class Floating(Number):
    def __init__(self, width_bytes: int) -> None:
        super().__init__(width_bytes)
        assert self.width_bytes in (4, 8)

    cname = property(lambda self: "float" if self.width_bytes == 4 else "double")
    abbr = property(lambda self: f"f{self.width_bits}")
    np_dtype = property(lambda self: np.float32 if self.width_bytes == 4 else np.float64)
    ctype = property(lambda self: ctypes.c_float if self.width_bytes == 4 else ctypes.c_double)

    def __repr__(self) -> str:
        return f"Floating({self.width_bits})"


class Structure(Value):
    """A Python ``@dataclass`` whose fields are all value types."""

    _ctype_cache: dict = {}

    def __init__(self, dataclass: type, name: str, elements: tuple) -> None:
        self.dataclass = dataclass
        self.name = name
        self.elements = tuple(elements)
        self.elements_map = dict(self.elements)

    def _key(self):
        return (self.name, self.elements)

    cname = property(lambda self: f"struct {self.name}")
    abbr = property(lambda self: f"st{self.name}")

    @property
    def np_dtype(self):
        return np.dtype([(n, t.np_dtype) for n, t in self.elements], align=True)

    @property
    def ctype(self):
        key = (self.name, self.elements)
        ct = Structure._ctype_cache.get(key)
        if ct is None:
            ct = type(f"st{self.name}", (ctypes.Structure,),
                      {"_fields_": [(n, t.ctype) for n, t in self.elements]})
            Structure._ctype_cache[key] = ct
        return ct

    def __repr__(self) -> str:
        return self.name


class Reference(BaseType):
    pass


class Pointer(Reference):
    def __init__(self, element: Value) -> None:
        self.element = element

    def _key(self):
        return (self.element,)

    cname = property(lambda self: f"{self.element.cname}*")

    @property
    def ctype(self):
        return ctypes.POINTER(self.element.ctype)

    def __repr__(self) -> str:
        return f"Pointer of {self.element!r}"


class Grid(Reference):
    """``grid[T, N]``: an N-dimensional time-ringed field of ``T``."""

    def __init__(self, element: Value, dimension: int) -> None:
        self.element = element
        self.dimension = int(dimension)

    def _key(self):
        return (self.element, self.dimension)

    @property
    def struct_name(self) -> str:
        return f"__Grid{self.dimension}d_{self.element.abbr}"

    def __repr__(self) -> str:
        return f"Grid({self.dimension}) of {self.element!r}"


# ---------------------------------------------------------------------------
# annotation markers: ``xgrid.grid[float, 2]`` and ``xgrid.ptr[int]``
# (xgrid/util/typing/annotation.py:21-28)
# ---------------------------------------------------------------------------
_L = TypeVar("_L")
_V = TypeVar("_V")


class Annotation:
    ...


class ptr(Annotation, Generic[_V]):
    def addr(self) -> int: ...


class grid(Annotation, Generic[_V, _L]):
    def __getitem__(self, key) -> Any: ...

    def __setitem__(self, key, value) -> Any: ...


C_INT = _struct.calcsize("i")


def parse_annotation(annotation, glbs: dict | None = None) -> BaseType | None:
    """Python annotation -> DSL type, ``None`` when it is not expressible.

    Follows xgrid/util/typing/annotation.py:31-68: ``None`` -> Void, ``Any`` ->
    Ignore, ``int`` is C ``int``, ``float`` follows ``init(precision=...)``,
    dataclasses become by-value structs, ``ptr[T]`` / ``grid[T, N]``.
    """
    if annotation is None:
        return Void()
    if isinstance(annotation, str):
        scope = glbs if glbs is not None else {}
        if annotation not in scope:
            return None
        annotation = scope[annotation]
    if annotation is Any:
        return Ignore()
    if annotation is int:
        return Integer(C_INT)
    if annotation is float:
        return Floating(get_config().fsize)
    if annotation is bool:
        return Boolean()
    if isinstance(annotation, type) and is_dataclass(annotation):
        elems = []
        for f in _dc_fields(annotation):
            t = parse_annotation(f.type, glbs)
            if not isinstance(t, Value):
                return None
            elems.append((f.name, t))
        return Structure(annotation, annotation.__name__, tuple(elems))
    origin, args = get_origin(annotation), get_args(annotation)
    if origin is None or len(args) not in (1, 2):
        return None
    elem = parse_annotation(args[0], glbs)
    if not isinstance(elem, Value):
        return None
    if origin is ptr and len(args) == 1:
        return Pointer(elem)
    if origin is grid and len(args) == 2 and type(args[1]) is int:
        return Grid(elem, args[1])
    return None
