"""Dataclass shapes for the type-layout conformance check: by-value structs must have
the reference's C layout (declaration order, natural alignment) and struct-element grids its NumPy record dtype."""
from dataclasses import dataclass


@dataclass
class A:
    x: int
    y: float


@dataclass
class B:
    flag: bool
    n: int
    ok: bool
    v: float
    last: bool


@dataclass
class C:
    a: A
    b: bool
    c: A


@dataclass
class D:
    f: float
    g: float
    h: float


@dataclass
class E:
    b1: bool
    b2: bool
    b3: bool


@dataclass
class F:
    inner: B
    tail: bool
    c: C


CASES = {"A": A, "B": B, "C": C, "D": D, "E": E, "F": F}


def describe(parse_annotation, parse_numpy_dtype, precision_tag: str) -> dict:
    """{case: {"size", "align", "offsets", "np_itemsize", "np_offsets"}} using one implementation's type system."""
    import ctypes
    import numpy as np
    out = {}
    for name, cls in CASES.items():
        t = parse_annotation(cls)
        ct = t.ctype
        dt = np.dtype(parse_numpy_dtype(t))
        out[f"{precision_tag}.{name}"] = {
            "size": ctypes.sizeof(ct), "align": ctypes.alignment(ct),
            "offsets": [getattr(ct, f).offset for f, _ in ct._fields_],
            "np_itemsize": dt.itemsize, "np_offsets": [dt.fields[f][1] for f in dt.names],
        }
    for name, ann in (("int", int), ("float", float), ("bool", bool)):
        t = parse_annotation(ann)
        out[f"{precision_tag}.{name}"] = {"size": ctypes.sizeof(t.ctype), "np": str(np.dtype(parse_numpy_dtype(t)))}
    return out
