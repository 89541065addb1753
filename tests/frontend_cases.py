"""Front-end conformance cases: small kernels, valid and invalid.

Each case is module text that defines a kernel ``k`` (and optionally ``CALL``, the arguments of one call of a
scalar-only kernel).  ``tests/golden/make_frontend_golden.py`` runs every case through the UNMODIFIED reference
(``import xgrid``) and records accept / reject, the ring depth and the returned value;
``tests/test_frontend_conformance.py`` runs the same text through this backend's front end
(``import xgrid_b200 as xgrid``) and compares.  ``DEVIATIONS`` lists the cases where this backend
deliberately differs, with the reason.
"""

HEADER = """\
IMPORT_LINE
from dataclasses import dataclass

f1 = xgrid.grid[float, 1]
f2 = xgrid.grid[float, 2]
i1 = xgrid.grid[int, 1]
GLOBAL_K = 7
GLOBAL_X = 0.5


@dataclass
class Vec:
    x: float
    y: float

    @xgrid.function(method=True)
    def dot(self, o: "Vec") -> float:
        return self.x * o.x + self.y * o.y


@dataclass
class IV:
    a: int
    b: int


@xgrid.function()
def helper(a: float, b: float) -> float:
    return a * b + 1.0


def _ext_check(args):
    return args[0]


@xgrid.external(typecheck_override=_ext_check)
def ext_twice(a: float) -> float:
    ...


v1 = xgrid.grid[Vec, 1]
pf = xgrid.ptr[float]
pi = xgrid.ptr[int]

"""

CASES = {
    # ------------------------------------------------------------------ scalar kernels (callable on the host)
    "chain_compare": """
@xgrid.kernel()
def k(a: int, b: int, c: int) -> int:
    return 1 if a < b < c else 0
CALL = (1, 5, 9)
""",
    "chain_compare_false": """
@xgrid.kernel()
def k(a: int, b: int, c: int) -> int:
    return 1 if a < b <= c else 0
CALL = (1, 5, 4)
""",
    "bool_ops": """
@xgrid.kernel()
def k(a: int, b: int) -> int:
    r = 0
    if a > 0 and b > 0 or not a == b:
        r = 3
    return r
CALL = (-1, 2)
""",
    "bool_arg_and_return": """
@xgrid.kernel()
def k(a: bool, b: int) -> bool:
    return a and b > 2
CALL = (True, 3)
""",
    "while_break_continue": """
@xgrid.kernel()
def k(n: int) -> int:
    i = 0
    acc = 0
    while i < n:
        i += 1
        if i % 3 == 0:
            continue
        if i > 20:
            break
        acc += i
    return acc
CALL = (50,)
""",
    "for_range_step": """
@xgrid.kernel()
def k(n: int) -> int:
    acc = 0
    for i in range(1, n, 3):
        for j in range(0, 2):
            acc = acc + i * (j + 1)
    return acc
CALL = (20,)
""",
    "for_range_single_argument": """
@xgrid.kernel()
def k(n: int) -> int:
    acc = 0
    for i in range(n):
        acc += i
    return acc
CALL = (10,)
""",
    "augassign_scalars": """
@xgrid.kernel()
def k(a: float, n: int) -> float:
    x = a
    x += 1.5
    x *= 2.0
    x -= a
    x /= 4.0
    m = n
    m %= 5
    return x + cast(float, m)
from typing import cast
CALL = (2.25, 13)
""",
    "casts": """
from typing import cast
@xgrid.kernel()
def k(a: float, n: int) -> int:
    return cast(int, a) + cast(int, cast(float, n) * 0.5) + cast(int, -a)
CALL = (3.75, 7)
""",
    "pow_general": """
@xgrid.kernel()
def k(a: float, b: float) -> float:
    return a ** b + a ** 2.0 + a ** 3.0 + a ** 0.5
CALL = (1.7, 2.3)
""",
    "int_division_and_modulo": """
@xgrid.kernel()
def k(a: int, b: int) -> int:
    return (a / b) * 1000 + (a % b) * 10 + ((0 - a) / b) + ((0 - a) % b)
CALL = (17, 5)
""",
    "ternary_nested": """
@xgrid.kernel()
def k(a: float) -> float:
    return 1.0 if a > 1.0 else (2.0 if a > 0.0 else 3.0)
CALL = (0.5,)
""",
    "unary_ops": """
@xgrid.kernel()
def k(a: float, n: int) -> float:
    return -a + (+a) * 2.0 + cast(float, -n)
from typing import cast
CALL = (1.25, 3)
""",
    "dataclass_attr_and_method": """
@xgrid.kernel()
def k(p: Vec, q: Vec) -> float:
    return p.dot(q) + p.x - q.y
CALL = (Vec(1.5, -2.0), Vec(0.25, 4.0))
""",
    "dataclass_constructor_local": """
@xgrid.kernel()
def k(a: float) -> float:
    v = Vec(a, a * 2.0)
    w = Vec(1.0, 1.0)
    return v.dot(w)
CALL = (1.5,)
""",
    "dataclass_return": """
@xgrid.kernel()
def k(a: int, b: int) -> IV:
    return IV(a + b, a - b)
CALL = (7, 3)
""",
    "function_operator_call": """
@xgrid.kernel()
def k(a: float, b: float) -> float:
    return helper(a, b) + helper(b, a * 2.0)
CALL = (1.5, 2.5)
""",
    "global_constants": """
@xgrid.kernel()
def k(a: int) -> float:
    return cast(float, a + GLOBAL_K) * GLOBAL_X
from typing import cast
CALL = (3,)
""",
    "local_redeclaration_same_type": """
@xgrid.kernel()
def k(a: float) -> float:
    x = a
    x = x * 2.0
    y = x
    return y
CALL = (1.5,)
""",
    "void_return_and_pass": """
@xgrid.kernel()
def k(a: int) -> None:
    if a > 0:
        return
    pass
CALL = (1,)
""",
    "float_literal_int_context_rejected": """
@xgrid.kernel()
def k(a: int) -> int:
    return a + 1.0
""",
    "int_float_mix_rejected": """
@xgrid.kernel()
def k(a: float, n: int) -> float:
    return a + n
""",
    "assignment_type_change_rejected": """
@xgrid.kernel()
def k(a: float) -> float:
    x = 1
    x = a
    return a
""",
    "return_type_mismatch_rejected": """
@xgrid.kernel()
def k(a: float) -> int:
    return a
""",
    "missing_return_annotation_rejected": """
@xgrid.kernel()
def k(a: float):
    return
""",
    "missing_argument_annotation_rejected": """
@xgrid.kernel()
def k(a) -> None:
    return
""",
    "default_argument_ignored": """
@xgrid.kernel()
def k(a: int = 3) -> int:
    return a
""",
    "keyword_only_argument_rejected": """
@xgrid.kernel()
def k(a: int, *, b: int) -> int:
    return a + b
""",
    "while_else_rejected": """
@xgrid.kernel()
def k(n: int) -> int:
    i = 0
    while i < n:
        i += 1
    else:
        i = 0
    return i
""",
    "for_over_non_range_rejected": """
@xgrid.kernel()
def k(n: int) -> int:
    acc = 0
    for i in [1, 2, 3]:
        acc += i
    return acc
""",
    "tuple_assignment_rejected": """
@xgrid.kernel()
def k(n: int) -> int:
    a, b = n, n
    return a + b
""",
    "unknown_name_rejected": """
@xgrid.kernel()
def k(n: int) -> int:
    return n + nowhere_defined
""",
    "augassign_type_mismatch_rejected": """
@xgrid.kernel()
def k(n: int) -> int:
    m = n
    m += 1.5
    return m
""",
    "call_arity_rejected": """
@xgrid.kernel()
def k(a: float) -> float:
    return helper(a)
""",
    "import_alias_rejected": """
@xgrid.kernel()
def k(a: int) -> int:
    import math as m
    return a
""",
    "condition_not_bool": """
@xgrid.kernel()
def k(a: int) -> int:
    r = 0
    if a:
        r = 1
    return r
""",
    "modulo_on_floats": """
@xgrid.kernel()
def k(a: float, b: float) -> float:
    return a % b
""",
    # ------------------------------------------------------------------ grid kernels (accept / reject + ring depth)
    "stencil_defaults": """
@xgrid.kernel()
def k(u: f1, c: float) -> None:
    u[0] = u[0] - c * (u[0] - u[-1])
    with xgrid.boundary(1):
        u[0] = 1.0
""",
    "stencil_explicit_times": """
@xgrid.kernel()
def k(u: f1) -> None:
    u[0] = 0.5 * u[0][1] + 0.25 * u[1][-2] + 0.25 * u[-1][2]
""",
    "stencil_store_time_index": """
@xgrid.kernel()
def k(u: f1) -> None:
    u[0][0] = u[0][1] * 2.0
""",
    "stencil_store_nonzero_offset": """
@xgrid.kernel()
def k(u: f1) -> None:
    u[1] = u[0] * 2.0
""",
    "stencil_implicit_jacobi": """
@xgrid.kernel()
def k(p: f2) -> None:
    for _ in range(0, 3):
        p[0, 0] = 0.25 * (p[0, 1][0] + p[0, -1][0] + p[1, 0][0] + p[-1, 0][0])
        with xgrid.boundary(1):
            p[0, 0] = p[0, 1][0]
""",
    "stencil_two_grids": """
@xgrid.kernel()
def k(u: f2, v: f2, a: float) -> None:
    u[0, 0] = u[0, 0] + a * v[1, 0]
    v[0, 0] = v[0, 0] - a * u[0, -1][0]
""",
    "stencil_shape_dimension": """
@xgrid.kernel()
def k(u: f2) -> None:
    n = xgrid.shape(u, 0) + xgrid.shape(u, 1) + xgrid.dimension(u)
    u[0, 0] = cast(float, n) * u[0, 0]
from typing import cast
""",
    "stencil_int_grid": """
@xgrid.kernel()
def k(a: i1) -> None:
    a[0] = a[-1] + 4
""",
    "stencil_nested_boundary": """
@xgrid.kernel()
def k(u: f1) -> None:
    with xgrid.boundary(1):
        u[0] = 1.0
        with xgrid.boundary(2):
            u[0] = 2.0
        u[0] = 3.0
""",
    "stencil_under_scalar_control_flow": """
@xgrid.kernel()
def k(u: f1, n: int) -> None:
    i = 0
    while i < n:
        if i % 2 == 0:
            u[0] = u[0] * 0.5
        else:
            u[0] = u[1] * 0.25
        i += 1
""",
    "stencil_variable_offset_is_zero": """
@xgrid.kernel()
def k(u: f1, i: int) -> None:
    u[0] = u[i]
""",
    "stencil_expression_offset_is_zero": """
@xgrid.kernel()
def k(u: f2) -> None:
    u[0, 0] = u[1 + 1, GLOBAL_K]
""",
    "stencil_unary_plus_offset_rejected": """
@xgrid.kernel()
def k(u: f1) -> None:
    u[0] = u[+1]
""",
    "stencil_float_offset_rejected": """
@xgrid.kernel()
def k(u: f1) -> None:
    u[0] = u[1.0]
""",
    "stencil_wrong_rank_rejected": """
@xgrid.kernel()
def k(u: f1) -> None:
    u[0, 0] = 1.0
""",
    "stencil_load_outside_assignment_rejected": """
@xgrid.kernel()
def k(u: f1) -> float:
    return u[0]
""",
    "stencil_load_in_scalar_assignment_rejected": """
@xgrid.kernel()
def k(u: f1) -> None:
    x = u[0]
    u[0] = x
""",
    "stencil_boundary_float_rejected": """
@xgrid.kernel()
def k(u: f1) -> None:
    with xgrid.boundary(1.5):
        u[0] = 1.0
""",
    "stencil_boundary_variable_rejected": """
@xgrid.kernel()
def k(u: f1, m: int) -> None:
    with xgrid.boundary(m):
        u[0] = 1.0
""",
    "stencil_element_type_mismatch_rejected": """
@xgrid.kernel()
def k(u: f1) -> None:
    u[0] = 1
""",
    "stencil_grid_augassign": """
@xgrid.kernel()
def k(u: f1) -> None:
    u[0] += 1.0
""",
    "stencil_two_argument_boundary": """
@xgrid.kernel()
def k(u: f1) -> None:
    with xgrid.boundary(u, 1):
        u[0] = 1.0
""",
    "stencil_time_index_variable_rejected": """
@xgrid.kernel()
def k(u: f1, t: int) -> None:
    u[0] = u[0][t]
""",
    "struct_grid_load_fields": """
@xgrid.kernel()
def k(g: v1, u: f1) -> None:
    u[0] = g[0].x + g[1].y * 2.0
""",
    "struct_grid_store_whole_element": """
@xgrid.kernel()
def k(g: v1, u: f1) -> None:
    g[0] = Vec(u[0], u[-1] * 2.0)
""",
    "struct_grid_store_field_rejected": """
@xgrid.kernel()
def k(g: v1) -> None:
    g[0].x = 1.0
""",
    "pointer_read_and_write": """
@xgrid.kernel()
def k(p: pf, n: pi) -> float:
    p = p * 2.0 + 1.0
    n = n + 1
    return p
""",
    "pointer_type_mismatch_rejected": """
@xgrid.kernel()
def k(p: pf) -> None:
    p = 1
""",
    "external_call_with_typecheck_override": """
@xgrid.kernel()
def k(a: float) -> float:
    return ext_twice(a) + 1.0
""",
    "inline_c_block": """
@xgrid.kernel()
def k(u: f1, a: float) -> None:
    u[0] = u[0] * a
    with xgrid.c():
        "/* inline text */"
""",
    "inline_c_non_string_rejected": """
@xgrid.kernel()
def k(a: int) -> int:
    with xgrid.c():
        a = a + 1
    return a
""",
    "import_as_include": """
@xgrid.kernel()
def k(a: int) -> int:
    import stdio
    import sys.types
    return a
""",
    "tick_call_inside_kernel": """
@xgrid.kernel()
def k(u: f1) -> None:
    u[0] = u[0] * 0.5
    xgrid.tick(u)
""",
    "shape_of_non_grid_rejected": """
@xgrid.kernel()
def k(a: int) -> int:
    return xgrid.shape(a, 0)
""",
    "dimension_in_scalar_kernel": """
@xgrid.kernel()
def k(u: f2) -> int:
    return xgrid.dimension(u) * 10 + 1
""",
    "with_unknown_context_rejected": """
@xgrid.kernel()
def k(a: int) -> int:
    with open("x"):
        a = a + 1
    return a
""",
    "nested_function_call_in_stencil": """
@xgrid.kernel()
def k(u: f1, a: float) -> None:
    u[0] = helper(u[0], a) + helper(a, u[1])
""",
    "method_call_in_stencil": """
@xgrid.kernel()
def k(u: f1, p: Vec) -> None:
    u[0] = p.dot(Vec(u[0], u[-1]))
""",
    "grid_as_return_type_rejected": """
@xgrid.kernel()
def k(u: f1) -> f1:
    return u
""",
}

# Cases where this backend deliberately differs from the reference's front end.
DEVIATIONS = {
    "tick_call_inside_kernel": "the reference parses an in-kernel xgrid.tick(u) but the C it generates "
                               "(`extern void tick(None grid);`) fails to compile at the first call (SURVEY.md F9); "
                               "this backend rejects it at parse time with a message that says so",
    "stencil_two_argument_boundary": "the stale two-argument form of test.py:217 is accepted as an alias so that "
                                     "test.py runs unmodified (SURVEY.md §8f rank 2); the reference rejects it (F3)",
}


def source_of(name: str, import_line: str) -> str:
    return HEADER.replace("IMPORT_LINE", import_line) + CASES[name]
