"""``with xgrid.c(): "text"`` -- inline native code inside a kernel.

The reference pastes the text verbatim into the generated C function (xgrid/lang/parser.py:200-204,
xgrid/lang/generator.py:366-374 -> ``ir.Inline``), where it runs on the calling host thread between the
OpenMP loop nests and sees the function's locals and the by-value grid structs.  On the B200 backend the
text is **CUDA C executed by ONE device thread**, enqueued on the backend stream in program order between
the surrounding sweeps (so it sees the results of earlier statements and later statements see its writes):

* every grid argument ``g`` is visible as a struct with the reference's member names
  (generator.py:139-147): ``g.time`` (ring depth), ``g.shape[d]``, ``g.data[level]`` (device pointers,
  ``data[0]`` = now, C-order) and ``g.boundary_mask`` -- a ``const uint8_t*`` here (the device mask is
  compiled to bytes; null while the mask is all zero), not the reference's ``int32_t*``;
* every scalar / dataclass argument and local of the kernel is visible under its own name as a local
  variable initialised with its current value; assignments to them stay inside the inline block (scalar
  state lives on the host in this backend) -- write results through a grid;
* ``macro=[...]`` lines of the decorator are pasted in front of the generated kernels
  (generator.py:223-224); ``printf`` works (device printf).
"""
from __future__ import annotations

import ctypes

from ..types import Grid as GridT, Pointer
from . import ir

PRELUDE = """
template <class T, int D, int DEPTH> struct xgb_inline_grid {
    int32_t time;
    int32_t shape[D];
    T *data[DEPTH];
    const uint8_t *boundary_mask;
};
"""


def collect(body: list) -> list:
    return [s for s in ir.walk_stmts(body) if isinstance(s, ir.Inline)]


class InlineKernel:
    def __init__(self, name: str, stmt, cls, grids: list, scalars: list) -> None:
        self.name, self.stmt, self.params_cls, self.grids, self.scalars = name, stmt, cls, grids, scalars


def emit(tag: str, index: int, stmt, scope: dict, depth: int, module) -> tuple[str, InlineKernel]:
    """CUDA C of the single-thread kernel that runs one inline block, and its launch description."""
    name = f"xg_{tag}_inline{index}"
    c_fields, py_fields, setup = [], [], []
    grids, scalars = [], []
    for vname, var in scope.items():
        t = var.type
        if isinstance(t, GridT):
            T, D = module.ctype(t.element), t.dimension
            c_fields += [f"    {T} *d_{vname}[{depth}];", f"    int32_t shape_{vname}[{D}];",
                         f"    const uint8_t *m_{vname};"]
            py_fields += [(f"d_{vname}", ctypes.c_void_p * depth), (f"shape_{vname}", ctypes.c_int32 * D),
                          (f"m_{vname}", ctypes.c_void_p)]
            setup.append(f"    xgb_inline_grid<{T}, {D}, {depth}> {vname};")
            setup.append(f"    {vname}.time = {depth}; {vname}.boundary_mask = p.m_{vname};")
            setup.append(f"    for (int i = 0; i < {D}; ++i) {vname}.shape[i] = p.shape_{vname}[i];")
            setup.append(f"    for (int i = 0; i < {depth}; ++i) {vname}.data[i] = p.d_{vname}[i];")
            grids.append(vname)
        elif isinstance(t, Pointer):
            continue                    # host pointers have no meaning on the device
        else:
            ct = module.ctype(t)
            c_fields.append(f"    {ct} u_{vname};")
            py_fields.append((f"u_{vname}", t.ctype))
            setup.append(f"    {ct} {vname} = p.u_{vname}; (void){vname};")
            scalars.append((vname, t))
    if not c_fields:
        c_fields.append("    int32_t unused;")
        py_fields.append(("unused", ctypes.c_int32))
    cls = type(f"{name}_P", (ctypes.Structure,), {"_fields_": py_fields})
    text = [f"struct {name}_P {{", *c_fields, "};",
            f'extern "C" __global__ void {name}(const __grid_constant__ {name}_P p)', "{",
            "    if (blockIdx.x != 0 || threadIdx.x != 0) return;", *setup,
            f'#line {stmt.location.line} "{stmt.location.file}"',
            "    {", stmt.source, "    }", "}", ""]
    return "\n".join(text), InlineKernel(name, stmt, cls, grids, scalars)
