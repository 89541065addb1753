"""`with xgrid.c()` inline text (xgrid/lang/parser.py:200-204): CUDA C run by one device thread in
program order between the sweeps (xgrid_b200/lang/inlinec.py)."""
from dataclasses import dataclass

import numpy as np
import pytest

import xgrid_b200 as xgrid
from xgrid_b200.lang.schedule import Program


@dataclass
class Pt:
    i: int
    w: float


def make():
    f1 = xgrid.grid[float, 1]

    @xgrid.kernel(macro=["#define TWICE(x) ((x) + (x))"])
    def poke(u: f1, a: float, at: Pt) -> None:
        u[0] = u[0] + a                       # sweep 1: level 0 = previous + a
        k = at.i + 1
        with xgrid.c():
            "u.data[0][k] = TWICE(at.w) + a + (double)u.shape[0] + (double)u.time;"
            "u.data[0][0] = u.data[1][0];"
        u[0] = u[0][0] * 2.0                  # sweep 2 (implicit): sees what the inline block wrote

    return poke


def test_inline_block_is_generated_and_compiles(tmp_path):
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"))
    prog = Program(make())
    assert len(prog.inlines) == 1
    src = prog.source
    assert "xgb_inline_grid<double, 1, 2> u;" in src and "#define TWICE(x)" in src
    assert "Pt at = p.u_at;" in src and "int32_t k = p.u_k;" in src
    assert prog.image()[:4] == b"\x7fELF"


@pytest.mark.gpu
def test_inline_block_runs_in_program_order(tmp_path):
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"))
    poke = make()
    n = 1000
    u = xgrid.Grid((n,), float)
    ic = np.arange(n, dtype=np.float64)
    u.now[...] = ic
    for call in range(3):                     # third call replays the recorded CUDA graph
        prev = u.now.copy()
        poke(u, 0.5, Pt(6, 1.25))
        want = prev + 0.5
        want[7] = 2 * 1.25 + 0.5 + n + 2
        want[0] = prev[0]
        want *= 2.0
        assert np.array_equal(u.now, want), call
