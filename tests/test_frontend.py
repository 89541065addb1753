"""Front end / scheduler facts, checked on CPU: the DSL surface of SURVEY.md
§8 a-1..a-3 (time defaults, masks, implicit detection, depth, errors), sweep
grouping, and that every workload lowers to CUDA C that NVRTC accepts for
sm_100a."""
import ctypes
from dataclasses import dataclass
from typing import cast

import numpy as np
import pytest

import xgrid_b200 as xgrid
from examples import workloads as W
from xgrid_b200.lang import ir
from xgrid_b200.lang.schedule import Program

TEMP = 10


@pytest.fixture(autouse=True)
def _init(tmp_path):
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"))


def sweeps(op):
    return [s for s in ir.walk_stmts(op.ir.body) if isinstance(s, ir.Assignment) and s.sweep is not None]


def test_time_defaults_and_depth():
    k = W.make_kernels()
    sw = sweeps(k["convection_1d"])
    assert [s.sweep.mask for s in sw] == [0, 1]
    store, loads = sw[0].sweep.store, sw[0].sweep.loads
    assert store.time_offset == 0 and all(l.time_offset == -1 for l in loads)      # parser.py:523
    assert sorted(l.space_offset for l in loads) == [(-1,), (0,), (0,)]
    assert k["convection_1d"].ir.depth == 2
    assert k["fill4"].ir.depth == 1                                                 # no loads -> one level
    assert not any(s.sweep.implicit for s in sw)


def test_explicit_time_sign_is_dropped():
    f1 = xgrid.grid[float, 1]

    @xgrid.kernel()
    def two_back(u: f1) -> None:
        u[0] = u[0][2] + u[1][-2]

    assert two_back.ir.depth == 3
    assert {l.level for l in sweeps(two_back)[0].sweep.loads} == {2}                # generator.py:428,432


def test_cavity_structure():
    k = W.make_kernels()
    sw = sweeps(k["cavity_kernel"])
    assert len(sw) == 16                      # b, p, 4 bc, (jacobi + 4 bc) in the loop, u, v, 3 bc
    implicit = [s for s in sw if s.sweep.implicit]
    assert len(implicit) == 1 and implicit[0].sweep.grid.name == "p"                # generator.py:67-68
    # boundary statements that read level 0 of their own grid are NOT implicit (mask != 0)
    neumann = [s for s in sw if s.sweep.mask in (1, 2, 3) and s.sweep.grid.name == "p"]
    assert neumann and not any(s.sweep.implicit for s in neumann)
    prog = Program(k["cavity_kernel"])
    kinds = [(len(g.stmts), g.implicit, g.sparse) for g in prog.groups]
    # b | p-first | 4 sparse BCs | Jacobi | 4 sparse BCs | fused tail (u, v, 3 Dirichlet BCs)
    assert kinds[0] == (1, False, False) and kinds[1] == (1, False, False)
    assert kinds[6] == (1, True, False)
    assert kinds[-1] == (5, False, False)
    assert sum(1 for kd in kinds if kd[2]) == 8


def test_fusion_rules():
    k = W.make_kernels()
    for name in ("convection_1d", "diffusion_1d", "convection_2d", "diffusion_2d", "heat_3d"):
        prog = Program(k[name])
        assert len(prog.groups) == 1 and len(prog.groups[0].stmts) == 2, name       # interior + Dirichlet BC fused
    f1 = xgrid.grid[float, 1]

    @xgrid.kernel()
    def dependent(u: f1, v: f1) -> None:
        u[0] = v[0] * 2.0
        v[0] = u[1][0]            # reads level 0 of u written above -> must not fuse

    assert len(Program(dependent).groups) == 2


def test_errors_are_plain_exceptions():
    f1 = xgrid.grid[float, 1]

    @xgrid.kernel()
    def bad_mix(u: f1, n: int) -> None:
        u[0] = u[0] + n                      # no implicit int -> float conversion (parser.py:374-376)

    with pytest.raises(Exception, match="Incompatible binary operator"):
        bad_mix.ir

    @xgrid.kernel()
    def bad_boundary(u: f1) -> None:
        with xgrid.boundary(1.5):
            u[0] = 1.0

    with pytest.raises(Exception, match="boundary"):
        bad_boundary.ir

    @xgrid.kernel()
    def load_outside(u: f1) -> float:
        return u[0]

    with pytest.raises(Exception, match="without stencil context"):
        load_outside.ir

    @xgrid.kernel()
    def wrong_rank(u: f1) -> None:
        u[0, 0] = 1.0

    with pytest.raises(Exception, match="subscript length"):
        wrong_rank.ir


def test_stale_two_argument_boundary_form_is_accepted():
    f1 = xgrid.grid[float, 1]

    @xgrid.kernel()
    def conv(u: f1, c: float) -> None:
        u[0] = u[0] - c * (u[0] - u[-1])
        with xgrid.boundary(u, 1):           # test.py:217 form
            u[0] = 1.0

    assert [s.sweep.mask for s in sweeps(conv)] == [0, 1]


def test_arity_is_type_error():
    k = W.make_kernels()
    with pytest.raises(TypeError):
        k["convection_1d"](xgrid.Grid((8,), float))


@dataclass
class Vector3i:
    x: int
    y: int
    z: int

    @xgrid.function(method=True)
    def dot(self, b: "Vector3i") -> int:
        return self.x * b.x + self.y * b.y + self.z * b.z


def test_scalar_kernels_match_reference(golden):
    """test.py:130-165 + a control-flow kernel; values from the real reference."""
    g = golden("scalars")

    @xgrid.kernel()
    def add3(a: int, b: int) -> int:
        return a + b + TEMP

    @xgrid.kernel()
    def vdot(a: Vector3i, b: Vector3i) -> int:
        return a.dot(b)

    @xgrid.kernel()
    def scal(a: float, b: float, n: int) -> float:
        acc = 0.0
        for i in range(0, n):
            if i % 2 == 0:
                acc = acc + a / b
            else:
                acc = acc - a * b ** 2.0
        return acc

    assert add3(123, 456) == int(g["add3"])
    assert vdot(Vector3i(1, -2, 3), Vector3i(4, 5, -6)) == int(g["vdot"])
    assert scal(1.7, 0.3, 9) == float(g["scal"])


@dataclass
class Pair2f:
    x: float
    y: float

    @xgrid.function(method=True)
    def bump(self, d: float) -> float:
        self.x = self.x + d          # the callee's own copy (C passes structs by value)
        return self.x


def test_struct_values_are_copied_like_c_structs():
    """Structs are passed, assigned and returned BY VALUE in the reference's generated C
    (xgrid/lang/generator.py:216-225, xgrid/util/typing/value.py:103-107).  Expected values were
    produced by running these three kernels through the unmodified reference (gcc 13.3, -O2):
    `alias` -> 1.0, `store` -> 7.0, `method` -> 4.5, and the caller's dataclass is never modified."""
    @xgrid.kernel()
    def alias(p: Pair2f) -> float:
        q = p
        q.x = 5.0
        return p.x

    @xgrid.kernel()
    def store(p: Pair2f) -> float:
        p.x = 7.0
        return p.x

    @xgrid.kernel()
    def method(p: Pair2f) -> float:
        t = p.bump(2.5)
        return t + p.x

    p = Pair2f(1.0, 2.0)
    assert alias(p) == 1.0 and p == Pair2f(1.0, 2.0)
    assert store(p) == 7.0 and p == Pair2f(1.0, 2.0)
    assert store(p) == 7.0 and p == Pair2f(1.0, 2.0)       # a repeated (replayable) call behaves the same
    assert method(p) == 4.5 and p == Pair2f(1.0, 2.0)


def test_c_integer_semantics():
    @xgrid.kernel()
    def idiv(a: int, b: int) -> int:
        return a / b + a % b

    assert idiv(-7, 2) == -3 + -1            # C truncation and dividend-signed remainder
    assert idiv(7, -2) == -3 + 1

    @xgrid.kernel()
    def casts(x: float) -> int:
        return cast(int, x) + cast(int, 0.0 - x)

    assert casts(2.75) == 2 + -2


def test_ptr_arguments():
    @xgrid.kernel()
    def bump(p: xgrid.ptr[int], by: int) -> None:
        p = p + by

    v = ctypes.c_int32(5)
    bump(ctypes.pointer(v), 3)
    assert v.value == 8


def test_every_workload_compiles_for_sm100a():
    for name, op in W.make_kernels().items():
        prog = Program(op)
        if not prog.groups:
            continue
        image = prog.image()
        assert image[:4] == b"\x7fELF", name
        for g in prog.groups:
            assert ctypes.sizeof(g.params_cls) <= 4096      # by-value kernel parameter limit


def test_fp32_mode_keeps_double_literals():
    xgrid.init(cacheroot=".xgrid_test_f32")          # precision="float"
    k = W.make_kernels()
    src = Program(k["diffusion_1d"]).source
    assert "float" in src and "2.0 *" in src and "2.0f" not in src
    assert "xgb::sq<float>" in src                                      # ** 2.0 -> product (F7), powf type
    import shutil
    shutil.rmtree(".xgrid_test_f32", ignore_errors=True)


def test_ring_semantics_on_host_only():
    """Ring rotation / truncation are host bookkeeping: no GPU needed."""
    g = xgrid.Grid((6,), float)
    g.now[:] = np.arange(6)
    g._op_invoke(2, True)
    assert len(g._ring) == 2 and not g.now.any() and np.array_equal(g._data[1], np.arange(6))
    g._op_invoke(3, True)
    assert len(g._ring) == 3
    g._op_invoke(1, False)
    assert len(g._ring) == 1


def test_struct_field_store_is_rejected_like_the_reference():
    """`g[0].x = ...`: Python gives the inner subscript a Load context, so the reference's
    StencilParser (xgrid/lang/generator.py:56-65) dies with this very message; same here."""
    from dataclasses import dataclass
    xgrid.init(precision="double", cacheroot=".xgrid")

    @dataclass
    class V2:
        x: float
        y: float

    globals()["V2"] = V2
    g1 = xgrid.grid[V2, 1]

    @xgrid.kernel()
    def store_field(v: g1) -> None:
        v[0].x = 1.0

    with pytest.raises(Exception, match="without stencil context"):
        Program(store_field)


def test_multistep_launch_plan():
    """A deferred run of 1-D calls = full T-step launches + one tail launch (multiple of S, >= 2*S) + single steps;
    every launch advances an even number of steps (ring order) and nothing is lost."""
    from xgrid_b200.lang.schedule import multistep_launches
    assert multistep_launches(10000 - 2 * 4096, 64, 4) == ([64] * 28 + [16], 0)
    assert multistep_launches(20, 64, 4) == ([20], 0)
    assert multistep_launches(70, 64, 4) == ([64], 6)            # 6 < 2*S: the remainder runs step-at-a-time
    assert multistep_launches(7, 64, 4) == ([], 7)
    assert multistep_launches(150, 64, 4, tail=False) == ([64, 64], 22)
    for T, S in ((64, 4), (32, 4), (16, 2), (6, 2)):
        for count in range(0, 200):
            steps, left = multistep_launches(count, T, S)
            assert sum(steps) + left == count and all(x % 2 == 0 and S * 2 <= x <= T and x % S == 0 for x in steps)
            assert left < 2 * S + S and steps.count(T) == count // T


def test_multistep_tail_variant_is_generated():
    prog = Program(W.make_kernels()["convection_1d"])
    src = prog.source
    assert "xg_convection_1d_g0_multistep_v1(" in src and "xg_convection_1d_g0_multistep_tail_v1(" in src
    assert "round < (int)p.opt0 / S" in src and "round < T / S" in src


def test_device_code_is_the_gpu_validated_one(tmp_path):
    """The CUDA C generated for every workload kernel, the template headers and the NVRTC flags are the ones
    that last went through the GPU suite on a B200 (tests/golden/codegen_hashes.json).  A session without GPU
    access must not change device code; a session with it regenerates the file after `pytest -m gpu`
    (tests/golden/make_codegen_hashes.py)."""
    import importlib.util
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_codegen_hashes", os.path.join(here, "golden", "make_codegen_hashes.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with open(os.path.join(here, "golden", "codegen_hashes.json")) as f:
        want = json.load(f)
    got = mod.hashes(str(tmp_path / "xg"))
    changed = sorted(k for k in set(want) | set(got) if want.get(k) != got.get(k))
    assert not changed, f"device code changed since its last GPU validation: {changed}"


def test_external_operator_compiles_against_the_named_cuda_header(tmp_path, monkeypatch):
    """`@xgrid.external(includes=[...])`: the reference emits an `extern` prototype and compiles the C file
    in (generator.py:96-97,208-212); here the header holds a __device__ function and is handed to NVRTC."""
    monkeypatch.chdir(tmp_path)
    (tmp_path / "bump.h").write_text("__device__ __forceinline__ double bump(double x, double k) { return x * x + k; }\n")
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"))

    @xgrid.external(includes=["bump.h"], typecheck_override=lambda args: args[0])
    def bump(x: float, k: float) -> float:
        ...

    @xgrid.external(typecheck_override=lambda args: args[0])
    def nowhere(x: float) -> float:
        ...

    f1 = xgrid.grid[float, 1]

    @xgrid.kernel()
    def apply(u: f1, k: float) -> None:
        u[0] = bump(u[0], k) + u[-1]

    @xgrid.kernel()
    def broken(u: f1) -> None:
        u[0] = nowhere(u[0])

    @xgrid.kernel()
    def scalar_use(k: float) -> float:
        return bump(k, k)

    assert '#include "bump.h"' in apply.src
    for names, image in apply._program().images():            # NVRTC, sm_100a, with the user header
        assert image[:4] == b"\x7fELF"
    with pytest.raises(Exception, match="includes="):
        broken.src
    with pytest.raises(Exception, match="CUDA header"):
        scalar_use(1.0)
