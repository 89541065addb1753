"""The C-ABI shared library loads on a machine without a GPU and exports every
symbol include/xgrid_b200.h declares; GPU entry points fail loudly (status +
xgb_last_error), never silently."""
import ctypes
import os
import re

import pytest

from xgrid_b200.runtime import shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "xgrid_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xgb_[a-z0-9_]+)\s*\(", text)))


def test_header_and_library_agree():
    names = declared_symbols()
    assert len(names) >= 45
    lib = ctypes.CDLL(shim.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/xgrid_b200.h but not exported"
    assert set(names) == set(shim.SIGNATURES), set(names) ^ set(shim.SIGNATURES)


def test_abi_version_and_loud_failure_without_init():
    l = shim.lib()
    assert l.xgb_abi_version() == 1
    p = ctypes.c_void_p()
    status = l.xgb_alloc(1024, ctypes.byref(p))
    import torch
    if not torch.cuda.is_available():
        assert status != 0 and b"xgb_init" in l.xgb_last_error()


def test_nvrtc_cross_compile_and_error_text():
    img, log = shim.compile_cuda('extern "C" __global__ void k(float* p) { p[threadIdx.x] = 1.f; }', "k.cu",
                                 ["--gpu-architecture=sm_100a"], {})
    assert img[:4] == b"\x7fELF"
    with pytest.raises(Exception, match="failed to compile"):
        shim.compile_cuda("innt main() {}", "bad.cu", ["--gpu-architecture=sm_100a"], {})


def test_generated_kernels_use_bulk_copy_engine(tmp_path):
    """SASS evidence that the tiled variant is Blackwell-native: UBLKCP (cp.async.bulk)
    + SYNCS (mbarrier) appear in the sm_100a cubin."""
    import subprocess
    import xgrid_b200 as xgrid
    from examples import workloads as W
    from xgrid_b200.lang.schedule import Program
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"))
    img = Program(W.make_kernels()["heat_3d"]).image()
    cubin = tmp_path / "heat3d.cubin"
    cubin.write_bytes(img)
    sass = subprocess.run(["cuobjdump", "-sass", str(cubin)], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "SYNCS" in sass


def test_no_generated_workload_kernel_spills(tmp_path):
    """Every variant of every workload kernel fits its registers: no local-memory spills in the sm_100a cubins
    (the small stacks belong to the out-of-line slow path of the exact division)."""
    import subprocess
    import xgrid_b200 as xgrid
    from examples import workloads as W
    from xgrid_b200.lang.schedule import Program
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"))
    seen = 0
    for name, op in W.make_kernels().items():
        prog = Program(op)
        if not prog.source:
            continue
        cubin = tmp_path / f"{name}.cubin"
        cubin.write_bytes(prog.image())
        out = subprocess.run(["cuobjdump", "--dump-resource-usage", str(cubin)], capture_output=True, text=True).stdout
        for m in re.finditer(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", out):
            regs, stack, _, local = map(int, m.groups())
            seen += 1
            assert local == 0 and regs <= 255 and stack <= 64, (name, m.group(0))
    assert seen >= 100


def test_per_kernel_units_give_the_same_machine_code_as_one_unit(tmp_path):
    """Lazy JIT compiles every generated kernel as its own NVRTC unit; the SASS of the hot kernels must be
    instruction for instruction what the whole program compiled as one unit gives."""
    import subprocess
    import xgrid_b200 as xgrid
    from examples import workloads as W
    from xgrid_b200.lang.schedule import Program
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"))

    def sass(path, fn):
        out = subprocess.run(["cuobjdump", "-sass", "-fun", fn, str(path)], capture_output=True, text=True).stdout
        return [re.sub(r"/\*[0-9a-f]+\*/", "", line).strip() for line in out.splitlines()
                if re.match(r"\s+/\*[0-9a-f]{4}\*/", line)]

    for workload, fn in (("convection_1d", "xg_convection_1d_g0_multistep_v1"), ("heat_3d", "xg_heat_3d_g0_tiled_v2")):
        prog = Program(W.make_kernels()[workload])
        whole, unit = tmp_path / "whole.cubin", tmp_path / "unit.cubin"
        whole.write_bytes(prog.image())
        prog._units()
        unit.write_bytes(prog._unit_image(prog._where[fn]))
        a, b = sass(whole, fn), sass(unit, fn)
        assert len(a) > 300 and a == b, fn
