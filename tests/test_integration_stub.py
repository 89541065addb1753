"""The ctypes binding INTEGRATION.md shows (the stub a reference maintainer would add as
xgrid/util/ffi_b200.py) is executed as written: on CPU it must load the library and resolve every entry
point it names; on a B200 it must compile, load, launch and copy through the C ABI with host buffers."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from xgrid_b200.runtime import shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_stub() -> dict:
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    code = next(b for b in blocks if "ffi_b200.py" in b)
    code = code.replace('C.CDLL("libxgrid_b200.so")', f'C.CDLL({shim.LIB_PATH!r})')
    ns: dict = {}
    exec(compile(code, "INTEGRATION.md:ffi_b200", "exec"), ns)
    return ns


def test_stub_executes_and_names_only_exported_symbols():
    ns = load_stub()
    assert ns["_l"].xgb_abi_version() == 1
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for sym in set(re.findall(r"_l\.(xgb_[a-z0-9_]+)", text)):
        assert hasattr(ns["_l"], sym), f"INTEGRATION.md uses {sym}, which the library does not export"
    for fn in ("init", "alloc", "h2d", "d2h", "sync", "compile", "function", "launch"):
        assert callable(ns[fn])


@pytest.mark.gpu
def test_stub_runs_a_kernel_from_host_buffers():
    ns = load_stub()
    ns["init"](0)
    src = r'''
struct P { double *x; double a; long long n; };
extern "C" __global__ void scale(const __grid_constant__ P p)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < p.n) p.x[i] = p.x[i] * p.a + 1.0;
}
'''
    mod = ns["compile"](src, "scale.cu", {})
    fn = ns["function"](mod, "scale")

    class P(C.Structure):
        _fields_ = [("x", C.c_void_p), ("a", C.c_double), ("n", C.c_longlong)]

    n = 100003
    host = np.random.default_rng(0).random(n)
    want = host * 2.5 + 1.0
    dev = ns["alloc"](host.nbytes)
    ns["h2d"](dev, host)
    ns["launch"](fn, ((n + 255) // 256, 1, 1), (256, 1, 1), P(dev, 2.5, n))
    out = np.empty_like(host)
    ns["d2h"](out, dev)
    ns["sync"]()
    assert np.array_equal(out, want)
    with pytest.raises(Exception, match="xgrid_b200"):
        ns["compile"]("innt main() {}", "bad.cu", {})      # Logger.dead semantics: Exception with the NVRTC log
