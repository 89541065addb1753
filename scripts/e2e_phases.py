"""Where does the end-to-end time of the default bench job go?  Host-side timers + cProfile around the
public-API job of bench.py's `e2e` leg (host IC + mask -> K calls -> .now).  Diagnostic only."""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                    # noqa: E402
import xgrid_b200 as xgrid                      # noqa: E402
from examples import workloads as W           # noqa: E402
from xgrid_b200.runtime.shim import Runtime     # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "conv1d"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
spec = bench.WORKLOADS[name]
shape = spec["shape"]
xgrid.init(precision="double", cacheroot=os.path.join(ROOT, ".xgrid"), device=0)
kern = W.make_kernels()[spec["kernel"]]
inputs, scalars = bench.build_inputs(name, shape)
rt = Runtime.get()


def job(tag):
    t = [time.perf_counter()]
    gs = [xgrid.Grid(shape, float) for _ in inputs]
    t.append(time.perf_counter())
    for g, (ic, mask) in zip(gs, inputs):
        bench._put(g.now, ic)
    t.append(time.perf_counter())
    for g, (ic, mask) in zip(gs, inputs):
        bench._put(g.boundary, mask)
    t.append(time.perf_counter())
    kern(*gs, *scalars)
    xgrid.flush()
    t.append(time.perf_counter())
    rt.sync()
    t.append(time.perf_counter())
    for _ in range(K - 1):
        kern(*gs, *scalars)
    t.append(time.perf_counter())
    xgrid.flush()
    t.append(time.perf_counter())
    rt.sync()
    t.append(time.perf_counter())
    outs = [g.now for g in gs]
    t.append(time.perf_counter())
    names = ["Grid()", "now[...]=ic", "boundary[...]=mask", "first call+flush (enqueue)", "sync after first call",
             f"{K - 1} calls (enqueue)", "flush (enqueue)", "sync", ".now (D2H)"]
    ms = [(b - a) * 1e3 for a, b in zip(t, t[1:])]
    print(f"[{tag}] total {sum(ms):.1f} ms: " + "; ".join(f"{n} {m:.1f}" for n, m in zip(names, ms)), flush=True)
    return outs


job("cold (JIT, first allocations)")
job("warm 1")
job("warm 2")
pr = cProfile.Profile()
pr.enable()
job("profiled")
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
