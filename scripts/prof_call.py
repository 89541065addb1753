import sys, time, cProfile, pstats; sys.path.insert(0, "/root/repo")
import numpy as np
import xgrid_b200 as xgrid
from examples import workloads as W
xgrid.init(precision="double", cacheroot="/tmp/xgp")
k = W.make_kernels()
a, b, r = (xgrid.Grid((10000,), float) for _ in range(3))
a.now[:] = 1.5; b.now[:] = 2.0
f = k["elementwise_mul"]
for _ in range(20): f(r, a, b)
xgrid.synchronize()
t0 = time.perf_counter()
for _ in range(5000): f(r, a, b)
t1 = time.perf_counter(); xgrid.synchronize(); t2 = time.perf_counter()
print("host issue %.2f us/call, incl. drain %.2f us/call" % ((t1-t0)/5000*1e6, (t2-t0)/5000*1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(5000): f(r, a, b)
pr.disable(); xgrid.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
