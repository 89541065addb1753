"""Control for `compute-sanitizer --tool racecheck` (scripts/sanitize_gpu.sh): the textbook SINGLE-STAGE bulk-copy
pattern -- thread 0 arms an mbarrier with expect_tx and issues `cp.async.bulk`, every thread waits on the barrier's
phase, reads the tile, and a `__syncthreads()` protects the tile before the next copy is issued.  It is correct by
construction (CUDA programming guide, "asynchronous data copies using the Tensor Memory Accelerator").  If racecheck
reports a hazard between the bulk copy's write and the reads HERE, the tool does not model completion through
`mbarrier::complete_tx`, and the same report on the pipeline kernels (which add only mbarrier-based release of a
stage) is a limitation of the tool, not a race."""
import ctypes as C

import numpy as np
import pytest

import xgrid_b200 as xgrid
from xgrid_b200.lang.schedule import template_headers
from xgrid_b200.runtime import shim

pytestmark = pytest.mark.gpu

SRC = r'''
#include "xgb_stencil.cuh"
struct ctl_params { const double *in; double *out; int tiles; };
extern "C" __global__ void __launch_bounds__(256) racecheck_control(const __grid_constant__ ctl_params p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    double *tile = reinterpret_cast<double *>(smem + 128);
    if (threadIdx.x == 0) { xgb::pipe::mbar_init(bar, 1); xgb::pipe::fence_barrier_init(); }
    __syncthreads();
    uint32_t phase = 0;
    for (int t = 0; t < p.tiles; ++t) {
        const int64_t base = ((int64_t)blockIdx.x * p.tiles + t) * 256;
        if (threadIdx.x == 0) {
            xgb::pipe::mbar_expect_tx(bar, 256 * sizeof(double));
            xgb::pipe::bulk_g2s(tile, p.in + base, 256 * sizeof(double), bar);
        }
        xgb::pipe::mbar_wait(bar, phase);
        phase ^= 1;
        const double left = tile[(threadIdx.x + 255) & 255], mid = tile[threadIdx.x];
        p.out[base + threadIdx.x] = mid + left;
        __syncthreads();                       // every thread has read the tile before it is overwritten
    }
}
'''


class Params(C.Structure):
    _fields_ = [("inp", C.c_void_p), ("out", C.c_void_p), ("tiles", C.c_int)]


def test_single_stage_bulk_copy_pattern(tmp_path):
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"))
    rt = shim.Runtime.get()
    image, _ = shim.compile_cuda(SRC, "racecheck_control.cu", ["--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"],
                                 template_headers())
    fn = rt.get_function(rt.module_load(image), "racecheck_control")
    blocks, tiles = 8, 6
    n = blocks * tiles * 256
    x = np.random.default_rng(0).random(n)
    din, dout = rt.alloc(n * 8), rt.alloc(n * 8)
    rt.h2d(din, x.ctypes.data, n * 8)
    rt.launch(fn, (blocks, 1, 1), (256, 1, 1), Params(din, dout, tiles), smem=128 + 256 * 8)
    y = np.empty(n)
    rt.d2h(y.ctypes.data, dout, n * 8)
    rt.sync()
    t = x.reshape(-1, 256)
    assert np.array_equal(y.reshape(-1, 256), t + np.roll(t, 1, axis=1))
    rt.free(din)
    rt.free(dout)


RING_SRC = r'''
#include "xgb_stencil.cuh"
struct ring_params { const double *in; double *out; int tiles; };
// Two-stage ring: warp 8 produces (bulk copy per tile, waits on the stage's "empty" barrier before re-using it),
// warps 0..7 consume.  ALL = every consumer thread arrives on "empty" (count 256); otherwise ONE elected lane per
// warp arrives after __syncwarp() (count 8) -- the release pattern of the pipeline kernels.
template <bool ALL>
__device__ __forceinline__ void ring_body(const ring_params &p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem), *empty = full + 2;
    double *tiles = reinterpret_cast<double *>(smem + 128);          // [2][256]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) { xgb::pipe::mbar_init(&full[s], 1); xgb::pipe::mbar_init(&empty[s], ALL ? 256 : 8); }
        xgb::pipe::fence_barrier_init();
    }
    __syncthreads();
    if (warp == 8) {
        int s = 0, eph = 1;
        for (int t = 0; t < p.tiles; ++t, ++s) {
            if (s == 2) { s = 0; eph ^= 1; }
            if (t >= 2) xgb::pipe::mbar_wait(&empty[s], eph);
            if (lane == 0) {
                xgb::pipe::mbar_expect_tx(&full[s], 256 * sizeof(double));
                xgb::pipe::bulk_g2s(tiles + s * 256, p.in + ((int64_t)blockIdx.x * p.tiles + t) * 256, 256 * sizeof(double), &full[s]);
            }
            __syncwarp();
        }
        return;
    }
    int s = 0, ph = 0;
    for (int t = 0; t < p.tiles; ++t) {
        xgb::pipe::mbar_wait(&full[s], ph);
        const double *tile = tiles + s * 256;
        const double v = tile[threadIdx.x] + tile[(threadIdx.x + 255) & 255];
        p.out[((int64_t)blockIdx.x * p.tiles + t) * 256 + threadIdx.x] = v;
        if (ALL) {
            xgb::pipe::mbar_arrive(&empty[s]);
        } else {
            __syncwarp();
            if (lane == 0) xgb::pipe::mbar_arrive(&empty[s]);
        }
        if (++s == 2) { s = 0; ph ^= 1; }
    }
}
extern "C" __global__ void __launch_bounds__(288) ring_elected_lane(const __grid_constant__ ring_params p) { ring_body<false>(p); }
extern "C" __global__ void __launch_bounds__(288) ring_every_thread(const __grid_constant__ ring_params p) { ring_body<true>(p); }
'''


@pytest.mark.parametrize("kernel", ["ring_elected_lane", "ring_every_thread"])
def test_two_stage_ring(tmp_path, kernel):
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"))
    rt = shim.Runtime.get()
    image, _ = shim.compile_cuda(RING_SRC, "racecheck_ring.cu", ["--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"],
                                 template_headers())
    fn = rt.get_function(rt.module_load(image), kernel)
    blocks, tiles = 6, 9
    n = blocks * tiles * 256
    x = np.random.default_rng(1).random(n)
    din, dout = rt.alloc(n * 8), rt.alloc(n * 8)
    rt.h2d(din, x.ctypes.data, n * 8)
    rt.launch(fn, (blocks, 1, 1), (288, 1, 1), Params(din, dout, tiles), smem=128 + 2 * 256 * 8)
    y = np.empty(n)
    rt.d2h(y.ctypes.data, dout, n * 8)
    rt.sync()
    t = x.reshape(-1, 256)
    assert np.array_equal(y.reshape(-1, 256), t + np.roll(t, 1, axis=1))
    rt.free(din)
    rt.free(dout)
