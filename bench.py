#!/usr/bin/env python
"""bench.py -- headline benchmark of the xgrid stencil hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One JSON line on stdout (rank 0).  A *step* is one kernel call = one timestep
of the workload (one pass of the hot path over the whole grid).

Workloads (BASELINE.json configs):
  heat3d   3-D 7-point, 256x2048x2048 fp64 per GPU, slab-sharded on axis 0 (config[4]) -- the MAIN line at
           every N (`--workload auto`): the largest single-GPU configuration, and the one north_star quotes
           the weak-scaling target on, so the driver's per-N values are comparable
  conv1d / conv1d_nl / diff1d   1-D kernels, 2^24 points fp64 (config[1])
  conv2d   2-D upwind convection 16384^2 fp64 (config[2]);  diff2d = 5-point variant
  cavity   lid-driven cavity 8192^2 fp64 (config[3]); value in Gpoint-updates/s, also timesteps/s
  ewmul    README elementwise_mul, 10 000 points (config[0]; launch-latency bound)
At N=1 the other configurations ride along as sub-records under "extra" (device-timed, a few seconds each):
one-pass and temporally blocked variants, cavity on the quiescent and on a developed-flow field, and the
FMA ("performance") build checked against the validation build at north_star's tolerance.

Metric: Gpoint-updates/s = grid points x interior (mask-0) statements per call x steps / time.
Roofline: algorithmic bytes per step / device time per step vs MEASURED_PEAKS.json.
"parity": after the timed region a few small cases run through the same (sharded, at N>1) path and are
compared bit for bit with the oracle -- the oracle is the checker only, nothing timed touches it.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# workload table: interior statements per call, algorithmic bytes/point/step (fp64), default size
WORKLOADS = {
    "conv1d":    dict(kernel="convection_1d", ndim=1, shape=(1 << 24,), stmts=1, bytes_pt=16, steps=10000, warmup=200),
    "conv1d_nl": dict(kernel="convection_1d_nonlinear", ndim=1, shape=(1 << 24,), stmts=1, bytes_pt=16, steps=10000, warmup=200),
    "diff1d":    dict(kernel="diffusion_1d", ndim=1, shape=(1 << 24,), stmts=1, bytes_pt=16, steps=10000, warmup=200),
    "conv2d":    dict(kernel="convection_2d", ndim=2, shape=(16384, 16384), stmts=1, bytes_pt=16, steps=200, warmup=10),
    "diff2d":    dict(kernel="diffusion_2d", ndim=2, shape=(16384, 16384), stmts=1, bytes_pt=16, steps=200, warmup=10),
    "cavity":    dict(kernel="cavity_kernel", ndim=2, shape=(8192, 8192), stmts=54, bytes_pt=1312, steps=10, warmup=5),
    "heat3d":    dict(kernel="heat_3d", ndim=3, shape=(256, 2048, 2048), stmts=1, bytes_pt=16, steps=20, warmup=5),
    "ewmul":     dict(kernel="elementwise_mul", ndim=1, shape=(10000,), stmts=1, bytes_pt=24, steps=2000, warmup=50),
}
MAIN = "heat3d"
SHARDABLE = ("heat3d", "conv1d", "conv1d_nl", "diff1d", "cavity", "conv2d", "diff2d")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def config_of(name: str, shape, world: int) -> dict:
    """The `config` object of the JSON line -- identical for both arms (`--impl reference` prints the same)."""
    spec = WORKLOADS[name]
    points = float(np.prod(shape))
    return {"workload": f"{name} {'x'.join(map(str, shape))} fp64" + (" per GPU, slab-sharded on axis 0" if world > 1 else ""),
            "kernel": spec["kernel"], "interior_statements_per_step": spec["stmts"],
            "l2": "working set (2 levels) larger than the 126 MB L2" if points * 16 > 126e6 else "L2-resident (small grid)",
            "validate_build": True}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md).  The sampler is
    started BEFORE the warm-up (nvidia-smi needs a few hundred ms to come up) and every row is time-stamped;
    `summary()` uses the rows that fall inside [mark_start, mark_end], or the nearest ones if the timed
    region was shorter than one sampling period."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int, enabled: bool = True) -> None:
        self.device, self.rows, self.proc = device, [], None
        self.t0 = self.t1 = None
        self.enabled = enabled          # rank 0 samples its GPU; eight pollers on one box only perturb it

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.06)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self) -> dict:
        good = [(t, r) for t, r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        t0 = self.t0 if self.t0 is not None else 0.0
        t1 = self.t1 if self.t1 is not None else float("inf")
        rows = [r for t, r in good if t0 <= t <= t1 + 0.03]
        window = "timed region"
        if not rows and good:
            mid = 0.5 * (t0 + min(t1, t0 + 1e9))
            rows = [r for _, r in sorted(good, key=lambda tr: abs(tr[0] - mid))[:2]]
            window = "nearest samples (timed region shorter than the sampling period)"
        sm = [float(r[0]) for r in rows]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[3:7]) if v == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "window": window, "reasons": reasons}


# --------------------------------------------------------------------------- workload construction
def build_inputs(name: str, shape, seed: int = 0, ic_mode: str = "default"):
    """Host-side synthetic inputs of SURVEY.md §8d: (list of (ic, mask) per grid arg, scalar args)."""
    from examples import workloads as W
    if name in ("conv1d", "conv1d_nl", "diff1d"):
        n = shape[0]
        ic, dx = W.ic_1d(n)
        mask = np.zeros(n, np.int32)
        mask[0] = 1
        if name == "conv1d":
            return [(ic, mask)], (1.0, 0.5 * dx, dx)
        if name == "conv1d_nl":
            return [(ic, mask)], (0.25 * dx, dx)
        mask[-1] = 1
        nu = 0.01
        return [(ic, mask)], (nu, 0.2 * dx * dx / nu, dx)
    if name in ("conv2d", "diff2d"):
        n0, n1 = shape
        dx = 2.0 / (n1 - 1)
        ic = np.ones(shape)
        ic[n0 // 4:n0 // 2, n1 // 4:n1 // 2] = 2.0
        if name == "conv2d":
            mask = np.zeros(shape, np.int32)
            mask[0, :] = 1
            mask[:, 0] = 1
            return [(ic, mask)], (1.0, 0.5 * dx, dx, dx)
        return [(ic, W.shell_mask(shape))], (0.2,)
    if name == "heat3d":
        return [(_fill_random(seed), _fill_shell(shape, 0, shape[0]))], (0.1,)
    if name == "cavity":
        n0, n1 = shape
        mb, mp, mu, mv = W.cavity_masks(n0, n1)
        dx, dy = 2.0 / (n1 - 1), 2.0 / (n0 - 1)
        dt = 1e-4 * (100.0 / (n1 - 1)) ** 2
        z = np.zeros(shape)
        if ic_mode == "random" or os.environ.get("XGB_BENCH_IC") == "random":
            # developed-flow stand-in: no field is exactly zero, so every fp64 divide takes its full path
            # instead of the zero-numerator shortcut the quiescent zero IC of SURVEY.md §8d allows
            rng = np.random.default_rng(seed)
            return [(1e-3 * rng.random(shape), m) for m in (mb, mp, mu, mv)], (W.Config(1.0, 0.1, dt, dx, dy),)
        return [(z, mb), (z, mp), (z, mu), (z, mv)], (W.Config(1.0, 0.1, dt, dx, dy),)
    if name == "ewmul":
        rng = np.random.default_rng(seed)
        n = shape[0]
        z = np.zeros(n, np.int32)
        return [(np.zeros(n), z), (rng.random(n), z), (rng.random(n), z)], ()
    raise SystemExit(f"unknown workload {name}")


def _fill_random(seed: int):
    """In-place initial condition for the big 3-D slab (8 GiB per level): U[0,1) written straight into the
    grid's own host array -- no second copy of the level in host memory."""
    def fill(out: np.ndarray) -> None:
        np.random.default_rng(seed).random(out=out.reshape(-1))
    return fill


def _fill_shell(global_shape, lo: int, hi: int):
    """In-place rows [lo, hi) of examples.workloads.shell_mask(global_shape)."""
    def fill(out: np.ndarray) -> None:
        out[...] = 1
        inner = tuple(slice(1, -1) for _ in global_shape[1:])
        a, b = max(lo, 1) - lo, min(hi, global_shape[0] - 1) - lo
        if b > a:
            out[(slice(a, b),) + inner] = 0
    return fill


def _put(dst, src) -> None:
    """Write an input (array, or in-place filler) into a grid's host array."""
    if callable(src):
        src(np.asarray(dst))
    else:
        dst[...] = src


def build_slab_inputs(name: str, shape, rank: int, world: int):
    """Weak scaling: every rank owns a `shape` slab of the (world*shape[0], ...) global grid."""
    from examples import workloads as W
    gshape = (shape[0] * world,) + tuple(shape[1:])
    lo, hi = rank * shape[0], (rank + 1) * shape[0]
    if name == "heat3d":
        return gshape, [(_fill_random(rank), _fill_shell(gshape, lo, hi))], (0.1,)
    if name in ("conv1d", "conv1d_nl", "diff1d"):
        n = gshape[0]
        dx = 2.0 / (n - 1)
        ic = np.ones(shape[0])
        a, b = int(.5 / dx), int(1 / dx + 1)              # global IC of test.py:195-198, local slice
        ic[max(a, lo) - lo:max(min(b, hi), lo) - lo] = 2.0
        mask = np.zeros(shape[0], np.int32)
        if rank == 0:
            mask[0] = 1
        if name == "conv1d":
            scalars = (1.0, 0.5 * dx, dx)
        elif name == "conv1d_nl":
            scalars = (0.25 * dx, dx)
        else:
            if rank == world - 1:
                mask[-1] = 1
            scalars = (0.01, 0.2 * dx * dx / 0.01, dx)
        return gshape, [(ic, mask)], scalars
    if name == "cavity":
        n0, n1 = gshape
        dx, dy = 2.0 / (n1 - 1), 2.0 / (n0 - 1)
        dt = 1e-4 * (100.0 / (n1 - 1)) ** 2
        z = np.zeros(shape)
        return gshape, [(z, m) for m in W.cavity_masks_slab(n0, n1, lo, hi)], (W.Config(1.0, 0.1, dt, dx, dy),)
    if name in ("conv2d", "diff2d"):
        n0, n1 = gshape
        dx = 2.0 / (n1 - 1)
        ic = np.ones(shape)
        a, b = max(n0 // 4, lo), min(n0 // 2, hi)          # the global square hat, local rows
        if b > a:
            ic[a - lo:b - lo, n1 // 4:n1 // 2] = 2.0
        mask = np.zeros(shape, np.int32)
        mask[:, 0] = 1
        if rank == 0:
            mask[0, :] = 1
        if name == "diff2d":
            mask[:, -1] = 1
            if rank == world - 1:
                mask[-1, :] = 1
            return gshape, [(ic, mask)], (0.2,)
        return gshape, [(ic, mask)], (1.0, 0.5 * dx, dx, dx)
    raise SystemExit("multi-GPU bench is defined for the slab-sharded workloads: " + ", ".join(SHARDABLE))


def oracle_stepper(name: str, grids, scalars):
    import oracle
    if name == "conv1d":
        return lambda: oracle.step_conv1d(grids[0], *scalars)
    if name == "conv1d_nl":
        return lambda: oracle.step_conv1d_nonlinear(grids[0], *scalars)
    if name == "diff1d":
        return lambda: oracle.step_diff1d(grids[0], *scalars)
    if name == "conv2d":
        return lambda: oracle.step_conv2d(grids[0], *scalars)
    if name == "diff2d":
        return lambda: oracle.step_diff2d(grids[0], *scalars)
    if name == "heat3d":
        return lambda: oracle.step_heat3d(grids[0], *scalars)
    if name == "cavity":
        c = scalars[0]
        cfg = oracle.Config(c.rho, c.nu, c.dt, c.dx, c.dy)
        return lambda: oracle.step_cavity(*grids, cfg)
    if name == "ewmul":
        return lambda: oracle.step_ewmul(*grids)
    raise SystemExit(name)


def cpu_sample_shape(name: str, shape):
    """Bounded sample of the workload for the CPU arm (same kernel, same row length, fewer planes)."""
    if name == "heat3d":
        return (min(64, shape[0]),) + tuple(shape[1:])
    return tuple(shape)


def cpu_arm(name: str, shape, budget_s: float, max_steps: int, sample_shape=None):
    """Time the reference's CPU implementation of the path, all host threads, on a bounded
    sample of the workload: the reference's OWN compiled kernels (oracle/_ref, built by
    oracle/make_ref.py from the unmodified reference; kind "reference") where they exist and
    are valid (1-D, square 2-D), else the C port (oracle/xgrid_oracle.c, gcc -O3 -fopenmp;
    kind "port").  Returns (Gpt/s, seconds per step, steps, shape, kind, description)."""
    import oracle
    from oracle import ref
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm always uses every core it may run on
    oracle.set_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    shape = tuple(sample_shape or shape)
    inputs, scalars = build_inputs(name, shape)
    grids = []
    for ic, mask in inputs:
        g = oracle.HostGrid(shape)
        _put(g.now, ic)
        _put(g.boundary, mask)
        grids.append(g)
    rk = ref.KERNEL_OF.get(name)
    use_ref = (rk is not None and ref.available(rk) and os.environ.get("XGB_CPU_ARM", "reference") != "port"
               and (len(shape) == 1 or (len(shape) == 2 and shape[0] == shape[1])))
    if use_ref:
        step = lambda: ref.call(rk, *grids, *scalars)      # noqa: E731
        kind = "reference"
        desc = f"oracle/_ref ({ref.manifest()['_meta']['cc'].lstrip('/').strip()}: the reference's generated C)"
    else:
        step = oracle_stepper(name, grids, scalars)
        kind, desc = "port", "oracle/xgrid_oracle.c gcc -O3 -fopenmp"
        if name == "heat3d":
            desc += " (the reference cannot run 3-D grids: its linear index is wrong there, SURVEY.md F1)"
    step()                                   # warm-up: first touch of the second ring level
    t0 = time.perf_counter()
    step()
    one = time.perf_counter() - t0
    steps = int(max(2, min(max_steps, budget_s / max(one, 1e-9))))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    pts = float(np.prod(shape)) * WORKLOADS[name]["stmts"]
    return pts * steps / dt / 1e9, dt / steps, steps, shape, kind, desc


# --------------------------------------------------------------------------- our arm
class Arm:
    """Our arm's process state: runtime, rank, init() variants."""

    def __init__(self, rank: int, world: int) -> None:
        import xgrid_b200 as xgrid
        from xgrid_b200.runtime.shim import Runtime
        self.xgrid, self.rank, self.world = xgrid, rank, world
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.mode = None
        self.kernels = None
        self.configure()
        self.rt = Runtime.get()

    def configure(self, validate: bool = True, temporal: bool = True) -> dict:
        """(Re-)initialise the backend; kernels are re-made because annotations and build flags follow init()."""
        from examples import workloads as W
        if self.mode != (validate, temporal):
            self.xgrid.init(precision="double", cacheroot=os.path.join(ROOT, ".xgrid"), device=self.local_rank,
                            distributed=self.world > 1, validate=validate, temporal=temporal,
                            graphs=os.environ.get("XGB_BENCH_GRAPHS", "1") != "0")
            self.kernels = W.make_kernels()
            self.mode = (validate, temporal)
        return self.kernels

    def barrier(self) -> None:
        if self.world > 1:
            import torch
            import torch.distributed as dist
            self.rt.device_sync()
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        import torch
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{self.local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks(self, x: float) -> list:
        if self.world == 1:
            return [x]
        import torch
        import torch.distributed as dist
        t = torch.zeros(self.world, dtype=torch.float64, device=f"cuda:{self.local_rank}")
        t[self.rank] = x
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    def grids_for(self, gshape, shape, inputs):
        out = []
        for ic, mask in inputs:
            g = self.xgrid.Grid(gshape, float)
            assert g.shape == tuple(shape), (g.shape, shape)
            _put(g.now, ic)
            _put(g.boundary, mask)
            out.append(g)
        return out


def device_leg(arm: Arm, kern, grids, scalars, K: int, Wm: int):
    """The device-timed leg: warm-up, then exactly K calls between two events on the backend stream.
    The warm-up executes the SAME launch plan as the timed region -- whole repetitions of "K calls, flush",
    at least Wm steps in total -- so that every buffer (spare / scratch levels, ghost layout, mask halos),
    module, CUDA graph and NCCL channel the timed K calls use exists before the first event.
    Returns (ms max over ranks, launches, clocks summary, warm-up steps)."""
    rt = arm.rt
    reps = max(1, -(-Wm // K))
    with ClockSampler(arm.local_rank, enabled=arm.rank == 0) as clocks:
        for _ in range(reps):
            for _ in range(K):
                kern(*grids, *scalars)
            rt.sync()                            # flushes the deferred-call queue, like the end of the timed region
        ev0, ev1 = rt.event_create(), rt.event_create()
        arm.barrier()
        rt.device_sync()
        n0 = rt.launch_count()
        clocks.mark_start()
        rt.event_record(ev0)
        for _ in range(K):
            kern(*grids, *scalars)
        rt.event_record(ev1)
        rt.event_sync(ev1)
        clocks.mark_end()
    rt.device_sync()
    arm.barrier()
    launches = rt.launch_count() - n0
    per_rank = arm.all_ranks(rt.event_elapsed_ms(ev0, ev1))
    summary = clocks.summary()
    if arm.world > 1:
        summary["ms_per_step_by_rank"] = [round(v / K, 4) for v in per_rank]     # `ms_per_step` is their maximum
    return max(per_rank), int(launches), summary, reps * K


def traffic_of(name: str, shape, launches: int, K: int, temporal: bool = True):
    """ncu DRAM bytes per launch of the dominant kernel of THIS run's variant (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            table = json.load(f)
    except OSError:
        return None, None
    key = name
    blocked = bool(launches and launches < K)
    if name in ("conv1d", "conv1d_nl", "diff1d"):
        if not blocked:
            key = name + ":onepass"
        elif K < 64:
            key = name + ":tail"                 # a short run executes through the multistep_tail variant
    elif name in ("conv2d", "diff2d") and not blocked:
        key = name + ":onepass"
    t = table.get(key)
    if not t or tuple(shape) != tuple(WORKLOADS[name]["shape"]):
        return None, None
    return t["bytes_per_launch"], (f"{t['kernel']}: DRAM read+write per launch ({t['steps_per_launch']} step(s)), "
                                   f"{t['source']}")


def record_of(arm: Arm, name: str, shape, K: int, ms: float, launches: int, clocks: dict,
              warm_steps: int, temporal: bool = True) -> dict:
    spec = WORKLOADS[name]
    points = float(np.prod(shape))
    ms_per_step = ms / K
    value = points * arm.world * spec["stmts"] * K / (ms * 1e-3) / 1e9
    peak, peak_src = measured_peaks()
    alg_bytes = points * spec["bytes_pt"]
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
    traffic, traffic_src = traffic_of(name, shape, launches, K, temporal)
    rec = {"value": value, "unit": "Gpoint-updates/s", "steps": K, "warmup": warm_steps, "ms_per_step": ms_per_step,
           "config": config_of(name, shape, arm.world),
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                        "peak_source": peak_src, "algorithmic_bytes_per_step": alg_bytes,
                        "frac_of_nominal_8TBs": achieved / 8000.0},
           "gpu_launches": launches, "clocks": clocks}
    if launches and launches < K:
        spl = K / launches
        rec["config"]["temporal_blocking"] = (f"runs of identical calls execute ~{spl:.0f} time steps per launch "
                                              "(bit-identical to step-at-a-time)")
        rec["roofline"]["note"] = ("achieved = one-pass algorithmic bytes / time; frac > 1 means the one-pass-per-step "
                                   "HBM roofline is exceeded by temporal blocking")
    if name == "cavity":
        rec["timesteps_per_s"] = 1e3 / ms_per_step
    return rec


def e2e_legs(arm: Arm, name: str, kern, gshape, shape, inputs, scalars, K: int) -> dict:
    """End to end through the public API: the job's inputs start in HOST buffers (the NumPy arrays behind
    Grid.now / Grid.boundary, filled before the clock starts); the timed region is the K kernel calls -- the
    first one uploads state and mask (H2D) -- and the read of every grid's newest level on the host (D2H).
    Device buffers come from the runtime's caching pool, warm like in any long-running program."""
    spec = WORKLOADS[name]
    points = float(np.prod(shape))
    world = arm.world
    out = {}
    arm.barrier()
    t0 = time.perf_counter()
    g2 = arm.grids_for(gshape, shape, inputs)      # host writes of IC + mask (pageable NumPy), not timed
    fill_ms = (time.perf_counter() - t0) * 1e3
    arm.barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        kern(*g2, *scalars)                        # first call uploads IC + mask (H2D)
    outs = [g.now for g in g2]                     # D2H of every grid's newest level (the rank's slab)
    dt = arm.max_over_ranks(time.perf_counter() - t0)
    h2d = sum(g.size * g.itemsize + g.size * 4 for g in g2)         # every grid's level + its int32 mask
    d2h = sum(o.nbytes for o in outs)
    out["e2e"] = {"value": points * world * spec["stmts"] * K / dt / 1e9, "unit": "Gpoint-updates/s",
                  "h2d_bytes_per_step": h2d * world / K, "d2h_bytes_per_step": d2h * world / K,
                  "ms_total": dt * 1e3, "host_fill_ms_not_timed": fill_ms,
                  "note": f"public API job: state + int32 mask in host buffers -> {K} kernel calls (the first uploads) -> "
                          ".now on host; wall clock, max over ranks; whole-state copies happen once per job, so the "
                          "per-step byte counts are the totals / steps (the mask crosses PCIe packed to one byte per "
                          "point; the count above is the int32 array the API holds)"}
    if world == 1:
        big = points * 8 > (1 << 30)
        # literal per-call offload: upload the state, one call, download the state, every step
        Ko = 3 if big else max(3, min(20, K))
        t0 = time.perf_counter()
        for _ in range(Ko):
            kern(*g2, *scalars)                    # the host owns every newest level (read below): uploads them
            outs = [g.now for g in g2]             # D2H; hands the levels back to the host
        dt = time.perf_counter() - t0
        out["e2e_offload"] = {"value": points * spec["stmts"] * Ko / dt / 1e9, "unit": "Gpoint-updates/s",
                              "h2d_bytes_per_step": float(sum(o.nbytes for o in outs)),
                              "d2h_bytes_per_step": float(sum(o.nbytes for o in outs)), "steps": Ko,
                              "note": "every step: H2D full state, one kernel call, D2H full state"}
        # strict per-step read-back: every step's result is observed on the host through element
        # indexing (forces one launch + one device->host read per step; no deferral, no batching)
        Kp = max(3, min(200 if big else 1000, K))
        probe = tuple(n // 2 for n in shape)
        t0 = time.perf_counter()
        acc = 0.0
        for _ in range(Kp):
            kern(*g2, *scalars)
            acc += float(g2[0][probe])
        dt = time.perf_counter() - t0
        out["e2e_probe"] = {"value": points * spec["stmts"] * Kp / dt / 1e9, "unit": "Gpoint-updates/s",
                            "h2d_bytes_per_step": 8.0 * len(scalars), "d2h_bytes_per_step": 8.0, "steps": Kp,
                            "note": "every step: one kernel call, then one element of the result read on the host"}
    del g2, outs
    return out


# --------------------------------------------------------------------------- parity field
def parity_cases(arm: Arm) -> dict:
    """Small cases through the SAME execution path as the timed region (slab-sharded over all ranks at N>1:
    halo exchange, edge-first overlap, multi-step launches), compared bit for bit with the oracle run on the
    whole domain.  Every rank checks its own slab; the verdict is the AND over ranks."""
    import oracle
    from examples import workloads as W
    xgrid, world = arm.xgrid, arm.world
    k = arm.configure()
    cases = []

    def slab(g):
        return slice(*g.row_range)

    def check(name, pairs):
        ok = all(np.array_equal(np.asarray(a), np.asarray(b)) for a, b in pairs)
        cases.append({"case": name, "ok": bool(ok)})

    def heat(gshape, steps, label):
        rng = np.random.default_rng(7)
        ic, mask = rng.random(gshape), W.shell_mask(gshape)
        u, h = xgrid.Grid(gshape, float), oracle.HostGrid(gshape)
        s = slab(u)
        u.now[...] = ic[s]
        u.boundary[...] = mask[s]
        h.now[...] = ic
        h.boundary[...] = mask
        for _ in range(steps):
            k["heat_3d"](u, 0.1)
            oracle.step_heat3d(h, 0.1)
        check(f"heat3d {'x'.join(map(str, gshape))} x{steps}{label}", [(u.now, h.now[s]), (u._data[1], h._data[1][s])])

    heat((8 * max(world, 2), 24, 256), 5, "")
    heat((16 * world, 16, 2048), 3, " (bulk-copy tiled variant = the timed kernel)")
    # 1-D diffusion, 150 deferred calls: multi-step launches + tail + single steps, masks next to slab cuts
    n1 = world * 40000 + 17
    ic1, dx1 = W.ic_1d(n1)
    ic1 = ic1 + 0.01 * np.random.default_rng(3).random(n1)
    m1 = np.zeros(n1, np.int32)
    m1[0] = m1[-1] = 1
    m1[n1 // 2] = 1
    m1[n1 // 2 + 3] = 7
    u1, h1 = xgrid.Grid((n1,), float), oracle.HostGrid((n1,))
    s = slab(u1)
    u1.now[...] = ic1[s]
    u1.boundary[...] = m1[s]
    h1.now[...] = ic1
    h1.boundary[...] = m1
    args1 = (0.01, 0.2 * dx1 * dx1 / 0.01, dx1)
    for _ in range(150):
        k["diffusion_1d"](u1, *args1)
        oracle.step_diff1d(h1, *args1)
    check(f"diff1d {n1} x150 (multi-step launches)", [(u1.now, h1.now[s]), (u1._data[1], h1._data[1][s])])
    # cavity: implicit Jacobi sweeps + Neumann statements that read level 0 across the slab cut
    nc = 96
    mb, mp, mu, mv = W.cavity_masks(nc, nc)
    dxc = 2.0 / (nc - 1)
    cfg = W.Config(1.0, 0.1, 1e-4, dxc, dxc)
    gs = [xgrid.Grid((nc, nc), float) for _ in range(4)]
    hs = [oracle.HostGrid((nc, nc)) for _ in range(4)]
    s = slab(gs[0])
    for gg, hh, m in zip(gs, hs, (mb, mp, mu, mv)):
        gg.boundary[...] = m[s]
        hh.boundary[...] = m
    for _ in range(2):
        k["cavity_kernel"](*gs, cfg)
        oracle.step_cavity(*hs, oracle.Config(cfg.rho, cfg.nu, cfg.dt, cfg.dx, cfg.dy))
    check(f"cavity {nc}x{nc} x2 (622 sweeps)",
          [(gg.now, hh.now[s]) for gg, hh in zip(gs, hs)] + [(gg._data[1], hh._data[1][s]) for gg, hh in zip(gs, hs)])
    # cavity wide enough for the fused Jacobi pairs: on slabs the fused pass covers the interior rows, the three rows
    # next to each cut run step-at-a-time on row bands (lang/launch.py::run_pair)
    from xgrid_b200.lang.launch import STATS
    n0c, n1c = 64 * max(world, 2), 640
    mb, mp, mu, mv = W.cavity_masks(n0c, n1c)
    cfg = W.Config(1.0, 0.1, 1e-4 * (100.0 / (n1c - 1)) ** 2, 2.0 / (n1c - 1), 2.0 / (n0c - 1))
    gs = [xgrid.Grid((n0c, n1c), float) for _ in range(4)]
    hs = [oracle.HostGrid((n0c, n1c)) for _ in range(4)]
    s = slab(gs[0])
    rngc = np.random.default_rng(5)
    for gg, hh, m in zip(gs, hs, (mb, mp, mu, mv)):
        icc = 1e-3 * rngc.random((n0c, n1c))
        gg.now[...] = icc[s]
        hh.now[...] = icc
        gg.boundary[...] = m[s]
        hh.boundary[...] = m
    fused_before = STATS.get("jacobi2", 0)
    for _ in range(3):
        k["cavity_kernel"](*gs, cfg)
        oracle.step_cavity(*hs, oracle.Config(cfg.rho, cfg.nu, cfg.dt, cfg.dx, cfg.dy))
    fused = STATS.get("jacobi2", 0) - fused_before
    check(f"cavity {n0c}x{n1c} x3 ({fused} fused Jacobi pairs)",
          [(gg.now, hh.now[s]) for gg, hh in zip(gs, hs)] + [(gg._data[1], hh._data[1][s]) for gg, hh in zip(gs, hs)]
          + [(fused, 72)])
    del gs, hs
    # 2-D diffusion, 9 deferred calls: four two-step passes + one single step; on slabs the rows next to a cut run
    # step-at-a-time on row bands beside the two-step pass of the interior (lang/schedule.py::_run_batch2)
    n0t, n1t = 80 * world + 1, 512
    ict, mt = np.random.default_rng(21).random((n0t, n1t)), W.shell_mask((n0t, n1t))
    ut, ht = xgrid.Grid((n0t, n1t), float), oracle.HostGrid((n0t, n1t))
    s = slab(ut)
    ut.now[...] = ict[s]
    ut.boundary[...] = mt[s]
    ht.now[...] = ict
    ht.boundary[...] = mt
    two_before = STATS.get("tiled2", 0)
    for _ in range(9):
        k["diffusion_2d"](ut, 0.2)
        oracle.step_diff2d(ht, 0.2)
    pairs_t = [(ut.now, ht.now[s]), (ut._data[1], ht._data[1][s])]
    two = STATS.get("tiled2", 0) - two_before
    check(f"diff2d {n0t}x{n1t} x9 ({two} two-step passes)", pairs_t + [(two, 4)])
    del ut, ht
    # overstep modes on slabs: "wrap" turns the ranks into a ring, "limit" clamps at the global ends only;
    # goldens produced by the unmodified reference (tests/golden/make_golden.py; square 32x32)
    from examples import workloads as W2
    for omode in ("wrap", "limit"):
        path = os.path.join(ROOT, "tests", "golden", f"diff2d_{omode}_f64.npz")
        if not os.path.exists(path):
            continue
        gd = np.load(path)
        xgrid.init(precision="double", cacheroot=os.path.join(ROOT, ".xgrid"), device=arm.local_rank,
                   distributed=world > 1, overstep=omode)
        arm.mode = None                            # the next configure() re-initialises
        ko = W2.make_kernels()["diffusion_2d_open"]
        uo = xgrid.Grid(gd["u_in"].shape, float)
        s = slab(uo)
        uo.now[...] = gd["u_in"][s]
        for _ in range(int(gd["steps"])):
            ko(uo, float(gd["params"][0]))
        check(f"diff2d overstep={omode} {'x'.join(map(str, gd['u_in'].shape))} x{int(gd['steps'])} (reference golden)",
              [(uo.now, gd["u.L0"][s]), (uo._data[1], gd["u.L1"][s])])
        del uo
    arm.configure()
    if world > 1:
        import torch
        import torch.distributed as dist
        t = torch.tensor([1 if c["ok"] else 0 for c in cases], device=f"cuda:{arm.local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        for c, v in zip(cases, t.tolist()):
            c["ok"] = bool(v)
    return {"ok": all(c["ok"] for c in cases), "ranks": world, "bit_exact": True, "cases": cases,
            "checker": "oracle/ (C restatement of the reference's generated loop nests, whole domain on every rank) and "
                       "reference-made goldens vs each rank's slab, all ring levels"}


# --------------------------------------------------------------------------- sub-records (N = 1)
def run_extra(arm: Arm, which: str, K: int, Wm: int, temporal: bool = True, validate: bool = True,
              ic_mode: str = "default", shape=None) -> dict:
    spec = WORKLOADS[which]
    shape = tuple(shape or spec["shape"])
    kern = arm.configure(validate=validate, temporal=temporal)[spec["kernel"]]
    inputs, scalars = build_inputs(which, shape, ic_mode=ic_mode)
    grids = arm.grids_for(shape, shape, inputs)
    del inputs
    ms, launches, clocks, warm = device_leg(arm, kern, grids, scalars, K, Wm)
    rec = record_of(arm, which, shape, K, ms, launches, clocks, warm, temporal)
    rec["config"]["validate_build"] = validate
    rec["config"]["temporal"] = temporal
    if ic_mode != "default":
        rec["config"]["ic"] = ic_mode
    del grids
    arm.rt.trim_pool()
    return rec


def fma_check(arm: Arm, which: str, shape, steps: int, tol: float) -> dict:
    """The performance build (init(validate=False): FMA contraction on) against the validation build on the
    same inputs: max relative error after `steps` steps, north_star's multi-step tolerance (1e-12 fp64)."""
    spec = WORKLOADS[which]
    results = []
    for validate in (True, False):
        kern = arm.configure(validate=validate)[spec["kernel"]]
        inputs, scalars = build_inputs(which, shape, ic_mode="random")
        grids = arm.grids_for(shape, shape, inputs)
        for _ in range(steps):
            kern(*grids, *scalars)
        results.append([np.array(g.now) for g in grids])
        del grids
    err = 0.0
    for a, b in zip(*results):
        scale = float(np.max(np.abs(a))) or 1.0
        err = max(err, float(np.max(np.abs(a - b))) / scale)
    arm.configure()
    return {"workload": f"{which} {'x'.join(map(str, shape))} x{steps}", "max_rel_err": err, "tolerance": tol,
            "ok": bool(err <= tol)}


EXTRA_PLAN = [
    ("conv1d_steps20", dict(which="conv1d", K=20, Wm=20)),
    ("conv1d_steps10000", dict(which="conv1d", K=10000, Wm=200)),
    ("conv1d_nl_steps10000", dict(which="conv1d_nl", K=10000, Wm=200)),
    ("diff1d_steps10000", dict(which="diff1d", K=10000, Wm=200)),
    ("conv1d_onepass", dict(which="conv1d", K=200, Wm=20, temporal=False)),
    ("conv2d_onepass", dict(which="conv2d", K=50, Wm=6, temporal=False)),
    ("conv2d_two_steps_per_pass", dict(which="conv2d", K=50, Wm=6)),
    ("diff2d_onepass", dict(which="diff2d", K=50, Wm=6, temporal=False)),
    ("diff2d_two_steps_per_pass", dict(which="diff2d", K=50, Wm=6)),
    ("cavity_zero_ic", dict(which="cavity", K=6, Wm=4)),
    ("cavity_developed_flow", dict(which="cavity", K=6, Wm=4, ic_mode="random")),
    ("ewmul_10000pts", dict(which="ewmul", K=2000, Wm=50)),
    ("heat3d_fma_build", dict(which="heat3d", K=20, Wm=5, validate=False)),
    ("cavity_developed_flow_fma_build", dict(which="cavity", K=6, Wm=4, validate=False, ic_mode="random")),
    ("conv1d_nl_steps10000_fma_build", dict(which="conv1d_nl", K=10000, Wm=200, validate=False)),
]


def run_extras(arm: Arm, select: str) -> dict:
    """The other BASELINE configurations at N=1, device-timed (seconds each)."""
    extra = {}
    only = set(select.split(",")) if select not in ("all", "") else None
    for key, kw in EXTRA_PLAN:
        if only is not None and key not in only:
            continue
        try:
            extra[key] = run_extra(arm, **kw)
        except Exception as e:                     # a sub-record must never take the main line down
            extra[key] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if only is None or "fma_parity" in only:
        try:
            extra["fma_parity"] = {"note": "performance build (init(validate=False): FMA contraction) vs validation "
                                           "build on the same inputs; north_star tolerance for multi-step fp64 solves",
                                   "cases": [fma_check(arm, "heat3d", (64, 64, 256), 50, 1e-12),
                                             fma_check(arm, "cavity", (256, 256), 5, 1e-12),
                                             fma_check(arm, "diff2d", (512, 2048), 100, 1e-12)]}
        except Exception as e:
            extra["fma_parity"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    arm.configure()
    return extra


def run_ours(args, rank: int, world: int):
    name = args.workload
    spec = WORKLOADS[name]
    shape = tuple(args.shape) if args.shape else spec["shape"]
    K = args.steps if args.steps is not None else spec["steps"]
    Wm = max(3, args.warmup if args.warmup is not None else spec["warmup"])
    arm = Arm(rank, world)
    temporal, validate = not getattr(args, "no_temporal", False), not getattr(args, "fma", False)
    kern = arm.configure(validate=validate, temporal=temporal)[spec["kernel"]]
    if world > 1:
        gshape, inputs, scalars = build_slab_inputs(name, shape, rank, world)
    else:
        gshape = shape
        inputs, scalars = build_inputs(name, shape, seed=rank)

    # ---- device-resident throughput (`value`) ---------------------------------
    grids = arm.grids_for(gshape, shape, inputs)
    ms, launches, clocks, warm = device_leg(arm, kern, grids, scalars, K, Wm)
    rec = record_of(arm, name, shape, K, ms, launches, clocks, warm, temporal)
    rec["config"]["validate_build"] = validate
    if not temporal:
        rec["config"]["temporal"] = False
    del grids

    # ---- end to end through the public API (`e2e`): host buffers -> K steps -> host result ----
    legs = {} if getattr(args, "no_e2e", False) else e2e_legs(arm, name, kern, gshape, shape, inputs, scalars, K)
    del inputs
    arm.rt.trim_pool()
    parity = None if getattr(args, "no_parity", False) else parity_cases(arm)
    select = getattr(args, "extra", "none")
    extra = run_extras(arm, select) if (world == 1 and select != "none") else None
    if world > 1 and select != "none" and 8192 % world == 0:
        # config[3] strong-scaled over the ranks: the cavity's 8192^2 grid in slabs of 8192 / N rows (fused Jacobi
        # pairs on the slab interiors, NCCL exchanges replayed from the recorded graph)
        try:
            cshape = (8192 // world, 8192)
            ck = arm.configure()["cavity_kernel"]
            cg, cin, csc = build_slab_inputs("cavity", cshape, rank, world)
            cgrids = arm.grids_for(cg, cshape, cin)
            cms, cl, cclk, cwarm = device_leg(arm, ck, cgrids, csc, 6, 4)
            crec = record_of(arm, "cavity", cshape, 6, cms, cl, cclk, cwarm)
            crec["config"]["workload"] = f"cavity 8192x8192 fp64 over {world} GPUs ({cshape[0]} rows each), slab-sharded on axis 0"
            crec["scaling"] = "strong"
            extra = {"cavity_8192_sharded": crec}
            del cgrids
        except Exception as e:                     # a sub-record must never take the main line down
            extra = {"cavity_8192_sharded": {"error": f"{type(e).__name__}: {e}"[:300]}}
    if world > 1 and select != "none":
        # config[2] weak-scaled: 16384^2 per GPU, two time steps per pass on the slab interiors
        extra = extra or {}
        for wl in ("diff2d", "conv2d"):
            key = f"{wl}_two_steps_per_pass_sharded"
            try:
                arm.rt.trim_pool()
                wk = arm.configure()[WORKLOADS[wl]["kernel"]]
                wshape = WORKLOADS[wl]["shape"]
                wg, win, wsc = build_slab_inputs(wl, wshape, rank, world)
                wgrids = arm.grids_for(wg, wshape, win)
                del win
                wms, wlaunch, wclk, wwarm = device_leg(arm, wk, wgrids, wsc, 50, 6)
                extra[key] = record_of(arm, wl, wshape, 50, wms, wlaunch, wclk, wwarm)
                del wgrids
            except Exception as e:
                extra[key] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if world > 1:
        from xgrid_b200 import dist as xdist
        xdist.quiesce()                 # no exchange kernel still stores into a neighbour's mailbox
    if rank != 0:
        return None
    line = {"metric": "stencil Gpoint-updates/s", "value": rec["value"], "unit": "Gpoint-updates/s",
            "n_gpus": world, "steps": K, "warmup": warm, "ms_per_step": rec["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "xgrid_b200", "config": rec["config"], "roofline": rec["roofline"],
            "gpu_launches": rec["gpu_launches"], "clocks": rec["clocks"]}
    if "timesteps_per_s" in rec:
        line["timesteps_per_s"] = rec["timesteps_per_s"]
    if world > 1:
        from xgrid_b200 import dist as xdist
        line["halo_transport"] = {"PeerTransport": "peer memory, one kernel per exchange (csrc/xgb_peer.cu)",
                                  "NcclTransport": "ncclSend/ncclRecv"}.get(type(xdist.transport()).__name__, "?")
    line.update(legs)
    if parity is not None:
        line["parity"] = parity
    if extra is not None:
        line["extra"] = extra
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--workload", default="auto", choices=["auto", *WORKLOADS])
    ap.add_argument("--shape", type=int, nargs="*", default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for the baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-temporal", action="store_true", help="one launch per step (init(temporal=False))")
    ap.add_argument("--fma", action="store_true", help="performance build (init(validate=False): FMA contraction)")
    ap.add_argument("--extra", default=None, help='sub-records at N=1: "all", "none" or a comma-separated list of keys '
                                                  '(default: all for --workload auto, none for a named workload)')
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.extra is None:
        args.extra = "all" if args.workload == "auto" else "none"
    if args.workload == "auto":
        args.workload = MAIN            # BASELINE.json configs[4] per GPU at every N (see the module docstring)
    spec = WORKLOADS[args.workload]

    if args.impl == "reference":
        # the reference's CPU implementation of the path: its own generated + compiled kernels
        # (oracle/_ref; the Python front end cannot travel to the GPU box, its output can), else
        # the oracle port; rank 0 only, every host thread, on a bounded sample of our arm's workload.
        if rank != 0:
            return
        import oracle
        shape = tuple(args.shape) if args.shape else spec["shape"]
        sample = cpu_sample_shape(args.workload, shape)
        K = args.steps if args.steps is not None else 10
        Wm = args.warmup if args.warmup is not None else 1
        n_gpus = max(world, args.gpus)
        gpts, sec, steps, sshape, kind, desc = cpu_arm(args.workload, shape, budget_s=1e9, max_steps=max(1, K),
                                                       sample_shape=sample)
        sample_txt = (f"{steps} steps of {args.workload} at {'x'.join(map(str, sshape))}, " + desc +
                      "; a throughput (points/s) of the same per-point work, so it compares with the aggregate of "
                      "any number of GPUs")
        line = {"metric": "stencil Gpoint-updates/s", "value": gpts, "unit": "Gpoint-updates/s",
                "n_gpus": n_gpus, "steps": steps, "warmup": Wm, "ms_per_step": sec * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "impl": "reference",
                "config": config_of(args.workload, shape, n_gpus),
                "cpu_baseline": {"value": gpts, "unit": "Gpoint-updates/s", "cores": oracle.threads(),
                                 "kind": kind, "sample": sample_txt},
                "e2e": {"value": gpts, "unit": "Gpoint-updates/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    line = run_ours(args, rank, world)
    if rank == 0:
        if not args.no_cpu and world == 1:
            import oracle
            shape = tuple(args.shape) if args.shape else spec["shape"]
            gpts, sec, steps, sshape, kind, desc = cpu_arm(args.workload, shape, args.cpu_budget, 200,
                                                           sample_shape=cpu_sample_shape(args.workload, shape))
            line["cpu_baseline"] = {"value": gpts, "unit": "Gpoint-updates/s", "cores": oracle.threads(),
                                    "kind": kind,
                                    "sample": f"{steps} steps of {args.workload} at {'x'.join(map(str, sshape))} "
                                              f"({sec * 1e3:.2f} ms/step), " + desc}
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
