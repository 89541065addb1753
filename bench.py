#!/usr/bin/env python
"""bench.py -- headline benchmark of the xgrid stencil hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One JSON line on stdout (rank 0).  A *step* is one kernel call = one timestep
of the workload (one pass of the hot path over the whole grid).

Workloads (BASELINE.json configs):
  conv1d   1-D linear convection, 2^24 points fp64 per GPU (config[1]) -- default at every N
  conv1d_nl / diff1d   the other two config[1] kernels
  conv2d   2-D upwind convection 16384^2 fp64 (config[2]);  diff2d = 5-point variant
  cavity   lid-driven cavity 8192^2 fp64 (config[3]); value in Gpoint-updates/s, also timesteps/s
  heat3d   3-D 7-point, 256x2048x2048 fp64 per GPU, slab-sharded (config[4]); use --workload heat3d
  ewmul    README elementwise_mul, 10 000 points (config[0]; launch-latency bound)

Metric: Gpoint-updates/s = grid points x interior (mask-0) statements per call x steps / time.
Roofline: algorithmic bytes per step / device time per step vs MEASURED_PEAKS.json.
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
    "heat3d":    dict(kernel="heat_3d", ndim=3, shape=(256, 2048, 2048), stmts=1, bytes_pt=16, steps=50, warmup=5),
    "ewmul":     dict(kernel="elementwise_mul", ndim=1, shape=(10000,), stmts=1, bytes_pt=24, steps=2000, warmup=50),
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md).  The sampler is
    started BEFORE the warm-up (nvidia-smi needs a few hundred ms to come up) and every row is time-stamped;
    `summary()` uses the rows that fall inside [mark_start, mark_end], or the nearest ones if the timed
    region was shorter than one sampling period."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int) -> None:
        self.device, self.rows, self.proc = device, [], None
        self.t0 = self.t1 = None

    def __enter__(self):
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
def build_inputs(name: str, shape, seed: int = 0):
    """Host-side synthetic inputs of SURVEY.md §8d: (list of (ic, mask) per grid arg, scalar args)."""
    from xgrid_b200 import workloads as W
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
        rng = np.random.default_rng(seed)
        ic = rng.random(shape)
        return [(ic, W.shell_mask(shape))], (0.1,)
    if name == "cavity":
        n0, n1 = shape
        mb, mp, mu, mv = W.cavity_masks(n0, n1)
        dx, dy = 2.0 / (n1 - 1), 2.0 / (n0 - 1)
        dt = 1e-4 * (100.0 / (n1 - 1)) ** 2
        z = np.zeros(shape)
        if os.environ.get("XGB_BENCH_IC") == "random":
            # developed-flow stand-in (diagnostic only): no field is exactly zero, so every fp64 divide
            # takes its full path instead of the zero-numerator shortcut the quiescent zero IC allows
            rng = np.random.default_rng(seed)
            return [(1e-3 * rng.random(shape), m) for m in (mb, mp, mu, mv)], (W.Config(1.0, 0.1, dt, dx, dy),)
        return [(z, mb), (z, mp), (z, mu), (z, mv)], (W.Config(1.0, 0.1, dt, dx, dy),)
    if name == "ewmul":
        rng = np.random.default_rng(seed)
        n = shape[0]
        z = np.zeros(n, np.int32)
        return [(np.zeros(n), z), (rng.random(n), z), (rng.random(n), z)], ()
    raise SystemExit(f"unknown workload {name}")


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


def cpu_arm(name: str, shape, budget_s: float, max_steps: int, sample_shape=None):
    """Time the reference's CPU implementation of the path, all host threads, on a bounded
    sample of the workload: the reference's OWN compiled kernels (oracle/_ref, built by
    oracle/make_ref.py from the unmodified reference; kind "reference") where they exist and
    are valid (1-D, square 2-D), else the C port (oracle/xgrid_oracle.c, gcc -O3 -fopenmp;
    kind "port").  Returns (Gpt/s, seconds per step, steps, shape, kind, description)."""
    import oracle
    from oracle import ref
    shape = tuple(sample_shape or shape)
    inputs, scalars = build_inputs(name, shape)
    grids = []
    for ic, mask in inputs:
        g = oracle.HostGrid(shape)
        g.now[...] = ic
        g.boundary[...] = mask
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
def run_ours(args, rank: int, world: int):
    import xgrid_b200 as xgrid
    from xgrid_b200 import workloads as W
    from xgrid_b200.runtime.shim import Runtime

    name = args.workload
    spec = WORKLOADS[name]
    shape = tuple(args.shape) if args.shape else spec["shape"]
    K = args.steps if args.steps is not None else spec["steps"]
    Wm = max(3, args.warmup if args.warmup is not None else spec["warmup"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    xgrid.init(precision="double", cacheroot=os.path.join(ROOT, ".xgrid"), device=local_rank,
               distributed=world > 1)
    kern = W.make_kernels()[spec["kernel"]]
    rt = Runtime.get()
    if world > 1:
        # weak scaling: every rank owns a `shape` slab of the (world*shape[0], ...) global grid
        gshape = (shape[0] * world,) + tuple(shape[1:])
        lo, hi = rank * shape[0], (rank + 1) * shape[0]
        if name == "heat3d":
            rng = np.random.default_rng(rank)
            inputs, scalars = [(rng.random(shape), W.shell_mask_slab(gshape, lo, hi))], (0.1,)
        elif name in ("conv1d", "conv1d_nl", "diff1d"):
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
            inputs = [(ic, mask)]
        else:
            raise SystemExit("multi-GPU bench is defined for the slab-sharded workloads: conv1d, conv1d_nl, "
                             "diff1d (axis-0 = the only axis) and heat3d")
    else:
        gshape = shape
        inputs, scalars = build_inputs(name, shape, seed=rank)

    def fresh_grids():
        out = []
        for ic, mask in inputs:
            g = xgrid.Grid(gshape, float)
            assert g.shape == tuple(shape), (g.shape, shape)
            g.now[...] = ic
            g.boundary[...] = mask
            out.append(g)
        return out

    def barrier():
        if world > 1:
            import torch
            import torch.distributed as dist
            rt.device_sync()
            dist.barrier()
            torch.cuda.synchronize()

    points = float(np.prod(shape))
    # ---- device-resident throughput (`value`) ---------------------------------
    grids = fresh_grids()
    with ClockSampler(local_rank) as clocks:
        for _ in range(Wm):
            kern(*grids, *scalars)
        rt.sync()
        ev0, ev1 = rt.event_create(), rt.event_create()
        barrier()
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
    barrier()
    launches = rt.launch_count() - n0
    ms = rt.event_elapsed_ms(ev0, ev1)
    if world > 1:
        import torch
        import torch.distributed as dist
        t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / K
    value = points * world * spec["stmts"] * K / (ms * 1e-3) / 1e9

    # ---- end to end through the public API (`e2e`): host buffers -> K steps -> host result ----
    # The job's inputs start in HOST buffers (the NumPy arrays behind Grid.now / Grid.boundary, filled
    # before the clock starts); the timed region is the K kernel calls -- the first one uploads state and
    # mask (H2D) -- and the read of every grid's newest level on the host (D2H).  Device buffers come from
    # the runtime's caching pool, warm like in any long-running program (the grids of the leg above are
    # dropped first).
    e2e = None
    offload = None
    probe_leg = None
    if not args.no_e2e:
        Ke = K
        del grids
        barrier()
        t0 = time.perf_counter()
        g2 = fresh_grids()                     # host writes of IC + mask (pageable NumPy), not timed
        fill_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            kern(*g2, *scalars)                # first call uploads IC + mask (H2D)
        outs = [g.now for g in g2]             # D2H of every grid's newest level (the rank's slab)
        dt = time.perf_counter() - t0
        if world > 1:
            import torch
            import torch.distributed as dist
            tt = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local_rank}")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        h2d = sum(ic.nbytes + mask.size * 4 for ic, mask in inputs)
        d2h = sum(o.nbytes for o in outs)
        e2e = {"value": points * world * spec["stmts"] * Ke / dt / 1e9, "unit": "Gpoint-updates/s",
               "h2d_bytes_per_step": h2d * world / Ke, "d2h_bytes_per_step": d2h * world / Ke,
               "ms_total": dt * 1e3, "host_fill_ms_not_timed": fill_ms,
               "note": f"public API job: state + mask in host buffers -> {Ke} kernel calls (the first uploads) -> "
                       ".now on host; wall clock, max over ranks; whole-state copies happen once per job, so the "
                       "per-step byte counts are the totals / steps"}
    if world == 1 and not args.no_e2e:
        # literal per-call offload: upload the state, one call, download the state, every step
        Ko = max(3, min(20, K))
        t0 = time.perf_counter()
        for _ in range(Ko):
            for g, o in zip(g2, outs):
                g.now[...] = o                 # host owns the level -> next call uploads it
            kern(*g2, *scalars)
            outs = [g.now for g in g2]
        dt = time.perf_counter() - t0
        offload = {"value": points * spec["stmts"] * Ko / dt / 1e9, "unit": "Gpoint-updates/s",
                   "h2d_bytes_per_step": float(sum(o.nbytes for o in outs)),
                   "d2h_bytes_per_step": float(sum(o.nbytes for o in outs)), "steps": Ko,
                   "note": "every step: H2D full state, one kernel call, D2H full state (page-locked NumPy mirrors)"}
        # strict per-step read-back: every step's result is observed on the host through element
        # indexing (forces one launch + one device->host read per step; no deferral, no batching)
        Kp = max(3, min(1000, K))
        probe = tuple(n // 2 for n in shape)
        t0 = time.perf_counter()
        acc = 0.0
        for _ in range(Kp):
            kern(*g2, *scalars)
            acc += float(g2[0][probe])
        dt = time.perf_counter() - t0
        probe_leg = {"value": points * spec["stmts"] * Kp / dt / 1e9, "unit": "Gpoint-updates/s",
                     "h2d_bytes_per_step": 8.0 * len(scalars), "d2h_bytes_per_step": 8.0, "steps": Kp,
                     "note": "every step: one kernel call, then one element of the result read on the host"}
        del g2, outs

    if rank != 0:
        return None
    peak, peak_src = measured_peaks()
    alg_bytes = points * spec["bytes_pt"]
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(name)
        if t and tuple(shape) == tuple(spec["shape"]):
            traffic = t["bytes_per_launch"]
            traffic_src = f"{t['kernel']}: DRAM read+write per launch ({t['steps_per_launch']} step(s)), {t['source']}"
    except OSError:
        pass
    line = {
        "metric": "stencil Gpoint-updates/s", "value": value, "unit": "Gpoint-updates/s",
        "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "xgrid_b200",
        "config": {"workload": f"{name} {'x'.join(map(str, shape))} fp64" + (" per GPU, slab-sharded on axis 0, NCCL halo exchange" if world > 1 else ""),
                   "kernel": spec["kernel"], "interior_statements_per_step": spec["stmts"],
                   "l2": "working set (2 levels) larger than the 126 MB L2" if points * 16 > 126e6 else "L2-resident (small grid)",
                   "validate_build": True},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_step": alg_bytes,
                     "frac_of_nominal_8TBs": achieved / 8000.0},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
    }
    if launches and launches < K:
        steps_per_launch = K / launches
        line["config"]["temporal_blocking"] = (
            f"runs of identical 1-D calls execute ~{steps_per_launch:.0f} time steps per launch from shared "
            "memory (bit-identical to step-at-a-time)")
        line["roofline"]["note"] = ("achieved = one-pass algorithmic bytes (16 B/pt/step) / time; frac > 1 means the "
                                    "one-pass-per-step HBM roofline is exceeded by temporal blocking "
                                    f"(real HBM traffic ~{32.0 / steps_per_launch:.2f} B/pt/step)")
    if name == "cavity":
        line["timesteps_per_s"] = 1e3 / ms_per_step
    if e2e is not None:
        line["e2e"] = e2e
        if offload is not None:
            line["e2e_offload"] = offload
            line["e2e_probe"] = probe_leg
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
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload == "auto":
        args.workload = "conv1d"        # BASELINE.json configs[1] at every N (sharded on its only axis)
    spec = WORKLOADS[args.workload]

    if args.impl == "reference":
        # the reference's CPU implementation of the path: its own generated + compiled kernels
        # (oracle/_ref; the Python front end cannot travel to the GPU box, its output can), else
        # the oracle port; rank 0 only.
        if rank != 0:
            return
        import oracle
        shape = tuple(args.shape) if args.shape else spec["shape"]
        sample = shape if args.workload != "heat3d" else (64, 2048, 2048)
        K = args.steps if args.steps is not None else 10
        Wm = args.warmup if args.warmup is not None else 1
        gpts, sec, steps, sshape, kind, desc = cpu_arm(args.workload, shape, budget_s=1e9, max_steps=max(1, K),
                                                       sample_shape=sample)
        line = {"metric": "stencil Gpoint-updates/s", "value": gpts, "unit": "Gpoint-updates/s",
                "n_gpus": max(world, args.gpus), "steps": steps, "warmup": Wm, "ms_per_step": sec * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "impl": "reference",
                "config": {"workload": f"{args.workload} {'x'.join(map(str, shape))} fp64",
                           "kernel": spec["kernel"]},
                "cpu_baseline": {"value": gpts, "unit": "Gpoint-updates/s", "cores": oracle.threads(),
                                 "kind": kind,
                                 "sample": f"{steps} steps of {args.workload} at {'x'.join(map(str, sshape))}, "
                                           + desc},
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
            sample = shape if args.workload != "heat3d" else (64, 2048, 2048)
            gpts, sec, steps, sshape, kind, desc = cpu_arm(args.workload, shape, args.cpu_budget, 200,
                                                           sample_shape=sample)
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
