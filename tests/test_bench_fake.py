"""bench.py's own arm dry-run on the recording fake runtime (CPU): the code path executes for every workload and
the JSON line carries every key of the measurement contract (values are meaningless here -- nothing ran)."""
import argparse
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench
import fake_runtime


class TimedFake(fake_runtime.FakeRuntime):
    def event_elapsed_ms(self, a, b):
        return 1.0


@pytest.mark.parametrize("workload,shape,steps", [
    ("conv1d", [1 << 16], 150), ("diff2d", [256, 2048], 9), ("cavity", [64, 64], 3), ("heat3d", [16, 32, 256], 4),
    ("ewmul", None, 20),
])
def test_bench_line_has_the_contract_keys(monkeypatch, tmp_path, workload, shape, steps):
    from xgrid_b200.runtime import shim
    rt = fake_runtime.install(monkeypatch)
    rt.__class__ = TimedFake
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))          # JIT cache and MEASURED_PEAKS lookup under tmp
    args = argparse.Namespace(workload=workload, shape=shape, steps=steps, warmup=3, no_e2e=False, no_cpu=True,
                              gpus=1, impl="ours", cpu_budget=1.0, no_parity=workload != "heat3d", extra="none")
    line = bench.run_ours(args, 0, 1)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "gpu_launches", "clocks", "e2e", "impl"):
        assert key in line, key
    assert line["steps"] == steps and line["warmup"] >= 3 and line["n_gpus"] == 1 and line["vs_baseline"] is None
    assert line["dtype"] == "f64" and line["data"] == "synthetic" and "workload" in line["config"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"])
    assert line["roofline"]["bound"] == "hbm" and line["roofline"]["unit"] == "GB/s"
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(line["clocks"])
    assert line["gpu_launches"] > 0 and shim.Runtime._instance is rt
    # the warm-up repeats the timed launch plan: whole repetitions of "K calls + flush", at least W steps
    assert line["warmup"] % steps == 0 and line["warmup"] >= 3
    if workload == "heat3d":
        # the parity field's code path runs (its verdict is meaningless on the fake runtime: nothing executes)
        assert {"ok", "ranks", "cases", "checker"} <= set(line["parity"]) and len(line["parity"]["cases"]) >= 5


def test_default_workload_is_the_3d_slab_at_every_n_and_both_arms_print_the_same_config():
    assert bench.MAIN == "heat3d" and bench.WORKLOADS["heat3d"]["shape"] == (256, 2048, 2048)
    for world in (1, 8):
        cfg = bench.config_of("heat3d", (256, 2048, 2048), world)
        assert cfg["workload"].startswith("heat3d 256x2048x2048 fp64") and ("per GPU" in cfg["workload"]) == (world > 1)


def test_sub_records_run_on_the_fake_runtime(monkeypatch, tmp_path):
    rt = fake_runtime.install(monkeypatch)
    rt.__class__ = TimedFake
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    arm = bench.Arm(0, 1)
    one = bench.run_extra(arm, "diff2d", K=8, Wm=3, temporal=False, shape=(256, 2048))
    two = bench.run_extra(arm, "diff2d", K=8, Wm=3, shape=(256, 2048))
    assert one["gpu_launches"] == 8 and two["gpu_launches"] == 4          # two steps per pass when deferred
    assert one["config"]["temporal"] is False and "temporal_blocking" in two["config"]
    fma = bench.run_extra(arm, "conv1d_nl", K=20, Wm=3, validate=False, shape=(1 << 16,))
    assert fma["config"]["validate_build"] is False and fma["gpu_launches"] == 1
    assert {k for k, _ in bench.EXTRA_PLAN} >= {"conv2d_onepass", "cavity_developed_flow", "heat3d_fma_build"}


def test_reference_arm_line(tmp_path):
    """`bench.py --impl reference` (the reference's own CPU kernels from oracle/_ref, else the oracle port) prints the
    contract's line with impl=reference, a cpu_baseline describing the run and an e2e that repeats its value."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "conv1d",
                          "--shape", "65536", "--steps", "3", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "stencil Gpoint-updates/s" and line["steps"] == 3
    assert line["config"] == bench.config_of("conv1d", (65536,), 1)      # the same object our arm prints
    # the default workload (the 3-D slab) under torchrun's environment: OMP_NUM_THREADS=1 must not stick
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", OMP_NUM_THREADS="1")
    out3 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                           "--shape", "8", "64", "256", "--steps", "2", "--warmup", "1"],
                          capture_output=True, text=True, timeout=300, cwd=str(tmp_path), env=env)
    assert out3.returncode == 0, out3.stderr[-2000:]
    l3 = json.loads(out3.stdout.strip().splitlines()[-1])
    assert l3["config"]["workload"].startswith("heat3d 8x64x256 fp64 per GPU") and l3["n_gpus"] == 2
    assert l3["cpu_baseline"]["kind"] == "port" and l3["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    # other ranks of a torchrun launch exit without work
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    quiet = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                           capture_output=True, text=True, timeout=120, cwd=str(tmp_path), env=env)
    assert quiet.returncode == 0 and quiet.stdout.strip() == ""
