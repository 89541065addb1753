"""Multi-rank path: partition / planning logic on CPU over gloo (world_size 2), and the
sharded CUDA path over NCCL on >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

from xgrid_b200 import dist as xdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_range_partitions_exactly():
    for n0 in (1, 7, 8, 23, 256, 2048):
        for world in (1, 2, 3, 4, 8):
            spans = [xdist.slab_range(n0, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n0
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_topology_chain():
    t = xdist.Topology(0, 4)
    assert (t.lo_rank, t.hi_rank) == (-1, 1)
    t = xdist.Topology(3, 4)
    assert (t.lo_rank, t.hi_rank) == (2, -1)
    assert not xdist.Topology(0, 1).sharded


def test_topology_ring_for_wrap():
    """overstep="wrap": the slabs form a ring, so rank 0's lower ghost rows are the last rows of the grid."""
    assert [(xdist.Topology(r, 4, ring=True).lo_rank, xdist.Topology(r, 4, ring=True).hi_rank) for r in range(4)] \
        == [(3, 1), (0, 2), (1, 3), (2, 0)]
    two = xdist.Topology(1, 2, ring=True)
    assert (two.lo_rank, two.hi_rank) == (0, 0)              # both neighbours are the same peer
    one = xdist.Topology(0, 1, ring=True)
    assert not one.ring and (one.lo_rank, one.hi_rank) == (-1, -1)      # a single rank wraps locally


def _run(mode, nproc, tmp_path, port, **extra_env):
    env = dict(os.environ, XG_CACHE=str(tmp_path / "xg"), OMP_NUM_THREADS="2", **extra_env)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "dist_worker.py"), mode]
    return subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600, cwd=str(tmp_path))


def test_two_ranks_gloo_cpu(tmp_path):
    r = _run("cpu", 2, tmp_path, 29611)
    assert "DIST_CPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
@pytest.mark.parametrize("halo, transport", [("peer", "PeerTransport"), ("nccl", "NcclTransport")])
def test_two_ranks_nccl_gpu(tmp_path, halo, transport):
    """Every sharded case of the worker, bit for bit against the single-domain oracle, once per halo transport:
    the one-kernel exchange over peer memory (the default) and ncclSend/ncclRecv."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = _run("gpu", 2, tmp_path, 29612 if halo == "peer" else 29613, XGB_HALO=halo, XGB_PEER_TIMEOUT_S="30")
    assert f"DIST_GPU_OK transport={transport}" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
