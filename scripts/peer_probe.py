"""Halo exchange alone (no sweep) on a heat3d slab: device time per blocking exchange of one level, per transport.
torchrun --nproc-per-node N scripts/peer_probe.py   (XGB_HALO=peer|nccl, XGB_PEER_CTAS=...)"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xgrid_b200 as xgrid                          # noqa: E402
from xgrid_b200 import dist as xdist                # noqa: E402
from xgrid_b200.runtime import shim                 # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    xgrid.init(precision="double", distributed=True, device=local)
    rt = shim.Runtime.get()
    for shape in ((64 * world, 2048, 2048), (256 * world, 8192), (64 * world, 256, 256)):
        u = xgrid.Grid(shape, float)
        u.now[...] = 1.0
        u._prepare_device(1)
        lv = u._ring[0]
        tr = xdist.transport()
        e0, e1 = rt.event_create(), rt.event_create()
        for reps in (5, 50):
            dist.barrier()
            rt.device_sync()
            rt.event_record_raw(e0, 0)
            for _ in range(reps):
                tr.exchange([(u, lv, 1)], 0)
            rt.event_record_raw(e1, 0)
            rt.device_sync()
            ms = rt.event_elapsed_ms(e0, e1) / reps
        face = int(np.prod(shape[1:])) * 8
        if rank == 0:
            print(f"{type(tr).__name__} ctas={os.environ.get('XGB_PEER_CTAS', '-')} face {face >> 10} KiB: "
                  f"{ms * 1e3:.1f} us per exchange = {face / ms / 1e6:.1f} GB/s per direction", flush=True)
        del u
    xdist.quiesce()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
