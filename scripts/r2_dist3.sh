#!/bin/bash
# Round 2, third multi-GPU session: two-step passes and diagonal-tap overhang on slabs (tests/dist_worker.py), then
# the driver's literal command at N = $1 (default 2) with the new parity case and the sharded 2-D sub-records.
N=${1:-2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
echo "== tests/test_dist.py on $N GPUs"
timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q 2>&1 | tail -15
echo "== our arm (driver command), N=$N"
( time timeout 1200 $TR bench.py --gpus $N --steps 20 --warmup 5 ) > $O/r2e_n${N}_bench.json 2> $O/r2e_n${N}_bench.err
tail -c 6000 $O/r2e_n${N}_bench.json; tail -4 $O/r2e_n${N}_bench.err
