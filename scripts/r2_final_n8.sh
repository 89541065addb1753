#!/bin/bash
# Round 2 final multi-GPU session (N = $1): the driver's commands; recorded-vs-direct comparison of the slab step.
N=${1:-8}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561"
echo "== reference arm N=$N"; timeout 300 $TR bench.py --impl reference --gpus $N --steps 20 --warmup 5 > $O/r2z_n${N}_reference.json 2> $O/r2z_n${N}_reference.err; tail -c 300 $O/r2z_n${N}_reference.json
echo "== our arm N=$N (driver command)"; ( time timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 ) > $O/r2z_n${N}_bench.json 2> $O/r2z_n${N}_bench.err; tail -c 1200 $O/r2z_n${N}_bench.json; tail -4 $O/r2z_n${N}_bench.err
echo "== same with the slab step recorded into a CUDA graph (XGB_SHARDED_GRAPH_MIN=0)"
XGB_SHARDED_GRAPH_MIN=0 timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-parity > $O/r2z_n${N}_bench_recorded.json 2> $O/r2z_n${N}_bench_recorded.err; tail -c 500 $O/r2z_n${N}_bench_recorded.json
