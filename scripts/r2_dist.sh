#!/bin/bash
# Round 2 multi-GPU session: the driver's literal commands at N = $1 (default 2).
N=${1:-2}
O=gpurun_out
mkdir -p $O
free -g | head -2 > $O/r2d_n${N}_host.txt; nproc >> $O/r2d_n${N}_host.txt; nvidia-smi -L >> $O/r2d_n${N}_host.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
echo "== reference arm, N=$N"
( time timeout 600 $TR bench.py --impl reference --gpus $N --steps 20 --warmup 5 ) > $O/r2d_n${N}_reference.json 2> $O/r2d_n${N}_reference.err
tail -c 700 $O/r2d_n${N}_reference.json; tail -3 $O/r2d_n${N}_reference.err
echo "== our arm (driver command), N=$N"
( time timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 ) > $O/r2d_n${N}_bench.json 2> $O/r2d_n${N}_bench.err
tail -c 2500 $O/r2d_n${N}_bench.json; tail -4 $O/r2d_n${N}_bench.err
echo "== conv1d sharded, 20 steps (round-1 collapse case), N=$N"
( time timeout 600 $TR bench.py --gpus $N --workload conv1d --steps 20 --warmup 5 ) > $O/r2d_n${N}_conv1d20.json 2> $O/r2d_n${N}_conv1d20.err
tail -c 1800 $O/r2d_n${N}_conv1d20.json; tail -4 $O/r2d_n${N}_conv1d20.err
echo "== cavity 8192^2 (config[3]) sharded over $N GPUs: $((8192 / N)) rows per GPU"
( time timeout 600 $TR bench.py --gpus $N --workload cavity --shape $((8192 / N)) 8192 --steps 6 --warmup 4 --no-e2e ) > $O/r2d_n${N}_cavity.json 2> $O/r2d_n${N}_cavity.err
tail -c 1500 $O/r2d_n${N}_cavity.json; tail -4 $O/r2d_n${N}_cavity.err
if [ "$N" = "2" ]; then
  echo "== N=1 for the efficiency denominator"
  timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --extra none --no-cpu > $O/r2d_n1_bench.json 2> $O/r2d_n1_bench.err; tail -c 600 $O/r2d_n1_bench.json
  echo "== tests/test_dist.py on 2 GPUs"
  timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q 2>&1 | tail -5
fi
