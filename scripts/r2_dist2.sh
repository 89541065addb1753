#!/bin/bash
# N=2: sharded CUDA-graph replay (tests/test_dist.py: heat3d, multi-step 1-D, cavity x7 with recorded + replayed calls,
# overstep modes), callee test, sharded cavity speed with and without graphs.
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
echo "== callee test (1 GPU)"; timeout 300 python -m pytest tests/test_callee_gpu.py -m gpu -q 2>&1 | tail -3
echo "== tests/test_dist.py on 2 GPUs"; timeout 300 python -m pytest tests/test_dist.py -m gpu -x -q 2>&1 | tail -5
echo "== driver command N=2"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2f_n2_bench.json 2> $O/r2f_n2_bench.err; tail -c 900 $O/r2f_n2_bench.json; tail -2 $O/r2f_n2_bench.err
echo "== sharded cavity 4096x8192 per GPU, graphs on"
timeout 300 $TR bench.py --gpus 2 --workload cavity --shape 4096 8192 --steps 6 --warmup 4 --no-e2e --no-parity > $O/r2f_n2_cavity.json 2> $O/r2f_n2_cavity.err; tail -c 700 $O/r2f_n2_cavity.json; tail -2 $O/r2f_n2_cavity.err
echo "== same, graphs off"
XGB_BENCH_GRAPHS=0 timeout 300 $TR bench.py --gpus 2 --workload cavity --shape 4096 8192 --steps 6 --warmup 4 --no-e2e --no-parity > $O/r2f_n2_cavity_nograph.json 2> $O/r2f_n2_cavity_nograph.err; tail -c 700 $O/r2f_n2_cavity_nograph.json; tail -2 $O/r2f_n2_cavity_nograph.err
echo "== single GPU cavity 4096x8192 for comparison"
timeout 300 python bench.py --workload cavity --shape 4096 8192 --steps 6 --warmup 4 --no-e2e --no-parity --no-cpu | tail -c 600
