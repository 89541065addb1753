#!/bin/bash
# Round 2, GPU session 1 (1 GPU): full gpu test suite, the default bench line (both arms), a short 1-D run.
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/r2a_gpu.txt 2>&1
nproc >> $O/r2a_gpu.txt; free -g >> $O/r2a_gpu.txt
echo "== pytest -m gpu"; ( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/r2a_gputests.log 2>&1; tail -5 $O/r2a_gputests.log
echo "== bench default"; ( time timeout 900 python bench.py ) > $O/r2a_bench_default.json 2> $O/r2a_bench_default.err; tail -c 600 $O/r2a_bench_default.json; tail -5 $O/r2a_bench_default.err
echo "== bench reference"; ( time timeout 300 python bench.py --impl reference --steps 20 --warmup 5 ) > $O/r2a_bench_reference.json 2> $O/r2a_bench_reference.err; tail -c 400 $O/r2a_bench_reference.json; tail -4 $O/r2a_bench_reference.err
echo "== bench conv1d 20"; timeout 300 python bench.py --workload conv1d --steps 20 --warmup 5 --no-cpu > $O/r2a_bench_conv1d20.json 2> $O/r2a_bench_conv1d20.err; tail -c 1500 $O/r2a_bench_conv1d20.json; tail -3 $O/r2a_bench_conv1d20.err
