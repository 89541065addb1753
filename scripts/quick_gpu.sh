#!/bin/bash
# Short GPU validation of this session's changes (multi-step tail variant, inline C, overstep index helpers,
# reference-output random programs) + two quick bench lines.  Everything is bounded by `timeout`.
mkdir -p gpurun_out
{
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== multistep / inline / overstep"; timeout 150 python -m pytest tests/test_parity_gpu.py tests/test_inline_c.py -q -x -m gpu -k "multistep or inline or overstep" 2>&1 | tail -5
echo "== bench --steps 20"; timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu 2>&1 | tail -1
echo "== random vs reference outputs"; timeout 120 python -m pytest tests/test_random_gpu.py -q -x -m gpu -k "against_reference" 2>&1 | tail -3
echo "== bench default"; timeout 150 python bench.py --no-cpu 2>&1 | tail -1
} > gpurun_out/quick.log 2>&1
tail -30 gpurun_out/quick.log
