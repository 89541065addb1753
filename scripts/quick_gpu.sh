#!/bin/bash
# Short GPU validation (bounded by `timeout`): mask upload / host-mirror / pool changes + two bench lines.
mkdir -p gpurun_out
{
echo "== bench default"; timeout 150 python bench.py --no-cpu 2>&1 | tail -1
echo "== bench --steps 20"; timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu 2>&1 | tail -1
echo "== edge cases + parity (no full sizes) + reference programs"; timeout 120 python -m pytest tests/test_edge_cases_gpu.py tests/test_parity_gpu.py tests/test_reference_programs_gpu.py tests/test_inline_c.py -q -x -m gpu -k "not full_size and not 2p24" 2>&1 | tail -4
echo "== jacobi2 + random"; timeout 150 python -m pytest tests/test_jacobi2.py tests/test_random_gpu.py -q -x -m gpu 2>&1 | tail -4
} > gpurun_out/quick3.log 2>&1
grep -v '^{' gpurun_out/quick3.log | tail -30
