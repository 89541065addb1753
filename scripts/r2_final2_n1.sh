#!/bin/bash
# Round 2, last single-GPU validation after the slab / peer-transport work: smoke, full gpu suite, driver bench command.
O=gpurun_out
mkdir -p $O
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu"; ( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/r2y_gputests.log 2>&1; grep -n "passed\|failed\|error" $O/r2y_gputests.log | tail -3
echo "== our arm (driver command)"; ( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $O/r2y_n1_bench.json 2> $O/r2y_n1_bench.err; python scripts/results_table.py $O/r2y_n1_bench.json | head -24; tail -4 $O/r2y_n1_bench.err
