#!/bin/bash
# Round 2 final single-GPU session: smoke, full gpu suite, the driver's two bench commands, launch list of the same command.
O=gpurun_out
mkdir -p $O
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu"; ( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r2z_gputests.log 2>&1; grep -n "passed\|failed" $O/r2z_gputests.log | tail -2
echo "== reference arm"; timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2z_n1_reference.json 2> $O/r2z_n1_reference.err; tail -c 300 $O/r2z_n1_reference.json
echo "== our arm (driver command)"; ( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $O/r2z_n1_bench.json 2> $O/r2z_n1_bench.err; tail -c 300 $O/r2z_n1_bench.json; tail -4 $O/r2z_n1_bench.err
echo "== launch list of the driver command (sub-records off)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_bench_launches.csv \
    python bench.py --gpus 1 --steps 20 --warmup 5 --extra none --no-cpu > $O/r2_bench_launches.out 2>&1
grep -c "xg_" $O/r2_bench_launches.csv
