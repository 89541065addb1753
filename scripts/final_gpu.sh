#!/bin/bash
# Round-end style validation on one B200 (run under gpurun): full GPU test suite, smoke, every bench
# workload, the reference arm, launch lists and ncu captures of the dominant kernels.
# Outputs land in gpurun_out/ with the prefix given as $1.
P=${1:-r1b}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15 > $O/${P}_gputests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${P}_smoke.log 2>&1
timeout 300 python bench.py > $O/${P}_bench_default.json 2> $O/${P}_bench_default.err
timeout 200 python bench.py --impl reference --steps 10 --warmup 1 > $O/${P}_bench_reference.json 2>/dev/null
: > $O/${P}_bench_all.jsonl
cat $O/${P}_bench_default.json >> $O/${P}_bench_all.jsonl
for w in conv1d_nl diff1d conv2d diff2d heat3d cavity ewmul; do
  timeout 300 python bench.py --workload $w --cpu-budget 4 >> $O/${P}_bench_all.jsonl 2> $O/${P}_err_$w.log
done
# host-path phases of the default job, with the driver's pageable copies and with the staged path
timeout 100 python scripts/e2e_phases.py conv1d 10000 > $O/${P}_e2e_phases.log 2>&1
XGB_STAGED_COPY=1 timeout 100 python scripts/e2e_phases.py conv1d 10000 > $O/${P}_e2e_phases_staged.log 2>&1
# launch lists (per-launch device time under ncu: cold cache, serialised -- compare shares)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/${P}_bench_launches.csv \
    python bench.py --steps 2000 --warmup 200 --no-cpu --no-e2e > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv --log-file $O/${P}_cavity_launches.csv \
    python bench.py --workload cavity --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
# full captures of the dominant kernels
timeout 200 ncu --set full --clock-control none --import-source on -k regex:jacobi2 -s 30 -c 1 -o $O/${P}_cavity_jacobi2 \
    python bench.py --workload cavity --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:tiled2 -s 4 -c 1 -o $O/${P}_diff2d_tiled2 \
    python bench.py --workload diff2d --steps 10 --warmup 10 --no-cpu --no-e2e > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:multistep -s 3 -c 1 -o $O/${P}_conv1d_nl_multistep \
    python bench.py --workload conv1d_nl --steps 1000 --warmup 200 --no-cpu --no-e2e > /dev/null 2>&1
tail -3 $O/${P}_gputests.log; cat $O/${P}_smoke.log | tail -1
python - "$P" <<'PY'
import json,sys
for l in open("gpurun_out/%s_bench_all.jsonl" % (sys.argv[1] if len(sys.argv)>1 else "r1b")):
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); print(d["config"]["workload"], round(d["value"],1), round(d["ms_per_step"],5), round(d["roofline"]["frac"],3), d.get("e2e",{}).get("value"), d.get("cpu_baseline",{}).get("kind"), d["clocks"]["reasons"])
PY
